"""Synthetic pitch clips and synthetic checkpoints (the reference's weights/*.pth are absent).

Everything here is bit-reproducible on any host: clips are built from integer arithmetic on
``torch.randint`` draws (mt19937, platform independent) and weights from ``torch.rand`` draws
(exact 24-bit mantissa conversion), so the committed golden logits in ``tests/golden`` stay
valid on the GPU box.

Clip layout follows the reference data path: ``dataloader.py:10-23`` yields uint8
``(B, T, H, W, C)`` batches and ``train.py:102-108`` turns them into ``(B, C, T, H, W)``
channels-last-3d views scaled by 1/255.

Why not default init: with constructor-default weights activations collapse to ~0 within a
few blocks (SURVEY.md finding 2), which makes every parity check vacuous.  ``synthetic_state_dict``
draws variance-preserving uniform weights instead; BatchNorm running statistics come from a
calibration fixture (``tests/golden/*_bnstats.pt``) produced by running the reference once.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Tuple

import torch

SEED_DATA = 0
SEED_LABELS = 1
SEED_WEIGHTS = 1234  # train.py:196 uses manual_seed(1234)
SEED_DROPOUT = 7


def synthetic_clips_u8(batch: int, frames: int = 16, height: int = 224, width: int = 224,
                       seed: int = SEED_DATA, device: str | torch.device = "cpu") -> torch.Tensor:
    """uint8 ``(B, T, H, W, 3)`` clips: flat background + moving bump + 5 % noise.

    Structured (not iid) on purpose: global pooling erases iid noise, so logits of iid clips
    are identical for every sample and argmax parity would be meaningless.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, T, H, W = batch, frames, height, width
    bg = torch.randint(30, 200, (B, 1, 1, 1, 3), generator=g)
    fg = torch.randint(0, 256, (B, 1, 1, 1, 3), generator=g)
    cx0 = torch.randint(0, W, (B, 1, 1, 1), generator=g)
    cy0 = torch.randint(0, H, (B, 1, 1, 1), generator=g)
    vx = torch.randint(-6, 7, (B, 1, 1, 1), generator=g)
    vy = torch.randint(-6, 7, (B, 1, 1, 1), generator=g)
    rad = torch.randint(max(4, H // 28), max(6, H // 9), (B, 1, 1, 1), generator=g)
    t = torch.arange(T).view(1, T, 1, 1)
    ys = torch.arange(H).view(1, 1, H, 1)
    xs = torch.arange(W).view(1, 1, 1, W)
    cx = cx0 + vx * t
    cy = cy0 + vy * t
    d2 = (xs - cx) ** 2 + (ys - cy) ** 2                      # (B,T,H,W) int64
    r2 = rad * rad
    bump = torch.clamp(r2 - d2, min=0) * 256 // r2            # 0..256, integer falloff
    bump = bump.unsqueeze(-1)
    img = bg + (fg - bg) * bump // 256
    noise = torch.randint(0, 13, (B, T, H, W, 3), generator=g)
    img = torch.clamp(img + noise - 6, 0, 255).to(torch.uint8)
    return img.to(device)


def synthetic_clips_u8_device(batch: int, frames: int, height: int, width: int, seed: int,
                               device: str | torch.device) -> torch.Tensor:
    """Same clip family as ``synthetic_clips_u8`` but built with int32 math on ``device`` (bench-sized
    batches: 64 clips are 154 MB of uint8).  Per-clip parameters still come from the CPU generator; only
    the pixel noise uses the device generator, so these clips are not bit-identical to the CPU ones."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, T, H, W = batch, frames, height, width
    i32 = dict(dtype=torch.int32, device=device)
    bg = torch.randint(30, 200, (B, 1, 1, 1, 3), generator=g).to(**i32)
    fg = torch.randint(0, 256, (B, 1, 1, 1, 3), generator=g).to(**i32)
    cx0 = torch.randint(0, W, (B, 1, 1, 1), generator=g).to(**i32)
    cy0 = torch.randint(0, H, (B, 1, 1, 1), generator=g).to(**i32)
    vx = torch.randint(-6, 7, (B, 1, 1, 1), generator=g).to(**i32)
    vy = torch.randint(-6, 7, (B, 1, 1, 1), generator=g).to(**i32)
    rad = torch.randint(max(4, H // 28), max(6, H // 9), (B, 1, 1, 1), generator=g).to(**i32)
    t = torch.arange(T, **i32).view(1, T, 1, 1)
    ys = torch.arange(H, **i32).view(1, 1, H, 1)
    xs = torch.arange(W, **i32).view(1, 1, 1, W)
    d2 = (xs - (cx0 + vx * t)) ** 2 + (ys - (cy0 + vy * t)) ** 2
    r2 = rad * rad
    bump = (torch.clamp(r2 - d2, min=0) * 256 // r2).unsqueeze(-1)
    img = bg + (fg - bg) * bump // 256
    gd = torch.Generator(device=device).manual_seed(seed)
    noise = torch.randint(0, 13, (B, T, H, W, 3), generator=gd, dtype=torch.int32, device=device)
    return torch.clamp(img + noise - 6, 0, 255).to(torch.uint8)


def synthetic_task_clips_u8(batch: int, frames: int = 8, height: int = 64, width: int = 64,
                            seed: int = SEED_DATA) -> Tuple[torch.Tensor, torch.Tensor]:
    """A separable two-class toy task in the clip family above (the argmax check needs a
    checkpoint whose decisions are not all the same class): class 1 = a warm (red-dominant) bump moving to the
    right, class 0 = a cold (blue-dominant) bump moving to the left, over a random grey background with 5 %
    noise.  Returns (uint8 (B,T,H,W,3), int64 labels (B,)).  Integer arithmetic only: bit-reproducible."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, T, H, W = batch, frames, height, width
    labels = torch.randint(0, 2, (B,), generator=g)
    lab = labels.view(B, 1, 1, 1)
    grey = torch.randint(60, 160, (B, 1, 1, 1, 1), generator=g)
    bg = grey + torch.randint(-12, 13, (B, 1, 1, 1, 3), generator=g)
    strong = torch.randint(170, 256, (B, 1, 1, 1), generator=g)
    weak = torch.randint(0, 90, (B, 1, 1, 1), generator=g)
    mid = torch.randint(40, 140, (B, 1, 1, 1), generator=g)
    red = torch.where(lab == 1, strong, weak)
    blue = torch.where(lab == 1, weak, strong)
    fg = torch.stack([red, mid, blue], dim=-1)                       # (B,1,1,1,3)
    speed = torch.randint(2, max(3, W // 10), (B, 1, 1, 1), generator=g)
    vx = torch.where(lab == 1, speed, -speed)
    vy = torch.randint(-2, 3, (B, 1, 1, 1), generator=g)
    cx0 = torch.randint(W // 4, 3 * W // 4, (B, 1, 1, 1), generator=g)
    cy0 = torch.randint(H // 4, 3 * H // 4, (B, 1, 1, 1), generator=g)
    rad = torch.randint(max(4, H // 8), max(6, H // 4), (B, 1, 1, 1), generator=g)
    t = torch.arange(T).view(1, T, 1, 1) - T // 2
    ys = torch.arange(H).view(1, 1, H, 1)
    xs = torch.arange(W).view(1, 1, 1, W)
    d2 = (xs - (cx0 + vx * t)) ** 2 + (ys - (cy0 + vy * t)) ** 2
    r2 = rad * rad
    bump = (torch.clamp(r2 - d2, min=0) * 256 // r2).unsqueeze(-1)
    img = bg + (fg - bg) * bump // 256
    noise = torch.randint(0, 13, (B, T, H, W, 3), generator=g)
    return torch.clamp(img + noise - 6, 0, 255).to(torch.uint8), labels


def clips_to_features(clips_u8: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """The reference's ``extract_features_labels`` (train.py:102-108): a ``(B,C,T,H,W)`` view with
    channels-last-3d strides, cast and divided by 255 (same op order, so bf16 values match)."""
    return clips_u8.permute(0, 4, 1, 2, 3).to(dtype) / 255


def synthetic_labels(batch: int, num_classes: int = 2, seed: int = SEED_LABELS,
                     device: str | torch.device = "cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, num_classes, (batch,), generator=g).to(device)


def _uniform(shape: Iterable[int], bound: float, g: torch.Generator) -> torch.Tensor:
    return (torch.rand(tuple(shape), generator=g) * 2.0 - 1.0) * bound


def synthetic_state_dict(template: Dict[str, torch.Tensor], seed: int = SEED_WEIGHTS,
                         bn_stats: Dict[str, torch.Tensor] | None = None) -> Dict[str, torch.Tensor]:
    """Fill a state_dict with the same keys/shapes/dtypes as ``template`` (any model of this
    family) with variance-preserving weights.  Keys are visited in sorted order so that the
    result does not depend on module registration order.

    * conv / linear weights: U(-a, a) with a = gain*sqrt(3/fan_in); gain 1 for the linear
      pw1 -> dw -> pw2 chain, 2 for squeeze-excite FCs (so the gate leaves 0.5), 3 for the
      classifier head (spreads the logits across samples so argmax parity is not vacuous)
    * biases: U(-0.1, 0.1); BN weight U(0.6, 1.4); BN bias U(-0.4, 0.4)
    * BN running stats: from ``bn_stats`` when given, else mean 0 / var 1
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    bn_prefixes = {k[: -len("running_mean")] for k in template if k.endswith("running_mean")}
    for key in sorted(template):
        ref = template[key]
        prefix = key[: key.rfind(".") + 1]
        leaf = key[key.rfind(".") + 1:]
        if leaf == "num_batches_tracked":
            val = torch.zeros((), dtype=torch.int64)
        elif prefix in bn_prefixes:
            if leaf == "weight":
                val = 1.0 + _uniform(ref.shape, 0.4, g)
            elif leaf == "bias":
                val = _uniform(ref.shape, 0.4, g)
            elif leaf == "running_mean":
                val = torch.zeros(ref.shape)
            else:
                val = torch.ones(ref.shape)
            if bn_stats is not None and key in bn_stats and leaf.startswith("running"):
                val = bn_stats[key].clone().to(torch.float32)
        elif leaf == "bias":
            val = _uniform(ref.shape, 0.1, g)
        else:
            fan_in = 1
            for d in ref.shape[1:]:
                fan_in *= int(d)
            gain = 2.0 if (".se." in key) else (3.0 if key.startswith("classifier.") else 1.0)
            val = _uniform(ref.shape, gain * math.sqrt(3.0 / fan_in), g)
        out[key] = val.to(ref.dtype) if leaf != "num_batches_tracked" else val
    return out


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> Tuple[int, float]:
    """(entry count, sum of |x|) -- cheap identity check used by the golden fixtures."""
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return len(sd), tot
