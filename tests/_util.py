"""Shared helpers for the tests: golden fixtures, synthetic checkpoints, error metrics."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from picklebot_b200 import synth  # noqa: E402

MODEL_NAMES = ("MobileNetLarge3D", "MobileNetSmall3D", "MoViNetA2")


def statedict_template(model: str):
    with open(os.path.join(GOLDEN, "statedict_keys.json")) as f:
        keys = json.load(f)[model]
    out = {}
    for k, shape in keys.items():
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.int64)
        else:
            out[k] = torch.zeros(shape, dtype=torch.float32)
    return out


def synthetic_checkpoint(model: str):
    """The synthetic stand-in for weights/<model>.pth: seeded weights + calibrated BN stats."""
    stats = torch.load(os.path.join(GOLDEN, f"{model}_bnstats.pt"))
    return synth.synthetic_state_dict(statedict_template(model), bn_stats=stats)


def golden(model: str):
    return torch.load(os.path.join(GOLDEN, f"{model}_golden.pt"))


def features(shape, seed=synth.SEED_DATA, dtype=torch.float32, device="cpu", channels_last=False):
    B, T, H, W = shape
    clips = synth.synthetic_clips_u8(B, T, H, W, seed=seed).to(device)
    x = synth.clips_to_features(clips, dtype)
    return x if channels_last else x.contiguous()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Norm-wise relative error ||a-b|| / ||b|| (many logits are ~1e-2)."""
    if a.is_cuda and b.is_cuda:              # large activations: reduce on the device, in fp64
        a, b = a.detach().double(), b.detach().double()
    else:
        a = a.detach().double().cpu()
        b = b.detach().double().cpu()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def hash_name(name: str) -> int:
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def probe_vector(n: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1


def grad_digest(name: str, g: torch.Tensor):
    flat = g.detach().double().flatten().cpu()
    pv = probe_vector(flat.numel(), (hash_name(name) % 100000) + 11)
    return float(flat.norm()), float((flat * pv).sum()), flat[:8].float().clone()
