"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A functional (stateless) restatement of the reference's 3D mobile CNN forward passes, written
against ``torch.nn.functional`` so that it can run on the GPU box where ``/root/reference`` does
not exist.  Gradients come from ``torch.autograd`` over these functions.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import this file.

Parity pinning: the reference ships no tests, golden vectors or usable checkpoints
(finding 1).  This oracle is pinned instead against outputs of the reference
modules themselves, executed in the build container by ``tests/golden/make_golden.py`` and
committed as ``tests/golden/*.pt`` (logits, loss, per-parameter gradient digests, updated BN
running statistics).  ``tests/test_oracle_golden.py`` replays them.

The arithmetic itself lives in PyTorch (third party, unpinned by the reference; this image has
torch 2.11.0+cu128): conv3d, batch_norm, adaptive_avg_pool3d, hardswish, hardsigmoid, relu,
leaky_relu, dropout3d, linear.

Each function cites the reference lines it follows.  State dicts use the reference's key layout
(SURVEY.md appendix C).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# (in, out, expanded, stride, use_se, k, act, p_drop) -- mobilenet.py:147-176 (Large)
LARGE_BLOCKS: Dict[str, List[Tuple[int, int, int, int, bool, int, str, float]]] = {
    "block2": [(16, 16, 16, 1, False, 3, "relu", 0.2), (16, 24, 64, 2, False, 3, "relu", 0.2),
               (24, 24, 72, 1, False, 3, "relu", 0.2)],
    "block3": [(24, 40, 72, 2, True, 5, "relu", 0.2), (40, 40, 120, 1, True, 5, "relu", 0.2),
               (40, 40, 120, 1, True, 5, "relu", 0.2)],
    "block4": [(40, 80, 240, 2, False, 3, "hswish", 0.2), (80, 80, 240, 1, False, 3, "hswish", 0.2),
               (80, 80, 184, 1, False, 3, "hswish", 0.2), (80, 80, 184, 1, False, 3, "hswish", 0.2),
               (80, 112, 480, 1, True, 3, "hswish", 0.2), (112, 112, 672, 1, True, 3, "hswish", 0.2)],
    "block5": [(112, 160, 672, 2, True, 5, "hswish", 0.2), (160, 160, 960, 1, True, 5, "hswish", 0.2),
               (160, 160, 960, 1, True, 5, "hswish", 0.2)],
}
# mobilenet.py:227-242 (Small)
SMALL_BLOCKS: Dict[str, List[Tuple[int, int, int, int, bool, int, str, float]]] = {
    "block2": [(16, 16, 16, 2, True, 3, "lrelu", 0.2), (16, 24, 72, 2, False, 3, "lrelu", 0.2),
               (24, 24, 88, 1, False, 3, "lrelu", 0.2)],
    "block3": [(24, 40, 96, 2, True, 5, "hswish", 0.2), (40, 40, 240, 1, True, 5, "hswish", 0.2),
               (40, 40, 240, 1, True, 5, "hswish", 0.2), (40, 48, 120, 1, True, 5, "hswish", 0.2),
               (48, 48, 144, 1, True, 5, "hswish", 0.2), (48, 96, 288, 2, True, 5, "hswish", 0.2),
               (96, 96, 576, 1, True, 5, "hswish", 0.2), (96, 96, 576, 1, True, 5, "hswish", 0.2)],
}
# (in, out, expanded, (kT,kH,kW), (sT,sH,sW), (pT,pH,pW)) -- movinet.py:98-137
MOVINET_A2_BLOCKS: Dict[str, List[Tuple[int, int, int, Tuple[int, int, int], Tuple[int, int, int], Tuple[int, int, int]]]] = {
    "block2": [(16, 16, 40, (1, 5, 5), (1, 2, 2), (0, 2, 2)), (16, 16, 40, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (16, 16, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))],
    "block3": [(16, 40, 96, (3, 3, 3), (1, 2, 2), (1, 1, 1)), (40, 40, 120, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (40, 40, 96, (3, 3, 3), (1, 1, 1), (1, 1, 1)), (40, 40, 96, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (40, 40, 120, (3, 3, 3), (1, 1, 1), (1, 1, 1))],
    "block4": [(40, 72, 240, (5, 3, 3), (1, 2, 2), (2, 1, 1)), (72, 72, 160, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (72, 72, 240, (3, 3, 3), (1, 1, 1), (1, 1, 1)), (72, 72, 192, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (72, 72, 240, (3, 3, 3), (1, 1, 1), (1, 1, 1))],
    "block5": [(72, 72, 240, (5, 3, 3), (1, 1, 1), (2, 1, 1)), (72, 72, 240, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (72, 72, 240, (3, 3, 3), (1, 1, 1), (1, 1, 1)), (72, 72, 240, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (72, 72, 144, (1, 5, 5), (1, 1, 1), (0, 2, 2)), (72, 72, 240, (3, 3, 3), (1, 1, 1), (1, 1, 1))],
    "block6": [(72, 144, 480, (5, 3, 3), (1, 2, 2), (2, 1, 1)), (144, 144, 384, (1, 5, 5), (1, 1, 1), (0, 2, 2)),
               (144, 144, 384, (1, 5, 5), (1, 1, 1), (0, 2, 2)), (144, 144, 480, (1, 5, 5), (1, 1, 1), (0, 2, 2)),
               (144, 144, 480, (1, 5, 5), (1, 1, 1), (0, 2, 2)), (144, 144, 480, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
               (144, 144, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1))],
}


def _act(x: torch.Tensor, kind: str) -> torch.Tensor:
    if kind == "relu":
        return F.relu(x)
    if kind == "hswish":
        return F.hardswish(x)
    if kind == "lrelu":
        return F.leaky_relu(x, 0.01)   # nn.LeakyReLU() default slope, mobilenet.py:228
    if kind == "none":
        return x
    raise ValueError(kind)


def _bn(sd: SD, prefix: str, x: torch.Tensor, train: bool, momentum: Optional[float] = 0.1) -> torch.Tensor:
    """nn.BatchNorm3d / BatchNorm1d semantics (eps 1e-5, momentum 0.1, running stats updated in place
    when training, num_batches_tracked incremented)."""
    rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    if train:
        nbt = sd.get(prefix + "num_batches_tracked")
        if nbt is not None:
            nbt += 1
        if momentum is None:  # cumulative moving average (used by the calibration script)
            momentum = 1.0 / float(nbt) if nbt is not None else 0.1
    return F.batch_norm(x, rm, rv, sd[prefix + "weight"], sd[prefix + "bias"], train,
                        momentum if momentum is not None else 0.1, 1e-5)


def _drop3d(x: torch.Tensor, p: float, train: bool, masks: Optional[list]) -> torch.Tensor:
    """nn.Dropout3d (mobilenet.py:82,92).  With ``masks`` (a list consumed front to back) the
    noise is injected instead of drawn, so tests can share it with the CUDA path."""
    if not train or p == 0.0:
        return x
    if masks is not None:
        m = masks.pop(0)
        return x * m.to(x.dtype).view(x.shape[0], x.shape[1], *([1] * (x.dim() - 2)))
    return F.dropout3d(x, p, True)


def se_block3d(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """SEBlock3D.forward, mobilenet.py:11-26 (prefix ends with 'se.')."""
    w = F.adaptive_avg_pool3d(x, 1)
    w = F.relu(F.conv3d(w, sd[prefix + "1.weight"], sd[prefix + "1.bias"]))
    w = F.hardsigmoid(F.conv3d(w, sd[prefix + "3.weight"], sd[prefix + "3.bias"]))
    return x * w


def bottleneck3d(sd: SD, prefix: str, x: torch.Tensor, stride: int, use_se: bool, k: int, act: str,
                 p_drop: float, train: bool, masks: Optional[list] = None,
                 momentum: Optional[float] = 0.1) -> torch.Tensor:
    """Bottleneck3D.forward, mobilenet.py:84-93.  Note the scalar stride/padding on a (1,k,k)
    kernel (mobilenet.py:67-75): time is padded by k//2 on both sides and strided too."""
    x = F.conv3d(x, sd[prefix + "pointwise_conv1.weight"])
    c = x.shape[1]
    x = F.conv3d(x, sd[prefix + "depthwise_conv.weight"], None, stride, k // 2, 1, c)
    if use_se:
        x = se_block3d(sd, prefix + "squeeze_excite.se.", x)
    x = F.conv3d(x, sd[prefix + "pointwise_conv2.weight"])
    x = _bn(sd, prefix + "batchnorm.", x, train, momentum)
    x = _act(x, act)
    return _drop3d(x, p_drop, train, masks)


def _mobilenet_stem(sd: SD, x: torch.Tensor, train: bool, momentum) -> torch.Tensor:
    """block1: Conv3d(3,16,k3,s2,p1)+bias -> BN -> Hardswish, mobilenet.py:140-144 / 220-224."""
    x = F.conv3d(x, sd["block1.0.weight"], sd["block1.0.bias"], 2, 1)
    x = _bn(sd, "block1.1.", x, train, momentum)
    return F.hardswish(x)


def mobilenet_large_tail(sd: SD, x: torch.Tensor, train: bool, momentum: Optional[float] = 0.1) -> torch.Tensor:
    """block6 + classifier of MobileNetLarge3D (mobilenet.py:178-190, 199-200) on a block5 output."""
    x = F.conv3d(x, sd["block6.0.weight"], sd["block6.0.bias"])
    x = F.hardswish(_bn(sd, "block6.1.", x, train, momentum))
    x = F.adaptive_avg_pool3d(x, 1)
    x = F.hardswish(F.conv3d(x, sd["classifier.1.weight"], sd["classifier.1.bias"]))
    x = F.conv3d(x, sd["classifier.3.weight"], sd["classifier.3.bias"])
    return x.view(x.shape[0], -1)


def mobilenet_large3d(sd: SD, x: torch.Tensor, train: bool = False, masks: Optional[list] = None,
                      momentum: Optional[float] = 0.1, taps: Optional[dict] = None) -> torch.Tensor:
    """MobileNetLarge3D.forward, mobilenet.py:192-201.  ``taps`` (optional dict) receives the
    output of every block for layer-by-layer comparisons."""
    x = _mobilenet_stem(sd, x, train, momentum)
    if taps is not None:
        taps["block1"] = x
    for blk, rows in LARGE_BLOCKS.items():
        for i, (_, _, _, s, se, k, act, p) in enumerate(rows):
            x = bottleneck3d(sd, f"{blk}.{i}.", x, s, se, k, act, p, train, masks, momentum)
            if taps is not None:
                taps[f"{blk}.{i}"] = x
    x = F.conv3d(x, sd["block6.0.weight"], sd["block6.0.bias"])          # mobilenet.py:178-182
    x = F.hardswish(_bn(sd, "block6.1.", x, train, momentum))
    x = F.adaptive_avg_pool3d(x, 1)                                       # mobilenet.py:185-190
    x = F.hardswish(F.conv3d(x, sd["classifier.1.weight"], sd["classifier.1.bias"]))
    x = F.conv3d(x, sd["classifier.3.weight"], sd["classifier.3.bias"])
    return x.view(x.shape[0], -1)


def mobilenet_small3d(sd: SD, x: torch.Tensor, train: bool = False, masks: Optional[list] = None,
                      momentum: Optional[float] = 0.1, taps: Optional[dict] = None) -> torch.Tensor:
    """MobileNetSmall3D.forward, mobilenet.py:258-265; block4 has SE between conv and BN
    (mobilenet.py:244-249)."""
    x = _mobilenet_stem(sd, x, train, momentum)
    if taps is not None:
        taps["block1"] = x
    for blk, rows in SMALL_BLOCKS.items():
        for i, (_, _, _, s, se, k, act, p) in enumerate(rows):
            x = bottleneck3d(sd, f"{blk}.{i}.", x, s, se, k, act, p, train, masks, momentum)
            if taps is not None:
                taps[f"{blk}.{i}"] = x
    x = F.conv3d(x, sd["block4.0.weight"], sd["block4.0.bias"])
    x = se_block3d(sd, "block4.1.se.", x)
    x = F.hardswish(_bn(sd, "block4.2.", x, train, momentum))
    x = F.adaptive_avg_pool3d(x, 1)                                       # mobilenet.py:251-256
    x = F.hardswish(F.conv3d(x, sd["classifier.1.weight"], sd["classifier.1.bias"]))
    x = F.conv3d(x, sd["classifier.3.weight"], sd["classifier.3.bias"])
    return x.view(x.shape[0], -1)


def movinet_bottleneck(sd: SD, prefix: str, x: torch.Tensor, kernel, stride, padding, train: bool,
                       momentum: Optional[float] = 0.1) -> torch.Tensor:
    """MoviNetBottleneck.forward, movinet.py:69-77: expand -> depthwise (kT,kH,kW) -> SE -> project
    -> BN -> Hardswish.  Its Dropout3d member is constructed but never applied."""
    x = F.conv3d(x, sd[prefix + "expand.weight"])
    x = F.conv3d(x, sd[prefix + "conv.weight"], None, stride, padding, 1, x.shape[1])
    x = se_block3d(sd, prefix + "squeeze_excite.se.", x)
    x = F.conv3d(x, sd[prefix + "project.weight"])
    return F.hardswish(_bn(sd, prefix + "batchnorm.", x, train, momentum))


def movinet_a2(sd: SD, x: torch.Tensor, train: bool = False, masks: Optional[list] = None,
               momentum: Optional[float] = 0.1, taps: Optional[dict] = None) -> torch.Tensor:
    """MoViNetA2.forward, movinet.py:156-165 (symmetric temporal padding: NOT causal).
    ``masks`` = [Dropout3d mask (B,640), Dropout mask (B,2048)] when injected."""
    x = F.conv3d(x, sd["block1.0.weight"], None, (1, 2, 2), (0, 1, 1))   # movinet.py:91-95
    x = F.hardswish(_bn(sd, "block1.1.", x, train, momentum))
    for blk, rows in MOVINET_A2_BLOCKS.items():
        for i, (_, _, _, k, s, p) in enumerate(rows):
            x = movinet_bottleneck(sd, f"{blk}.{i}.", x, k, s, p, train, momentum)
            if taps is not None:
                taps[f"{blk}.{i}"] = x
    x = F.conv3d(x, sd["conv.0.weight"])                                  # movinet.py:139-144
    x = F.hardswish(_bn(sd, "conv.1.", x, train, momentum))
    x = _drop3d(x, 0.2, train, masks)
    x = F.adaptive_avg_pool3d(x, 1).flatten(1)                            # movinet.py:146-154
    x = F.linear(x, sd["classifier.2.weight"], sd["classifier.2.bias"])
    x = F.hardswish(_bn(sd, "classifier.3.", x, train, momentum))
    if train:
        if masks is not None:
            x = x * masks.pop(0).to(x.dtype)
        else:
            x = F.dropout(x, 0.2, True)
    return F.linear(x, sd["classifier.6.weight"], sd["classifier.6.bias"])


def causal_conv3d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride=1,
                  fill: float = 0.0, groups: int = 1, padding=0, dilation=1) -> torch.Tensor:
    """CausalConv3d.forward, movinet.py:34-39.  The constructor (movinet.py:23-28) sets the left pad to
    p_left + p_right, which is kT-1 for odd and for even kT alike; the pad value is the scalar
    ``stream_buffer`` (default 0).  ``padding``/``dilation`` are the **kwargs handed to nn.Conv3d."""
    kt = weight.shape[2]
    x = F.pad(x, (0, 0, 0, 0, kt - 1, 0), "constant", fill)
    return F.conv3d(x, weight, bias, stride, padding, dilation, groups)


def movinet_a2_stream(sd: SD, chunks: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Causal, chunked MoViNetA2 inference -- PARITY UNPINNED BY THE REFERENCE (movinet.py never wires
    CausalConv3d or its buffers into MoViNetA2; SURVEY.md finding 5).  This restates the specification the
    build adopted (picklebot_b200/movinet.py docstring) with reference ops so the CUDA path has a checker:
      * every depthwise conv is a CausalConv3d (movinet.py:23-39): left pad kT-1, and across chunks the pad
        frames are the previous chunk's last kT-1 *input* frames (zeros before the first chunk);
      * squeeze-excite means and the classifier pool are cumulative over all frames seen so far;
      * BatchNorm in eval mode, dropouts off.
    Returns the logits after each chunk."""
    blocks_ = [(f"{blk}.{i}.", row) for blk, rows in MOVINET_A2_BLOCKS.items() for i, row in enumerate(rows)]
    state = [dict() for _ in blocks_]
    head_sum, head_n, outs = None, 0, []
    for x in chunks:
        x = F.conv3d(x, sd["block1.0.weight"], None, (1, 2, 2), (0, 1, 1))
        x = F.hardswish(_bn(sd, "block1.1.", x, False))
        for (prefix, (_, _, _, k, s, p)), st in zip(blocks_, state):
            x = F.conv3d(x, sd[prefix + "expand.weight"])
            if k[0] > 1:
                prev = st.get("buf")
                if prev is None:
                    prev = x.new_zeros(x.shape[0], x.shape[1], k[0] - 1, x.shape[3], x.shape[4])
                xin = torch.cat([prev, x], 2)
                st["buf"] = xin[:, :, xin.shape[2] - (k[0] - 1):].clone()
            else:
                xin = x
            x = F.conv3d(xin, sd[prefix + "conv.weight"], None, s, (0, p[1], p[2]), 1, x.shape[1])
            n = x.shape[2] * x.shape[3] * x.shape[4]
            ssum = x.sum((2, 3, 4))
            st["sum"] = ssum if "sum" not in st else st["sum"] + ssum
            st["n"] = st.get("n", 0) + n
            w = (st["sum"] / st["n"]).view(x.shape[0], -1, 1, 1, 1)
            se = prefix + "squeeze_excite.se."
            w = F.relu(F.conv3d(w, sd[se + "1.weight"], sd[se + "1.bias"]))
            w = F.hardsigmoid(F.conv3d(w, sd[se + "3.weight"], sd[se + "3.bias"]))
            x = F.conv3d(x * w, sd[prefix + "project.weight"])
            x = F.hardswish(_bn(sd, prefix + "batchnorm.", x, False))
        x = F.hardswish(_bn(sd, "conv.1.", F.conv3d(x, sd["conv.0.weight"]), False))
        n = x.shape[2] * x.shape[3] * x.shape[4]
        hsum = x.sum((2, 3, 4))
        head_sum = hsum if head_sum is None else head_sum + hsum
        head_n += n
        f = head_sum / head_n
        f = F.linear(f, sd["classifier.2.weight"], sd["classifier.2.bias"])
        f = F.hardswish(_bn(sd, "classifier.3.", f, False))
        outs.append(F.linear(f, sd["classifier.6.weight"], sd["classifier.6.bias"]))
    return outs


MODELS = {
    "MobileNetLarge3D": mobilenet_large3d,
    "MobileNetSmall3D": mobilenet_small3d,
    "MoViNetA2": movinet_a2,
}


def clone_state(sd: SD, requires_grad: bool = False, device=None, dtype=None) -> SD:
    """Deep copy of a state dict; float tensors optionally become autograd leaves."""
    out: SD = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if device is not None:
            t = t.to(device)
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            if requires_grad and not k.endswith(("running_mean", "running_var")):
                t.requires_grad_(True)
        out[k] = t
    return out


def train_step(model: str, sd: SD, x: torch.Tensor, labels: torch.Tensor,
               masks: Optional[list] = None) -> Tuple[torch.Tensor, torch.Tensor, SD]:
    """One forward + CrossEntropyLoss + backward (train.py:264-269 without the GradScaler, whose
    power-of-two scale is exact).  Returns (logits, loss, grads by key); ``sd`` must hold leaves
    from ``clone_state(..., requires_grad=True)``; running stats are updated in place."""
    logits = MODELS[model](sd, x, True, masks)
    loss = F.cross_entropy(logits.float(), labels)
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads))
