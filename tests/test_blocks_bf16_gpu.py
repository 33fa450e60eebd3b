"""bf16 parity of the PRODUCTION kernels (tcgen05 GEMMs, TMA-tiled depthwise, fused BN epilogues), block by block.

Teacher forcing: one fp32 pass of the oracle over MobileNetLarge3D on 16 full-size clips (3x16x224x224, train
mode, fixed Dropout3d masks) yields, for every block, its input activation and the gradient arriving at its
output.  Both are rounded to bf16 once; then every block is run in isolation
  * by the fp32 oracle on those rounded tensors (torch ops on the GPU, TF32 off) -- the truth,
  * by picklebot_b200 under autocast(bf16) -- the production path (asserted through the path counters),
  * by the oracle ops under autocast(bf16) -- what the reference's own GPU path does (train.py:264-269), printed
    for context only,
Bars (norm-wise relative error against the fp32 truth; measured values: profiles/r02_parity.md):
  * block OUTPUT: 1e-2, absolute -- the north_star's bf16 tolerance, met by every block;
  * input gradient and parameter gradients: bf16 STORAGE of the intermediate gradients alone puts the reference's
    own autocast path at 1-6e-2 per block (train-mode BatchNorm backward subtracts the two dominant components of
    the upstream gradient, which amplifies the rounding of what is left; squeeze-excite parameter gradients are sums
    with heavy cancellation: 3-25e-2).  1e-2 is therefore not a property any bf16 implementation of these blocks
    has, the reference's included.  The bar here is: within 1e-2, or no further from the truth than
    2 x the reference's own bf16 path measured in the same test on the same tensors -- and in any case under the
    explicit ceilings GRAD_CEIL (a hard stop that does not depend on torch).  Squeeze-excite parameters get the
    ceiling alone: ReLU / Hardsigmoid have kinks, a unit that sits on one flips with a 1e-3 change of the pooled mean
    and moves the whole (tiny) gradient -- ours and torch's path land on different sides at random (block3.1: both
    0.12; block3.2: 0.24 vs 0.03; block5.2: 0.008 vs 0.047).
Errors cannot compound across blocks here, so a miss is the block's own arithmetic.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from _util import rel_err, synthetic_checkpoint
from oracle import picklebot_oracle as O
from picklebot_b200 import _lib, synth

pytestmark = pytest.mark.gpu

TOL = 1e-2
# hard ceilings for gradients, whatever torch's bf16 path does (measured maxima over the 15 blocks, round 2:
# dx 5.4e-2, conv weights 5.4e-2, BN affine 2.2e-2, squeeze-excite FC parameters 2.4e-1)
GRAD_CEIL = {"dx": 8e-2, "conv": 8e-2, "batchnorm": 4e-2, "squeeze_excite": 3e-1}
MODEL = "MobileNetLarge3D"


def _grad_bar(name, torch_err):
    kind = "dx" if name == "dx" else "squeeze_excite" if "squeeze_excite" in name else \
        "batchnorm" if "batchnorm" in name or name.endswith((".1.weight", ".1.bias")) else "conv"
    if kind == "squeeze_excite":
        return GRAD_CEIL[kind]
    return max(TOL, min(2.0 * torch_err, GRAD_CEIL[kind]))
B = 16
CLIP = (16, 224, 224)


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _masks(batch, seed=synth.SEED_DROPOUT):
    g = torch.Generator().manual_seed(seed)
    return [(torch.empty(batch, row[1]).bernoulli_(0.8, generator=g) / 0.8).cuda()
            for rows in O.LARGE_BLOCKS.values() for row in rows]


@pytest.fixture(scope="module")
def teacher():
    """fp32 oracle pass: per-block inputs and upstream gradients (bf16-rounded), masks, checkpoint."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = O.clone_state(synthetic_checkpoint(MODEL), requires_grad=True, device="cuda")
    clips = synth.synthetic_clips_u8(B, *CLIP).cuda()
    x = synth.clips_to_features(clips, torch.float32)
    labels = synth.synthetic_labels(B, 2).cuda()
    masks = _masks(B)
    taps = {}
    logits = O.mobilenet_large3d(sd, x, True, [m.clone() for m in masks], taps=taps)
    for t in taps.values():
        t.retain_grad()
    F.cross_entropy(logits, labels).backward()
    names = ["block1"] + [f"{blk}.{i}" for blk, rows in O.LARGE_BLOCKS.items() for i in range(len(rows))]
    acts = {n: taps[n].detach().to(torch.bfloat16) for n in names}
    # the loss gradient is tiny (1/B scale); bring every upstream gradient to unit RMS, as GradScaler would
    grads = {}
    for n in names:
        g = taps[n].grad.detach()
        grads[n] = (g / g.float().pow(2).mean().sqrt().clamp_min(1e-30)).to(torch.bfloat16)
    del taps, logits
    torch.cuda.empty_cache()
    return {"clips": clips, "labels": labels, "masks": masks, "acts": acts, "grads": grads, "names": names}


def _ours():
    import picklebot_b200 as pb
    m = pb.MobileNetLarge3D(num_classes=2)
    m.load_state_dict(synthetic_checkpoint(MODEL))
    return m.cuda().train()


def _report(tag, rows):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "blocks_bf16_parity.jsonl"), "a") as f:
        f.write(json.dumps({"block": tag, "errors": rows}) + "\n")


BLOCKS = [(blk, i) for blk, rows in O.LARGE_BLOCKS.items() for i in range(len(rows))]


@pytest.mark.parametrize("blk,i", BLOCKS, ids=[f"{b}.{i}" for b, i in BLOCKS])
def test_bottleneck_bf16_teacher_forced(teacher, blk, i):
    names = teacher["names"]
    name = f"{blk}.{i}"
    pos = names.index(name)
    x_b = teacher["acts"][names[pos - 1]]                 # (B,C,T,H,W) bf16, channels-last-3d or not: any strides
    g_b = teacher["grads"][name]
    mask = teacher["masks"][pos - 1]
    row = O.LARGE_BLOCKS[blk][i]
    _, _, _, stride, use_se, k, act, p_drop = row
    prefix = name + "."
    ck = synthetic_checkpoint(MODEL)

    def oracle(autocast):
        sd = O.clone_state({kk[len(prefix):]: v for kk, v in ck.items() if kk.startswith(prefix)},
                           requires_grad=True, device="cuda")
        xin = (x_b.clone() if autocast else x_b.float()).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = O.bottleneck3d(sd, "", xin, stride, use_se, k, act, p_drop, True, [mask.clone()])
        out.backward(g_b.to(out.dtype))
        grads = {kk: v.grad.detach().float() for kk, v in sd.items() if v.requires_grad}
        return out.detach().float(), xin.grad.detach().float(), grads

    t_out, t_dx, t_g = oracle(False)
    r_out, r_dx, r_g = oracle(True)

    m = _ours()
    mod = getattr(m, blk)[i]
    _lib.path_reset()
    xin = x_b.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = mod(xin, mask.clone())
    out.backward(g_b)
    torch.cuda.synchronize()
    paths = _lib.path_counts()
    # the production kernels served this block: TMA depthwise in all three directions, tcgen05 GEMMs and wgrads
    assert paths["dw_fwd_generic"] == 0 and paths["dw_dgrad_generic"] == 0 and paths["dw_wgrad_generic"] == 0, paths
    assert paths["dw_fwd_tma"] == 1 and paths["dw_dgrad_tma"] + paths["dw_bwd_fused_tma"] >= 1, paths
    assert paths["gemm_simt"] == 0 and paths["wgrad_simt"] == 0, paths
    assert paths["gemm_tc"] >= 4 and paths["wgrad_tc"] == 2, paths

    mine_g = {kk: p.grad.detach().float() for kk, p in mod.named_parameters()}
    assert set(mine_g) == set(t_g)
    rows = {"out": (rel_err(out.float(), t_out), rel_err(r_out, t_out)),
            "dx": (rel_err(xin.grad.float(), t_dx), rel_err(r_dx, t_dx))}
    gscale = max(float(v.norm()) for v in t_g.values())
    for kk, tv in t_g.items():
        den = max(float(tv.norm()), 1e-3 * gscale)
        rows[kk] = (float((mine_g[kk] - tv).norm()) / den, float((r_g[kk] - tv).norm()) / den)
    _report(name, rows)
    print(f"\n{name}: " + ", ".join(f"{kk} {a:.1e}/{b:.1e}" for kk, (a, b) in rows.items()) + "   (ours/torch-autocast)")
    assert rows["out"][0] < TOL, rows["out"]
    bad = {kk: v for kk, v in rows.items() if kk != "out" and not v[0] < _grad_bar(kk, v[1])}
    assert not bad, f"{name}: {bad}"


def test_stem_bf16_teacher_forced(teacher):
    ck = synthetic_checkpoint(MODEL)
    clips = teacher["clips"]
    g_b = teacher["grads"]["block1"]
    x_b = synth.clips_to_features(clips, torch.bfloat16)            # train.py:106

    def oracle(autocast):
        sd = O.clone_state({k: v for k, v in ck.items() if k.startswith("block1.")}, requires_grad=True, device="cuda")
        xin = x_b if autocast else x_b.float()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = O._mobilenet_stem(sd, xin, True, 0.1)
        out.backward(g_b.to(out.dtype))
        return out.detach().float(), {k: v.grad.detach().float() for k, v in sd.items() if v.requires_grad}

    t_out, t_g = oracle(False)
    r_out, r_g = oracle(True)
    m = _ours()
    _lib.path_reset()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m._stem(clips.permute(0, 4, 1, 2, 3), torch.bfloat16)            # raw uint8 clip, /255 fused
    out.backward(g_b)
    torch.cuda.synchronize()
    paths = _lib.path_counts()
    assert paths["stem_tc"] == 2 and paths["stem_simt"] == 0, paths
    rows = {"out": (rel_err(out.float(), t_out), rel_err(r_out, t_out))}
    mine = {k: p.grad.detach().float() for k, p in m.named_parameters() if p.grad is not None}
    gscale = max(float(v.norm()) for v in t_g.values())
    for k, tv in t_g.items():
        if k == "block1.0.bias":       # a bias in front of a train-mode BatchNorm: its true gradient is exactly zero
            assert float(mine[k].norm()) < 2e-2 * gscale
            continue
        den = max(float(tv.norm()), 1e-3 * gscale)
        rows[k] = (float((mine[k] - tv).norm()) / den, float((r_g[k] - tv).norm()) / den)
    _report("block1", rows)
    print("\nblock1: " + ", ".join(f"{k} {a:.1e}/{b:.1e}" for k, (a, b) in rows.items()))
    assert rows["out"][0] < TOL, rows["out"]
    bad = {k: v for k, v in rows.items() if k != "out" and not v[0] < _grad_bar(k, v[1])}
    assert not bad, bad


def test_tail_bf16_teacher_forced(teacher):
    ck = synthetic_checkpoint(MODEL)
    x_b = teacher["acts"]["block5.2"]
    labels = teacher["labels"]
    keys = ("block6.", "classifier.")

    def oracle(autocast):
        sd = O.clone_state({k: v for k, v in ck.items() if k.startswith(keys)}, requires_grad=True, device="cuda")
        xin = (x_b.clone() if autocast else x_b.float()).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits = O.mobilenet_large_tail(sd, xin, True)
        F.cross_entropy(logits.float(), labels).backward()
        return (logits.detach().float(), xin.grad.detach().float(),
                {k: v.grad.detach().float() for k, v in sd.items() if v.requires_grad})

    t_out, t_dx, t_g = oracle(False)
    r_out, r_dx, r_g = oracle(True)
    m = _ours()
    from picklebot_b200.blocks import MobileNetTailFn
    from picklebot_b200.mobilenet import _bn_args
    conv, se, bn, fc1, fc2 = m._tail_modules()
    eps, mom, rm, rv, nbt = _bn_args(bn)
    xin = x_b.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = MobileNetTailFn.apply(xin, m._cache, True, eps, mom, False, rm, rv, nbt, conv.weight, conv.bias,
                                       bn.weight, bn.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias,
                                       None, None, None, None)
    F.cross_entropy(logits.float(), labels).backward()
    rows = {"logits": (rel_err(logits.float(), t_out), rel_err(r_out, t_out)),
            "dx": (rel_err(xin.grad.float(), t_dx), rel_err(r_dx, t_dx))}
    mine = {k: p.grad.detach().float() for k, p in m.named_parameters() if p.grad is not None}
    gscale = max(float(v.norm()) for v in t_g.values())
    for k, tv in t_g.items():
        if k == "block6.0.bias":       # a bias in front of a train-mode BatchNorm: its true gradient is exactly zero
            assert float(mine[k].norm()) < 2e-2 * gscale
            continue
        den = max(float(tv.norm()), 1e-3 * gscale)
        rows[k] = (float((mine[k] - tv).norm()) / den, float((r_g[k] - tv).norm()) / den)
    _report("tail", rows)
    print("\ntail: " + ", ".join(f"{k} {a:.1e}/{b:.1e}" for k, (a, b) in rows.items()))
    assert rows["logits"][0] < TOL, rows["logits"]
    bad = {k: v for k, v in rows.items() if k != "logits" and not v[0] < _grad_bar(k, v[1])}
    assert not bad, bad
