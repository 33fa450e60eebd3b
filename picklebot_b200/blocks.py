"""Fused forward/backward of the building blocks, as torch.autograd.Functions over the CUDA kernels.

One autograd node per bottleneck (not per op): the block owns its intermediate tensors, so what is
saved for backward and how ops are fused is decided here, and DDP's bucketed all-reduce still overlaps
with the backward of earlier blocks because parameter gradients become ready block by block.

Tensors crossing a block boundary are ordinary ``(B, C, T, H, W)`` torch tensors with channels-last-3d
strides (a zero-copy permute of the NDHWC buffer the kernels work on), so the modules stay drop-in.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import os

import torch

from . import ops
from . import gemm_tc          # noqa: F401  (import enables the tcgen05 GEMM path)
from .ops import ACT_CODES


# ---------------------------------------------------------------------------------------------
# layout helpers
# ---------------------------------------------------------------------------------------------
def to_ndhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,T,H,W) any strides -> contiguous (B,T,H,W,C).  Zero-copy when x is channels-last-3d."""
    y = x.permute(0, 2, 3, 4, 1)
    return y if y.is_contiguous() else y.contiguous()


def from_ndhwc(y: torch.Tensor) -> torch.Tensor:
    return y.permute(0, 4, 1, 2, 3)


def compute_dtype(x: torch.Tensor) -> torch.dtype:
    """Activation storage type: the autocast dtype when autocast is on (train.py:264), else x's own."""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        if dt != torch.bfloat16:
            raise RuntimeError(f"picklebot_b200 supports bf16 autocast only, got {dt}")
        return dt
    if x.dtype in (torch.float32, torch.bfloat16):
        return x.dtype
    if x.dtype == torch.uint8:
        return torch.bfloat16
    raise TypeError(f"unsupported input dtype {x.dtype}")


class WeightCache:
    """Repacked / cast shadow copies of parameters, refreshed when the parameter changes
    (optimizer steps bump ``_version``; ``.to()`` / load_state_dict change data_ptr or version)."""

    def __init__(self):
        self._store = {}

    def get(self, key, param: torch.Tensor, maker):
        tag = (param.data_ptr(), param._version, param.device)
        hit = self._store.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        val = maker()
        self._store[key] = (tag, val)
        return val

    def clear(self):
        self._store.clear()


@dataclass(frozen=True)
class BlockCfg:
    k: Tuple[int, int, int]
    s: Tuple[int, int, int]
    p: Tuple[int, int, int]
    act: int
    slope: float
    use_se: bool
    p_drop: float
    eps: float
    momentum: float


def draw_dropout3d_mask(B: int, C: int, p: float, dtype: torch.dtype, device) -> torch.Tensor:
    """Same generator calls as aten::feature_dropout behind nn.Dropout3d (mobilenet.py:82,92): noise of
    shape (B,C,1,1,1) in the activation dtype, bernoulli_(1-p) then div_(1-p).  Returned as fp32 [B][C]."""
    noise = torch.empty((B, C, 1, 1, 1), dtype=dtype, device=device).bernoulli_(1 - p).div_(1 - p)
    return noise.view(B, C).float()


# ---------------------------------------------------------------------------------------------
# GEMM front-ends (pick tcgen05 or CUDA-core kernel)
# ---------------------------------------------------------------------------------------------
def _w2d(w: torch.Tensor) -> torch.Tensor:
    return w.detach().reshape(w.shape[0], -1)


def _row_fold(rows: int, K: int) -> int:
    """Row-fold factor for a GEMM whose reduction dimension K is 16..32 channels wide.  A TMA box row costs the
    same whether it carries 32 or 128 bytes, so X[rows][K] is streamed as X'[rows/F][F*K] against the block-diagonal
    weight diag(W, ..., W); the output C'[rows/F][F*N] is C[rows][N] byte for byte."""
    F = 4 if K <= 16 else 2 if K <= 32 else 1
    while F > 1 and rows % F:
        F //= 2
    return F


def pw_fwd(A: torch.Tensor, w: torch.Tensor, cache: WeightCache, key: str, bias=None, gate=None, Bt: int = 1,
           want_stats: bool = False, fuse_stats: bool = True):
    """[rows][K] x W[N][K]^T (+bias); gate [Bt][K] scales A's columns per sample (squeeze-excite).
    want_stats: returns (C, sums) where sums are the BatchNorm column sums of C if the tensor-core epilogue could
    accumulate them (else None and the caller runs the statistics pass)."""
    W = _w2d(w)
    N, K = W.shape
    C = sums = None
    if ops.use_tc(A.dtype, K, N):
        fuse = want_stats and fuse_stats and bias is None and N <= 256
        if gate is not None:
            Wb = ops.fold_gate(W, gate)
            C = gemm_tc.gemm(A, Wb, N, K, Bw=Bt, Bt=Bt, bias=bias, stat_mod=N if fuse else 0)
        else:
            Wb = cache.get((key, "bf16"), w, lambda: ops.cast_matrix(W, N, K, torch.bfloat16))
            F = _row_fold(A.numel() // K, K) if bias is None else 1
            if F > 1 and (not fuse or N * F <= 256):
                Wf = cache.get((key, "bf16", F), w, lambda: ops.block_diag(Wb, F))
                C = gemm_tc.gemm(A, Wf, N * F, K * F, stat_mod=N if fuse else 0)
            else:
                C = gemm_tc.gemm(A, Wb, N, K, Bw=1, Bt=1, bias=bias, stat_mod=N if fuse else 0)
        if fuse:
            C, sums = C
        C = C.view(-1, N)
    else:
        C = ops.gemm_simt(A, W, N, K, K, 1, bias=bias, ascale=gate, Bt=Bt)
    return (C, sums) if want_stats else C


def pw_dgrad(dC: torch.Tensor, w: torch.Tensor, cache: WeightCache, key: str) -> torch.Tensor:
    """dA[rows][K] = dC[rows][N] x W[N][K]."""
    W = _w2d(w)
    N, K = W.shape
    if ops.use_tc(dC.dtype, N, K):
        Wt = cache.get((key, "bf16_t"), w, lambda: ops.cast_matrix(W, N, K, torch.bfloat16, transpose=True))
        F = _row_fold(dC.numel() // N, N)
        if F > 1:
            Wf = cache.get((key, "bf16_t", F), w, lambda: ops.block_diag(Wt, F))
            return gemm_tc.gemm(dC, Wf, K * F, N * F).view(-1, K)
        return gemm_tc.gemm(dC, Wt, K, N, Bw=1, Bt=1)
    return ops.gemm_simt(dC, W, K, N, 1, K)


def pw_wgrad(A: torch.Tensor, dC: torch.Tensor, w: torch.Tensor, gate=None, Bt: int = 1, want_bias: bool = False):
    W = _w2d(w)
    N, K = W.shape
    if ops.use_tc(A.dtype, K, N) and gemm_tc.wgrad_ready():
        dW, _ = gemm_tc.wgrad(A, dC, K, N, gate=gate, Bt=Bt)
        db = ops.colstats(dC, N)[0].float() if want_bias else None
        return dW.view(w.shape), db
    dW, db = ops.wgrad_simt(A, dC, K, N, ascale=gate, Bt=Bt, want_bias=want_bias)
    return dW.view(w.shape), db


def se_pw2_backward(dz: torch.Tensor, y2: torch.Tensor, w2: torch.Tensor, cache: WeightCache, gate, pooled, hidden,
                    se_w1, se_w2, B: int, R_out: int):
    """Backward of  z = (y2 * gate) W2^T  with gate = SE(mean(y2)):  returns (dW2, dy2, SE parameter grads).
    tcgen05 path: the per-sample products P_b = dz_b^T y2_b give dW2 AND dgate, and the input-gradient GEMM
    applies gate / adds dmean in its epilogue, so y2 and g are each touched once."""
    W = _w2d(w2)
    Cout, Cexp = W.shape
    if ops.use_tc(dz.dtype, Cexp, Cout) and gemm_tc.wgrad_ready():
        dW2, dgate = gemm_tc.wgrad(y2.view(-1, Cexp), dz, Cexp, Cout, gate=gate, W=W.contiguous(), Bt=B,
                                   want_dgate=True)
        dmean, dW1s, db1s, dW2s, db2s = ops.se_fc_bwd(dgate, pooled, hidden, gate, _w2d(se_w1), _w2d(se_w2),
                                                      1.0 / float(R_out))
        # dy2 = (dz W2) * gate + dmean: the gate goes into per-sample weights (rows of W2^T scaled), dmean pre-loads
        # the accumulator, so the GEMM runs its plain bf16 epilogue (the fp32 epilogue-vector path is write-starved:
        # 2.0 TB/s of stores against 3.7 for the plain one on the 112 -> 672 layer)
        Wt32 = cache.get(("w2", "f32_t"), w2, lambda: ops.cast_matrix(W, Cout, Cexp, torch.float32, transpose=True))
        Wtg = ops.fold_rows(Wt32, gate)
        dy2 = gemm_tc.gemm(dz, Wtg, Cexp, Cout, Bw=B, Bt=B, coladd=dmean)
    else:
        dW2, _ = ops.wgrad_simt(y2.view(-1, Cexp), dz, Cexp, Cout, ascale=gate, Bt=B)
        g = ops.gemm_simt(dz, W, Cexp, Cout, 1, Cexp)
        dgate = ops.rowdot(g, y2, B, Cexp)
        dmean, dW1s, db1s, dW2s, db2s = ops.se_fc_bwd(dgate, pooled, hidden, gate, _w2d(se_w1), _w2d(se_w2),
                                                      1.0 / float(R_out))
        dy2 = ops.scale_add_(g, gate, dmean, B, Cexp)
    return dW2.view(w2.shape), dy2, (dW1s.view(se_w1.shape), db1s, dW2s.view(se_w2.shape), db2s)


# ---------------------------------------------------------------------------------------------
# BatchNorm(+act+dropout) helper shared by every block
# ---------------------------------------------------------------------------------------------
def bn_forward(z: torch.Tensor, B: int, C: int, gamma, beta, rmean, rvar, nbt, training: bool, eps: float,
               momentum: float, act: int, slope: float, mask, sums=None):
    """z: NDHWC/2-D activations with C channels.  Returns (out, (scale, shift, mean, invstd)).
    sums: batch statistics already accumulated by the producing GEMM's epilogue (else a pass over z)."""
    M = z.numel() // C
    use_batch = training or rmean is None
    if not use_batch:
        sums = None
    elif sums is None:
        sums = ops.colstats(z, C)
    scale, shift, mean, invstd = ops.bn_finalize(sums, M, gamma, beta, rmean, rvar, use_batch, momentum, eps, C,
                                                 z.device, nbt=nbt if use_batch else None)
    out = ops.bn_act_fwd(z, scale, shift, mask, B, C, act, slope)
    return out, (scale, shift, mean, invstd)


# ---------------------------------------------------------------------------------------------
# inference fast paths (no autograd): the eval-mode BatchNorm that follows a convolution is folded into it --
# scale into the bf16 weights, shift into a bias that pre-loads the TMEM accumulator, activation in the GEMM / stem
# epilogue -- so the BN/activation pass over the block's output (one read + one write) and its launch disappear.
# Only for bf16 activations on the tensor-core kernels; everything else takes the general path.
# ---------------------------------------------------------------------------------------------
def eval_fold_ok(x_dtype: torch.dtype, training: bool, rmean) -> bool:
    return (not training and rmean is not None and not torch.is_grad_enabled() and x_dtype == torch.bfloat16
            and ops._TC_READY and not ops._FORCE_SIMT and not os.environ.get("PB_NO_EVAL_FOLD"))


def pw_fwd_folded(A: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, act: int, slope: float,
                  gate=None, Bt: int = 1) -> torch.Tensor:
    """act((A * gate) W^T * scale + shift): conv -> BN(eval) -> act as ONE GEMM."""
    W = _w2d(w).detach().contiguous()
    N, K = W.shape
    rows = A.numel() // K
    Wb = ops.fold_scaled(W, gate, scale)                          # [Bt or 1][N][K]
    if gate is None:
        F = _row_fold(rows, K)
        if F > 1 and N * F <= 256:
            Wf = ops.block_diag(Wb.view(N, K), F)
            return gemm_tc.gemm(A, Wf, N * F, K * F, bias=shift.repeat(F).contiguous(), act=act, slope=slope).view(-1, N)
        return gemm_tc.gemm(A, Wb.view(N, K), N, K, Bw=1, Bt=1, bias=shift, act=act, slope=slope).view(-1, N)
    return gemm_tc.gemm(A, Wb, N, K, Bw=Bt, Bt=Bt, bias=shift, act=act, slope=slope).view(-1, N)


def bottleneck_eval(x, cfg: BlockCfg, cache: WeightCache, rmean, rvar, w1, wdw, w2, gamma, beta,
                    se_w1, se_b1, se_w2, se_b2):
    """Bottleneck3D.forward in eval mode (mobilenet.py:84-93) with the BatchNorm folded into the projection."""
    x5 = to_ndhwc(x)
    B, T, H, W, Cin = x5.shape
    Cexp, Cout = w1.shape[0], w2.shape[0]
    dt = x5.dtype
    y1 = pw_fwd(x5.view(-1, Cin), w1, cache, "w1").view(B, T, H, W, Cexp)
    wdw_tc = cache.get(("wdw", dt), wdw, lambda: ops.dw_weight_tapmajor(wdw, dt))
    gate = None
    if cfg.use_se:
        y2, pooled = ops.dwconv_fwd_pool(y1, wdw_tc, cfg.k, cfg.s, cfg.p)
        _, gate = ops.se_fc_fwd(pooled, _w2d(se_w1), se_b1.detach(), _w2d(se_w2), se_b2.detach())
    else:
        y2 = ops.dwconv_fwd(y1, wdw_tc, cfg.k, cfg.s, cfg.p)
    _, To, Ho, Wo, _ = y2.shape
    scale, shift, _, _ = ops.bn_finalize(None, 1, gamma.detach(), beta.detach(), rmean, rvar, False, 0.0, cfg.eps, Cout,
                                         x5.device)
    out = pw_fwd_folded(y2.view(-1, Cexp), w2, scale, shift, cfg.act, cfg.slope, gate=gate, Bt=B if gate is not None else 1)
    return from_ndhwc(out.view(B, To, Ho, Wo, Cout))


def stem_eval(x, k, s, p, dt, eps, rmean, rvar, w, bias, gamma, beta):
    """block1 (Conv3d 3->16 + BatchNorm3d(eval) + Hardswish, mobilenet.py:141-143) as one kernel."""
    Cout = w.shape[0]
    scale, shift, _, _ = ops.bn_finalize(None, 1, gamma.detach(), beta.detach(), rmean, rvar, False, 0.0, eps, Cout, x.device)
    wf = (w.detach().float() * scale.view(Cout, 1, 1, 1, 1)).contiguous()
    bf = shift if bias is None else torch.addcmul(shift, bias.detach().float(), scale)
    return from_ndhwc(ops.stem_fwd(x, wf, bf.contiguous(), k, s, p, dt, act=ACT_CODES["hswish"]))


# ---------------------------------------------------------------------------------------------
# inverted-residual bottleneck: pw1 -> depthwise -> [SE] -> pw2 -> BN -> act -> Dropout3d
# (Bottleneck3D.forward mobilenet.py:84-93; MoviNetBottleneck.forward movinet.py:69-77)
# ---------------------------------------------------------------------------------------------
class BottleneckFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg: BlockCfg, cache: WeightCache, training: bool, mask, rmean, rvar, nbt,
                w1, wdw, w2, gamma, beta, se_w1, se_b1, se_w2, se_b2):
        x5 = to_ndhwc(x)
        B, T, H, W, Cin = x5.shape
        Cexp, Cout = w1.shape[0], w2.shape[0]
        dt = x5.dtype
        y1 = pw_fwd(x5.view(-1, Cin), w1, cache, "w1").view(B, T, H, W, Cexp)
        wdw_tc = cache.get(("wdw", dt), wdw, lambda: ops.dw_weight_tapmajor(wdw, dt))
        pooled = hidden = gate = None
        if cfg.use_se:      # the squeeze (average pool) rides on the depthwise kernel's output stores
            y2, pooled = ops.dwconv_fwd_pool(y1, wdw_tc, cfg.k, cfg.s, cfg.p)
        else:
            y2 = ops.dwconv_fwd(y1, wdw_tc, cfg.k, cfg.s, cfg.p)
        _, To, Ho, Wo, _ = y2.shape
        if cfg.use_se:
            hidden, gate = ops.se_fc_fwd(pooled, _w2d(se_w1), se_b1.detach(), _w2d(se_w2), se_b2.detach())
        z, zsums = pw_fwd(y2.view(-1, Cexp), w2, cache, "w2", gate=gate, Bt=B if gate is not None else 1,
                          want_stats=True, fuse_stats=training or rmean is None)
        out, bn_state = bn_forward(z, B, Cout, gamma.detach(), beta.detach(), rmean, rvar, nbt, training, cfg.eps,
                                   cfg.momentum, cfg.act, cfg.slope, mask, sums=zsums)
        ctx.cfg, ctx.cache, ctx.training = cfg, cache, training
        ctx.shapes = (B, T, H, W, Cin, Cexp, Cout, To, Ho, Wo)
        ctx.save_for_backward(x5, y1, y2, z, mask, pooled, hidden, gate, *bn_state,
                              w1, wdw, w2, se_w1, se_w2, wdw_tc)
        return from_ndhwc(out.view(B, To, Ho, Wo, Cout))

    @staticmethod
    def backward(ctx, dout):
        (x5, y1, y2, z, mask, pooled, hidden, gate, scale, shift, mean, invstd,
         w1, wdw, w2, se_w1, se_w2, wdw_tc) = ctx.saved_tensors
        cfg, cache = ctx.cfg, ctx.cache
        B, T, H, W, Cin, Cexp, Cout, To, Ho, Wo = ctx.shapes
        d5 = to_ndhwc(dout)
        if d5.dtype != z.dtype:
            d5 = d5.to(z.dtype)
        dz, dgamma, dbeta = ops.bn_act_bwd(d5, False, z, scale, shift, mean, invstd, mask, B, Cout, cfg.act,
                                           ctx.training, cfg.slope)
        dse = (None, None, None, None)
        if cfg.use_se:
            dw2, dy2, dse = se_pw2_backward(dz, y2, w2, cache, gate, pooled, hidden, se_w1, se_w2, B, To * Ho * Wo)
        else:
            dw2, _ = pw_wgrad(y2.view(-1, Cexp), dz, w2)
            dy2 = pw_dgrad(dz, w2, cache, "w2")
        dy2 = dy2.view(B, To, Ho, Wo, Cexp)
        dwdw_tc = ops.dwconv_wgrad(y1, dy2, cfg.k, cfg.s, cfg.p)
        dwdw = ops.dw_weight_grad_from_tapmajor(dwdw_tc, wdw.shape)
        dy1 = ops.dwconv_dgrad(dy2, wdw_tc, y1.shape, cfg.k, cfg.s, cfg.p)
        dw1, _ = pw_wgrad(x5.view(-1, Cin), dy1.view(-1, Cexp), w1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = from_ndhwc(pw_dgrad(dy1.view(-1, Cexp), w1, cache, "w1").view(B, T, H, W, Cin))
        return (dx, None, None, None, None, None, None, None,
                dw1, dwdw, dw2, dgamma, dbeta, *dse)


# ---------------------------------------------------------------------------------------------
# stem: Conv3d(3->16) [+bias] -> BN -> Hardswish   (mobilenet.py:140-144; movinet.py:91-95)
# ---------------------------------------------------------------------------------------------
class StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p, dt, training, eps, momentum, rmean, rvar, nbt, w, bias, gamma, beta):
        B = x.shape[0]
        Cout = w.shape[0]
        if x.dtype not in (torch.uint8, torch.float32, torch.bfloat16):
            raise TypeError(f"stem: unsupported clip dtype {x.dtype}")
        z = ops.stem_fwd(x, w.detach().contiguous(), None if bias is None else bias.detach(), k, s, p, dt)
        out, bn_state = bn_forward(z, B, Cout, gamma.detach(), beta.detach(), rmean, rvar, nbt, training, eps,
                                   momentum, ACT_CODES["hswish"], 0.0, None)
        ctx.conv = (k, s, p)
        ctx.training = training
        ctx.has_bias = bias is not None
        ctx.wshape = w.shape
        ctx.save_for_backward(x, z, *bn_state)
        return from_ndhwc(out)

    @staticmethod
    def backward(ctx, dout):
        x, z, scale, shift, mean, invstd = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("picklebot_b200: gradient w.r.t. the input clip is not implemented")
        k, s, p = ctx.conv
        B, Cout = z.shape[0], z.shape[-1]
        d5 = to_ndhwc(dout)
        if d5.dtype != z.dtype:
            d5 = d5.to(z.dtype)
        dz, dgamma, dbeta = ops.bn_act_bwd(d5, False, z, scale, shift, mean, invstd, None, B, Cout,
                                           ACT_CODES["hswish"], ctx.training)
        dw, db = ops.stem_wgrad(x, dz, ctx.wshape, k, s, p, ctx.has_bias)
        return (None,) * 11 + (dw, db, dgamma, dbeta)


# ---------------------------------------------------------------------------------------------
# tail of the MobileNets: 1x1x1 conv (+bias) [-> SE] -> BN -> Hardswish -> global pool -> FC -> Hardswish -> FC
# (mobilenet.py:178-190 Large; 244-256 Small, where SE sits between the conv and the BN)
# ---------------------------------------------------------------------------------------------
class MobileNetTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cache: WeightCache, training, eps, momentum, use_se, rmean, rvar, nbt,
                wc, bc, gamma, beta, wf1, bf1, wf2, bf2, se_w1, se_b1, se_w2, se_b2):
        x5 = to_ndhwc(x)
        B, T, H, W, Cin = x5.shape
        Cmid = wc.shape[0]
        R = T * H * W
        hs = ACT_CODES["hswish"]
        z0 = pw_fwd(x5.view(-1, Cin), wc, cache, "tail_conv", bias=bc.detach())
        pooled = hidden = gate = None
        z = z0
        if use_se:
            pooled = ops.pool_fwd(z0, B, Cmid)
            hidden, gate = ops.se_fc_fwd(pooled, _w2d(se_w1), se_b1.detach(), _w2d(se_w2), se_b2.detach())
            z = ops.rowscale(z0, gate, B, Cmid)
        a, bn_state = bn_forward(z, B, Cmid, gamma.detach(), beta.detach(), rmean, rvar, nbt, training, eps, momentum,
                                 hs, 0.0, None)
        feat = ops.pool_fwd(a, B, Cmid)                                   # fp32 [B][Cmid]
        F1, NC = wf1.shape[0], wf2.shape[0]
        u1 = ops.fc_fwd(feat, _w2d(wf1), bf1.detach())
        ones = torch.ones(F1, dtype=torch.float32, device=x.device)
        zeros = torch.zeros(F1, dtype=torch.float32, device=x.device)
        h1 = ops.bn_act_fwd(u1, ones, zeros, None, B, F1, hs)
        logits = ops.fc_fwd(h1, _w2d(wf2), bf2.detach())
        ctx.cache, ctx.training, ctx.use_se = cache, training, use_se
        ctx.shapes = (B, T, H, W, Cin, Cmid, R, F1, NC)
        ctx.save_for_backward(x5, z0, z, pooled, hidden, gate, *bn_state, feat, u1, h1, ones, zeros,
                              wc, wf1, wf2, se_w1, se_w2)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        (x5, z0, z, pooled, hidden, gate, scale, shift, mean, invstd, feat, u1, h1, ones, zeros,
         wc, wf1, wf2, se_w1, se_w2) = ctx.saved_tensors
        cache = ctx.cache
        B, T, H, W, Cin, Cmid, R, F1, NC = ctx.shapes
        hs = ACT_CODES["hswish"]
        dl = dlogits.contiguous().float()
        dwf2, dbf2 = ops.wgrad_simt(h1, dl, F1, NC, want_bias=True)
        dh1 = ops.fc_dgrad(dl, _w2d(wf2))
        du1, _, _ = ops.bn_act_bwd(dh1, False, u1, ones, zeros, zeros, ones, None, B, F1, hs, False)
        dwf1, dbf1 = ops.wgrad_simt(feat, du1, Cmid, F1, want_bias=True)
        dfeat = ops.fc_dgrad(du1, _w2d(wf1), 1.0 / R)                      # [B][Cmid] fp32, gradient of the mean
        dz, dgamma, dbeta = ops.bn_act_bwd(dfeat, True, z, scale, shift, mean, invstd, None, B, Cmid, hs,
                                           ctx.training)
        dse = (None, None, None, None)
        if ctx.use_se:
            dgate = ops.rowdot(dz, z0, B, Cmid)
            dmean, dW1, db1, dW2, db2 = ops.se_fc_bwd(dgate, pooled, hidden, gate, _w2d(se_w1), _w2d(se_w2), 1.0 / R)
            ops.scale_add_(dz, gate, dmean, B, Cmid)                       # dz0 = dz*gate + dmean/R
            dse = (dW1.view(se_w1.shape), db1, dW2.view(se_w2.shape), db2)
        dwc, dbc = pw_wgrad(x5.view(-1, Cin), dz, wc, want_bias=True)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = from_ndhwc(pw_dgrad(dz, wc, cache, "tail_conv").view(B, T, H, W, Cin))
        return (dx, None, None, None, None, None, None, None, None,
                dwc, dbc, dgamma, dbeta, dwf1.view(wf1.shape), dbf1, dwf2.view(wf2.shape), dbf2, *dse)


# ---------------------------------------------------------------------------------------------
# tail of MoViNetA2: conv 144->640 -> BN -> Hardswish -> Dropout3d -> pool -> Linear -> BN1d -> Hardswish
# -> Dropout -> Linear   (movinet.py:139-154)
# ---------------------------------------------------------------------------------------------
class MoViNetTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cache: WeightCache, training, eps, momentum, mask3d, mask1d,
                rmean, rvar, nbt, rmean1, rvar1, nbt1,
                wc, gamma, beta, wf1, bf1, gamma1, beta1, wf2, bf2):
        x5 = to_ndhwc(x)
        B, T, H, W, Cin = x5.shape
        Cmid = wc.shape[0]
        R = T * H * W
        hs = ACT_CODES["hswish"]
        # (BatchNorm3d, BatchNorm1d) each with its own mode, eps and momentum (movinet.py:141,150)
        (training, training1), (eps, eps1), (momentum, momentum1) = training, eps, momentum
        z = pw_fwd(x5.view(-1, Cin), wc, cache, "tail_conv")
        a, bn_state = bn_forward(z, B, Cmid, gamma.detach(), beta.detach(), rmean, rvar, nbt, training, eps, momentum,
                                 hs, 0.0, mask3d)
        feat = ops.pool_fwd(a, B, Cmid)
        F1, NC = wf1.shape[0], wf2.shape[0]
        u1 = ops.fc_fwd(feat, wf1.detach(), bf1.detach())
        h1, bn1_state = bn_forward(u1, B, F1, gamma1.detach(), beta1.detach(), rmean1, rvar1, nbt1, training1, eps1,
                                   momentum1, hs, 0.0, mask1d)
        logits = ops.fc_fwd(h1, wf2.detach(), bf2.detach())
        ctx.cache, ctx.training, ctx.training1 = cache, training, training1
        ctx.shapes = (B, T, H, W, Cin, Cmid, R, F1, NC)
        ctx.save_for_backward(x5, z, mask3d, mask1d, *bn_state, feat, u1, *bn1_state, h1, wc, wf1, wf2)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        (x5, z, mask3d, mask1d, scale, shift, mean, invstd, feat, u1, scale1, shift1, mean1, invstd1, h1,
         wc, wf1, wf2) = ctx.saved_tensors
        cache = ctx.cache
        B, T, H, W, Cin, Cmid, R, F1, NC = ctx.shapes
        hs = ACT_CODES["hswish"]
        dl = dlogits.contiguous().float()
        dwf2, dbf2 = ops.wgrad_simt(h1, dl, F1, NC, want_bias=True)
        dh1 = ops.fc_dgrad(dl, wf2.detach())
        du1, dgamma1, dbeta1 = ops.bn_act_bwd(dh1, False, u1, scale1, shift1, mean1, invstd1, mask1d, B, F1, hs,
                                              ctx.training1)
        dwf1, dbf1 = ops.wgrad_simt(feat, du1, Cmid, F1, want_bias=True)
        dfeat = ops.fc_dgrad(du1, wf1.detach(), 1.0 / R)
        dz, dgamma, dbeta = ops.bn_act_bwd(dfeat, True, z, scale, shift, mean, invstd, mask3d, B, Cmid, hs,
                                           ctx.training)
        dwc, _ = pw_wgrad(x5.view(-1, Cin), dz, wc)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = from_ndhwc(pw_dgrad(dz, wc, cache, "tail_conv").view(B, T, H, W, Cin))
        return (dx,) + (None,) * 12 + (dwc, dgamma, dbeta, dwf1, dbf1, dgamma1, dbeta1, dwf2, dbf2)
