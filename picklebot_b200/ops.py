"""Tensor-level launchers over the C ABI.  torch is plumbing here: it owns device memory and the
current stream; every function below only enqueues kernels of libpicklebot_b200.so on that stream.

All activations are contiguous NDHWC tensors: ``(B, T, H, W, C)`` or their 2-D/3-D views
``(M, C)`` / ``(B, R, C)``.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_HSIGMOID, ACT_HSWISH, ACT_LRELU, ACT_NONE, ACT_RELU, PB_BF16, PB_F32, PB_F32_RBF16,
                   PB_U8, STAT_REPLICAS)


def call(name, *args, nbytes=0, tag="", wbytes=0):
    """nbytes: algorithmic bytes of the launch (read + written), wbytes: the written part (a write-only stream tops
    out at ~3.9 TB/s on B200 against ~6.5 TB/s for a copy, so the split matters for the roofline bound)."""
    if _lib.PROFILER is not None and not tag:
        tag = ",".join(str(a) for a in args if isinstance(a, int) and not isinstance(a, bool) and 0 < a < 100000)
    _lib.call(name, *args, nbytes=nbytes, tag=tag, wbytes=wbytes)

ACT_CODES = {"none": ACT_NONE, "relu": ACT_RELU, "hswish": ACT_HSWISH, "lrelu": ACT_LRELU,
             "hsigmoid": ACT_HSIGMOID}

# PB_GEMM=simt forces the CUDA-core GEMM for bf16 too (bring-up / debugging of the tcgen05 path).
_FORCE_SIMT = os.environ.get("PB_GEMM", "").lower() == "simt"


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return PB_F32
    if t.dtype == torch.bfloat16:
        return PB_BF16
    if t.dtype == torch.uint8:
        return PB_U8
    raise TypeError(f"unsupported activation dtype {t.dtype}")


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: picklebot_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor, got strides {t.stride()}")
    return t


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected fp32, got {t.dtype}")
    return _chk(t, name)


def conv_out(n: int, k: int, s: int, p: int) -> int:
    return (n + 2 * p - k) // s + 1


# ---------------------------------------------------------------------------------------------
# weight repacking
# ---------------------------------------------------------------------------------------------
def cast_matrix(src: torch.Tensor, rows: int, cols: int, dst_dtype: torch.dtype, transpose: bool = False,
                round_bf16: bool = False) -> torch.Tensor:
    """fp32 [rows][cols] -> [cols][rows] (transpose) or [rows][cols] in dst_dtype."""
    _f32(src, "cast_matrix.src")
    shape = (cols, rows) if transpose else (rows, cols)
    dst = torch.empty(shape, dtype=dst_dtype, device=src.device)
    code = PB_BF16 if dst_dtype == torch.bfloat16 else (PB_F32_RBF16 if round_bf16 else PB_F32)
    call("pb_cast_matrix", src.data_ptr(), dst.data_ptr(), code, rows, cols, int(transpose), _st())
    return dst


def dw_weight_tapmajor(w: torch.Tensor, act_dtype: torch.dtype) -> torch.Tensor:
    """(C,1,kT,kH,kW) fp32 parameter -> [taps][C] fp32, values rounded to the activation dtype."""
    C = w.shape[0]
    taps = w.numel() // C
    return cast_matrix(w.detach().contiguous(), C, taps, torch.float32, transpose=True,
                       round_bf16=(act_dtype == torch.bfloat16))


def dw_weight_grad_from_tapmajor(dw_tc: torch.Tensor, shape) -> torch.Tensor:
    C = shape[0]
    taps = dw_tc.numel() // C
    return cast_matrix(dw_tc, taps, C, torch.float32, transpose=True).view(shape)


def block_diag(Wb: torch.Tensor, F: int) -> torch.Tensor:
    """bf16 [N][K] -> bf16 [F*N][F*K] = diag(W, ..., W) (weights of a row-folded GEMM)."""
    _chk(Wb, "block_diag.W")
    N, K = Wb.shape
    dst = torch.empty((F * N, F * K), dtype=torch.bfloat16, device=Wb.device)
    call("pb_block_diag_bf16", Wb.data_ptr(), dst.data_ptr(), F, N, K, _st())
    return dst


def fold_gate(W: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """bf16 [B][N][K] = W[n][k] * gate[b][k]."""
    N, K = W.shape[0], W.shape[1]
    B = gate.shape[0]
    dst = torch.empty((B, N, K), dtype=torch.bfloat16, device=W.device)
    call("pb_fold_gate_bf16", W.data_ptr(), gate.data_ptr(), dst.data_ptr(), B, N, K, _st())
    return dst


def fold_scaled(W: torch.Tensor, gate: Optional[torch.Tensor], rowscale: Optional[torch.Tensor]) -> torch.Tensor:
    """bf16 [B][N][K] = W[n][k] * gate[b][k] * rowscale[n] (either factor optional; B = 1 without a gate): inference
    weights with the squeeze-excite gate and the following eval-mode BatchNorm's scale folded in."""
    _f32(W, "fold_scaled.W"); _f32(gate, "fold_scaled.gate"); _f32(rowscale, "fold_scaled.rowscale")
    N, K = W.shape[0], W.shape[1]
    B = gate.shape[0] if gate is not None else 1
    dst = torch.empty((B, N, K), dtype=torch.bfloat16, device=W.device)
    call("pb_fold_scaled_bf16", W.data_ptr(), _p(gate), _p(rowscale), dst.data_ptr(), B, N, K, _st())
    return dst


def fold_rows(Wt: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """bf16 [B][R][C] = Wt[r][c] * gate[b][r] for an fp32 weight that is already transposed ([R][C]): what ``fold_gate_t``
    computes from the untransposed weight, with coalesced accesses only."""
    _f32(Wt, "fold_rows.Wt"); _f32(gate, "fold_rows.gate")
    R, C = Wt.shape[0], Wt.shape[1]
    B = gate.shape[0]
    dst = torch.empty((B, R, C), dtype=torch.bfloat16, device=Wt.device)
    call("pb_fold_rows_bf16", Wt.data_ptr(), gate.data_ptr(), dst.data_ptr(), B, R, C, _st())
    return dst


def fold_gate_t(W: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """bf16 [B][K][N] = W[n][k] * gate[b][k]: the per-sample weights of the input-gradient GEMM  dy2 = (dz W) * gate."""
    _f32(W, "fold_gate_t.W"); _f32(gate, "fold_gate_t.gate")
    N, K = W.shape[0], W.shape[1]
    B = gate.shape[0]
    dst = torch.empty((B, K, N), dtype=torch.bfloat16, device=W.device)
    call("pb_fold_gate_t_bf16", W.data_ptr(), gate.data_ptr(), dst.data_ptr(), B, N, K, _st())
    return dst


# ---------------------------------------------------------------------------------------------
# depthwise conv
# ---------------------------------------------------------------------------------------------
def _dw_dims(x_shape, k, s, p):
    B, T, H, W, C = x_shape
    To, Ho, Wo = conv_out(T, k[0], s[0], p[0]), conv_out(H, k[1], s[1], p[1]), conv_out(W, k[2], s[2], p[2])
    return (B, C, T, H, W, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], To, Ho, Wo)


def dwconv_fwd(x: torch.Tensor, w_tc: torch.Tensor, k, s, p) -> torch.Tensor:
    _chk(x, "dwconv_fwd.x")
    d = _dw_dims(x.shape, k, s, p)
    y = torch.empty((d[0], d[14], d[15], d[16], d[1]), dtype=x.dtype, device=x.device)
    call("pb_dwconv3d_fwd", x.data_ptr(), w_tc.data_ptr(), y.data_ptr(), _dt(x), *d, _st(),
         nbytes=(x.numel() + y.numel()) * x.element_size() + w_tc.numel() * x.element_size(),
         wbytes=y.numel() * y.element_size())
    return y


def dwconv_fwd_pool(x: torch.Tensor, w_tc: torch.Tensor, k, s, p) -> Tuple[torch.Tensor, torch.Tensor]:
    """Depthwise forward + the global average pool of its output, fp32 [B][C], in the same pass (squeeze-excite)."""
    _chk(x, "dwconv_fwd_pool.x")
    d = _dw_dims(x.shape, k, s, p)
    y = torch.empty((d[0], d[14], d[15], d[16], d[1]), dtype=x.dtype, device=x.device)
    pooled = torch.empty((d[0], d[1]), dtype=torch.float32, device=x.device)
    call("pb_dwconv3d_fwd_pool", x.data_ptr(), w_tc.data_ptr(), y.data_ptr(), pooled.data_ptr(), _dt(x), *d, _st(),
         nbytes=(x.numel() + y.numel()) * x.element_size() + w_tc.numel() * x.element_size(),
         wbytes=y.numel() * y.element_size())
    return y, pooled


def dwconv_dgrad(dy: torch.Tensor, w_tc: torch.Tensor, x_shape, k, s, p) -> torch.Tensor:
    _chk(dy, "dwconv_dgrad.dy")
    d = _dw_dims(x_shape, k, s, p)
    dx = torch.empty(tuple(x_shape), dtype=dy.dtype, device=dy.device)
    call("pb_dwconv3d_dgrad", dy.data_ptr(), w_tc.data_ptr(), dx.data_ptr(), _dt(dy), *d, _st(),
         nbytes=(dx.numel() + dy.numel()) * dy.element_size() + w_tc.numel() * dy.element_size(),
         wbytes=dx.numel() * dx.element_size())
    return dx


def dwconv_wgrad(x: torch.Tensor, dy: torch.Tensor, k, s, p) -> torch.Tensor:
    """Returns the tap-major [taps][C] fp32 gradient."""
    _chk(x, "dwconv_wgrad.x"); _chk(dy, "dwconv_wgrad.dy")
    d = _dw_dims(x.shape, k, s, p)
    dw_tc = torch.empty((k[0] * k[1] * k[2], x.shape[-1]), dtype=torch.float32, device=x.device)
    call("pb_dwconv3d_wgrad", x.data_ptr(), dy.data_ptr(), dw_tc.data_ptr(), _dt(x), *d, _st(),
         nbytes=(x.numel() + dy.numel()) * x.element_size() + dw_tc.numel() * 4)
    return dw_tc


def stream_dwconv_fwd(x: torch.Tensor, sbuf: Optional[torch.Tensor], w_tc: torch.Tensor, k, s, p,
                      inplace: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Causal chunked depthwise conv; returns (y, new stream buffer)."""
    _chk(x, "stream_dwconv_fwd.x")
    B, T, H, W, C = x.shape
    Ho, Wo = conv_out(H, k[1], s[1], p[1]), conv_out(W, k[2], s[2], p[2])
    y = torch.empty((B, T, Ho, Wo, C), dtype=x.dtype, device=x.device)
    new_buf = None
    if k[0] > 1:
        if sbuf is None:
            sbuf = torch.zeros((B, k[0] - 1, H, W, C), dtype=x.dtype, device=x.device)
        # in place when the chunk is at least as long as the history (the tail then comes from x alone)
        new_buf = sbuf if (inplace and T >= k[0] - 1) else torch.empty_like(sbuf)
    hist = 0 if sbuf is None else sbuf.numel()
    call("pb_stream_dwconv3d_fwd", x.data_ptr(), _p(sbuf), w_tc.data_ptr(), y.data_ptr(), _p(new_buf), _dt(x),
         B, C, T, H, W, k[0], k[1], k[2], s[1], s[2], p[1], p[2], Ho, Wo, _st(),
         nbytes=(x.numel() + y.numel() + 2 * hist) * x.element_size())   # chunk in, out, history read + rewritten
    return y, new_buf


# ---------------------------------------------------------------------------------------------
# pointwise GEMMs
# ---------------------------------------------------------------------------------------------
def use_tc(dtype: torch.dtype, K: int, N: int) -> bool:
    return dtype == torch.bfloat16 and not _FORCE_SIMT and _TC_READY and K % 8 == 0 and N % 8 == 0


_TC_READY = False   # flipped by gemm_tc.py once the tcgen05 kernel is validated on the device


def gemm_simt(A: torch.Tensor, W: torch.Tensor, N: int, K: int, w_sn: int, w_sk: int,
              bias=None, ascale=None, colscale=None, coladd=None, Bt: int = 1) -> torch.Tensor:
    """A: (..., K) contiguous, viewed as [Bt][R][K].  W fp32 addressed as w[n*w_sn + k*w_sk]."""
    _chk(A, "gemm.A")
    rows = A.numel() // K
    R = rows // Bt
    C = torch.empty((rows, N), dtype=A.dtype, device=A.device)
    call("pb_pw_gemm_simt", A.data_ptr(), W.data_ptr(), w_sn, w_sk, _p(bias), _p(ascale), _p(colscale),
         _p(coladd), C.data_ptr(), _dt(A), Bt, R, K, N, _st(),
         nbytes=(A.numel() + C.numel() + N * K) * A.element_size(), wbytes=C.numel() * C.element_size())
    return C


def wgrad_simt(A: torch.Tensor, dC: torch.Tensor, K: int, N: int, ascale=None, Bt: int = 1,
               want_bias: bool = False):
    _chk(A, "wgrad.A"); _chk(dC, "wgrad.dC")
    rows = A.numel() // K
    R = rows // Bt
    dW = torch.empty((N, K), dtype=torch.float32, device=A.device)
    db = torch.empty((N,), dtype=torch.float32, device=A.device) if want_bias else None
    call("pb_pw_wgrad_simt", A.data_ptr(), dC.data_ptr(), _p(ascale), dW.data_ptr(), _p(db), _dt(A), Bt, R, K, N, _st(),
         nbytes=(A.numel() + dC.numel()) * A.element_size() + N * K * 4)
    return dW, db


# ---------------------------------------------------------------------------------------------
# batch norm + activation + dropout mask
# ---------------------------------------------------------------------------------------------
def colstats(x: torch.Tensor, C: int) -> torch.Tensor:
    _chk(x, "colstats.x")
    M = x.numel() // C
    sums = torch.empty((STAT_REPLICAS, 2, C), dtype=torch.float64, device=x.device)
    call("pb_colstats", x.data_ptr(), _dt(x), M, C, sums.data_ptr(), _st(), nbytes=x.numel() * x.element_size())
    return sums


def bn_finalize(sums, M: int, gamma, beta, rmean, rvar, training: bool, momentum: float, eps: float, C: int,
                device, nbt=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """nbt: nn.BatchNorm's num_batches_tracked (int64 device scalar), incremented by the kernel when training."""
    out = torch.empty((4, C), dtype=torch.float32, device=device)
    if nbt is not None:
        assert nbt.dtype == torch.int64 and nbt.is_cuda
    call("pb_bn_finalize", _p(sums), M, _p(gamma), _p(beta), _p(rmean), _p(rvar), int(training), momentum, eps,
         out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), _p(nbt), C, _st())
    return out[0], out[1], out[2], out[3]


def bn_act_fwd(z: torch.Tensor, scale, shift, mask, B: int, C: int, act: int, slope: float = 0.01) -> torch.Tensor:
    _chk(z, "bn_act_fwd.z")
    R = z.numel() // (B * C)
    out = torch.empty_like(z)
    call("pb_bn_act_fwd", z.data_ptr(), scale.data_ptr(), shift.data_ptr(), _p(mask), out.data_ptr(), _dt(z),
         B, R, C, act, slope, _st(), nbytes=2 * z.numel() * z.element_size(), wbytes=z.numel() * z.element_size())
    return out


def bn_act_bwd(dout: torch.Tensor, dout_bcast: bool, z: torch.Tensor, scale, shift, mean, invstd, mask,
               B: int, C: int, act: int, training: bool, slope: float = 0.01, want_param_grads: bool = True):
    """Returns (dz, dgamma, dbeta)."""
    _chk(z, "bn_act_bwd.z"); _chk(dout, "bn_act_bwd.dout")
    R = z.numel() // (B * C)
    M = B * R
    dev = z.device
    sums = torch.empty((STAT_REPLICAS, 2, C), dtype=torch.float64, device=dev)
    call("pb_bn_act_bwd_reduce", dout.data_ptr(), int(dout_bcast), z.data_ptr(), scale.data_ptr(), shift.data_ptr(),
         mean.data_ptr(), invstd.data_ptr(), _p(mask), sums.data_ptr(), _dt(z), B, R, C, act, slope, _st(),
         nbytes=(z.numel() + (0 if dout_bcast else dout.numel())) * z.element_size())
    small = torch.empty((4, C), dtype=torch.float32, device=dev)   # dgamma | dbeta | coef0 | coef1
    call("pb_bn_bwd_finalize", sums.data_ptr(), M, int(training), small[0].data_ptr(), small[1].data_ptr(),
         small[2].data_ptr(), C, _st())
    dz = torch.empty_like(z)
    call("pb_bn_act_bwd_apply", dout.data_ptr(), int(dout_bcast), z.data_ptr(), scale.data_ptr(), shift.data_ptr(),
         mean.data_ptr(), invstd.data_ptr(), _p(mask), small[2].data_ptr(), dz.data_ptr(), _dt(z), B, R, C, act,
         slope, _st(), nbytes=(2 * z.numel() + (0 if dout_bcast else dout.numel())) * z.element_size(),
         wbytes=z.numel() * z.element_size())
    return dz, small[0], small[1]


# ---------------------------------------------------------------------------------------------
# squeeze-excite / pooling
# ---------------------------------------------------------------------------------------------
def pool_fwd(x: torch.Tensor, B: int, C: int) -> torch.Tensor:
    _chk(x, "pool_fwd.x")
    R = x.numel() // (B * C)
    mean = torch.empty((B, C), dtype=torch.float32, device=x.device)
    call("pb_pool_fwd", x.data_ptr(), _dt(x), B, R, C, mean.data_ptr(), _st(), nbytes=x.numel() * x.element_size())
    return mean


def stream_pool_update(chunk_mean: torch.Tensor, R: int, ssum: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """Cumulative mean over all chunks seen so far (in-place update of the stream state ``ssum`` / ``rows``)."""
    B, C = chunk_mean.shape
    mean = torch.empty_like(chunk_mean)
    call("pb_stream_pool_update", chunk_mean.data_ptr(), R, ssum.data_ptr(), rows.data_ptr(), mean.data_ptr(), B, C, _st())
    return mean


def fc_fwd(X: torch.Tensor, W: torch.Tensor, bias=None) -> torch.Tensor:
    """nn.Linear forward for a small batch: fp32 X [B][K], W [N][K] -> [B][N]."""
    _chk(X, "fc_fwd.X"); _chk(W, "fc_fwd.W")
    assert X.dtype == torch.float32 and W.dtype == torch.float32
    B, K = X.shape
    N = W.shape[0]
    Y = torch.empty((B, N), dtype=torch.float32, device=X.device)
    call("pb_fc_fwd", X.data_ptr(), W.data_ptr(), _p(bias), Y.data_ptr(), B, N, K, _st(), nbytes=(N * K + B * (N + K)) * 4)
    return Y


def fc_dgrad(dY: torch.Tensor, W: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """dX [B][K] = scale * dY [B][N] x W [N][K]  (fp32)."""
    _chk(dY, "fc_dgrad.dY"); _chk(W, "fc_dgrad.W")
    assert dY.dtype == torch.float32 and W.dtype == torch.float32
    B, N = dY.shape
    K = W.shape[1]
    dX = torch.empty((B, K), dtype=torch.float32, device=dY.device)
    call("pb_fc_dgrad", dY.data_ptr(), W.data_ptr(), dX.data_ptr(), B, N, K, float(scale), _st(),
         nbytes=(N * K + B * (N + K)) * 4)
    return dX


def se_fc_fwd(mean, W1, b1, W2, b2) -> Tuple[torch.Tensor, torch.Tensor]:
    B, C = mean.shape
    Ch = W1.shape[0]
    hidden = torch.empty((B, Ch), dtype=torch.float32, device=mean.device)
    gate = torch.empty((B, C), dtype=torch.float32, device=mean.device)
    call("pb_se_fc_fwd", mean.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
         hidden.data_ptr(), gate.data_ptr(), B, C, Ch, _st())
    return hidden, gate


def se_fc_bwd(dgate, mean, hidden, gate, W1, W2, inv_R: float):
    B, C = mean.shape
    Ch = W1.shape[0]
    dev = mean.device
    dmean = torch.empty((B, C), dtype=torch.float32, device=dev)
    work = torch.empty((B * (C + Ch),), dtype=torch.float32, device=dev)
    dW1 = torch.empty((Ch, C), dtype=torch.float32, device=dev)
    db1 = torch.empty((Ch,), dtype=torch.float32, device=dev)
    dW2 = torch.empty((C, Ch), dtype=torch.float32, device=dev)
    db2 = torch.empty((C,), dtype=torch.float32, device=dev)
    call("pb_se_fc_bwd", dgate.data_ptr(), mean.data_ptr(), hidden.data_ptr(), gate.data_ptr(), W1.data_ptr(),
         W2.data_ptr(), inv_R, dmean.data_ptr(), work.data_ptr(), dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(),
         db2.data_ptr(), B, C, Ch, _st())
    return dmean, dW1, db1, dW2, db2


def rowscale(x: torch.Tensor, gate: torch.Tensor, B: int, C: int) -> torch.Tensor:
    _chk(x, "rowscale.x")
    R = x.numel() // (B * C)
    y = torch.empty_like(x)
    call("pb_rowscale", x.data_ptr(), gate.data_ptr(), y.data_ptr(), _dt(x), B, R, C, _st(),
         nbytes=2 * x.numel() * x.element_size())
    return y


def rowdot(g: torch.Tensor, y: torch.Tensor, B: int, C: int) -> torch.Tensor:
    _chk(g, "rowdot.g"); _chk(y, "rowdot.y")
    R = g.numel() // (B * C)
    out = torch.empty((B, C), dtype=torch.float32, device=g.device)
    call("pb_rowdot", g.data_ptr(), y.data_ptr(), _dt(g), B, R, C, out.data_ptr(), _st(),
         nbytes=2 * g.numel() * g.element_size())
    return out


def scale_add_(g: torch.Tensor, gate: torch.Tensor, add: torch.Tensor, B: int, C: int) -> torch.Tensor:
    _chk(g, "scale_add.g")
    R = g.numel() // (B * C)
    call("pb_scale_add", g.data_ptr(), gate.data_ptr(), add.data_ptr(), _dt(g), B, R, C, _st(),
         nbytes=2 * g.numel() * g.element_size())
    return g


# ---------------------------------------------------------------------------------------------
# stem
# ---------------------------------------------------------------------------------------------
def _stem_args(x: torch.Tensor, k, s, p, Cout: int):
    B, Cin, T, H, W = x.shape
    To, Ho, Wo = conv_out(T, k[0], s[0], p[0]), conv_out(H, k[1], s[1], p[1]), conv_out(W, k[2], s[2], p[2])
    sb, sc, st, sh, sw = x.stride()
    dims = (B, Cin, T, H, W, Cout, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], To, Ho, Wo)
    return (sb, sc, st, sh, sw), dims, (B, To, Ho, Wo, Cout)


def stem_fwd(x: torch.Tensor, w: torch.Tensor, bias, k, s, p, out_dtype: torch.dtype, act: int = 0,
             slope: float = 0.0) -> torch.Tensor:
    """x: logical (B,Cin,T,H,W) with ANY strides, dtype uint8 (divided by 255) / fp32 / bf16.
    act != 0: y = act(conv + bias) (inference with the BatchNorm folded into w / bias; tensor-core kernels only)."""
    if not x.is_cuda:
        raise RuntimeError("stem_fwd: picklebot_b200 kernels need CUDA tensors (there is no CPU fallback)")
    strides, dims, oshape = _stem_args(x, k, s, p, w.shape[0])
    y = torch.empty(oshape, dtype=out_dtype, device=x.device)
    if act:
        call("pb_stem_conv_fwd_act", x.data_ptr(), _dt(x), *strides, 255.0, w.data_ptr(), _p(bias), y.data_ptr(), _dt(y),
             *dims, act, float(slope), _st(), nbytes=x.numel() * x.element_size() + y.numel() * y.element_size())
        return y
    call("pb_stem_conv_fwd", x.data_ptr(), _dt(x), *strides, 255.0, w.data_ptr(), _p(bias), y.data_ptr(), _dt(y),
         *dims, _st(), nbytes=x.numel() * x.element_size() + y.numel() * y.element_size())
    return y


def stem_wgrad(x: torch.Tensor, dy: torch.Tensor, w_shape, k, s, p, want_bias: bool):
    _chk(dy, "stem_wgrad.dy")
    strides, dims, _ = _stem_args(x, k, s, p, w_shape[0])
    dw = torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device)
    db = torch.empty((w_shape[0],), dtype=torch.float32, device=x.device) if want_bias else None
    call("pb_stem_conv_wgrad", x.data_ptr(), _dt(x), *strides, 255.0, dy.data_ptr(), _dt(dy), dw.data_ptr(), _p(db),
         *dims, _st(), nbytes=x.numel() * x.element_size() + dy.numel() * dy.element_size())
    return dw, db
