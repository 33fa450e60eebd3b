"""Launchers for the tcgen05 / TMEM / TMA GEMM kernels (csrc/pwgemm_tc.cu, csrc/pwwgrad_tc.cu)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib, ops
from .ops import _chk, _p, _st, call

# The tensor-core path is the production path for bf16; PB_GEMM=simt (see ops.py) turns it off.
ops._TC_READY = os.environ.get("PB_TC", "1") != "0"
_WGRAD_READY = os.environ.get("PB_TC_WGRAD", "1") != "0"


def gemm(A: torch.Tensor, Wb: torch.Tensor, N: int, K: int, Bw: int = 1, Bt: int = 1, bias=None, colscale=None,
         coladd=None, stat_mod: int = 0, act: int = 0, slope: float = 0.0):
    """A bf16 [Bt][R][K] (contiguous), Wb bf16 [Bw][N][K] -> bf16 [Bt*R][N];
    C = (A W^T + bias) * colscale[b] + coladd[b].
    stat_mod > 0: also returns the BatchNorm sums of C ([STAT_REPLICAS][2][stat_mod] fp64, channel = column % stat_mod),
    accumulated by the GEMM epilogue (N <= 256, no epilogue vectors)."""
    _chk(A, "gemm_tc.A"); _chk(Wb, "gemm_tc.W")
    assert A.dtype == torch.bfloat16 and Wb.dtype == torch.bfloat16
    rows = A.numel() // K
    R = rows // Bt
    C = torch.empty((rows, N), dtype=torch.bfloat16, device=A.device)
    sums = torch.empty((_lib.STAT_REPLICAS, 2, stat_mod), dtype=torch.float64, device=A.device) if stat_mod else None
    if act:     # C = act(A W^T + bias): inference with the eval-mode BatchNorm folded into Wb / bias
        call("pb_pw_gemm_tc_act", A.data_ptr(), Wb.data_ptr(), Bw, _p(bias), _p(colscale), _p(coladd), C.data_ptr(),
             _p(sums), stat_mod, Bt, R, K, N, act, float(slope), _st(),
             nbytes=(A.numel() + C.numel() + Wb.numel()) * 2, wbytes=C.numel() * 2)
        return (C, sums) if stat_mod else C
    call("pb_pw_gemm_tc", A.data_ptr(), Wb.data_ptr(), Bw, _p(bias), _p(colscale), _p(coladd), C.data_ptr(),
         _p(sums), stat_mod, Bt, R, K, N, _st(), nbytes=(A.numel() + C.numel() + Wb.numel()) * 2, wbytes=C.numel() * 2)
    return (C, sums) if stat_mod else C


def wgrad_ready() -> bool:
    return _WGRAD_READY and ops._TC_READY


def wgrad(A: torch.Tensor, dC: torch.Tensor, K: int, N: int, gate: Optional[torch.Tensor] = None,
          W: Optional[torch.Tensor] = None, Bt: int = 1, want_dgate: bool = False
          ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """dW[n][k] = sum_b gate[b][k] * sum_r dC[b][r][n] A[b][r][k]  (fp32 [N][K]);
    optionally dgate[b][k] = sum_n W[n][k] * (sum_r dC A)[b]."""
    _chk(A, "wgrad_tc.A"); _chk(dC, "wgrad_tc.dC")
    assert A.dtype == torch.bfloat16 and dC.dtype == torch.bfloat16
    rows = A.numel() // K
    R = rows // Bt
    nbytes_ws = int(_lib.lib().pb_pw_wgrad_tc_workspace_bytes(Bt, R, K, N))
    ws = torch.empty((nbytes_ws // 4,), dtype=torch.float32, device=A.device)
    dW = torch.empty((N, K), dtype=torch.float32, device=A.device)
    dgate = torch.empty((Bt, K), dtype=torch.float32, device=A.device) if want_dgate else None
    call("pb_pw_wgrad_tc", A.data_ptr(), dC.data_ptr(), _p(gate), _p(W), ws.data_ptr(), dW.data_ptr(), _p(dgate),
         Bt, R, K, N, _st(), nbytes=(A.numel() + dC.numel()) * 2 + N * K * 4)
    return dW, dgate
