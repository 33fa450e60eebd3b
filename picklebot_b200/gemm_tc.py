"""Launchers for the tcgen05 / TMEM / TMA GEMM kernels (csrc/pwgemm_tc.cu)."""
from __future__ import annotations

import os

import torch

from . import ops
from .ops import _chk, _p, _st, call

# The tensor-core path is the production path for bf16; PB_GEMM=simt (see ops.py) turns it off.
ops._TC_READY = os.environ.get("PB_TC", "1") != "0"
_WGRAD_READY = os.environ.get("PB_TC_WGRAD", "1") != "0"


def gemm(A: torch.Tensor, Wb: torch.Tensor, N: int, K: int, Bw: int = 1, Bt: int = 1, bias=None, colscale=None,
         coladd=None) -> torch.Tensor:
    """A bf16 [Bt][R][K] (contiguous), Wb bf16 [Bw][N][K] -> bf16 [Bt*R][N]."""
    _chk(A, "gemm_tc.A"); _chk(Wb, "gemm_tc.W")
    assert A.dtype == torch.bfloat16 and Wb.dtype == torch.bfloat16
    rows = A.numel() // K
    R = rows // Bt
    C = torch.empty((rows, N), dtype=torch.bfloat16, device=A.device)
    call("pb_pw_gemm_tc", A.data_ptr(), Wb.data_ptr(), Bw, _p(bias), _p(colscale), _p(coladd), C.data_ptr(),
         Bt, R, K, N, _st(), nbytes=(A.numel() + C.numel() + Wb.numel()) * 2)
    return C


def wgrad_ready() -> bool:
    return False
