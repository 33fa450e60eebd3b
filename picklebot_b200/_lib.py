"""ctypes binding of libpicklebot_b200.so (C ABI declared in include/picklebot_b200.h).

The product path has no CPU or library fallback: if the shared library is missing, or a call
returns a non-zero code, this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpicklebot_b200.so")

PB_OK = 0
PB_F32, PB_BF16, PB_U8, PB_F32_RBF16 = 0, 1, 2, 16
STAT_REPLICAS = 16   # PB_STAT_REPLICAS in include/picklebot_b200.h
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_LRELU, ACT_HSIGMOID = 0, 1, 2, 3, 4

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
_T = {"p": _P, "i": _I, "l": _L, "f": _F}

# name -> argument type string (p = pointer, i = int, l = long long, f = float); all return int
_DW = "pppi" + "i" * 17 + "p"
SIGNATURES = {
    "pb_dwconv3d_fwd": _DW,
    "pb_dwconv3d_fwd_pool": "ppppi" + "i" * 17 + "p",
    "pb_dwconv3d_dgrad": _DW,
    "pb_dwconv3d_wgrad": _DW,
    "pb_stream_dwconv3d_fwd": "pppppi" + "i" * 14 + "p",
    "pb_pw_gemm_simt": "ppllpppppiiliip",
    "pb_pw_gemm_tc": "ppipppp" + "pi" + "iliip",
    "pb_pw_gemm_tc_act": "ppipppp" + "pi" + "ilii" + "if" + "p",
    "pb_pw_wgrad_simt": "pppppiiliip",
    "pb_pw_wgrad_tc": "pppppppiliip",
    "pb_cast_matrix": "ppiiiip",
    "pb_fold_gate_bf16": "pppiiip",
    "pb_fold_gate_t_bf16": "pppiiip",
    "pb_fold_rows_bf16": "pppiiip",
    "pb_fold_scaled_bf16": "ppppiiip",
    "pb_block_diag_bf16": "ppiiip",
    "pb_colstats": "pilipp",
    "pb_bn_finalize": "plppppiffpppppip",
    "pb_bn_act_fwd": "pppppiiliifp",
    "pb_bn_act_bwd_reduce": "pippppppp" + "iiliifp",
    "pb_bn_bwd_finalize": "plipppip",
    "pb_bn_act_bwd_apply": "pipppppppp" + "iiliifp",
    "pb_pool_fwd": "piilipp",
    "pb_stream_pool_update": "plpppiip",
    "pb_fc_fwd": "ppppiiip",
    "pb_fc_dgrad": "pppiiifp",
    "pb_se_fc_fwd": "pppppppiiip",
    "pb_se_fc_bwd": "ppppppfpppppp" + "iiip",
    "pb_rowscale": "pppiilip",
    "pb_rowdot": "ppiilipp",
    "pb_scale_add": "pppiilip",
    "pb_stem_conv_fwd": "pi" + "lllll" + "f" + "pppi" + "i" * 18 + "p",
    "pb_stem_conv_fwd_act": "pi" + "lllll" + "f" + "pppi" + "i" * 18 + "if" + "p",
    "pb_stem_conv_wgrad": "pi" + "lllll" + "f" + "pipp" + "i" * 18 + "p",
    "pb_adamw_step": "pppp" + "ii" + "ffffffff" + "p",
    "pb_ce_loss": "ppppp" + "iif" + "p",
}
PLAIN = ("pb_abi_version", "pb_last_error_string", "pb_launch_count", "pb_device_check",
         "pb_pw_wgrad_tc_workspace_bytes", "pb_adamw_chunk_elems", "pb_path_count", "pb_path_reset")
# PB_PATH_* of include/picklebot_b200.h, in enum order
PATHS = ("dw_fwd_tma", "dw_fwd_generic", "dw_dgrad_tma", "dw_dgrad_generic", "dw_wgrad_tma", "dw_wgrad_generic",
         "gemm_tc", "gemm_simt", "wgrad_tc", "wgrad_simt", "stem_tc", "stem_simt", "dw_bwd_fused_tma",
         "dw_stream_tma", "dw_stream_generic", "stem_tma")
EXPORTS = tuple(SIGNATURES) + PLAIN


class PicklebotKernelError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python picklebot_b200/csrc/build.py` (or __graft_entry__.build()). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, sig in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = [_T[c] for c in sig]
        fn.restype = _I
    lib.pb_abi_version.restype = _I
    lib.pb_last_error_string.restype = ctypes.c_char_p
    lib.pb_launch_count.restype = _L
    lib.pb_device_check.restype = _I
    lib.pb_pw_wgrad_tc_workspace_bytes.argtypes = [_I, _L, _I, _I]
    lib.pb_pw_wgrad_tc_workspace_bytes.restype = _L
    lib.pb_adamw_chunk_elems.restype = _I
    lib.pb_path_count.argtypes = [_I]
    lib.pb_path_count.restype = _L
    lib.pb_path_reset.argtypes = []
    lib.pb_path_reset.restype = None
    return lib


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


class KernelProfiler:
    """Optional per-launch timing: CUDA events recorded on the launching (current torch) stream around
    every C-ABI call, with the algorithmic bytes the caller attributes to it.  Used by bench.py for the
    live roofline numbers; off (None) by default so the hot path pays nothing."""

    def __init__(self):
        self.records = []          # (name, start_event, end_event, algorithmic_bytes, tag, written_bytes)

    def summary(self, by_tag: bool = False):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, nbytes, tag, wbytes in self.records:
            d = out.setdefault(name if not by_tag else f"{name}|{tag}",
                               {"launches": 0, "ms": 0.0, "bytes": 0, "wbytes": 0, "per_launch": []})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
            d["wbytes"] += wbytes
            d["per_launch"].append((nbytes, wbytes))
        return out


PROFILER = None


def call(name: str, *args, nbytes: int = 0, tag: str = "", wbytes: int = 0) -> None:
    prof = PROFILER
    if prof is not None:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib(), name)(*args)
    if rc != PB_OK:
        msg = lib().pb_last_error_string()
        raise PicklebotKernelError(f"{name} failed (code {rc}): {msg.decode() if msg else '?'}")
    if prof is not None:
        e1.record()
        prof.records.append((name, e0, e1, nbytes, tag, wbytes))


def launch_count() -> int:
    return int(lib().pb_launch_count())


def path_counts() -> dict:
    """Calls served by each kernel family since load / the last ``path_reset()`` (``PATHS`` names):
    tests use it to assert that the TMA / tcgen05 production kernels ran, not the CUDA-core correctness paths."""
    return {name: int(lib().pb_path_count(i)) for i, name in enumerate(PATHS)}


def path_reset() -> None:
    lib().pb_path_reset()
