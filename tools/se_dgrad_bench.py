"""Squeeze-excite input-gradient GEMM dy2 = (dz W2) * gate + dmean on MobileNetLarge3D's shapes at 64 clips: per-sample
folded weights with / without the accumulator pre-load, against the shared-weight GEMM of the same size.
usage: python tools/se_dgrad_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from picklebot_b200 import gemm_tc, ops

B = 64
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    return sorted(ts)[len(ts) // 2]


for R, Cout, Cexp in ((3528, 112, 672), (3136, 112, 480), (10976, 40, 120), (931, 160, 960), (735, 160, 960), (7840, 40, 120)):
    dz = torch.randn(B * R, Cout, device="cuda").bfloat16()
    W = torch.randn(Cout, Cexp, device="cuda") * 0.1
    gate = torch.rand(B, Cexp, device="cuda")
    dmean = torch.randn(B, Cexp, device="cuda")
    Wtg = ops.fold_gate_t(W, gate)
    Wt = W.t().contiguous().bfloat16()
    nbytes = (dz.numel() + B * R * Cexp) * 2
    t_shared = timeit(lambda: gemm_tc.gemm(dz, Wt, Cexp, Cout))
    t_ps = timeit(lambda: gemm_tc.gemm(dz, Wtg, Cexp, Cout, Bw=B, Bt=B))
    t_ps_add = timeit(lambda: gemm_tc.gemm(dz, Wtg, Cexp, Cout, Bw=B, Bt=B, coladd=dmean))
    t_sh_bt = timeit(lambda: gemm_tc.gemm(dz, Wt, Cexp, Cout, Bw=1, Bt=B))
    print(f"R={R:6d} {Cout:4d}->{Cexp:4d} {nbytes/1e6:6.1f}MB | shared Bt=1 {t_shared:6.1f}us {nbytes/t_shared/1e3:5.0f}GB/s | shared Bt=64 {t_sh_bt:6.1f}us | "
          f"per-sample {t_ps:6.1f}us | per-sample + preload {t_ps_add:6.1f}us {nbytes/t_ps_add/1e3:5.0f}GB/s")
