"""Projection GEMMs with and without the BatchNorm statistics in the epilogue (MobileNetLarge3D shapes, 64 clips).
usage: python tools/gemm_stats_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from picklebot_b200 import gemm_tc

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


for Bt, R, K, N in ((1, 1204224, 64, 24), (1, 1204224, 72, 24), (1, 702464, 120, 40), (1, 702464, 240, 40), (1, 175616, 184, 80),
                    (1, 175616, 240, 80), (1, 225792, 480, 112), (64, 3528, 672, 112), (64, 931, 960, 160)):
    A = torch.randn(Bt * R, K, device="cuda").bfloat16()
    W = (torch.randn(Bt, N, K, device="cuda") * 0.1).bfloat16() if Bt > 1 else (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    nbytes = A.numel() * 2 + Bt * R * N * 2
    t0 = timeit(lambda: gemm_tc.gemm(A, W, N, K, Bw=Bt, Bt=Bt))
    t1 = timeit(lambda: gemm_tc.gemm(A, W, N, K, Bw=Bt, Bt=Bt, stat_mod=N))
    print(f"Bt={Bt:2d} R={R:8d} {K:4d}->{N:4d} {nbytes/1e6:6.1f}MB | plain {t0:6.1f}us {nbytes/t0/1e3:5.0f}GB/s | with statistics {t1:6.1f}us {nbytes/t1/1e3:5.0f}GB/s (+{100*(t1/t0-1):.0f} %)")
