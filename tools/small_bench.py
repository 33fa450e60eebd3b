"""Warm-cache timing of the small latency-bound kernels of a squeeze-excite block and of the weight-gradient
partial reduction (the shapes of MobileNetLarge3D at 64 clips).  usage: python tools/small_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from picklebot_b200 import gemm_tc, ops

dev = "cuda"
B = 64


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / reps


for C in (72, 120, 480, 672, 960):
    Ch = C // 4
    pooled = torch.rand(B, C, device=dev)
    W1, b1 = torch.randn(Ch, C, device=dev) * 0.1, torch.zeros(Ch, device=dev)
    W2, b2 = torch.randn(C, Ch, device=dev) * 0.1, torch.zeros(C, device=dev)
    t_f = timeit(lambda: ops.se_fc_fwd(pooled, W1, b1, W2, b2))
    hidden, gate = ops.se_fc_fwd(pooled, W1, b1, W2, b2)
    dgate = torch.randn(B, C, device=dev)
    t_b = timeit(lambda: ops.se_fc_bwd(dgate, pooled, hidden, gate, W1, W2, 1.0 / 100))
    print(f"SE C={C:4d}: se_fc_fwd (2 kernels) {t_f:6.1f} us   se_fc_bwd (4 kernels) {t_b:6.1f} us")
for (R, K, N, Bt) in ((702464 // 64 * 64, 40, 240, 1), (64 * 3528, 112, 672, 1), (64 * 931, 160, 960, 1), (931, 960, 160, 64),
                      (3528, 672, 112, 64)):
    A = torch.randn(Bt * R if Bt > 1 else R, K, device=dev).bfloat16()
    dC = torch.randn(A.shape[0], N, device=dev).bfloat16()
    gate = torch.rand(Bt, K, device=dev) if Bt > 1 else None
    W = torch.randn(N, K, device=dev) if Bt > 1 else None
    fn = (lambda: gemm_tc.wgrad(A, dC, K, N, gate=gate, W=W, Bt=Bt, want_dgate=True)) if Bt > 1 else (lambda: gemm_tc.wgrad(A, dC, K, N))
    t = timeit(fn, reps=20)
    print(f"wgrad_tc rows={A.shape[0]:7d} K={K:4d} N={N:4d} Bt={Bt:2d}: {t:7.1f} us total (wgrad + reduce{' + dgate' if Bt > 1 else ''}), "
          f"{(A.numel() + dC.numel()) * 2 / t / 1e3:6.0f} GB/s")
