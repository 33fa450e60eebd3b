"""Micro-benchmark of the pointwise-conv kernels (tcgen05 GEMM and weight gradient) on given shapes.
usage: python tools/pw_bench.py [wgrad|gemm] Bt,R,K,N [Bt,R,K,N ...] [--reps N]
R is rows per batch entry (Bt=1: all rows).  Inputs are flushed from L2 between repetitions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from picklebot_b200 import gemm_tc

# MobileNetLarge3D at 64 clips: the (Bt, R, K, N) list of the per-block weight gradients
LARGE_WGRAD = [(1, 6422528, 16, 16), (1, 6422528, 16, 64), (1, 1204224, 64, 24), (1, 1204224, 24, 72),
               (1, 1204224, 72, 24), (1, 301056, 72, 40), (1, 702464, 40, 120), (1, 702464, 120, 40),
               (1, 702464, 40, 240), (1, 175616, 240, 80), (1, 175616, 80, 184), (1, 175616, 184, 80),
               (1, 225792, 80, 480), (1, 225792, 112, 672), (64, 3528, 672, 112), (64, 931, 960, 160)]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = 10
    for i, a in enumerate(sys.argv):
        if a == "--reps":
            reps = int(sys.argv[i + 1])
            args.remove(sys.argv[i + 1])
    kind = args[0] if args and args[0] in ("wgrad", "gemm") else "wgrad"
    shapes = [tuple(int(v) for v in a.split(",")) for a in args if "," in a] or LARGE_WGRAD
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for Bt, R, K, N in shapes:
        A = torch.randn(Bt * R, K, device="cuda").bfloat16()
        if kind == "wgrad":
            dC = torch.randn(Bt * R, N, device="cuda").bfloat16()
            fn = lambda: gemm_tc.wgrad(A, dC, K, N, Bt=Bt)
            nbytes = (A.numel() + dC.numel()) * 2
        else:
            W = torch.randn(N, K, device="cuda") * 0.1
            Wb = W.bfloat16()
            fn = lambda: gemm_tc.gemm(A, Wb, N, K)
            nbytes = (A.numel() + Bt * R * N) * 2
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"{kind} Bt={Bt:3d} R={R:8d} K={K:4d} N={N:4d} {nbytes/1e6:7.1f}MB {t*1000:8.1f}us {nbytes/t/1e6:6.0f}GB/s")


if __name__ == "__main__":
    main()
