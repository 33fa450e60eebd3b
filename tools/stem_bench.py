"""Micro-benchmark of the stem convolution (tcgen05 implicit im2col) on the benchmark's clip shape.
usage: python tools/stem_bench.py [--reps N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from picklebot_b200 import ops, synth

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
B, T, H, W = 64, 16, 224, 224
x = synth.synthetic_clips_u8_device(B, T, H, W, seed=0, device="cuda").permute(0, 4, 1, 2, 3)   # (B,3,T,H,W) view
w = torch.randn(16, 3, 3, 3, 3, device="cuda") * 0.2
b = torch.randn(16, device="cuda")
k, s, p = (3, 3, 3), (2, 2, 2), (1, 1, 1)
y = ops.stem_fwd(x, w, b, k, s, p, torch.bfloat16)
dy = torch.randn_like(y)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
nbytes = x.numel() + y.numel() * 2
for name, fn in (("fwd", lambda: ops.stem_fwd(x, w, b, k, s, p, torch.bfloat16)),
                 ("wgrad", lambda: ops.stem_wgrad(x, dy, w.shape, k, s, p, True))):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"stem {name:5s} {nbytes/1e6:7.1f}MB {t*1000:8.1f}us {nbytes/t/1e6:6.0f}GB/s")
