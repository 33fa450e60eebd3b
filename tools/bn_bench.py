"""Micro-benchmark of the BatchNorm row-streaming kernels on MobileNetLarge3D's shapes at 64 clips (L2 flushed
between launches).  usage: python tools/bn_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from picklebot_b200 import ops

B = 64
# (rows per clip R, channels C, act): BN after pw2 of each block + stem + tail
SHAPES = [(8 * 112 * 112, 16, "hswish"), (10 * 112 * 112, 16, "relu"), (6 * 56 * 56, 24, "relu"), (8 * 56 * 56, 24, "relu"),
          (8 * 28 * 28, 40, "relu"), (14 * 28 * 28, 40, "relu"), (10 * 14 * 14, 80, "hswish"), (14 * 14 * 14, 80, "hswish"),
          (16 * 14 * 14, 112, "hswish"), (18 * 14 * 14, 112, "hswish"), (11 * 7 * 7, 160, "hswish"), (19 * 7 * 7, 160, "hswish"),
          (19 * 7 * 7, 960, "hswish")]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    return sorted(ts)[len(ts) // 2]


tot = [0.0, 0.0, 0.0]
for R, C, act in SHAPES:
    M = B * R
    z = torch.randn(M, C, device="cuda").bfloat16()
    dout = torch.randn(M, C, device="cuda").bfloat16()
    scale, shift = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    mean, invstd = torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5
    mask = (torch.rand(B, C, device="cuda") > 0.2).float() * 1.25
    code = ops.ACT_CODES[act]
    lib = ops._lib.lib()
    sums = torch.empty(ops._lib.STAT_REPLICAS * 2 * C, dtype=torch.float64, device="cuda")
    coef = torch.rand(2 * C, device="cuda") * 0.01
    out = torch.empty_like(z)
    st = torch.cuda.current_stream().cuda_stream
    f_fwd = lambda: ops.call("pb_bn_act_fwd", z.data_ptr(), scale.data_ptr(), shift.data_ptr(), mask.data_ptr(), out.data_ptr(), ops.PB_BF16, B, R, C, code, 0.0, st)
    f_red = lambda: ops.call("pb_bn_act_bwd_reduce", dout.data_ptr(), 0, z.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), mask.data_ptr(), sums.data_ptr(), ops.PB_BF16, B, R, C, code, 0.0, st)
    f_app = lambda: ops.call("pb_bn_act_bwd_apply", dout.data_ptr(), 0, z.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), mask.data_ptr(), coef.data_ptr(), out.data_ptr(), ops.PB_BF16, B, R, C, code, 0.0, st)
    t = [timeit(f) for f in (f_fwd, f_red, f_app)]
    nb = [2 * z.numel() * 2, 2 * z.numel() * 2, 3 * z.numel() * 2]
    for i in range(3):
        tot[i] += t[i]
    print(f"C={C:4d} rows={M:8d} {z.numel()*2/1e6:7.1f}MB/tensor | fwd {t[0]:7.1f}us {nb[0]/t[0]/1e3:6.0f}GB/s | reduce {t[1]:7.1f}us {nb[1]/t[1]/1e3:6.0f}GB/s | apply {t[2]:7.1f}us {nb[2]/t[2]/1e3:6.0f}GB/s")
print(f"total: fwd {tot[0]:.0f} us, reduce {tot[1]:.0f} us, apply {tot[2]:.0f} us")
