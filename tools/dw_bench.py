"""Micro-benchmark of the depthwise kernels on MobileNetLarge3D's layer shapes (B=64 clips).
usage: python tools/dw_bench.py [layer-substring] [--reps N] [--only fwd|dgrad|wgrad]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from picklebot_b200 import ops

LAYERS = {  # name: (C, T, H, W, k, s)
    "2.0": (16, 8, 112, 112, 3, 1), "2.1": (64, 10, 112, 112, 3, 2), "2.2": (72, 6, 56, 56, 3, 1),
    "3.0": (72, 8, 56, 56, 5, 2), "3.1": (120, 6, 28, 28, 5, 1), "3.2": (120, 10, 28, 28, 5, 1),
    "4.0": (240, 14, 28, 28, 3, 2), "4.1": (240, 8, 14, 14, 3, 1), "4.2": (184, 10, 14, 14, 3, 1),
    "4.3": (184, 12, 14, 14, 3, 1), "4.4": (480, 14, 14, 14, 3, 1), "4.5": (672, 16, 14, 14, 3, 1),
    "5.0": (672, 18, 14, 14, 5, 2), "5.1": (960, 11, 7, 7, 5, 1), "5.2": (960, 15, 7, 7, 5, 1),
}

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = 10
    only = None
    for i, a in enumerate(sys.argv):
        if a == "--reps": reps = int(sys.argv[i + 1])
        if a == "--only": only = sys.argv[i + 1]
    args = [a for a in args if not a.isdigit() and a not in ("fwd", "dgrad", "wgrad")]
    sel = args[0] if args else ""
    B = 64
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    tot = {"fwd": [0, 0], "dgrad": [0, 0], "wgrad": [0, 0]}
    for name, (C, T, H, W, k, s) in LAYERS.items():
        if sel and sel not in name:
            continue
        K, S, P = (1, k, k), (s, s, s), (k // 2,) * 3
        x = torch.randn(B, T, H, W, C, device="cuda").bfloat16()
        w = torch.randn(C, 1, 1, k, k, device="cuda") * 0.2
        w_tc = ops.dw_weight_tapmajor(w, torch.bfloat16)
        y = ops.dwconv_fwd(x, w_tc, K, S, P)
        dy = torch.randn_like(y)
        nbytes = (x.numel() + y.numel()) * 2
        fns = {"fwd": lambda: ops.dwconv_fwd(x, w_tc, K, S, P),
               "dgrad": lambda: ops.dwconv_dgrad(dy, w_tc, x.shape, K, S, P),
               "wgrad": lambda: ops.dwconv_wgrad(x, dy, K, S, P)}
        line = f"{name:4s} C={C:4d} {T:2d}x{H:3d}x{W:3d} k{k}s{s} {nbytes/1e6:7.1f}MB "
        for kind, fn in fns.items():
            if only and kind != only:
                continue
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            tot[kind][0] += t; tot[kind][1] += nbytes
            line += f"| {kind} {t*1000:7.1f}us {nbytes/t/1e6:6.0f}GB/s "
        print(line)
    for kind, (t, b) in tot.items():
        if t:
            print(f"total {kind}: {t:.3f} ms, {b/t/1e6:.0f} GB/s")

if __name__ == "__main__":
    main()
