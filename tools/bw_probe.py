"""HBM probes on the bench box: copy (the MEASURED_PEAKS.json method), pure write (fill), pure read (sum), all on 2 GiB
bf16 buffers, best of 10, CUDA events.  Context for the write-dominated expand GEMMs and the depthwise kernels."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)


def best(fn, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


t = best(lambda: b.copy_(a));           print(f"copy  (read+write) {2 * n * 2 / t / 1e6:8.0f} GB/s")
t = best(lambda: b.zero_());            print(f"fill  (write only) {n * 2 / t / 1e6:8.0f} GB/s")
t = best(lambda: b.fill_(1.5));         print(f"fill  (write only) {n * 2 / t / 1e6:8.0f} GB/s")
t = best(lambda: a.float().sum() if False else torch.sum(a, dtype=torch.float32)); print(f"sum   (read only)  {n * 2 / t / 1e6:8.0f} GB/s")
t = best(lambda: torch.add(a, 1.0, out=b)); print(f"add   (read+write) {2 * n * 2 / t / 1e6:8.0f} GB/s")
