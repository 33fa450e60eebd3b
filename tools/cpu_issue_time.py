"""How long does the host need to ISSUE one micro-batch (forward + backward) compared with the GPU's time to run it?
usage: python tools/cpu_issue_time.py [micro_batches]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import picklebot_b200 as pb
from picklebot_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
m = pb.MobileNetLarge3D(num_classes=2).to(dev)
m.train()
x = synth.synthetic_clips_u8_device(64, 16, 224, 224, seed=0, device=dev)
y = synth.synthetic_labels(64, 2, seed=1).to(dev)
xv = x.permute(0, 4, 1, 2, 3)


def micro():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.cross_entropy(m(xv), y)
    loss.backward()


for _ in range(3):
    micro()
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        micro()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{n} micro-batches: host issue {1e3 * (t1 - t0) / n:.2f} ms each, GPU {e0.elapsed_time(e1) / n:.2f} ms each, "
          f"host wait at the end {1e3 * (t2 - t1):.1f} ms")
