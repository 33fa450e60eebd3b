"""Kernel timeline of one captured MobileNetLarge3D micro-batch (graph replay) from CUPTI via torch.profiler:
how much of the replay is kernels running back to back, how much is gaps, and which kernels the gaps follow.
usage: python tools/graph_timeline.py [--model MobileNetLarge3D] > gpurun_out/graph_timeline.txt"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import picklebot_b200 as pb
from picklebot_b200 import loss as pbloss, synth
from picklebot_b200.graph import GraphedTrainStep

name = sys.argv[sys.argv.index("--model") + 1] if "--model" in sys.argv else "MobileNetLarge3D"
dev = torch.device("cuda")
torch.manual_seed(0)
model = getattr(pb, name)(num_classes=2).to(dev).train()
B = 64
clips = synth.synthetic_clips_u8_device(B, 16, 224, 224, seed=1, device=dev)
labels = synth.synthetic_labels(B, 2, seed=2).to(dev)
step = GraphedTrainStep(model, clips.permute(0, 4, 1, 2, 3), labels,
                        loss_fn=lambda lg, y: pbloss.cross_entropy(lg, y, scale=0.125))
for _ in range(3):
    step(clips.permute(0, 4, 1, 2, 3), labels, weights_changed=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step(clips.permute(0, 4, 1, 2, 3), labels, weights_changed=False)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "Memcpy" not in e.name
      and "Memset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)
# second replay only
half = len(ev) // 2
ev = ev[half:]
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = 0.0
cur_end = t0
gaps = collections.defaultdict(lambda: [0, 0.0])
durs = collections.defaultdict(lambda: [0, 0.0])
overlap = 0.0
for i, e in enumerate(ev):
    s, en = e.time_range.start, e.time_range.end
    short = e.name.split("(")[0].split("<")[0].replace("void ", "").replace("pb::", "").replace("tc::", "")[:40]
    durs[short][0] += 1
    durs[short][1] += en - s
    if s > cur_end:
        prev = ev[i - 1].name.split("(")[0].split("<")[0].replace("void ", "").replace("pb::", "").replace("tc::", "")[:40] if i else "-"
        gaps[prev + " -> " + short][0] += 1
        gaps[prev + " -> " + short][1] += s - cur_end
    else:
        overlap += min(en, cur_end) - s
    busy += max(0.0, en - max(s, cur_end))
    cur_end = max(cur_end, en)
span = t1 - t0
print(f"{name}: {len(ev)} kernels in one replay, span {span/1000:.3f} ms, union of kernel time {busy/1000:.3f} ms, "
      f"gaps {(span-busy)/1000:.3f} ms ({100*(span-busy)/span:.1f} %), sum of durations {sum(v[1] for v in durs.values())/1000:.3f} ms, "
      f"overlap (PDL prologues) {overlap/1000:.3f} ms")
print("\nkernel families by time:")
for k, v in sorted(durs.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {v[1]/1000:8.3f} ms {v[0]:4d} x {v[1]/v[0]:8.1f} us  {k}")
print("\nlargest gap sources (previous kernel -> next kernel):")
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {v[1]:8.1f} us {v[0]:4d} x {v[1]/v[0]:6.2f} us  {k}")
