"""N-rank check of the data-parallel training step on real GPUs (NCCL): the gradients after one optimizer step's
worth of graph replays + the bucketed SUM all-reduce (1/world folded into the loss scale, GradientBuckets
average=False -- what bench.py runs) must equal the single-process emulation of SURVEY section 8(e): every
(rank, micro-batch) shard through the same model separately, gradients averaged.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picklebot_b200 import dp, synth                      # noqa: E402
from picklebot_b200 import loss as pbloss                 # noqa: E402
from picklebot_b200.graph import GraphedTrainStep         # noqa: E402
from picklebot_b200.mobilenet import MobileNetSmall3D     # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    micro, accum, shape = 4, 2, (8, 64, 64)
    torch.manual_seed(11 + rank)                          # ranks start different: the broadcast must fix that
    model = MobileNetSmall3D(num_classes=2).to(dev).train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout3d) or isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                     # the emulation below must see the same function
    dp.broadcast_module(model)
    start = {k: v.detach().clone() for k, v in model.state_dict().items()}
    clips = {(r, a): synth.synthetic_clips_u8_device(micro, *shape, seed=1000 * r + a, device=dev)
             for r in range(world) for a in range(accum)}
    labels = {(r, a): synth.synthetic_labels(micro, 2, seed=77 + 1000 * r + a).to(dev)
              for r in range(world) for a in range(accum)}
    scale = 1.0 / accum / world
    worst = 0.0
    # one capture per process, like bench.py (a second GraphedTrainStep after an NCCL collective trips CUDA's
    # capture check on the legacy stream inside autograd's worker thread; not a path the product takes)
    for autocast in (torch.bfloat16,):
        model.load_state_dict(start)
        buckets = dp.GradientBuckets(model.parameters(), grad_as_bucket_view=True, average=False,
                                     bucket_cap_mb=2.0, first_bucket_mb=0.25)
        with buckets.no_sync():
            step = GraphedTrainStep(model, clips[(rank, 0)].permute(0, 4, 1, 2, 3), labels[(rank, 0)],
                                    loss_fn=lambda lg, y: pbloss.cross_entropy(lg, y, scale=scale),
                                    autocast_dtype=autocast)
        model.load_state_dict(start)                      # capture warm-ups moved nothing but the BN statistics
        buckets.zero_grad()
        for a in range(accum):
            step(clips[(rank, a)].permute(0, 4, 1, 2, 3), labels[(rank, a)])
        buckets.reduce_all()
        buckets.finish()
        got = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
        ref = got.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(got, ref), "ranks hold different gradients after the all-reduce"
        buckets.remove()
        for p in model.parameters():
            p.grad = None
        # emulation on this rank: all shards, eager, same kernels
        model.load_state_dict(start)
        for r in range(world):
            for a in range(accum):
                model.load_state_dict({k: v for k, v in start.items()})       # BN statistics do not matter in train mode
                with torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
                    l = pbloss.cross_entropy(model(clips[(r, a)].permute(0, 4, 1, 2, 3)), labels[(r, a)], scale=scale)
                l.backward()
        want = torch.cat([p.grad.flatten() for p in model.parameters()])
        err = float((got - want).norm() / want.norm())
        if rank == 0:
            print(f"dp_check world={world} autocast={autocast}: gradient rel err vs shard emulation {err:.2e} "
                  f"({len(buckets.buckets)} buckets)")
        # same kernels, same inputs; 4-clip micro-batches of 64x64 clips leave BatchNorm a handful of values per
        # channel in the last stages, which amplifies summation-order noise to ~1e-2 (two eager runs of the
        # emulation differ by that much).  A wrong 1/world factor or a missing shard shows up as >= 0.3.
        assert err < 5e-2, err
        worst = max(worst, err)
        for p in model.parameters():
            p.grad = None
    dist.barrier()
    if rank == 0:
        print("dp_check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
