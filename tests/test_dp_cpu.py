"""World-size-2 tests of the data-parallel host logic on CPU (gloo): sharding, bucketed asynchronous gradient
averaging with accumulation, parameter/buffer broadcast.  Emulation oracle: split the
batch into N shards, run each through the same model separately, average the gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from picklebot_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(12, 300), torch.nn.BatchNorm1d(300), torch.nn.ReLU(),
                               torch.nn.Linear(300, 700), torch.nn.ReLU(), torch.nn.Linear(700, 5))


def _data(n=16):
    g = torch.Generator().manual_seed(5)
    return torch.rand(n, 12, generator=g), torch.randint(0, 5, (n,), generator=g)


def _worker(rank, world, port, accum, out, views=False, deferred=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        model = _model()
        if rank != 0:                                   # ranks start different; broadcast must fix that
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        dp.broadcast_module(model)
        x, y = _data()
        micro = 16 // (world * accum)
        buckets = dp.GradientBuckets(model.parameters(), bucket_cap_mb=0.5, first_bucket_mb=0.01,
                                     grad_as_bucket_view=views)
        assert len(buckets.buckets) >= 3
        ranges = dp.shard_range(rank, world, 16, micro)
        for step in range(2 if views else 1):           # second step: zero_grad() keeps the views and re-arms
            buckets.zero_grad()
            for i, (a, b) in enumerate(ranges):
                loss = torch.nn.functional.cross_entropy(model(x[a:b]), y[a:b]) / accum
                if deferred or i + 1 < len(ranges):     # deferred: gradients produced without the hooks firing
                    with buckets.no_sync():
                        loss.backward()
                else:
                    loss.backward()
            if deferred:
                buckets.reduce_all()
            buckets.finish()
        if views:
            for b, flat in zip(buckets.buckets, buckets._flat):
                assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in b)
        if rank == 0:
            torch.save({k: p.grad.clone() for k, p in model.named_parameters()}, out)
        # a second step must work too (buckets re-armed), and all ranks must hold identical gradients
        flat = torch.cat([p.grad.flatten() for p in model.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(flat, ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("accum,views,deferred", [(1, False, False), (2, False, False), (2, True, False), (1, True, True),
                                                  (2, True, True)])
def test_bucketed_gradient_average_matches_shard_emulation(tmp_path, accum, views, deferred):
    """views: .grad tensors are views into the bucket buffers (no pack/unpack); deferred: all buckets reduced by
    reduce_all() after the backward passes (how the captured-graph training step exchanges gradients)."""
    world = 2
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(world, _free_port(), accum, out, views, deferred), nprocs=world, join=True)
    got = torch.load(out)
    # emulation: every (rank, micro-batch) shard through the same model separately, local mean loss / accum,
    # gradients summed over micro-batches and averaged over ranks (BatchNorm statistics are per shard)
    x, y = _data()
    micro = 16 // (world * accum)
    want = None
    for rank in range(world):
        model = _model()
        for a, b in dp.shard_range(rank, world, 16, micro):
            (torch.nn.functional.cross_entropy(model(x[a:b]), y[a:b]) / accum).backward()
        g = {k: p.grad.clone() / world for k, p in model.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-7), k


def test_shard_range():
    assert dp.shard_range(0, 1, 512, 64) == [(i, i + 64) for i in range(0, 512, 64)]
    assert dp.shard_range(3, 8, 512, 64) == [(192, 256)]
    assert dp.shard_range(1, 2, 512, 64) == [(256 + i, 320 + i) for i in range(0, 256, 64)]
    covered = sorted(r for rank in range(4) for r in dp.shard_range(rank, 4, 512, 64))
    assert covered == [(i, i + 64) for i in range(0, 512, 64)]
    with pytest.raises(ValueError):
        dp.shard_range(0, 3, 512, 64)
    with pytest.raises(ValueError):
        dp.shard_range(0, 2, 512, 96)


def _eval_worker(rank, world, port, out):
    import torch.nn as nn
    import torch.nn.functional as F
    from picklebot_b200.evalloop import estimate_loss
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        torch.manual_seed(0)
        model = nn.Sequential(nn.Flatten(), nn.Linear(3 * 2 * 4 * 4, 5))            # stands in for the CUDA models
        g = torch.Generator().manual_seed(1)
        data = [(torch.randint(0, 256, (6, 2, 4, 4, 3), generator=g, dtype=torch.uint8),
                 torch.randint(0, 5, (6, 1), generator=g)) for _ in range(4)]
        shard = data[rank::world]                                                    # DistributedSampler-like split
        kw = dict(use_autocast=False, feature_dtype=torch.float32)
        local = estimate_loss(model, shard, F.cross_entropy, "cpu", all_reduce=False, **kw)
        glob = estimate_loss(model, shard, F.cross_entropy, "cpu", all_reduce=True, **kw)
        full = estimate_loss(model, data, F.cross_entropy, "cpu", all_reduce=False, **kw)
        torch.save((local, glob, full), f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_estimate_loss_all_reduced_over_ranks(tmp_path):
    """evalloop.estimate_loss: per-rank numbers are the reference's (train.py:123-153 on that rank's shard); the
    all-reduced numbers equal a single-process pass over the whole validation set."""
    out = str(tmp_path / "eval")
    mp.spawn(_eval_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (l0, g0, f0), (l1, g1, f1) = torch.load(out + ".0"), torch.load(out + ".1")
    assert g0 == pytest.approx(g1) and g0 == pytest.approx(f0, rel=1e-6)
    assert l0 != pytest.approx(l1)                       # the shards differ, so do the reference-style numbers
    assert g0[0] == pytest.approx((l0[0] + l1[0]) / 2, rel=1e-6)
