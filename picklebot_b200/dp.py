"""Data-parallel plumbing for the training step (one process per GPU, torch.distributed over NCCL).

The hot path shards by clips: every rank runs the same model on its own micro-batches, BatchNorm statistics
stay per rank (the reference uses plain BN under DDP, train.py:203-204), and the only exchange is the
average of the fp32 parameter gradients.  This module is the host-side logic of that exchange:

* ``shard_range``        which clips of a global batch a rank owns (DistributedSampler-like, train.py:59-60)
* ``GradientBuckets``    DDP-style bucketed, asynchronous all-reduce driven by post-accumulate-grad hooks, so
                         the reduction of the last blocks' gradients overlaps the backward of the first blocks
                         (our autograd nodes are per bottleneck, so gradients become ready block by block)
* ``broadcast_module``   rank 0's parameters and buffers to everybody (DDP constructor / broadcast_buffers)

It is backend agnostic: the tests run it with gloo on CPU (world size 2); bench.py can use it on NCCL
(``--dp buckets``) instead of ``torch.nn.parallel.DistributedDataParallel``.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(rank: int, world: int, global_batch: int, micro_batch: int) -> List[Tuple[int, int]]:
    """Clip index ranges [start, stop) of this rank's micro-batches inside one global batch.
    Rank r owns a contiguous slab of global_batch/world clips, cut into micro-batches."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not divide over {world} ranks")
    per_rank = global_batch // world
    if per_rank % micro_batch:
        raise ValueError(f"per-rank batch {per_rank} is not a multiple of the micro-batch {micro_batch}")
    base = rank * per_rank
    return [(base + i, base + i + micro_batch) for i in range(0, per_rank, micro_batch)]


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers (what DDP does at construction and,
    for buffers, before each forward with broadcast_buffers=True)."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


class GradientBuckets:
    """Bucketed gradient averaging.

    Parameters are assigned to buckets in reverse registration order (the order in which backward produces
    their gradients): a small first bucket (1 MiB) so communication starts early, then ``bucket_cap_mb``
    buckets, as torch DDP does.  A hook on every parameter fires after its gradient was accumulated; when a
    bucket is complete its gradients are copied into one flat fp32 buffer and an asynchronous all-reduce is
    launched.  ``finish()`` waits, divides by the world size and copies the averages back into ``.grad``.

    ``grad_as_bucket_view=True`` (what torch DDP calls gradient_as_bucket_view) makes every ``.grad`` a view into
    its bucket's flat buffer for good: autograd accumulates straight into the bucket, the all-reduce runs in
    place and nothing is packed or unpacked (~340 small copy kernels per step for MobileNetLarge3D otherwise).
    Use ``zero_grad()`` of this object then (one fill per bucket, and the views survive), not
    ``optimizer.zero_grad(set_to_none=True)``.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, bucket_cap_mb: float = 25.0,
                 first_bucket_mb: float = 1.0, grad_as_bucket_view: bool = False, average: bool = True):
        """``average=False``: the caller already folded 1/world into the loss scale, so the SUM all-reduce IS the
        average and ``finish()`` launches no division kernels."""
        self.average = average
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cap = int(first_bucket_mb * 2 ** 20)
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > cap:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
                cap = int(bucket_cap_mb * 2 ** 20)
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._work: List[Optional[object]] = [None] * len(self.buckets)
        self._sync = True
        self.grad_as_bucket_view = grad_as_bucket_view
        if grad_as_bucket_view:
            for i, bucket in enumerate(self.buckets):
                flat = torch.zeros(sum(p.numel() for p in bucket), dtype=torch.float32, device=bucket[0].device)
                self._flat[i] = flat
                off = 0
                for p in bucket:
                    view = flat[off:off + p.numel()].view_as(p)
                    if p.grad is not None:
                        view.copy_(p.grad)
                    p.grad = view
                    off += p.numel()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.reset()

    def zero_grad(self) -> None:
        """Zero all gradients; with bucket views that is one fill per bucket."""
        if self.grad_as_bucket_view:
            for flat in self._flat:
                flat.zero_()
        else:
            for p in self.params:
                if p.grad is not None:
                    p.grad.zero_()

    def reduce_all(self) -> None:
        """Launch the all-reduce of every bucket now, for callers that produced the gradients without the hooks
        firing (under ``no_sync()``, e.g. by replaying a captured CUDA graph).  Follow with ``finish()``."""
        for i in range(len(self.buckets)):
            if self._work[i] is None:
                self._pending[i] = 0
                self._launch(i)

    def reset(self) -> None:
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    @contextmanager
    def no_sync(self):
        """Gradient accumulation: micro-batches inside this context only accumulate locally."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self._sync:
            return
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def _launch(self, i: int) -> None:
        bucket = self.buckets[i]
        n = sum(p.numel() for p in bucket)
        flat = self._flat[i]
        if self.grad_as_bucket_view:
            # the views must still be what autograd accumulated into: optimizer.zero_grad(set_to_none=True) or a
            # ``p.grad = ...`` assignment silently detaches them, and the all-reduce below would then average stale
            # buffers.  Re-attach (copy the fresh gradient into the view) instead of losing the synchronisation.
            off = 0
            for p in bucket:
                view = flat[off:off + p.numel()]
                g = p.grad
                if g is None:
                    view.zero_()
                    p.grad = view.view_as(p)
                elif g.data_ptr() != view.data_ptr():
                    view.copy_(g.reshape(-1))
                    p.grad = view.view_as(p)
                off += p.numel()
        else:
            if flat is None or flat.device != bucket[0].device:
                flat = torch.empty(n, dtype=torch.float32, device=bucket[0].device)
                self._flat[i] = flat
            off = 0
            for p in bucket:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
        if self.world > 1:
            self._work[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self._work[i] = "local"

    def finish(self) -> None:
        """Wait for all buckets, write the averaged gradients back, re-arm for the next step."""
        for i, bucket in enumerate(self.buckets):
            if self._work[i] is None:
                if self._pending[i] != len(bucket) and self._pending[i] != 0:
                    raise RuntimeError("GradientBuckets.finish(): a bucket is only partially ready "
                                       "(a parameter received no gradient this step)")
                if self._pending[i] == len(bucket):
                    continue                      # nothing was reduced for this bucket (e.g. frozen part)
            w = self._work[i]
            if w is not None and w != "local":
                w.wait()
            flat = self._flat[i]
            if self.world > 1 and self.average:
                flat.div_(self.world)
            if not self.grad_as_bucket_view:
                off = 0
                for p in bucket:
                    p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                    off += p.numel()
        self.reset()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
