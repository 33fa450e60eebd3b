"""CUDA-graph capture of one training micro-batch (forward, loss, backward).

A MobileNetLarge3D micro-batch is ~370 kernel launches that run for ~14 ms on a B200, and issuing them from
Python (ctypes calls, ``torch.empty``, autograd bookkeeping) costs the host ~11 ms: one GPU is barely kept busy,
eight processes sharing one box's cores are not.  Replaying a captured graph costs the host ~0.1 ms.

Everything the modules launch is capture-safe: the C-ABI functions only enqueue kernels / memsets on the
current stream (tensor maps are encoded on the host and baked into the kernel parameters, so all buffers must
keep their addresses -- torch's graph memory pool guarantees that), the Dropout3d noise comes from torch's
graph-aware Philox generator, BatchNorm running statistics and ``num_batches_tracked`` are updated in place.

Gradients: every parameter gets a persistent ``.grad`` BEFORE capture and the captured backward adds into it in
place (one multi-tensor add), which makes replays compose with gradient accumulation; the caller zeroes the
gradients at the start of an optimizer step (``GradientBuckets.zero_grad()`` or ``optimizer.zero_grad(
set_to_none=False)`` -- never ``set_to_none=True``, the graph holds the addresses).  Hooks do not fire on replay:
under data parallelism capture inside ``GradientBuckets.no_sync()`` and exchange with ``reduce_all()`` +
``finish()`` after the last micro-batch.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F

from . import _lib


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, example_x, example_y)``; then ``loss = step(x, y)`` per micro-batch.

    ``example_x`` fixes shape, dtype and strides of the clips (e.g. the uint8 ``(B,3,T,H,W)`` view of a
    ``(B,T,H,W,3)`` batch), ``example_y`` the labels.  ``step(x, y)`` copies both into the graph's static inputs
    (device-to-device or straight from pinned host memory) and replays; the returned loss tensor is static too
    (read it before the next replay).  ``launches`` is the number of this library's kernels in the graph.

    Two graphs are captured into one memory pool.  The first re-derives the shadow copies of the weights (bf16
    casts, transposes, block-diagonal forms: ~90 small kernels for MobileNetLarge3D) and runs the pass; the second
    was captured with those copies already in place and only runs the pass (``launches_warm`` kernels).
    ``step(x, y)`` replays the first -- always correct; ``step(x, y, weights_changed=False)`` replays the second
    and is for the 2nd..nth micro-batch of a gradient-accumulation step, where no optimizer step intervened.
    """

    def __init__(self, model: torch.nn.Module, example_x: torch.Tensor, example_y: torch.Tensor,
                 loss_fn: Callable = F.cross_entropy, autocast_dtype: Optional[torch.dtype] = torch.bfloat16,
                 warmup: int = 3):
        if not example_x.is_cuda:
            raise ValueError("GraphedTrainStep: inputs must live on the GPU")
        self.model, self.loss_fn, self.autocast_dtype = model, loss_fn, autocast_dtype
        self.x = torch.empty_strided(example_x.shape, example_x.stride(), dtype=example_x.dtype, device=example_x.device)
        self.y = torch.empty_like(example_y)
        self.x.copy_(example_x)
        self.y.copy_(example_y)
        for p in model.parameters():
            if p.requires_grad and p.grad is None:
                p.grad = torch.zeros_like(p)
        self._params = [p for p in model.parameters() if p.requires_grad]
        # the warm-up passes are real training passes: put the BatchNorm running statistics back afterwards
        buffers = [(b, b.detach().clone()) for b in model.buffers()]
        # warm-up on a side stream (torch.cuda.graph's rule): lazy one-time initialisation happens here
        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        # shadow copies of the weights (bf16 casts, transposes, block-diagonal forms) are cached per parameter
        # version; start the capture cold so the casts are part of the graph and follow every optimizer step
        for mod in model.modules():
            cache = getattr(mod, "_cache", None)
            if cache is not None and hasattr(cache, "clear"):
                cache.clear()
        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.launches = _lib.launch_count() - before
        # second capture, same pool: the shadow copies made above are cached (the caches hold them, so their
        # addresses stay put) and every replay of the first graph rewrites them in place
        before = _lib.launch_count()
        self.graph_warm = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_warm, pool=self.graph.pool()):
            self.loss_warm = self._eager()
        self.launches_warm = _lib.launch_count() - before
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)
            for p in model.parameters():      # warm-up passes accumulated into the gradients
                if p.grad is not None:
                    p.grad.zero_()

    def _eager(self) -> torch.Tensor:
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                loss = self.loss_fn(self.model(self.x), self.y)
        else:
            loss = self.loss_fn(self.model(self.x), self.y)
        # gradients are taken with autograd.grad and added to .grad with one multi-tensor kernel: loss.backward()
        # would launch one small accumulation kernel per parameter (~170 for MobileNetLarge3D)
        grads = torch.autograd.grad(loss, self._params, allow_unused=True)
        pairs = [(p.grad, g) for p, g in zip(self._params, grads) if g is not None]
        torch._foreach_add_([a for a, _ in pairs], [g if g.dtype == a.dtype else g.to(a.dtype) for a, g in pairs])
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor, weights_changed: bool = True) -> torch.Tensor:
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        if weights_changed:
            self.graph.replay()
            return self.loss
        self.graph_warm.replay()
        return self.loss_warm


class GraphedForward:
    """Inference twin of ``GraphedTrainStep``: ``fwd = GraphedForward(model.eval(), example_x)``; ``logits = fwd(x)``
    copies ``x`` into the graph's static input (device-to-device or from pinned host memory) and replays the captured
    eval-mode forward under bf16 autocast.  The returned logits tensor is static (read it before the next replay)."""

    def __init__(self, model: torch.nn.Module, example_x: torch.Tensor,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 2):
        if not example_x.is_cuda:
            raise ValueError("GraphedForward: inputs must live on the GPU")
        if model.training:
            raise ValueError("GraphedForward captures an inference pass: call model.eval() first")
        self.model, self.autocast_dtype = model, autocast_dtype
        self.x = torch.empty_strided(example_x.shape, example_x.stride(), dtype=example_x.dtype, device=example_x.device)
        self.x.copy_(example_x)
        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = self._eager()
        self.launches = _lib.launch_count() - before

    @torch.no_grad()
    def _eager(self) -> torch.Tensor:
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                return self.model(self.x)
        return self.model(self.x)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.logits


class GraphedStream:
    """Causal streaming inference of ``MoViNetA2`` (``forward_stream``), one CUDA graph per chunk shape: the stream
    state (tail frames of every temporal depthwise conv, cumulative squeeze-excite / head sums and counts) lives in
    HBM and is only ever updated in place by the kernels, so ONE captured chunk step serves every chunk of every clip.

        stream = GraphedStream(model.eval(), example_chunk)      # (B,3,Tc,H,W), e.g. the uint8 view of 8 frames
        stream.reset()                                           # new clip: zero the state in place
        for chunk in clip_chunks: logits = stream(chunk)         # logits after the frames seen so far (static tensor)
    """

    def __init__(self, model: torch.nn.Module, example_chunk: torch.Tensor,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16):
        if model.training:
            raise ValueError("GraphedStream captures an inference pass: call model.eval() first")
        self.model, self.autocast_dtype = model, autocast_dtype
        self.x = torch.empty_strided(example_chunk.shape, example_chunk.stride(), dtype=example_chunk.dtype,
                                     device=example_chunk.device)
        self.x.copy_(example_chunk)
        self.state = model.init_stream_state()
        side = torch.cuda.Stream(device=example_chunk.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                       # allocates the state tensors and warms the weight caches
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = self._eager()
        self.launches = _lib.launch_count() - before
        self.reset()

    @torch.no_grad()
    def _eager(self) -> torch.Tensor:
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                return self.model.forward_stream(self.x, self.state)[0]
        return self.model.forward_stream(self.x, self.state)[0]

    def reset(self) -> None:
        self.model.reset_stream_state(self.state)

    def __call__(self, chunk: torch.Tensor) -> torch.Tensor:
        if chunk.data_ptr() != self.x.data_ptr():
            self.x.copy_(chunk, non_blocking=True)
        self.graph.replay()
        return self.logits
