"""Multi-tensor AdamW: the optimiser step that follows backward in the reference's
loop (train.py:208-212, 283-289), as ONE kernel launch over all parameters (``pb_adamw_step``).

Same arithmetic and state layout as ``torch.optim.AdamW`` (decoupled weight decay, bias-corrected moments, fp32
``exp_avg`` / ``exp_avg_sq``, ``step``), same ``torch.optim.Optimizer`` interface (``param_groups``,
``state_dict`` / ``load_state_dict``, ``zero_grad``), so it drops into the reference's loop; ``grad_scale``
covers the reference's ``GradScaler.unscale_`` for fp16 runs.  The reference's own choice, bitsandbytes'
``AdamW8bit``, is not installable here and is lossy by design; this is the fp32-state optimiser it approximates.

Device tables (tensor addresses, sizes, chunk list) are built once per parameter group and refreshed only when an
address changes (e.g. after ``zero_grad(set_to_none=True)`` allocated new gradients).
"""
from __future__ import annotations

import math
from typing import Iterable

import torch

from . import _lib
from .ops import call, _st


class AdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("picklebot_b200.optim.AdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}          # group index -> (address tuple, device tensors)

    def _group_tables(self, gi: int, plist):
        addrs = []
        for p in plist:
            st = self.state[p]
            addrs += [p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()]
        key = tuple(addrs)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        dev = plist[0].device
        n = len(plist)
        chunk = int(_lib.lib().pb_adamw_chunk_elems())
        ptrs = [0] * (4 * n)
        sizes, chunk_tensor, chunk_start = [], [], []
        for i, p in enumerate(plist):
            st = self.state[p]
            ptrs[i], ptrs[n + i] = p.data_ptr(), p.grad.data_ptr()
            ptrs[2 * n + i], ptrs[3 * n + i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            sizes.append(p.numel())
            for s in range(0, p.numel(), chunk):
                chunk_tensor.append(i)
                chunk_start.append(s)
        tabs = (torch.tensor(ptrs, dtype=torch.int64, device=dev), torch.tensor(sizes, dtype=torch.int64, device=dev),
                torch.tensor(chunk_tensor, dtype=torch.int32, device=dev),
                torch.tensor(chunk_start, dtype=torch.int64, device=dev), n, len(chunk_tensor))
        self._tables[gi] = (key, tabs)
        return tabs

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("picklebot_b200.optim.AdamW needs fp32 CUDA parameters and gradients "
                                       "(there is no CPU fallback)")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("picklebot_b200.optim.AdamW: parameters and gradients must be contiguous")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            # one step counter per group (all tensors of a group advance together)
            steps = {int(self.state[p]["step"]) for p in plist}
            if len(steps) != 1:
                raise RuntimeError("picklebot_b200.optim.AdamW: parameters of one group have different step counts")
            t = steps.pop() + 1
            for p in plist:
                self.state[p]["step"] = t
            beta1, beta2 = group["betas"]
            ptrs, sizes, chunk_tensor, chunk_start, n, n_chunks = self._group_tables(gi, plist)
            call("pb_adamw_step", ptrs.data_ptr(), sizes.data_ptr(), chunk_tensor.data_ptr(), chunk_start.data_ptr(),
                 n, n_chunks, float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                 float(group["weight_decay"]), float(1.0 - beta1 ** t), float(math.sqrt(1.0 - beta2 ** t)),
                 float(grad_scale), _st())
            # The kernel wrote the parameters through raw pointers: tell torch (and blocks.WeightCache, which keys the
            # bf16 / transposed / tap-major shadow copies on Parameter._version) that they changed.
            torch.autograd.graph.increment_version(plist)
        return loss
