"""Cross-entropy criterion with the accuracy count: ``nn.CrossEntropyLoss`` of
train.py:214,266-267 and ``calculate_accuracy`` of train.py:110-114 in one kernel, no host synchronisation.

``loss = cross_entropy(logits, labels)`` is differentiable w.r.t. the logits (the kernel already produced the
gradient); ``cross_entropy_with_accuracy`` additionally returns the int32 device scalar of correct argmax calls,
so the training loop can accumulate accuracy without the two ``.item()`` syncs per micro-batch of the reference.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .ops import _st, call


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, scale: float):
        if not logits.is_cuda:
            raise RuntimeError("picklebot_b200.loss: CUDA tensors only (there is no CPU fallback)")
        if logits.dim() != 2 or labels.shape != (logits.shape[0],) or labels.dtype != torch.int64:
            raise ValueError("picklebot_b200.loss: logits must be [B][classes], labels int64 [B]")
        lg = logits.detach().float().contiguous()
        B, NC = lg.shape
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        dlogits = torch.empty_like(lg)
        correct = torch.empty((), dtype=torch.int32, device=lg.device)
        call("pb_ce_loss", lg.data_ptr(), labels.contiguous().data_ptr(), loss.data_ptr(), dlogits.data_ptr(),
             correct.data_ptr(), B, NC, float(scale), _st())
        ctx.save_for_backward(dlogits)
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(correct)
        return loss, correct

    @staticmethod
    def backward(ctx, grad_loss, _grad_correct):
        (dlogits,) = ctx.saved_tensors
        return (dlogits * grad_loss).to(ctx.in_dtype), None, None


def cross_entropy_with_accuracy(logits: torch.Tensor, labels: torch.Tensor, scale: float = 1.0
                                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scale * mean cross-entropy, number of correct argmax calls as an int32 device scalar)."""
    return _CrossEntropyFn.apply(logits, labels, scale)


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    return _CrossEntropyFn.apply(logits, labels, scale)[0]
