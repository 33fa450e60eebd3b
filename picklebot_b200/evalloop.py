"""Validation loop: ``estimate_loss`` of the reference (train.py:123-153) without its
per-batch host synchronisations, and with the all-reduce the reference lacks.

The reference accumulates ``criterion(outputs, labels).item()`` and a correct-call count batch by batch (one
``.item()`` sync per batch), divides the loss sum by ``len(val_loader)`` and -- under DDP -- reports rank 0's shard
only (train.py:305-313: "val loss"/"val accuracy" are per-rank numbers).  Here the three accumulators (loss sum,
correct calls, samples) stay on the device, are summed over the ranks with ONE all-reduce at the end
(``all_reduce=True``), and only then read back.  With ``all_reduce=False`` and a CUDA model the numbers are the
reference's, rank by rank.

``criterion`` may be ``picklebot_b200.loss.cross_entropy_with_accuracy`` (loss and correct calls from one kernel,
``pb_ce_loss``) or any callable ``(logits, labels) -> loss`` (the accuracy is then counted with ``torch.max`` like
train.py:110-114).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def extract_features_labels(batch, device, dtype: Optional[torch.dtype] = None):
    """train.py:102-108: uint8 ``(B,T,H,W,C)`` clips -> ``(B,C,T,H,W)`` view on ``device``.  With ``dtype=None`` the
    raw uint8 view is returned: the stem kernels of this package divide by 255 themselves (one read of 2.4 MB per
    clip instead of a 4.8 MB bf16 copy); with a dtype it is the reference's ``.to(dtype) / 255``."""
    features = batch[0].to(device, non_blocking=True).permute(0, -1, 1, 2, 3)
    if dtype is not None:
        features = features.to(dtype) / 255
    labels = batch[1].to(device, non_blocking=True).to(torch.long).view(-1)
    return features, labels


@torch.no_grad()
def estimate_loss(model: torch.nn.Module, val_loader: Iterable, criterion: Callable, device,
                  use_autocast: bool = True, dtype: torch.dtype = torch.bfloat16, all_reduce: bool = True,
                  group=None, feature_dtype: Optional[torch.dtype] = None) -> Tuple[float, float]:
    """Returns ``(val_loss, val_accuracy)``: mean of the per-batch losses and correct / samples, over this rank's
    shard (``all_reduce=False``, the reference's numbers) or over all ranks' shards."""
    was_training = model.training
    model.eval()
    dev = torch.device(device)
    acc = torch.zeros(4, dtype=torch.float64, device=dev)          # loss sum, correct, samples, batches
    ac_type = dev.type
    for batch in val_loader:
        features, labels = extract_features_labels(batch, dev, feature_dtype)
        with torch.autocast(ac_type, dtype=dtype, enabled=use_autocast):
            outputs = model(features)
            res = criterion(outputs, labels)
        if isinstance(res, tuple):                                  # (loss, correct) from pb_ce_loss
            loss, correct = res
        else:
            loss = res
            correct = (outputs.argmax(dim=1) == labels).sum()       # calculate_accuracy, train.py:110-114
        acc[0] += loss.detach().double()
        acc[1] += correct.double()
        acc[2] += labels.shape[0]
        acc[3] += 1
    if all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    loss_sum, correct, samples, batches = acc.tolist()              # the only host synchronisation
    if was_training:
        model.train()
    return loss_sum / max(batches, 1.0), correct / max(samples, 1.0)
