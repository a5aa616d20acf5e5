"""Drop-ins for the reference's IoU helpers (src/util.py:237-308) plus exact int64 forms.

``batch_intersectionAndUnionGPU`` / ``intersectionAndUnionGPU`` keep the reference's signatures and
return float32 tensors of the same shapes; the ``*_int`` variants return the int64 counts the
kernels produce (the reference's float32 accumulation loses integers above 2**24 over a long sweep).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


def batch_intersection_union_int(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = 255):
    """logits [n_task,shot,2,h,w]; target [n_task,shot,H,W] -> int64 counts [n_task,shot,2,3] (I,U,T)
    and float64 ce [n_task,shot,2] (sum of -log p[y], #valid)."""
    n_task, shots, ncls, h, w = logits.shape
    H, W = target.shape[-2:]
    counts, ce = ops.upsample_argmax_iou(logits.reshape(n_task * shots, ncls, h, w),
                                         target.reshape(n_task * shots, H, W), ignore_index)
    return counts.view(n_task, shots, ncls, 3), ce.view(n_task, shots, 2)


def batch_intersectionAndUnionGPU(logits: torch.Tensor, target: torch.Tensor, num_classes: int,
                                  ignore_index=255) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same contract as src/util.py:237-277: returns (area_intersection, area_union, area_target),
    each float32 [n_task, shot, num_class]. Up-sampling, argmax and the histograms run in one kernel;
    nothing is materialised at H x W."""
    if logits.shape[2] != num_classes:
        raise ValueError(f"logits has {logits.shape[2]} classes, num_classes={num_classes}")
    counts, _ = batch_intersection_union_int(logits, target, ignore_index)
    c = counts.to(torch.float32)
    return c[..., 0], c[..., 1], c[..., 2]


def intersection_union_int(preds: torch.Tensor, target: torch.Tensor, num_classes: int, ignore_index: int = 255):
    assert preds.dim() in [1, 2, 3]                # src/util.py:297
    assert preds.shape == target.shape             # src/util.py:298
    p = preds.reshape(1, -1)
    t = target.reshape(1, -1)
    if p.dtype not in (torch.uint8, torch.int64):
        p = p.long()
    return ops.intersection_union(p, t.to(p.dtype), num_classes, ignore_index)[0]


def intersectionAndUnionGPU(preds: torch.Tensor, target: torch.Tensor, num_classes: int,
                            ignore_index=255) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same contract as src/util.py:280-308 (float32 [num_class] x3), including the reference's side
    effect on the caller's tensor: ``preds[target == ignore_index] = ignore_index``."""
    counts = intersection_union_int(preds, target, num_classes, ignore_index)
    preds.view(-1)[target.reshape(-1) == ignore_index] = ignore_index        # src/util.py:301
    c = counts.to(torch.float32)
    return c[:, 0], c[:, 1], c[:, 2]
