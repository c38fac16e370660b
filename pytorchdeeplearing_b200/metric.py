"""Drop-in per-step accuracy functions of the reference ``model/metric.py:146-181`` (called every training /
validation step: model/modelVNet.py:582, modelUnet.py:136,166): ONE pass over ``probs`` + labels producing the
per-sample threshold counts (``b200seg_metric_partials``) and a scalar finalize, instead of ~8 elementwise / reduce
launches and an int64 one-hot tensor.  (Inside a training step the same numbers come for free from the loss pass:
``lossfn.last_dice()`` / ``GraphedStep.dice`` -- SURVEY.md 8f-2.)
"""
from __future__ import annotations

import torch

from . import runtime


def _channels_last(t: torch.Tensor) -> torch.Tensor:
    """(N, C, *spatial) -> contiguous (N, *spatial, C) fp32 (free for the tensors the drop-in networks return)"""
    perm = (0,) + tuple(range(2, t.dim())) + (1,)
    t = t.detach().permute(*perm)
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _sums(input: torch.Tensor, target: torch.Tensor, binary: bool) -> torch.Tensor:
    if binary:
        n = target.size(0)
        p = input.detach().reshape(n, -1, 1)                     # dice_coeff views (num, -1): metric.py:150
        if p.dtype != torch.float32:
            p = p.float()
        p = p.contiguous()
        t = target.detach().reshape(n, -1)
        t = t.float() if t.is_floating_point() else t.long()
        c = 1
    else:
        p = _channels_last(input)
        n, c = p.shape[0], p.shape[-1]
        t = target.detach().long()
    t = t.contiguous()
    metric = torch.zeros(n, c, 3, dtype=torch.float64, device=p.device)
    runtime.get_backend(p).metric_partials(p, t, 0.5, metric)
    return metric


def _finish(metric: torch.Tensor, which: int) -> torch.Tensor:
    out = torch.empty(2, dtype=torch.float32, device=metric.device)
    runtime.get_backend(metric).metric_finalize(metric, out)
    return out[which]


def dice_coeff(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """reference model/metric.py:146-155: per-sample Dice of ``input > 0.5`` against ``target``, batch mean"""
    return _finish(_sums(input, target, True), 0)


def iou_coeff(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """reference model/metric.py:158-167"""
    return _finish(_sums(input, target, True), 1)


def multiclass_dice_coeff(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """reference model/metric.py:170-181: mean over the non-background classes of ``dice_coeff`` on each class's
    probability map (threshold 0.5) against the one-hot labels"""
    if input.shape[1] < 2:
        raise ValueError("multiclass_dice_coeff needs at least two classes")
    return _finish(_sums(input, target, False), 0)


def multiclass_iou_coeff(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """the IoU analogue (reference model/metric.py:205-216; its ``assert input.size() == target.size()`` makes the
    original unusable with index labels -- the formula is the one of ``iou_coeff`` per class)"""
    if input.shape[1] < 2:
        raise ValueError("multiclass_iou_coeff needs at least two classes")
    return _finish(_sums(input, target, False), 1)


__all__ = ["dice_coeff", "iou_coeff", "multiclass_dice_coeff", "multiclass_iou_coeff"]
