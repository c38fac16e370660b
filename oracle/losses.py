"""Closed-form CPU restatement of the eight hot-path losses of ``model/losses.py``
(test infrastructure only; formulas follow SURVEY.md App. C).

All losses cast logits to fp32 (``.float()``, model/losses.py:47,143,165,253,276,310) unless
an explicit ``dtype`` (fp64 noise-floor runs) is given.  Sums with no index run over batch
AND all spatial positions.  ``z`` = logits ``(N, C, *spatial)``, ``t`` = integer labels
``(N, *spatial)``.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SMOOTH = 1e-5   # model/losses.py:40,312
EPS = 1e-7      # model/losses.py:41,313


def _flat_binary(z: Tensor, t: Tensor, dtype):
    n, c = t.shape[0], z.shape[1]
    return z.to(dtype).reshape(n, c, -1), t.to(dtype).reshape(n, c, -1)    # losses.py:45-48


def binary_dice(z: Tensor, t: Tensor, dtype=torch.float32) -> Tensor:
    """BinaryDiceLoss.forward, model/losses.py:43-53."""
    zf, tf = _flat_binary(z, t, dtype)
    p = torch.sigmoid(zf)
    inter = (p * tf).sum()
    den = (p.sum() + tf.sum() + SMOOTH).clamp_min(EPS)
    return 1.0 - (2.0 * inter + SMOOTH) / den


def binary_bce(z: Tensor, t: Tensor, dtype=torch.float32) -> Tensor:
    """BinaryCrossEntropyLoss.forward, model/losses.py:141-147: mean(max(z,0) - z t + log1p(exp(-|z|)))."""
    zf, tf = _flat_binary(z, t, dtype)
    e = zf.clamp_min(0) - zf * tf + torch.log1p(torch.exp(-zf.abs()))
    return e.mean()


def binary_focal(z: Tensor, t: Tensor, alpha: float = 0.25, gamma: float = 2, dtype=torch.float32) -> Tensor:
    """BinaryFocalLoss.forward, model/losses.py:160-181."""
    zf, tf = _flat_binary(z, t, dtype)
    b = zf.clamp_min(0) - zf * tf + torch.log1p(torch.exp(-zf.abs()))
    pt = torch.exp(-b)
    return (alpha * (1 - pt) ** gamma * b).mean()


def binary_bce_dice(z: Tensor, t: Tensor, dtype=torch.float32) -> Tensor:
    """BinaryCrossEntropyDiceLoss.forward, model/losses.py:192-197."""
    return binary_bce(z, t, dtype) + binary_dice(z, t, dtype)


def binary_dice_focal(z: Tensor, t: Tensor, dtype=torch.float32) -> Tensor:
    """Config 5 "Dice+focal" := BinaryDiceLoss + BinaryFocalLoss()."""
    return binary_dice(z, t, dtype) + binary_focal(z, t, dtype=dtype)


def _multi_parts(z: Tensor, t: Tensor, dtype):
    n, c = z.shape[0], z.shape[1]
    zf = z.to(dtype).reshape(n, c, -1)
    tl = t.long().reshape(n, -1)
    onehot = F.one_hot(tl, c).permute(0, 2, 1)                 # losses.py:254-255, 311-312
    present = onehot.sum((0, 2)) > 0                           # losses.py:256, 323
    return zf, tl, onehot, present


def multi_dice(z: Tensor, t: Tensor, alpha: Tensor, dtype=torch.float32) -> Tensor:
    """MutilDiceLoss.forward, model/losses.py:301-325 (result is negative: range [-1, 0])."""
    zf, tl, onehot, present = _multi_parts(z, t, dtype)
    p = torch.softmax(zf, dim=1)
    oh = onehot.to(dtype)
    inter = (oh * p).sum((0, 2))
    den = (oh + p).sum((0, 2))
    d = ((2.0 * inter + SMOOTH) / (den + SMOOTH)).clamp_min(EPS)
    per_class = -d * present.to(dtype)
    return (per_class * alpha.to(dtype)).sum() / torch.count_nonzero(present)


def multi_ce(z: Tensor, t: Tensor, alpha: Optional[Tensor] = None, dtype=torch.float32) -> Tensor:
    """MutilCrossEntropyLoss.forward, model/losses.py:252-260: CE weighted by the present-class
    mask == plain mean NLL (absent classes never occur as targets); ``alpha`` unused."""
    zf, tl, onehot, present = _multi_parts(z, t, dtype)
    logp = torch.log_softmax(zf, dim=1)
    nll = -logp.gather(1, tl.unsqueeze(1)).squeeze(1)
    w = present.to(dtype)[tl]
    return (w * nll).sum() / w.sum()


def multi_focal(z: Tensor, t: Tensor, alpha: Optional[Tensor] = None, gamma: float = 2, dtype=torch.float32) -> Tensor:
    """MutilFocalLoss.forward, model/losses.py:273-285; ``alpha`` unused."""
    zf, tl, onehot, present = _multi_parts(z, t, dtype)
    logp = torch.log_softmax(zf, dim=1)
    ce = -logp.gather(1, tl.unsqueeze(1)).squeeze(1) * present.to(dtype)[tl]
    pt = torch.exp(-ce)
    return ((1 - pt) ** gamma * ce).mean()


def multi_ce_dice(z: Tensor, t: Tensor, alpha: Tensor, dtype=torch.float32) -> Tensor:
    """MutilCrossEntropyDiceLoss.forward, model/losses.py:337-342."""
    return multi_ce(z, t, alpha, dtype) + multi_dice(z, t, alpha, dtype)


LOSSES: Dict[str, Callable] = {
    "BinaryDiceLoss": binary_dice,
    "BinaryCrossEntropyLoss": binary_bce,
    "BinaryFocalLoss": binary_focal,
    "BinaryCrossEntropyDiceLoss": binary_bce_dice,
    "BinaryDiceFocalLoss": binary_dice_focal,
    "MutilDiceLoss": multi_dice,
    "MutilCrossEntropyLoss": multi_ce,
    "MutilFocalLoss": multi_focal,
    "MutilCrossEntropyDiceLoss": multi_ce_dice,
}


def loss_forward(name: str, z: Tensor, t: Tensor, alpha: Optional[Tensor] = None, gamma: Optional[float] = None,
                 dtype=torch.float32) -> Tensor:
    fn = LOSSES[name]
    if name.startswith("Binary"):
        return fn(z, t, dtype=dtype)
    if name == "MutilFocalLoss":
        return fn(z, t, alpha, gamma if gamma is not None else 2, dtype=dtype)
    return fn(z, t, alpha, dtype=dtype)
