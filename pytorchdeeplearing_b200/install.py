"""Patch the drop-in into an imported reference tree.

The reference wrappers bind the classes at import time (``from networks.VNet3d import VNet3d``
model/modelVNet.py:2-4, model/modelUnet.py:2-4; losses model/modelVNet.py:7-8), so ``install``
rebinds those names in whichever of ``networks``, ``networks.*``, ``model.modelVNet``,
``model.modelUnet``, ``model.losses``, ``model.metric`` are already imported (and in ``sys.modules`` entries
imported later via the same names).  ``uninstall`` restores the originals.
"""
from __future__ import annotations

import sys

_SAVED = []

_NET_NAMES = ("VNet3d", "VNet2d", "UNet3d", "UNet2d")
_METRIC_NAMES = ("dice_coeff", "iou_coeff", "multiclass_dice_coeff", "multiclass_iou_coeff")
_LOSS_NAMES = ("BinaryDiceLoss", "BinaryCrossEntropyLoss", "BinaryFocalLoss", "BinaryCrossEntropyDiceLoss",
               "MutilCrossEntropyLoss", "MutilFocalLoss", "MutilDiceLoss", "MutilCrossEntropyDiceLoss")
_MODULES = ("networks", "networks.VNet3d", "networks.VNet2d", "networks.Unet3d", "networks.Unet2d", "model",
            "model.losses", "model.metric", "model.modelVNet", "model.modelUnet")


def install() -> int:
    """Rebind the reference's names to the B200 implementations. Returns the number of bindings changed."""
    from . import networks as nets, losses, metric
    table = {n: getattr(nets, n) for n in _NET_NAMES}
    table.update({n: getattr(losses, n) for n in _LOSS_NAMES})
    table.update({n: getattr(metric, n) for n in _METRIC_NAMES})     # per-step accuracy, model/modelVNet.py:582
    count = 0
    for modname in _MODULES:
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for name, obj in table.items():
            if hasattr(mod, name) and getattr(mod, name) is not obj:
                _SAVED.append((mod, name, getattr(mod, name)))
                setattr(mod, name, obj)
                count += 1
    return count


def uninstall() -> None:
    while _SAVED:
        mod, name, obj = _SAVED.pop()
        setattr(mod, name, obj)
