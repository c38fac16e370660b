"""pytorchdeeplearing_b200 -- B200-native (sm_100a) drop-in for the segmentation hot path of
junqiangchen/PytorchDeepLearing: ``networks/VNet3d.py``, ``networks/Unet3d.py``,
``networks/Unet2d.py`` forward/backward and the Dice / BCE / focal / CE reductions of
``model/losses.py``.

    from pytorchdeeplearing_b200 import VNet3d, MutilDiceLoss      # same ctor signatures
    import pytorchdeeplearing_b200 as b200; b200.install()         # or patch the reference in place

Host code is Python/PyTorch (device memory, streams, torch.distributed); every FLOP runs in
hand-written CUDA kernels behind the C ABI of ``include/b200seg.h``.
"""
from .runtime import (set_precision, get_precision, enable_data_parallel, disable_data_parallel)
from .networks import VNet3d, VNet2d, UNet3d, UNet2d, initialize_weights
from .losses import (BinaryDiceLoss, BinaryCrossEntropyLoss, BinaryFocalLoss, BinaryCrossEntropyDiceLoss,
                     BinaryDiceFocalLoss, MutilCrossEntropyLoss, MutilFocalLoss, MutilDiceLoss,
                     MutilCrossEntropyDiceLoss)
from .install import install, uninstall
from .graphed import GraphedStep
from .optim import FusedAdamW, FusedAdam
from .metric import dice_coeff, iou_coeff, multiclass_dice_coeff, multiclass_iou_coeff
from .inference import predict, predict_mask, sliding_window_mask
from .staging import InputStager, stage_batch

__all__ = ["VNet3d", "VNet2d", "UNet3d", "UNet2d", "initialize_weights", "FusedAdamW", "FusedAdam", "dice_coeff",
           "iou_coeff", "multiclass_dice_coeff", "multiclass_iou_coeff", "predict", "predict_mask",
           "sliding_window_mask", "InputStager", "stage_batch", "set_precision", "get_precision",
           "enable_data_parallel", "disable_data_parallel", "install", "uninstall", "GraphedStep",
           "BinaryDiceLoss", "BinaryCrossEntropyLoss", "BinaryFocalLoss", "BinaryCrossEntropyDiceLoss",
           "BinaryDiceFocalLoss", "MutilCrossEntropyLoss", "MutilFocalLoss", "MutilDiceLoss",
           "MutilCrossEntropyDiceLoss"]
