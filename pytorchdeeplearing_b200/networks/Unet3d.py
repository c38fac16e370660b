"""Drop-in ``UNet3d`` (reference networks/Unet3d.py:6-86): same ctor, same 64-tensor state_dict."""
from ._unet import _UNetBase


class UNet3d(_UNetBase):
    _dims = 3
