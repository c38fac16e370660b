"""Drop-in ``UNet2d`` (reference networks/Unet2d.py:6-85): same ctor, same 64-tensor state_dict."""
from ._unet import _UNetBase


class UNet2d(_UNetBase):
    _dims = 2
