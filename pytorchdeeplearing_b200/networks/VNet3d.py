"""Drop-in ``VNet3d`` (reference networks/VNet3d.py:102-158); the reference's ``self.feature`` typo
(VNet3d.py:127) is not reproduced: the drop-in constructs."""
from ._vnet import _VNetBase


class VNet3d(_VNetBase):
    _dims = 3
