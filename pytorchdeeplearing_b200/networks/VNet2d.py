"""Drop-in ``VNet2d`` (reference networks/VNet2d.py:102-160; wrappers model/modelVNet.py:25-466) -- SURVEY 8f-4."""
from ._vnet import _VNetBase


class VNet2d(_VNetBase):
    _dims = 2
