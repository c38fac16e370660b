"""Development aid: %globaltimer stamps of CTA 0 of the weight-streaming halo conv (see conv_halo_ws.cu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
os.environ["B200SEG_WS_DBG"] = hex(dbg.data_ptr())
from pytorchdeeplearing_b200._abi import CudaBackend  # noqa: E402
from pytorchdeeplearing_b200.engine import K3  # noqa: E402

be = CudaBackend()
names = ["start", "setup done", "slices landed", "MMAs issued", "tfull seen", "tiles drained", "end"]
for ci, s in ((64, 24), (128, 12)):
    n = 2
    w = torch.randn((ci, ci, 3, 3, 3), device="cuda") * 0.05
    wp = be.pack_weight(w, K3, "fwd", torch.bfloat16, 3, vox=s ** 3)
    x = torch.randn((n, s, s, s, ci), device="cuda").bfloat16()
    y = torch.empty_like(x)
    bias = torch.zeros(ci, device="cuda")
    stats = torch.zeros((n, ci, 2), dtype=torch.float64, device="cuda")
    for _ in range(3):
        be.conv(K3, 3, x, wp, bias, y, stats, None)
    torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    print(f"{ci}ch @{s}^3:", ", ".join(f"{names[i]} +{(t[i] - t[0]) / 1e3:.2f}us" for i in range(1, 7)))
