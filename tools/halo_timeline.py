#!/usr/bin/env python
"""Development aid for the halo-staged 3x3x3 kernels: device time of one layer at the VNet3d shapes (CUDA events,
warm), optionally with the per-role timeline of CTA 0 (B200SEG_HALO_DBG=1, conv_halo3 only).

    B200SEG_HALO3=0 B200SEG_HALO_LOADER=r python tools/halo_timeline.py     # old kernel, register-staged loader
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorchdeeplearing_b200 import runtime  # noqa: E402

be = runtime.cuda_backend()
K3 = 0
tag = f"HALO3={os.environ.get('B200SEG_HALO3', '1')} LOADER={os.environ.get('B200SEG_HALO_LOADER', 'cp.async')}"
shapes = ((2, (96, 96, 96), 16), (2, (48, 48, 48), 32), (1, (128, 128, 128), 16))
if os.environ.get("HALO_SHAPES") == "d":
    shapes = ((2, (8, 96, 96), 16), (2, (32, 96, 96), 16), (2, (96, 96, 96), 16), (2, (192, 96, 96), 16))
for (n, sp, c) in shapes:
    w = torch.randn(c, c, 3, 3, 3, device="cuda") * 0.05
    x = torch.randn((n,) + sp + (c,), device="cuda").bfloat16()
    y = torch.empty_like(x)
    st = torch.zeros(n, c, 2, dtype=torch.float64, device="cuda")
    wp = be.pack_weight(w, K3, "fwd", torch.bfloat16, 3, vox=10 ** 9)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        be.conv(K3, 3, x, wp, None, y, st, None)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    by = 2 * x.numel() * 2
    print(f"[{tag}] {c}->{c} @{sp} n={n}: {min(ts[1:]):7.1f} us  ({by / min(ts[1:]) / 1e3:6.1f} GB/s algorithmic, "
          f"{2 * 27 * c * c * x.numel() / c / min(ts[1:]) / 1e6:6.1f} TFLOP/s)", flush=True)
