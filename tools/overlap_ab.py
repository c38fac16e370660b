#!/usr/bin/env python
"""A/B of the host-side scheduling options (B200SEG_OVERLAP bit mask, engine.py) on ONE box and in ONE process:
for every mask a fresh GraphedStep is captured on the same model / inputs / dropout draw, its gradients are compared
with mask 0 (same kernels on the same data: only fp32/fp64 atomic ordering may differ) and the replay is timed like
bench.py times it (device events, 256 MiB L2 flush before every step, median).

  python tools/overlap_ab.py --masks 0,15,1,2,4,8,0,15 [--workload vnet3d|unet2d|unet3d] [--steps 30] >> gpurun_out/overlap_ab.jsonl
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import pytorchdeeplearing_b200 as b200  # noqa: E402
from pytorchdeeplearing_b200.graphed import GraphedStep  # noqa: E402
import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--masks", default="0,15")
    ap.add_argument("--workload", default="vnet3d")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    b200.set_precision("bf16")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if args.workload == "vnet3d":
        model, ncls, shape = b200.VNet3d(1, 2), 2, (2, 1, 96, 96, 96)
        lossfn = b200.MutilDiceLoss(torch.ones(2, device=dev))
    elif args.workload == "unet3d":
        model, ncls, shape = b200.UNet3d(1, 4), 4, (1, 1, 128, 128, 128)
        lossfn = b200.MutilCrossEntropyDiceLoss(torch.ones(4, device=dev))
    else:
        model, ncls, shape = b200.UNet2d(1, 1), 1, (8, 1, 512, 512)
        lossfn = b200.BinaryDiceFocalLoss()
    model.apply(b200.initialize_weights)
    model = model.to(dev).train()
    x, y = oracle.make_inputs(shape[0], shape[1], shape[2:], ncls)
    x, y = x.to(dev), y.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ref = None
    for mask in [int(m) for m in args.masks.split(",")]:
        os.environ["B200SEG_OVERLAP"] = str(mask)
        step = GraphedStep(model, lossfn, x, y, warmup=2)
        for _ in range(5):
            step()
        # one replay on a fixed generator state: loss + gradients to compare across masks
        torch.cuda.manual_seed(1234)
        loss = float(step())
        torch.cuda.synchronize()
        flat = step._flat.detach().clone()
        ms = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            ms.append((e0, e1))
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ms]
        rec = {"workload": args.workload, "tag": args.tag, "mask": mask, "ms_median": statistics.median(ms),
               "ms_min": min(ms), "ms_max": max(ms), "loss": loss, "grad_norm": float(flat.norm()),
               "finite": bool(torch.isfinite(flat).all()), "graphs": len(step.graphs),
               "env": {k: v for k, v in os.environ.items() if k.startswith("B200SEG_")}}
        if ref is None:
            ref = (loss, flat)
        else:
            rec["loss_diff_vs_first"] = abs(loss - ref[0])
            rec["grad_rel_l2_vs_first"] = float((flat - ref[1]).norm() / ref[1].norm())
            rec["grad_max_abs_diff_vs_first"] = float((flat - ref[1]).abs().max())
        print(json.dumps(rec), flush=True)
        del step
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
