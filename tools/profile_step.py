#!/usr/bin/env python
"""One eager training step of the bench workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc \
      -c 3 -o gpurun_out/prof_conv_tc python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import pytorchdeeplearing_b200 as b200  # noqa: E402
import oracle  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "vnet3d"
b200.set_precision(os.environ.get("B200SEG_PRECISION", "bf16"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if arch == "vnet3d":
    model, ncls, shape = b200.VNet3d(1, 2), 2, (2, 1, 96, 96, 96)
    lossfn = b200.MutilDiceLoss(torch.ones(2, device=dev))
elif arch == "unet3d":
    model, ncls, shape = b200.UNet3d(1, 4), 4, (1, 1, 128, 128, 128)
    lossfn = b200.MutilCrossEntropyDiceLoss(torch.ones(4, device=dev))
else:
    model, ncls, shape = b200.UNet2d(1, 1), 1, (8, 1, 512, 512)
    lossfn = b200.BinaryDiceFocalLoss()
model.apply(b200.initialize_weights)
model = model.to(dev).train()
x, y = oracle.make_inputs(shape[0], shape[1], shape[2:], ncls)
x, y = x.to(dev), y.to(dev)


def step():
    for p in model.parameters():
        p.grad = None
    logits, _ = model(x)
    loss = lossfn(logits, y)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step of", arch)
