"""CUDA-graph capture of one training step (forward + loss + backward [+ gradient all-reduce])
through the public nn.Module / loss API.

The hot path is ~400 short kernel launches per step issued from Python through ctypes; at a
roofline step time of ~0.5 ms the launch path would dominate.  ``GraphedStep`` captures the
launches once (static shapes, static input/output buffers) and replays them with one
``cudaGraphLaunch`` -- "CUDA streams and graphs instead of a tracing compiler".

    step = GraphedStep(model, lossfn, x_example, y_example)
    loss = step(x, y)            # copies x, y into the static buffers, replays, returns the loss tensor
    # model.parameters() .grad now hold this step's gradients (static tensors, overwritten per replay)
"""
from __future__ import annotations

from typing import Optional

import torch


class GraphedStep:
    def __init__(self, model: torch.nn.Module, lossfn, x: torch.Tensor, y: torch.Tensor, warmup: int = 2,
                 optimizer: Optional[torch.optim.Optimizer] = None):
        if not x.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors (no CPU fallback)")
        self.model, self.lossfn, self.optimizer = model, lossfn, optimizer
        self.x = torch.empty_like(x)
        self.y = torch.empty_like(y)
        self.x.copy_(x)
        self.y.copy_(y)
        self.params = [p for p in model.parameters() if p.requires_grad]
        # warm-up on a side stream (allocator + lazy init), as torch.cuda.graph requires
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        torch.cuda.synchronize()

    def _eager_step(self):
        for p in self.params:
            p.grad = None
        logits, _ = self.model(self.x)
        loss = self.lossfn(logits, self.y)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
