"""CUDA-graph capture of one training step (forward + loss + backward [+ gradient all-reduce])
through the public nn.Module / loss API.

The hot path is ~200 short kernel launches per step issued from Python through ctypes; at a
roofline step time of ~0.5 ms the launch path would dominate.  ``GraphedStep`` captures the
launches once (static shapes, static input/output buffers) and replays them with one
``cudaGraphLaunch`` -- "CUDA streams and graphs instead of a tracing compiler".

    step = GraphedStep(model, lossfn, x_example, y_example)
    loss = step(x, y)            # copies x, y into the static buffers, replays, returns the loss tensor
    step.prefetch(x2, y2); loss = step(prefetched=True)   # pipelined form: the H2D copy overlaps the previous step
    # model.parameters() .grad now hold this step's gradients (static tensors, overwritten per replay)

Single GPU: ONE graph holds ``model(x) -> lossfn -> loss.backward()`` exactly as a user writes it.

Data parallel (``enable_data_parallel()`` active): collectives are kept OUT of the graphs.  The step is
split into graph A (forward + loss partial sums), an eager NCCL all-reduce of the partial sums, graph B
(loss finalize + d loss/d logits + network backward into the flat gradient bucket) and an eager NCCL
SUM all-reduce of the bucket.  The split mode drives the same engine / kernels
as the autograd path; ``tests/test_engine_cpu.py`` checks that both produce the same gradients.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import runtime
from .engine import Engine
from .losses import loss_spec


class GraphedStep:
    def __init__(self, model: torch.nn.Module, lossfn, x: torch.Tensor, y: torch.Tensor, warmup: int = 2,
                 optimizer: Optional[torch.optim.Optimizer] = None, use_graph: bool = True):
        self.model, self.lossfn, self.optimizer = model, lossfn, optimizer
        self.x = torch.empty_like(x)
        self.y = torch.empty_like(y)
        self.x.copy_(x)
        self.y.copy_(y)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.split = runtime.dp_state()[0]
        self.use_graph = use_graph and x.is_cuda
        self.graph = self.graph_a = self.graph_b = None
        if self.split:
            self._init_split(warmup)
        else:
            self._init_single(warmup)

    # ------------------------------------------------------------------ single graph (no collectives)
    def _eager_step(self):
        for p in self.params:
            p.grad = None
        logits, _ = self.model(self.x)
        loss = self.lossfn(logits, self.y)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    def _init_single(self, warmup):
        if not self.use_graph:
            return
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

    # ------------------------------------------------------------------ split phases (data parallel)
    def _phase_a(self):
        """forward + loss partial sums (no grad mode, engine driven directly)."""
        m = self.model
        be = runtime.get_backend(self.x)
        eng = Engine(be, runtime.act_dtype(), m._dims)
        names = m._pnames
        P = dict(zip(names, [p.detach() for _, p in m.named_parameters()]))
        masks = m._draw_masks(self.x)
        x = self.x if self.x.dtype == torch.float32 else self.x.float()
        logits, probs = getattr(eng, m._arch + "_forward")(P, x, masks, True)
        perm = (0, 2, 3, 4, 1) if logits.dim() == 5 else (0, 2, 3, 1)
        z = logits.permute(*perm)
        if not z.is_contiguous():
            z = z.contiguous()
        c = z.shape[-1]
        t = self.y if self.y.dtype == torch.int64 else self.y.long()
        part = torch.zeros(3 * c + 3 if c > 1 else 6, dtype=torch.float64, device=z.device)
        be.loss_partials(z, t, self._gamma, self._alpha_f, part)
        self._eng, self._z, self._t, self._part, self._probs = eng, z, t, part, probs

    def _phase_b(self):
        """loss finalize + d loss / d logits + network backward -> flat gradient bucket."""
        be = runtime.get_backend(self.x)
        z, t, part, eng = self._z, self._t, self._part, self._eng
        c = z.shape[-1]
        dev = z.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lcoef = torch.empty(2 * c + 3 if c > 1 else 5, dtype=torch.float32, device=dev)
        alpha = self._alpha
        alpha = torch.ones(c, dtype=torch.float32, device=dev) if alpha is None else \
            torch.as_tensor(alpha, dtype=torch.float32, device=dev)
        be.loss_finalize(part, c, self._terms, alpha, self._gamma, self._alpha_f, loss, lcoef)
        dz = torch.empty_like(z)
        be.loss_bwd(z, t, lcoef, self._one, dz)
        g = dz if dz.dim() == 5 else dz.unsqueeze(1)
        flat = getattr(eng, self.model._arch + "_backward")(g)
        self.loss, self._flat = loss, flat
        self._grads = [eng.grads[n] for n in self.model._pnames]

    def _split_step_eager(self):
        import torch.distributed as dist
        _, group = runtime.dp_state()
        with torch.no_grad():
            self._phase_a()
            dist.all_reduce(self._part, op=dist.ReduceOp.SUM, group=group)
            self._phase_b()
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group)
        for p, g in zip(self.model.parameters(), self._grads):
            p.grad = g
        return self.loss

    def _init_split(self, warmup):
        self._terms, self._alpha, self._gamma, self._alpha_f = loss_spec(self.lossfn)
        self._one = torch.ones(1, dtype=torch.float32, device=self.x.device)
        for _ in range(max(1, warmup)):
            self._split_step_eager()
        if not self.use_graph:
            return
        torch.cuda.synchronize()
        self.graph_a, self.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.no_grad():
            with torch.cuda.graph(self.graph_a):
                self._phase_a()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                self._phase_b()
        torch.cuda.synchronize()
        for p, g in zip(self.model.parameters(), self._grads):
            p.grad = g

    # ------------------------------------------------------------------ input prefetch
    def prefetch(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Start the host->device copy of the NEXT step's inputs on a copy stream (pinned host tensors) while the
        current step is still running; ``step(prefetched=True)`` then only moves them device-to-device into the
        static buffers.  This is the input pipeline a training loop would run around the graph."""
        if not self.x.is_cuda:
            self._staged_host = (x, y)
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.x.device)
            self._sx, self._sy = torch.empty_like(self.x), torch.empty_like(self.y)
            self._staged_ev = torch.cuda.Event()
            self._free_ev = torch.cuda.Event()
            self._free_ev.record(torch.cuda.current_stream(self.x.device))
        cs = self._copy_stream
        cs.wait_event(self._free_ev)                 # the previous step has drained the staging buffers
        with torch.cuda.stream(cs):
            self._sx.copy_(x, non_blocking=True)
            self._sy.copy_(y, non_blocking=True)
            self._staged_ev.record(cs)

    def _take_prefetched(self) -> None:
        if not self.x.is_cuda:
            x, y = self._staged_host
            self.x.copy_(x)
            self.y.copy_(y)
            return
        main = torch.cuda.current_stream(self.x.device)
        main.wait_event(self._staged_ev)
        self.x.copy_(self._sx, non_blocking=True)
        self.y.copy_(self._sy, non_blocking=True)
        self._free_ev.record(main)

    # ------------------------------------------------------------------ replay
    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None,
                 prefetched: bool = False) -> torch.Tensor:
        if prefetched:
            self._take_prefetched()
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if self.split:
            if self.graph_a is None:
                return self._split_step_eager()
            import torch.distributed as dist
            _, group = runtime.dp_state()
            self.graph_a.replay()
            dist.all_reduce(self._part, op=dist.ReduceOp.SUM, group=group)
            self.graph_b.replay()
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group)
            return self.loss
        if self.graph is None:
            return self._eager_step()
        self.graph.replay()
        return self.loss
