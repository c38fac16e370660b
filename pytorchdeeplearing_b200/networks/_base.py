"""Shared plumbing of the drop-in networks: one ``autograd.Function`` per forward that hands
the whole net to the layer program (``engine.Engine``), i.e. to the sm_100a kernels."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import runtime
from ..engine import Engine, P_DROP


class _SegNetFunction(torch.autograd.Function):
    """forward: (x, *params) -> (logits, probs); backward: d logits -> d params.
    Backward runs on the autograd thread; every kernel launch takes device+stream explicitly."""

    @staticmethod
    def forward(ctx, mod, need_grad, x, masks, *params):
        be = runtime.get_backend(x)
        eng = Engine(be, runtime.act_dtype(), mod._dims)
        P = dict(zip(mod._pnames, [p.detach() for p in params]))
        xx = x.detach()
        if xx.dtype != torch.float32:
            xx = xx.float()
        logits, probs = getattr(eng, mod._arch + "_forward")(P, xx, masks, need_grad)
        ctx.eng = eng
        ctx.arch = mod._arch
        ctx.names = mod._pnames
        ctx.mark_non_differentiable(probs)
        return logits, probs

    @staticmethod
    def backward(ctx, g_logits, g_probs):
        eng: Engine = ctx.eng
        g = g_logits.permute(0, 2, 3, 4, 1) if g_logits.dim() == 5 else g_logits.permute(0, 2, 3, 1).unsqueeze(1)
        if not g.is_contiguous():
            g = g.contiguous()
        if g.dtype != torch.float32:
            g = g.float()
        flat = getattr(eng, ctx.arch + "_backward")(g)
        enabled, group = runtime.dp_state()
        if enabled:
            import torch.distributed as dist
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        grads = [eng.grads[n] for n in ctx.names]
        ctx.eng = None
        return (None, None, None, None) + tuple(grads)


class SegNetBase(nn.Module):
    """Common forward of VNet3d / UNet3d / UNet2d.  Subclasses build the parameter tree with the
    reference's attribute names (so ``state_dict`` / ``load_state_dict`` / ``apply(initialize_weights)``
    behave identically) and set ``_arch`` / ``_dims`` / ``_mask_channels``."""

    _arch = ""
    _dims = 3
    dropout_masks: Optional[List[torch.Tensor]] = None     # test hook: inject (N,C) scales

    def _finish_init(self):
        self._pnames = [n for n, _ in self.named_parameters()]

    def _mask_channels(self) -> List[int]:
        raise NotImplementedError

    def _draw_masks(self, x: torch.Tensor) -> Optional[List[torch.Tensor]]:
        """nn.Dropout3d/2d(p=0.2) contract: per call, in module-call
        order, ``x.new_empty((N,C,1,1,1)).bernoulli_(1-p).div_(1-p)`` from the default generator."""
        if not self.training:
            return None
        if self.dropout_masks is not None:
            return [m.to(device=x.device, dtype=torch.float32) for m in self.dropout_masks]
        # One flat buffer: every mask is drawn by its own ``bernoulli_`` call on an (N,C,1,..) view -- the same
        # generator consumption as the reference's per-module draws -- and the 1/(1-p) scaling of all of them
        # is a single multiply instead of one launch per layer.
        n = x.shape[0]
        ones = (1,) * self._dims
        chans = self._mask_channels()
        flat = x.new_empty((n * sum(chans),), dtype=torch.float32)
        masks, off = [], 0
        for c in chans:
            flat[off:off + n * c].view((n, c) + ones).bernoulli_(1 - P_DROP)
            masks.append(flat[off:off + n * c].view(n, c))
            off += n * c
        flat.div_(1 - P_DROP)
        return masks

    def forward(self, x):
        if x.dim() != self._dims + 2:
            raise RuntimeError(f"expected a {self._dims + 2}-D input (N,C,{'D,' if self._dims == 3 else ''}H,W), "
                               f"got {tuple(x.shape)}")
        for s in x.shape[2:]:
            if s % 16 != 0:
                raise RuntimeError("spatial sizes must be multiples of 16 (four stride-2 stages), got "
                                   f"{tuple(x.shape[2:])}")
        params = [p for _, p in self.named_parameters()]
        masks = self._draw_masks(x)
        # (grad mode is always off inside Function.forward, so the decision is taken here)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _SegNetFunction.apply(self, need_grad, x, masks, *params)


class _Holder(nn.Module):
    """Parameter container with the reference's attribute names; never called directly."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-blocks of the B200 drop-in are parameter holders; call the network itself")
