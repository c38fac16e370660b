"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py
(x ~ N(0,1); labels Bernoulli(0.3) for binary / 2-class heads,
uniform over classes otherwise; generator seed 1234, x drawn before labels)."""
from __future__ import annotations

import torch


def make_inputs(n: int, cin: int, spatial, numclass: int, seed: int = 1234, absent_class: int | None = None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, cin) + tuple(spatial), generator=g, dtype=torch.float32)
    if numclass <= 2:
        y = (torch.rand((n,) + tuple(spatial), generator=g) > 0.7).long()
    else:
        y = torch.randint(0, numclass, (n,) + tuple(spatial), generator=g)
    if absent_class is not None:
        y[y == absent_class] = (absent_class + 1) % numclass
    return x, y
