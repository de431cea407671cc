"""Torch-eager restatement of the reference's fake-quant op chain, for timing
the reference's CPU path on the GPU box's host cores (TEST/BENCH INFRASTRUCTURE
ONLY — see oracle/quant_oracle.py's header for who may import this).

``/root/reference`` cannot travel to the GPU box, so ``bench.py``'s
``cpu_baseline`` and ``--impl reference`` legs run this port instead.  It issues
the same ATen ops in the same order as the reference (abs, max(dim), expand,
add, reciprocal*Q, mul, round, add, div ... — utils_quant.py:50-72, 110-147 —
and clone, ge, le, 2x masked-fill — :83-87), so its CPU time is the
reference's.  tests/test_torch_chain.py checks it bit-for-bit against the
golden vectors produced by the live reference.
"""
from __future__ import annotations

import torch


def _row_stat(t: torch.Tensor, layerwise: bool, fn):
    if layerwise:
        return fn(t).expand_as(t)
    if t.dim() <= 3:
        return fn(t, dim=-1, keepdim=True)[0].expand_as(t)
    if t.dim() == 4:
        flat = t.view(t.shape[0], t.shape[1], -1)
        return fn(flat, dim=-1, keepdim=True)[0].unsqueeze(-1).expand_as(t)
    raise ValueError


def sym_forward(x: torch.Tensor, num_bits: int, layerwise: bool = False) -> torch.Tensor:
    """utils_quant.py:50-72."""
    m = _row_stat(torch.abs(x), layerwise, torch.max).detach()
    s = (2 ** (num_bits - 1) - 1) / (m + 1e-6)
    return torch.round(x * s).div(s + 1e-6)


def asym_forward(x: torch.Tensor, num_bits: int, layerwise: bool = False) -> torch.Tensor:
    """utils_quant.py:110-147 (min is reduced twice, as the reference does)."""
    alpha = (_row_stat(x, layerwise, torch.max) - _row_stat(x, layerwise, torch.min)).detach()
    beta = _row_stat(x, layerwise, torch.min).detach()
    n = (x - beta) / (alpha + 1e-8)
    s = 2 ** num_bits - 1
    q = torch.round(n * s).div(s)
    return q * (alpha + 1e-8) + beta


def ste_backward(g: torch.Tensor, x: torch.Tensor, clip_val: torch.Tensor) -> torch.Tensor:
    """utils_quant.py:83-87 / 158-162."""
    gx = g.clone()
    gx[x.ge(clip_val[1])] = 0
    gx[x.le(clip_val[0])] = 0
    return gx


def fwd_bwd(x, g, num_bits, symmetric=True, clip=(-2.0, 2.0)):
    clip_val = torch.tensor(list(clip))
    y = (sym_forward if symmetric else asym_forward)(x, num_bits)
    return y, ste_backward(g, x, clip_val)
