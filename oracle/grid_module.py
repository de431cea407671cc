"""QuantizeLinear restated on the INTEGER GRID, in torch — the checker for K4.

TEST INFRASTRUCTURE ONLY (see oracle/quant_oracle.py's header for who may
import this); never the product.

The reference computes ``F.linear(x_q, W_q)`` on the two *dequantized* tensors
(/root/reference/models/utils_quant.py:250): each operand ``q/e`` is rounded to
bf16 before the GEMM.  The product's tcgen05 kernel contracts the integer codes
exactly and applies the two row scales afterwards, which is the
same mathematical value without those two operand roundings — a ~1e-3 relative
difference per linear.  Inside a decoder layer the following 4-bit K/V and
8-bit activation quantizers turn any such perturbation into whole-step code
flips for a fraction of a percent of the elements, so "layer output vs the
reference layer" has a noise floor of a few percent no matter how exact the
kernel is.  This module states the grid arithmetic itself:

    qx = rint(fl(x * s_x))  (saturated to the int8 range, as the GEMM feed is)
    out[t, n] = fl_dt( fl32( fl32(float(sum_k qx[t,k] qw[n,k]) * fl32(1/e_x[t])) * fl32(1/e_w[n]) ) )

with s, e from the reference's own chain (oracle/torch_chain.py), the integer
dot product taken in float64 (exact), and the reference's backward on the
dequantized operands with its STE masks.  The product must match THIS to bf16
rounding, and the harness tests assert that; the looser comparison against
oracle/ref_module.py documents the operand-rounding noise floor.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import torch_chain as tc
from .ref_module import AsymQuantizer, SymQuantizer  # noqa: F401  (same quantizers as the reference)


def sym_codes(x: torch.Tensor, num_bits: int):
    """(codes, e) of utils_quant.py:53-72 in x's dtype: q = round(x*s), e = s + 1e-6."""
    m = torch.abs(x).max(dim=-1, keepdim=True)[0]
    s = (2 ** (num_bits - 1) - 1) / (m + 1e-6)
    return torch.round(x * s), s + 1e-6


class _GridLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, w_bits, a_bits):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        qx, ex = sym_codes(x2, a_bits)
        qw, ew = sym_codes(w, w_bits)
        # int8 feed: a bf16 A8 code can reach +-128; the product carries -128
        # exactly and saturates +128 to 127
        qx, qw = qx.clamp(-128, 127), qw.clamp(-128, 127)
        ctx.save_for_backward(x, w, qx.div(ex), qw.div(ew))          # the reference's fake-quant tensors
        dot = qx.double() @ qw.double().t()
        rx, rw = 1.0 / ex.float(), 1.0 / ew.float()
        out = ((dot.float() * rx) * rw.t()).to(x.dtype)
        return out.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x, w, xq, wq = ctx.saved_tensors
        clip = torch.tensor([-2.0, 2.0])
        g2 = g.reshape(-1, g.shape[-1])
        gx = tc.ste_backward((g2 @ wq).view(x.shape), x, clip)        # utils_quant.py:250 then :83-87
        gw = tc.ste_backward(g2.t() @ xq, w, clip)
        return gx, gw, None, None


class QuantizeLinear(nn.Linear):
    """Main path only (symmetric, 3 <= bits <= 8, row-wise scales) — the path K4 serves."""

    def __init__(self, *kargs, symmetric=True, bias=False, w_bits=32, a_bits=32, act_layerwise=False,
                 weight_layerwise=False):
        super().__init__(*kargs, bias=False)
        assert symmetric and 3 <= w_bits <= 8 and 3 <= a_bits <= 8 and not act_layerwise and not weight_layerwise
        self.w_bits, self.a_bits = w_bits, a_bits

    def forward(self, input_):
        return _GridLinear.apply(input_, self.weight, self.w_bits, self.a_bits)
