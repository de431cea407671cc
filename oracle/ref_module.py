"""The reference's L1 (`models/utils_quant.py`) restated on top of
oracle/torch_chain.py as a module with the same three names, so that the
harness can run the reference's eager path on the GPU box (where
/root/reference does not exist).  TEST/BENCH INFRASTRUCTURE ONLY: it is the
comparator, never the product.  tests/test_harness.py checks it against the
live reference module bit for bit.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import torch_chain as tc


class SymQuantizer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, clip_val, num_bits, layerwise):  # utils_quant.py:37-74
        ctx.save_for_backward(input, clip_val)
        return tc.sym_forward(input, num_bits, layerwise)

    @staticmethod
    def backward(ctx, grad_output):  # utils_quant.py:77-87
        input, clip_val = ctx.saved_tensors
        return tc.ste_backward(grad_output, input, clip_val), None, None, None


class AsymQuantizer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, clip_val, num_bits, layerwise):  # utils_quant.py:96-149
        ctx.save_for_backward(input, clip_val)
        return tc.asym_forward(input, num_bits, layerwise)

    @staticmethod
    def backward(ctx, grad_output):  # utils_quant.py:152-162
        input, clip_val = ctx.saved_tensors
        return tc.ste_backward(grad_output, input, clip_val), None, None, None


class QuantizeLinear(nn.Linear):  # utils_quant.py:165-254
    def __init__(self, *kargs, symmetric=True, bias=False, w_bits=32, a_bits=32, act_layerwise=False,
                 weight_layerwise=False):
        super().__init__(*kargs, bias=False)
        self.w_bits, self.a_bits = w_bits, a_bits
        self.act_layerwise, self.weight_layerwise = act_layerwise, weight_layerwise
        if 2 < a_bits < 32:
            self.act_quantizer = SymQuantizer if symmetric else AsymQuantizer

    def forward(self, input_):
        assert self.weight.dim() == 2
        w = self.weight
        if self.w_bits >= 32:
            weight = w
        elif self.w_bits >= 3:
            weight = SymQuantizer.apply(w, torch.tensor([-2.0, 2.0]), self.w_bits, self.weight_layerwise)
        else:
            if self.w_bits == 1:
                sf = (w.abs().mean() if self.weight_layerwise else w.abs().mean(dim=1, keepdim=True)).detach()
                q = sf * torch.sign(w / sf)
            else:
                levels, clip = 2 ** (self.w_bits - 1), 1 - 1e-2
                sf = 2 * (w.abs().mean() if self.weight_layerwise else w.abs().mean(dim=1, keepdim=True)).detach()
                q = sf * (torch.round(torch.clamp(w / sf, -clip, clip) * levels - 0.5) + 0.5) / levels
            weight = q.detach() - w.detach() + w
        if 2 < self.a_bits < 32:
            input_ = self.act_quantizer.apply(input_, torch.tensor([-2.0, 2.0]), self.a_bits, self.act_layerwise)
        return nn.functional.linear(input_, weight)
