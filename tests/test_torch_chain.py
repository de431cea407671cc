"""CPU: the torch-eager port used for the CPU baseline reproduces the live
reference's outputs (golden vectors) bit for bit."""
import pytest
import torch

import qat_testutil as U
from oracle import torch_chain as tc


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_chain_matches_reference_goldens(dtype):
    g = U.golden(dtype)
    for q, key, bits, lw in U.quant_cases(g):
        x = U.bits_to_tensor(g[f"in/{key}"], dtype)
        y = (tc.sym_forward if q == "sym" else tc.asym_forward)(x, bits, lw)
        ref = g[f"y/{q}/{key}/b{bits}/{'lw' if lw else 'row'}"]
        assert U.mismatches(U.tensor_bits(y), ref, dtype) == 0, (q, key, bits, lw)
    for q, key, lw, lo, hi, gk in U.clip_cases(g):
        x = U.bits_to_tensor(g[f"in/{key}"], dtype)
        gr = U.bits_to_tensor(g[f"grad/{key}"], dtype)
        gx = tc.ste_backward(gr, x, torch.tensor([lo, hi]))
        assert U.mismatches(U.tensor_bits(gx), g[gk], dtype) == 0, gk
