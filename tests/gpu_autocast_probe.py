#!/usr/bin/env python
"""GPU-box tool: what torch.autocast(bf16) — the context HF's Trainer runs the step in
(kd_trainer.py:106) — does to the reference's quantizer chains on bf16 tensors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import torch_chain as tc
g = torch.Generator().manual_seed(0)
x = torch.randn(64, 512, generator=g).bfloat16().cuda()
for name, fn in (("sym", tc.sym_forward), ("asym", tc.asym_forward)):
    y0 = fn(x, 8)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y1 = fn(x, 8)
        m = torch.abs(x).max(dim=-1, keepdim=True)[0]
        d = m + 1e-6
        r = d.reciprocal()
    diff = (y0.float() != y1.float()).float().mean().item()
    print(f"{name}: plain -> {y0.dtype}, under autocast -> {y1.dtype}; elements that differ: {diff:.4f}; "
          f"(max+1e-6).dtype={d.dtype} reciprocal.dtype={r.dtype}")
