#!/usr/bin/env python
"""GPU-box tool: the dequant kernel alone at [11008, 4096] (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from llm_qat_b200.utils_quant import dequant_codes
g = torch.Generator().manual_seed(0)
c = [torch.randint(-7, 8, (11008, 4096), generator=g, dtype=torch.int8).cuda() for _ in range(3)]
e = (torch.rand(11008, generator=g) * 300 + 10).bfloat16().float().cuda()
for i in range(9):
    out = dequant_codes(c[i % 3], e, torch.bfloat16)
torch.cuda.synchronize()
print("ok")
