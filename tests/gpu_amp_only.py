#!/usr/bin/env python
"""GPU-box tool: the autocast-variant K1 alone at [8192, 4096] (fp32 y, then the GEMM feed), for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from llm_qat_b200._lib import CODES_I8
from llm_qat_b200.utils_quant import fake_quant_forward
g = torch.Generator().manual_seed(0)
xs = [(torch.randn(8192, 4096, generator=g) * 0.5).bfloat16().cuda() for _ in range(3)]
for i in range(6):
    y = fake_quant_forward(xs[i % 3], 8, False, True, amp=True)[0]
for i in range(6):
    r = fake_quant_forward(xs[i % 3], 8, False, True, want_y=False, codes_kind=CODES_I8, want_scales=True,
                           mask_clip=(-2.0, 2.0), amp=True)
torch.cuda.synchronize()
print("ok")
