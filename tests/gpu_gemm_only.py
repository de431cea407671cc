#!/usr/bin/env python
"""GPU-box tool: K4 alone at BASELINE config 2 (T=8192, K=4096, N=11008) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from llm_qat_b200.utils_quant import qlinear_i8
T, K, N = 8192, 4096, 11008
g = torch.Generator().manual_seed(0)
qx = torch.randint(-127, 128, (T, K), generator=g, dtype=torch.int8).cuda()
qw = torch.randint(-7, 8, (N, K), generator=g, dtype=torch.int8).cuda()
ex = (torch.rand(T, generator=g) * 50 + 1).cuda()
ew = (torch.rand(N, generator=g) * 300 + 10).cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(3):
    out = qlinear_i8(qx, qw, ex, ew, torch.bfloat16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = qlinear_i8(qx, qw, ex, ew, torch.bfloat16)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"qlinear_i8 {T}x{N}x{K}: {ms*1e3:.1f} us  {2*T*N*K/ms/1e9:.0f} TOP/s")
