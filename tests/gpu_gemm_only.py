#!/usr/bin/env python
"""GPU-box tool: K4 alone at BASELINE config 2 (T=8192, K=4096, N=11008) and the
LLaMA-7B layer shapes at T=2048, for both tile plans.  argv: [reps] [cta_group]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from llm_qat_b200 import _lib
from llm_qat_b200.utils_quant import qlinear_i8
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
groups = [int(sys.argv[2])] if len(sys.argv) > 2 else [1, 2]
shapes = [(8192, 4096, 11008)] if len(sys.argv) > 2 else [(8192, 4096, 11008), (2048, 4096, 4096),
                                                            (2048, 4096, 11008), (2048, 11008, 4096)]
g = torch.Generator().manual_seed(0)
for T, K, N in shapes:
    qx = torch.randint(-127, 128, (T, K), generator=g, dtype=torch.int8).cuda()
    qw = torch.randint(-7, 8, (N, K), generator=g, dtype=torch.int8).cuda()
    ex = (torch.rand(T, generator=g) * 50 + 1).cuda()
    ew = (torch.rand(N, generator=g) * 300 + 10).cuda()
    for cg in groups:
        _lib.check(_lib.lib().qat_set_gemm_cta_group(cg))
        for _ in range(3):
            out = qlinear_i8(qx, qw, ex, ew, torch.bfloat16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = qlinear_i8(qx, qw, ex, ew, torch.bfloat16)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"qlinear_i8 cta_group={cg} {T}x{N}x{K}: {ms*1e3:.1f} us  {2*T*N*K/ms/1e9:.0f} TOP/s", flush=True)
