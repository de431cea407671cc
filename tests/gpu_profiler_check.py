#!/usr/bin/env python
"""GPU-box tool: the library's launches under torch.profiler (CUPTI tracing), with and without
programmatic dependent launch — a training job must be profilable without hanging."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from llm_qat_b200 import QuantizeLinear, SymQuantizer

BIG = "big" in sys.argv          # LLaMA-7B shapes: K4 takes the CTA-pair (cluster) plan
AUTOCAST = "autocast" in sys.argv
import contextlib
K, N, T = (4096, 11008, 2048) if BIG else (1024, 1024, 512)
lin = QuantizeLinear(K, N, w_bits=4, a_bits=8).bfloat16().cuda()
x = torch.randn(T, K, device="cuda").bfloat16().requires_grad_(True)
clip = torch.tensor([-2.0, 2.0])
def step():
    with (torch.autocast("cuda", dtype=torch.bfloat16) if AUTOCAST else contextlib.nullcontext()):
        y = SymQuantizer.apply(lin(x), clip, 4, False)
    y.float().sum().backward()
    x.grad = None; lin.weight.grad = None
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.time()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
n = sum(1 for e in prof.key_averages() if "qat::" in e.key)
print(f"{sys.argv[1:]} PDL={os.environ.get('QAT_B200_PDL', '1')}: profiled 3 steps in {time.time() - t0:.2f} s, {n} distinct library kernels seen", flush=True)
