#!/usr/bin/env python
"""GPU-box tool: host-side cost per call of the Python boundary (tiny tensors, so the
GPU is never the bound): QuantizeLinear fwd / fwd+bwd, SymQuantizer.apply fwd+bwd,
against nn.Linear and the reference's eager chain (oracle/ref_module.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import llm_qat_b200
from llm_qat_b200 import QuantizeLinear, SymQuantizer
from oracle import ref_module as R

dev = "cuda"
x = torch.randn(4, 16, 256, device=dev).bfloat16().requires_grad_(True)
go = torch.randn(4, 16, 256, device=dev).bfloat16()
clip = torch.tensor([-2.0, 2.0])
N = 2000

def timeit(name, fn):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:48s} {(t1 - t0) / N * 1e6:8.1f} us/call (host enqueue)", flush=True)

mine = QuantizeLinear(256, 256, w_bits=4, a_bits=8).bfloat16().to(dev)
ref = R.QuantizeLinear(256, 256, w_bits=4, a_bits=8).bfloat16().to(dev)
plain = torch.nn.Linear(256, 256, bias=False).bfloat16().to(dev)
def fwd(m):
    with torch.no_grad(): m(x)
def fb(m):
    y = m(x); y.backward(go); x.grad = None; m.weight.grad = None
for name, m in (("nn.Linear", plain), ("QuantizeLinear (product, fused)", mine), ("QuantizeLinear (reference eager chain)", ref)):
    timeit(name + " fwd", lambda: fwd(m))
    timeit(name + " fwd+bwd", lambda: fb(m))
os.environ["QAT_B200_FUSED_LINEAR"] = "0"
timeit("QuantizeLinear (product, unfused) fwd", lambda: fwd(mine))
timeit("QuantizeLinear (product, unfused) fwd+bwd", lambda: fb(mine))
os.environ.pop("QAT_B200_FUSED_LINEAR")
def q(Q):
    y = Q.apply(x, clip, 4, False); y.backward(go); x.grad = None
timeit("SymQuantizer.apply fwd+bwd (product)", lambda: q(SymQuantizer))
timeit("SymQuantizer.apply fwd+bwd (reference chain)", lambda: q(R.SymQuantizer))
