#!/usr/bin/env python
"""GPU-box tool (one process, >= 2 GPUs): modules placed on a device that is not the current one
(the reference's non-FSDP path is device_map="auto", train.py:61 — layers spread over GPUs in one
process).  Every entry point must launch on the tensor's device and stream, and the per-device
caches must not leak across devices."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from llm_qat_b200 import AsymQuantizer, QuantizeLinear, SymQuantizer

assert torch.cuda.device_count() >= 2
gen = torch.Generator().manual_seed(0)
x = torch.randn(3, 70, 256, generator=gen).bfloat16()
w = (torch.randn(384, 256, generator=gen) * 0.05).bfloat16()
go = torch.randn(3, 70, 384, generator=gen).bfloat16()
clip = torch.tensor([-2.0, 2.0])
res = {}
torch.cuda.set_device(0)      # current device stays 0 throughout
for dev in ("cuda:0", "cuda:1"):
    lin = QuantizeLinear(256, 384, w_bits=4, a_bits=8).bfloat16().to(dev)
    with torch.no_grad():
        lin.weight.copy_(w.to(dev))
    outs = []
    for _ in range(2):        # second round: caches populated by the first
        xi = x.to(dev).requires_grad_(True)
        k = SymQuantizer.apply(lin(xi), clip, 4, False)
        v = AsymQuantizer.apply(xi, clip, 8, False)
        (k.float().mul(go.to(dev).float()).sum() + v.float().sum()).backward()
        outs.append((k.detach().cpu(), v.detach().cpu(), xi.grad.cpu(), lin.weight.grad.cpu().clone()))
        lin.weight.grad = None
    s = torch.cuda.Stream(device=dev)      # and on a side stream of that device
    with torch.cuda.stream(s):
        xi = x.to(dev).requires_grad_(True)
        k = SymQuantizer.apply(lin(xi), clip, 4, False)
    s.synchronize()
    outs.append((k.detach().cpu(),))
    res[dev] = outs
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
ok = True
for a, b in zip(res["cuda:0"], res["cuda:1"]):
    ok = ok and all(torch.equal(u, v) for u, v in zip(a, b))
ok = ok and all(torch.equal(u, v) for u, v in zip(res["cuda:0"][0], res["cuda:0"][1]))
ok = ok and torch.equal(res["cuda:1"][2][0], res["cuda:1"][0][0])
print("MULTI-DEVICE CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
