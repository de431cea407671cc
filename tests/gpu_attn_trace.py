#!/usr/bin/env python
"""GPU-box tool: per-tile phase timeline (clock64) of ONE CTA of the fused attention forward, the heaviest
query tile at [1, 2048, 32, 128] — where a tile's ~N thousand cycles actually go."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import llm_qat_b200  # noqa: E402
from llm_qat_b200 import _lib  # noqa: E402
from llm_qat_b200.fused_ops import causal_attention  # noqa: E402

L = _lib.lib()
B, S, H = 1, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 32
q, k, v = (torch.randn(B, S, H, 128, device="cuda").bfloat16() for _ in range(3))
for _ in range(3):
    causal_attention(q, k, v)
buf = torch.zeros(2 * 3 * 16 * 8, dtype=torch.int64, device="cuda")
qg, kg, vg = (x.clone().requires_grad_(True) for x in (q, k, v))
L.qat_attn_debug_trace(buf.data_ptr())
o = causal_attention(qg, kg, vg)
o.backward(torch.randn_like(o))
torch.cuda.synchronize()
L.qat_attn_debug_trace(0)
t = buf[:384].view(3, 16, 8).cpu()
t0 = int(t[t > 0].min())
names = {0: ["qk:start", "Kfull", "Sfree", "qk:issued", "pv:start", "Vfull", "Pfull", "pv:issued"],
         1: ["start", "Sfull", "ld done", "max done", "exp done", "PVdone+resc", "sts done", "arrived"]}
for role, nm in ((0, "MMA thread"), (1, "warpgroup 0"), (2, "warpgroup 1")):
    print(f"== {nm}: cycles since first event; columns = {names[min(role, 1)]}")
    for j in range(16):
        row = [int(x) - t0 if x > 0 else -1 for x in t[role, j]]
        if max(row) >= 0:
            print(f"tile {j:2d}: " + " ".join(f"{x:7d}" for x in row))

t = buf[384:].view(3, 16, 8).cpu()
t0 = int(t[t > 0].min())
names = {0: ["sp:start", "K,Vfull", "SPfree", "sp:issued", "dq:start", "DSfull", "dq:issued", "-"],
         1: ["start", "SPfull", "ld done", "ds done", "DSfree", "arrived", "-", "-"]}
print("==== attn_bwd_dq_kernel, CTA (0,0,0) = heaviest query tile, 64-key steps")
for role, nm in ((0, "MMA thread"), (1, "warpgroup 0"), (2, "warpgroup 1")):
    print(f"== {nm}: cycles since first event; columns = {names[min(role, 1)]}")
    for j in range(16):
        row = [int(x) - t0 if x > 0 else -1 for x in t[role, j]]
        if max(row) >= 0:
            print(f"step {j:2d}: " + " ".join(f"{x:7d}" for x in row))
