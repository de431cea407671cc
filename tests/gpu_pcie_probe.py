#!/usr/bin/env python
"""GPU-box tool: what the host link can do — pinned H2D, D2H and both at once —
the ceiling of bench.py's e2e figure (314.6 MB each way per step)."""
import torch
n = 314572800
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
for name, a, b in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both directions at once", 1, 1)):
    run(a, b, 2)
    ms = run(a, b)
    print(f"{name}: {ms:.2f} ms per 314.6 MB -> {n/ms/1e6:.1f} GB/s per direction", flush=True)
