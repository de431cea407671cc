#!/usr/bin/env python
"""GPU-box tool: timeline of a chunked H2D | kernel | D2H pipeline (torch streams), to see
where the host-buffer path loses time against the raw bidirectional link rate."""
import sys, torch
rows, cols = 11008, 4096
chunk_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dep = (sys.argv[2] != "nodep") if len(sys.argv) > 2 else True
hx = torch.randn(rows, cols).bfloat16().pin_memory(); hg = torch.randn(rows, cols).bfloat16().pin_memory()
hy = torch.empty_like(hx).pin_memory(); hgx = torch.empty_like(hx).pin_memory()
dx, dg, dy, dgx = [torch.empty(rows, cols, dtype=torch.bfloat16, device="cuda") for _ in range(4)]
s_in, s_out, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
def run(trace=False):
    ev = []
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    s_in.wait_stream(s_k); s_out.wait_stream(s_k)
    for r0 in range(0, rows, chunk_rows):
        r1 = min(rows, r0 + chunk_rows)
        with torch.cuda.stream(s_in):
            dx[r0:r1].copy_(hx[r0:r1], non_blocking=True); dg[r0:r1].copy_(hg[r0:r1], non_blocking=True)
            e_in = torch.cuda.Event(enable_timing=trace); e_in.record()
        if dep: s_k.wait_event(e_in)
        torch.mul(dx[r0:r1], 2.0, out=dy[r0:r1]); torch.mul(dg[r0:r1], 2.0, out=dgx[r0:r1])
        e_k = torch.cuda.Event(enable_timing=trace); e_k.record()
        with torch.cuda.stream(s_out):
            if dep: s_out.wait_event(e_k)
            hy[r0:r1].copy_(dy[r0:r1], non_blocking=True); hgx[r0:r1].copy_(dgx[r0:r1], non_blocking=True)
            e_out = torch.cuda.Event(enable_timing=trace); e_out.record()
        ev.append((e_in, e_k, e_out))
    s_k.wait_stream(s_out); s_k.wait_stream(s_in)
    t1 = torch.cuda.Event(enable_timing=True); t1.record(); t1.synchronize()
    return t0, t1, ev
for _ in range(3): run()
t0, t1, ev = run(True)
tot = t0.elapsed_time(t1)
mb = rows * cols * 2 / 1e6
print(f"chunk_rows={chunk_rows} dep={dep}: total {tot:.2f} ms for {mb:.0f} MB x2 each way -> {2*mb/tot:.1f} GB/s per direction")
for i, (a, b, c) in enumerate(ev):
    print(f"  chunk {i:2d}: in done {t0.elapsed_time(a):6.2f}  kernel done {t0.elapsed_time(b):6.2f}  out done {t0.elapsed_time(c):6.2f}")
