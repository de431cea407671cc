#!/usr/bin/env python
"""GPU-box tool (not a pytest file): the reference's eager op chain
(oracle/torch_chain.py port) on CUDA tensors versus (a) the same chain on CPU —
SURVEY.md section 7's "possible CPU != CUDA divergence inside the reference" —
and (b) our kernels, with CUDA-event timings of both.  Writes
gpurun_out/eager_compare.json.  Test/bench infrastructure: it may import oracle/.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import torch_chain as tc  # noqa: E402


def bits(t):
    return t.contiguous().view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32)


def nmis(a, b):
    a, b = a.cpu(), b.cpu()
    both_nan = torch.isnan(a) & torch.isnan(b)
    return int(((bits(a) != bits(b)) & ~both_nan).sum())


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from llm_qat_b200 import AsymQuantizer, SymQuantizer
    from llm_qat_b200.utils_quant import fake_quant_forward, ste_backward

    out = {"divergence": {}, "timing_ms": {}}
    clip = torch.tensor([-2.0, 2.0])
    for dname, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        g = torch.Generator().manual_seed(1234)
        x = (torch.randn(2048, 4096, generator=g) * 0.5).to(dt)
        gr = torch.randn(2048, 4096, generator=g).to(dt)
        xc, gc = x.cuda(), gr.cuda()
        for sym in (True, False):
            for nbits in (4, 8):
                fwd = tc.sym_forward if sym else tc.asym_forward
                y_cpu, y_gpu = fwd(x, nbits), fwd(xc, nbits)
                ours = (SymQuantizer if sym else AsymQuantizer).apply(xc, clip, nbits, False)
                key = f"{'sym' if sym else 'asym'}{nbits}_{dname}"
                out["divergence"][key] = {
                    "eager_cuda_vs_eager_cpu": nmis(y_gpu, y_cpu),
                    "ours_vs_eager_cpu": nmis(ours, y_cpu),
                    "ours_vs_eager_cuda": nmis(ours, y_gpu),
                    "numel": x.numel(),
                }
        gx_cpu = tc.ste_backward(gr, x, clip)
        gx_gpu = tc.ste_backward(gc, xc, clip)
        out["divergence"][f"ste_{dname}"] = {"eager_cuda_vs_eager_cpu": nmis(gx_gpu, gx_cpu),
                                              "ours_vs_eager_cpu": nmis(ste_backward(gc, xc, clip), gx_cpu)}
    # timings at BASELINE config 1 / config 2 operand shapes
    for dname, dt, shape in (("fp32", torch.float32, (8192, 4096)), ("bf16", torch.bfloat16, (8192, 4096)),
                             ("bf16", torch.bfloat16, (11008, 4096))):
        g = torch.Generator().manual_seed(1)
        xc = (torch.randn(*shape, generator=g) * 0.5).to(dt).cuda()
        gc = torch.randn(*shape, generator=g).to(dt).cuda()
        for sym in (True, False):
            for nbits in (4, 8):
                key = f"{'sym' if sym else 'asym'}{nbits}_{dname}_{shape[0]}x{shape[1]}"
                e_f = timeit(lambda: (tc.sym_forward if sym else tc.asym_forward)(xc, nbits))
                e_b = timeit(lambda: tc.ste_backward(gc, xc, clip))
                o_f = timeit(lambda: fake_quant_forward(xc, nbits, False, sym))
                o_b = timeit(lambda: ste_backward(gc, xc, clip))
                esz = 4 if dt == torch.float32 else 2
                out["timing_ms"][key] = {
                    "eager_cuda_fwd": round(e_f, 4), "eager_cuda_bwd": round(e_b, 4),
                    "ours_fwd": round(o_f, 4), "ours_bwd": round(o_b, 4),
                    "speedup_fwd_bwd": round((e_f + e_b) / (o_f + o_b), 2),
                    "ours_fwd_bwd_GBps": round(xc.numel() * 5 * esz / ((o_f + o_b) * 1e6), 1),
                    "eager_fwd_bwd_GBps": round(xc.numel() * 5 * esz / ((e_f + e_b) * 1e6), 1),
                }
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "eager_compare.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["divergence"], indent=1))
    for k, v in out["timing_ms"].items():
        print(k, v)


if __name__ == "__main__":
    main()
