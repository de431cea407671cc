#!/usr/bin/env python
"""GPU-box tool: tile / grid sweep of the two grid-stride streaming kernels — `dequant_codes_kernel`
(QAT_B200_DEQUANT_UNROLL x QAT_B200_DEQUANT_CTAS) and `ste_bwd_kernel` (QAT_B200_STE_CTAS; 0 = one CTA per
tile) — at the LLaMA-7B operand shapes.  Every setting's output is compared bit for bit with the default
setting's.  CUDA events over 30 back-to-back launches, three rotating buffer sets (> L2).

    python tests/gpu_dequant_tune.py > gpurun_out/dequant_tune.json
"""
import json
import os
import sys

os.environ["QAT_B200_DEQUANT_TUNE"] = "1"
os.environ["QAT_B200_STE_TUNE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from llm_qat_b200 import _lib  # noqa: E402

PEAK = 6459.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
NBUF, REPS = 3, 30


def timed(once):
    for i in range(NBUF):
        once(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(REPS):
        once(k % NBUF)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / REPS * 1e3


out = {"what": __doc__.split("\n\n")[0].replace("\n", " "), "peak_GBps": PEAK, "dequant": {}, "ste": {}}
g = torch.Generator().manual_seed(0)

# ------------------------------------------------------------------ dequant_codes (bf16 out: 3 B/elem)
for rows, cols in ((11008, 4096), (8192, 4096), (4096, 4096), (2048, 11008), (2048, 4096)):
    n = rows * cols
    cs = [torch.randint(-7, 8, (rows, cols), generator=g, dtype=torch.int8).cuda() for _ in range(NBUF)]
    outs = [torch.empty(rows, cols, dtype=torch.bfloat16, device="cuda") for _ in range(NBUF)]
    for ekind in ("bf16_divisor", "fp32_divisor"):
        e = torch.rand(rows, generator=g) * 300 + 10
        e = (e.bfloat16().float() if ekind == "bf16_divisor" else e).cuda()
        res = {}
        ref = None
        for unroll in (4, 2, 8):
            for ctas in (8, 0, 4, 5, 6, 12, 16):
                os.environ["QAT_B200_DEQUANT_UNROLL"] = str(unroll)
                os.environ["QAT_B200_DEQUANT_CTAS"] = str(ctas)

                def once(i):
                    _lib.check(L.qat_dequant_codes(cs[i].data_ptr(), e.data_ptr(), outs[i].data_ptr(), rows, cols, 1, st))
                us = timed(once)
                if ref is None:
                    ref = outs[0].clone()
                same = bool(torch.equal(outs[0].view(torch.int16), ref.view(torch.int16)))
                res[f"unroll{unroll}_ctas{ctas}"] = {"us": round(us, 2), "frac": round(n * 3 / us / 1e3 / PEAK, 4),
                                                     "bit_equal_to_default": same}
        out["dequant"][f"bf16[{rows},{cols}] {ekind}"] = res
    del cs, outs

# ------------------------------------------------------------------ ste_bwd (3e B/elem from x, 2e + 1/8 from a mask)
for dt_name, dt, tdt, esz in (("bf16", 1, torch.bfloat16, 2), ("fp32", 0, torch.float32, 4)):
    for rows, cols in ((8192, 4096), (11008, 4096), (2048, 4096)):
        n = rows * cols
        xs = [(torch.randn(rows, cols, generator=g) * 1.5).to(tdt).cuda() for _ in range(NBUF)]
        gs = [torch.randn(rows, cols, generator=g).to(tdt).cuda() for _ in range(NBUF)]
        ds = [torch.empty_like(t) for t in xs]
        ms = [torch.empty(n // 8, dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
        for mode in ("from_x", "from_x_mask_out", "from_mask"):
            res = {}
            ref = refm = None
            for ctas in (8, 0, 4, 5, 6, 12, 16, 32):
                os.environ["QAT_B200_STE_CTAS"] = str(ctas)

                def once(i):
                    if mode == "from_mask":
                        rc = L.qat_ste_bwd_from_mask(gs[i].data_ptr(), ms[i].data_ptr(), ds[i].data_ptr(), n, dt, st)
                    else:
                        rc = L.qat_ste_bwd(gs[i].data_ptr(), xs[i].data_ptr(), ds[i].data_ptr(),
                                           ms[i].data_ptr() if mode == "from_x_mask_out" else 0, -2.0, 2.0, n, dt, st)
                    _lib.check(rc)
                us = timed(once)
                if ref is None:
                    ref, refm = ds[0].clone(), ms[0].clone()
                same = bool(torch.equal(ds[0].view(torch.uint8), ref.view(torch.uint8)) and torch.equal(ms[0], refm))
                nbytes = n * esz * 3 if mode == "from_x" else n * esz * 3 + n // 8 if mode == "from_x_mask_out" \
                    else n * esz * 2 + n // 8
                res[f"ctas{ctas}"] = {"us": round(us, 2), "frac": round(nbytes / us / 1e3 / PEAK, 4),
                                      "bit_equal_to_default": same}
            out["ste"][f"{dt_name}[{rows},{cols}] {mode}"] = res
        del xs, gs, ds, ms

print(json.dumps(out, indent=1))
