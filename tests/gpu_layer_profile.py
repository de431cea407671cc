#!/usr/bin/env python
"""GPU-box tool: where does a LLaMA-7B decoder layer's fwd+bwd time go?
torch.profiler kernel table for the product (fused) and the reference eager path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import faulthandler
# one run of this tool under torch.profiler (CUPTI tracing) with programmatic dependent launch on did not
# return (not reproduced in five later runs, tests/gpu_profiler_check.py passes): trace with plain stream order
faulthandler.dump_traceback_later(int(os.environ.get('PROFILE_TOOL_TIMEOUT', '240')), exit=True)   # a stuck run reports where
import torch
from torch.profiler import profile, ProfilerActivity
from harness import llama_qat as H
from oracle import ref_module as R
import llm_qat_b200

cfg = H.QatConfig.llama_7b(w_bits=4, a_bits=8, kv_bits=4)
seq, bsz = 2048, 1
AUTOCAST = len(sys.argv) > 1 and sys.argv[1] == "autocast"   # the recipe's context (kd_trainer.py:106)
import contextlib
ONLY = [a for a in sys.argv[1:] if a != "autocast"]
for name, quant in (("b200_fused_model", llm_qat_b200.utils_quant), ("b200_fused", llm_qat_b200.utils_quant),
                    ("reference_eager", R)):
    if ONLY and name not in ONLY:
        continue
    torch.manual_seed(0)
    layer = H.DecoderLayer(cfg, quant).bfloat16().cuda()
    if name == "b200_fused_model":
        llm_qat_b200.fuse_model(layer)
    x = torch.randn(bsz, seq, cfg.hidden_size).bfloat16().cuda().requires_grad_(True)
    go = torch.randn(bsz, seq, cfg.hidden_size).bfloat16().cuda()
    mask = H.causal_mask(bsz, seq, torch.bfloat16, "cuda")
    if name == "b200_fused_model":
        llm_qat_b200.mark_causal_mask(mask)
    pos = torch.arange(seq, device="cuda")[None].expand(bsz, seq)
    def step():
        with (torch.autocast("cuda", dtype=torch.bfloat16) if AUTOCAST else contextlib.nullcontext()):
            y = layer(x, mask, pos)
        y.backward(go.to(y.dtype)); x.grad = None
        for p in layer.parameters(): p.grad = None
    for _ in range(3): step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(10): step()
    t_cpu = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 10 * 1e3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    print(f"==== {name}: cpu-enqueue {t_cpu:.2f} ms/step, wall {t_all:.2f} ms/step")
    ka = [e for e in prof.key_averages() if getattr(e, "device_type", None) is not None
          and "CUDA" in str(e.device_type) and getattr(e, "self_device_time_total", 0) > 0]
    tot = sum(e.self_device_time_total for e in ka)
    ours = sum(e.self_device_time_total for e in ka if "qat::" in e.key or "qat" in e.key.split("(")[0])
    gemm = sum(e.self_device_time_total for e in ka if "qat::" not in e.key and
               ("nvjet" in e.key or "gemm" in e.key.lower() or "cutlass" in e.key.lower()))
    print(f"device kernel time per step: {tot/3/1e3:.3f} ms = libqat_b200 {ours/3/1e3:.3f} + library GEMM {gemm/3/1e3:.3f} "
          f"+ ATen (attention, norms, residuals, copies) {(tot-ours-gemm)/3/1e3:.3f}; {sum(e.count for e in ka)//3} launches")
    for e in sorted(ka, key=lambda e: -e.self_device_time_total)[:45]:
        print(f"{e.self_device_time_total/3/1e3:8.3f} ms  x{e.count//3:4d}  {e.key[:110]}")
