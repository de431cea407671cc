#!/usr/bin/env python
"""GPU-box tool: the e2e leg of bench.py alone (host buffers -> qat_sym_fwd_bwd_host -> host buffers, configs[1]
operands), for one pipeline setting per process: QAT_B200_HOST_SCHEDULE=split|chunk, QAT_B200_HOST_CHUNK_MB=n.
`python tests/gpu_e2e_probe.py sweep` runs the settings below in subprocesses and prints one line each."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch

    import bench
    from llm_qat_b200.host_api import fake_quant_fwd_bwd_host

    dev = torch.device("cuda", 0)
    hx, hw, hgx, hgw = [t.pin_memory() for t in bench.make_inputs(1234)]
    yx, dxh, yw, dwh = [torch.empty_like(t).pin_memory() for t in (hx, hx, hw, hw)]

    def step():
        fake_quant_fwd_bwd_host(hx, hgx, bench.CLIP, bench.A_BITS, symmetric=True, device=dev, y=yx, gx=dxh)
        fake_quant_fwd_bwd_host(hw, hgw, bench.CLIP, bench.W_BITS, symmetric=True, device=dev, y=yw, gx=dwh)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    best, tot, n = 1e9, 0.0, 12
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    print(json.dumps({"schedule": os.environ.get("QAT_B200_HOST_SCHEDULE", "split"),
                      "chunk_mb": int(os.environ.get("QAT_B200_HOST_CHUNK_MB", "8")),
                      "ms_per_step_mean": round(tot / n, 3), "ms_per_step_best": round(best, 3),
                      "GBps_mean": round(bench.STEP_BYTES / (tot / n) / 1e6, 1),
                      "GBps_per_direction": round(bench.ELEMS * 4 / (tot / n) / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        for sched, mb in (("chunk", 8), ("split", 8), ("split", 16), ("split", 32)):
            env = dict(os.environ, QAT_B200_HOST_SCHEDULE=sched, QAT_B200_HOST_CHUNK_MB=str(mb))
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
    else:
        one()
