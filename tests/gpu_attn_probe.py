"""GPU probe for the fused attention kernels (csrc/attention.cu): forward and backward against the
reference's eager op chain (modeling_llama_quant.py:352-377) and an fp32 statement of it; timing at the
LLaMA-7B shape.  Run on the GPU box:  python tests/gpu_attn_probe.py"""
from __future__ import annotations

import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llm_qat_b200  # noqa: E402,F401
from llm_qat_b200.fused_ops import causal_attention  # noqa: E402

dev = torch.device("cuda", 0)


def eager(q, k, v, causal=True):
    """the reference chain on [B,S,H,D] inputs (transposed to [B,H,S,D] like the model does)"""
    B, S, H, D = q.shape
    qh, kh, vh = (t.transpose(1, 2) for t in (q, k, v))
    w = torch.matmul(qh, kh.transpose(2, 3)) / math.sqrt(D)
    if causal:
        m = torch.full((S, S), torch.finfo(w.dtype).min, device=w.device, dtype=w.dtype).triu(1)
        w = w + m[None, None]
        w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, device=w.device))
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(qh.dtype)
    return torch.matmul(w, vh).transpose(1, 2)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(B, S, H, causal, seed=0, scale_in=1.0):
    g = torch.Generator().manual_seed(seed)
    mk = lambda: (torch.randn(B, S, H, 128, generator=g) * scale_in).bfloat16().to(dev).requires_grad_(True)  # noqa: E731
    q, k, v = mk(), mk(), mk()
    go = torch.randn(B, S, H, 128, generator=g).bfloat16().to(dev)
    out = {"B": B, "S": S, "H": H, "causal": causal}
    try:
        o = causal_attention(q, k, v, causal=causal)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        out["fwd_exc"] = repr(e)[:200]
        return out
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    of = eager(qf, kf, vf, causal)
    of.backward(go.float())
    q2, k2, v2 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    oe = eager(q2, k2, v2, causal)
    oe.backward(go)
    out["fwd_vs_fp32"] = rel(o, of)
    out["eager_vs_fp32"] = rel(oe, of)
    out["fwd_finite"] = bool(torch.isfinite(o).all())
    if os.environ.get("PROBE_FWD_ONLY"):
        return out
    try:
        o.backward(go)
        torch.cuda.synchronize()
        for n, a, e_, f in (("dq", q.grad, q2.grad, qf.grad), ("dk", k.grad, k2.grad, kf.grad),
                            ("dv", v.grad, v2.grad, vf.grad)):
            out[n + "_vs_fp32"] = rel(a, f)
            out[n + "_eager_vs_fp32"] = rel(e_, f)
    except Exception as e:  # noqa: BLE001
        out["bwd_exc"] = repr(e)[:200]
    return out


def main():
    ok = True
    for (B, S, H, causal) in [(1, 128, 1, True), (1, 128, 1, False), (1, 256, 2, True), (2, 512, 3, True),
                              (1, 200, 2, True), (1, 72, 1, True), (1, 1024, 4, False), (1, 2048, 4, True)]:
        r = run(B, S, H, causal, scale_in=1.0)
        keys = ("fwd_vs_fp32",) if os.environ.get("PROBE_FWD_ONLY") else ("fwd_vs_fp32", "dq_vs_fp32", "dk_vs_fp32", "dv_vs_fp32")
        good = all(r.get(k_, 1.0) < 2e-2 for k_ in keys) and \
            "fwd_exc" not in r and "bwd_exc" not in r
        r["ok"] = good
        ok &= good
        print(json.dumps(r), flush=True)
    # large-magnitude scores (exercises the lazy rescale) at one shape
    r = run(1, 512, 2, True, seed=3, scale_in=4.0)
    r["big"] = True
    print(json.dumps(r), flush=True)

    if os.environ.get("PROBE_FWD_ONLY"):
        print("PROBE", "OK" if ok else "FAILED", flush=True)
        sys.exit(0 if ok else 1)
    # timing, LLaMA-7B layer: B=1, S=2048, H=32
    B, S, H = 1, 2048, 32
    q, k, v = (torch.randn(B, S, H, 128, device=dev).bfloat16().requires_grad_(True) for _ in range(3))
    go = torch.randn(B, S, H, 128, device=dev).bfloat16()

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    fl_f = 4.0 * S * S * 128 * H * B / 2
    t_f = timeit(lambda: causal_attention(q, k, v))
    o = causal_attention(q, k, v)
    t_b = timeit(lambda: torch.autograd.grad(o, (q, k, v), go, retain_graph=True))
    t_ef = timeit(lambda: eager(q, k, v))
    oe = eager(q, k, v)
    t_eb = timeit(lambda: torch.autograd.grad(oe, (q, k, v), go, retain_graph=True))
    t_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(
        q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True))
    print(json.dumps({"shape": [B, S, H, 128], "fwd_us": round(t_f, 1), "fwd_TF": round(fl_f / t_f / 1e6, 1),
                      "bwd_us": round(t_b, 1), "bwd_TF": round(2.5 * fl_f / t_b / 1e6, 1),
                      "eager_fwd_us": round(t_ef, 1), "eager_bwd_us": round(t_eb, 1),
                      "sdpa_fwd_us": round(t_sdpa, 1)}), flush=True)
    print("PROBE", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
