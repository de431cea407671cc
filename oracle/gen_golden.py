#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference and pin the oracle to it.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/gen_golden.py``

What it does
  1. imports the unmodified reference module /root/reference/models/utils_quant.py
     (read-only; no bytecode written);
  2. builds seeded inputs (SURVEY.md section 8d distributions + appendix A.4
     edge rows) in fp32 and bf16;
  3. runs the reference's SymQuantizer / AsymQuantizer forward+backward and
     QuantizeLinear forward+backward on CPU;
  4. checks oracle/quant_oracle.py against those outputs bit-for-bit (exits
     non-zero on any mismatch), and
  5. stores inputs and reference outputs as raw bit patterns under
     tests/golden/ so the CPU and GPU test-suites can run without the reference.

Test infrastructure only — nothing in the product imports this.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402

from oracle import quant_oracle as qo  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(os.cpu_count() or 1)


def load_reference():
    import importlib

    mod = importlib.import_module("models.utils_quant")
    assert mod.__file__.startswith(REF), mod.__file__
    return mod


# ------------------------------------------------------------------ inputs
def edge_rows(cols: int) -> torch.Tensor:
    """Appendix A.4 rows: +-2.0 boundaries, signed zeros, outlier, all-zero,
    huge, tiny, NaN, +inf, -inf."""
    g = torch.Generator().manual_seed(99)
    base = torch.randn(11, cols, generator=g) * 0.5
    r = base.clone()
    r[0, ::3] = 2.0
    r[0, 1::3] = -2.0
    r[1, ::2] = 0.0
    r[1, 1::2] = -0.0
    r[2, cols // 2] = 6.0
    r[3, :] = 0.0
    r[4, :] = base[4] * 1e30
    r[5, :] = base[5] * 1e-30
    r[6, cols // 3] = float("nan")
    r[7, cols // 4] = float("inf")
    r[8, cols // 5] = float("-inf")
    r[9, :] = 1.9999999          # just inside the clip (fp32) / rounds to 2.0 (bf16)
    r[10, :] = -2.0000002
    return r


def make_inputs():
    g = torch.Generator().manual_seed(1234)
    inp = {}
    body = torch.randn(21, 172, generator=g) * 0.5
    inp["r2d_172"] = torch.cat([body, edge_rows(172)], 0)           # [32,172] odd width
    inp["r2d_4096"] = torch.cat([torch.randn(5, 4096, generator=g) * 0.5, edge_rows(4096)], 0)
    x = torch.randn(6, 11008, generator=g) * 0.02                   # weight-like, 11008 cols
    x[1, 77] = 0.6
    inp["r2d_11008"] = x
    inp["r1d_300"] = torch.randn(300, generator=g)
    inp["r3d"] = torch.randn(2, 5, 256, generator=g) * 1.5          # K/V-like [b, s, h]
    inp["r4d"] = torch.randn(2, 3, 4, 40, generator=g)              # per-(b,h)
    inp["tiny"] = torch.tensor([[0.3]])
    inp["cols7"] = torch.randn(9, 7, generator=g)
    grads = {k: torch.randn(v.shape, generator=g) for k, v in inp.items()}
    return inp, grads


CASES = [
    # (input key, layerwise)
    ("r2d_172", False), ("r2d_172", True),
    ("r2d_4096", False),
    ("r2d_11008", False),
    ("r1d_300", False),
    ("r3d", False), ("r3d", True),
    ("r4d", False), ("r4d", True),
    ("tiny", False),
    ("cols7", False),
]
BITS = {"sym": (2, 3, 4, 6, 8), "asym": (3, 4, 8)}
CLIPS = [(-2.0, 2.0), (-1.3, 2.3)]


def tbits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    return t.contiguous().view(torch.int32).numpy().view(np.uint32)


def tf32(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().numpy()


# ------------------------------------------------------------------ main
def main() -> int:
    ref = load_reference()
    inp32, grad32 = make_inputs()
    os.makedirs(GOLD, exist_ok=True)
    bad = 0
    checked = 0
    for dtype, tdt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        store = {}
        for k in inp32:
            store[f"in/{k}"] = tbits(inp32[k].to(tdt))
            store[f"grad/{k}"] = tbits(grad32[k].to(tdt))
        for qname, cls, ofn in (("sym", ref.SymQuantizer, qo.sym_forward),
                                ("asym", ref.AsymQuantizer, qo.asym_forward)):
            for key, layerwise in CASES:
                x = inp32[key].to(tdt)
                g = grad32[key].to(tdt)
                for bits in BITS[qname]:
                    xr = x.clone().requires_grad_(True)
                    y = cls.apply(xr, torch.tensor([-2.0, 2.0]), bits, layerwise)
                    y.backward(g)
                    tag = f"{qname}/{key}/b{bits}/{'lw' if layerwise else 'row'}"
                    store[f"y/{tag}"] = tbits(y.detach())
                    o = ofn(tf32(x), bits, layerwise, dtype)
                    nm = qo.count_mismatch(o["y"], tf32(y))
                    checked += 1
                    if nm:
                        bad += 1
                        print(f"MISMATCH fwd {dtype} {tag}: {nm}/{y.numel()}")
                    if bits == BITS[qname][0]:
                        for lo, hi in CLIPS:
                            xr = x.clone().requires_grad_(True)
                            y2 = cls.apply(xr, torch.tensor([lo, hi]), bits, layerwise)
                            y2.backward(g)
                            ctag = f"{qname}/{key}/{'lw' if layerwise else 'row'}/clip{lo}_{hi}"
                            store[f"gx/{ctag}"] = tbits(xr.grad)
                            ob = qo.ste_backward(tf32(g), tf32(x), lo, hi, dtype)
                            nm = qo.count_mismatch(ob["gx"], tf32(xr.grad))
                            checked += 1
                            if nm:
                                bad += 1
                                print(f"MISMATCH bwd {dtype} {ctag}: {nm}/{y.numel()}")
        # ---- QuantizeLinear (utils_quant.py:165-254) -----------------------
        g = torch.Generator().manual_seed(4321)
        xl = (torch.randn(3, 10, 192, generator=g)).to(tdt)
        xl.view(-1)[::517] *= 20.0
        wl = (torch.randn(80, 192, generator=g) * 0.02).to(tdt)
        gl = torch.randn(3, 10, 80, generator=g).to(tdt)
        store["lin/x"], store["lin/w"], store["lin/g"] = tbits(xl), tbits(wl), tbits(gl)
        for w_bits, a_bits, sym in ((4, 8, True), (8, 8, True), (4, 8, False), (4, 32, True),
                                    (32, 8, True), (1, 8, True), (2, 8, True), (2, 32, True)):
            for wlw in (False, True):
                if wlw and w_bits not in (1, 2, 4):
                    continue
                lin = ref.QuantizeLinear(192, 80, symmetric=sym, w_bits=w_bits, a_bits=a_bits,
                                         weight_layerwise=wlw).to(tdt)
                with torch.no_grad():
                    lin.weight.copy_(wl)
                xr = xl.clone().requires_grad_(True)
                out = lin(xr)
                out.backward(gl)
                tag = f"w{w_bits}a{a_bits}{'s' if sym else 'a'}{'_wlw' if wlw else ''}"
                store[f"lin/out/{tag}"] = tbits(out.detach())
                store[f"lin/gx/{tag}"] = tbits(xr.grad)
                store[f"lin/gw/{tag}"] = tbits(lin.weight.grad)
                o = qo.qlinear_forward(tf32(xl), tf32(wl), w_bits, a_bits, False, wlw, sym, dtype)
                ref_out = tf32(out)
                denom = np.linalg.norm(ref_out.astype(np.float64)) + 1e-30
                rel = np.linalg.norm((o["out"] - ref_out).astype(np.float64)) / denom
                tol = 2e-6 if dtype == "fp32" else 4e-3
                checked += 1
                if not rel <= tol:
                    bad += 1
                    print(f"MISMATCH qlinear {dtype} {tag}: rel={rel:.3e} > {tol}")
        # ---- low-bit weight path (utils_quant.py:202-242) --------------------
        # The reference never exposes the effective weight; feeding the identity
        # through an a_bits=32 layer returns it exactly (one non-zero term per sum).
        g = torch.Generator().manual_seed(777)
        wlb = torch.randn(40, 192, generator=g) * 0.02
        wlb[3] = 0.0
        wlb[6, ::2] = 0.0
        wlb = wlb.to(tdt)
        store["lowbit/w"] = tbits(wlb)
        for w_bits in (1, 2):
            for wlw in (False, True):
                lin = ref.QuantizeLinear(192, 40, w_bits=w_bits, a_bits=32, weight_layerwise=wlw).to(tdt)
                with torch.no_grad():
                    lin.weight.copy_(wlb)
                weff = lin(torch.eye(192, dtype=tdt)).detach().t().contiguous()
                tag = f"w{w_bits}{'_lw' if wlw else ''}"
                store[f"lowbit/weff/{tag}"] = tbits(weff)
                o = qo.lowbit_weight(tf32(wlb), w_bits, wlw, dtype)["w_eff"]
                checked += 1
                nm = qo.count_mismatch(o, tf32(weff))   # bit for bit in both dtypes (mean|w| in torch's own order)
                if nm:
                    bad += 1
                    print(f"MISMATCH lowbit {dtype} {tag}: {nm}/{weff.numel()}")
        path = os.path.join(GOLD, f"quant_{dtype}.npz")
        np.savez_compressed(path, **store)
        print(f"wrote {path}: {len(store)} arrays, {os.path.getsize(path)/1e6:.2f} MB")
    print(f"oracle vs live reference: {checked} comparisons, {bad} mismatching")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
