"""CPU oracle for LLM-QAT's fake-quantization hot path (TEST INFRASTRUCTURE ONLY).

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``llm-qat_b200``) never
imports anything under ``oracle/`` and has no CPU fallback.

It restates, in numpy with one IEEE rounding per line, the arithmetic that the
reference executes through eager PyTorch:

* ``sym_forward``      <- /root/reference/models/utils_quant.py:37-74   (SymQuantizer.forward)
* ``asym_forward``     <- /root/reference/models/utils_quant.py:96-149  (AsymQuantizer.forward)
* ``ste_backward``     <- /root/reference/models/utils_quant.py:77-87, 152-162 (both backward()s)
* ``lowbit_weight``    <- /root/reference/models/utils_quant.py:202-242 (w_bits in {1,2})
* ``qlinear_forward``  <- /root/reference/models/utils_quant.py:190-254 (QuantizeLinear.forward)
* ``kv_fake_quant``    <- /root/reference/models/modeling_llama_quant.py:320-327

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the pin is the live reference itself:
``oracle/gen_golden.py`` imports ``/root/reference/models/utils_quant.py`` in
the build container, runs it on seeded inputs, checks this restatement against
it bit-for-bit and commits the input/output vectors under ``tests/golden/``.
``tests/test_oracle.py`` re-checks the restatement against those fixtures.

The arithmetic is an *op-order contract* (SURVEY.md appendix A): ``Q / t`` is
``reciprocal(t) * Q`` in torch (two roundings), the dequant divisor is
``s + 1e-6``, Asym's ``q / S * a + b`` is three separately rounded ops, and in
bf16 every intermediate is rounded to bf16.  ``fl()`` below is that rounding.

Unlike the reference, the oracle also returns the integer codes and the STE
mask, which the reference never materialises; they are what the CUDA kernels'
optional ``codes`` / ``mask`` outputs are compared with.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPS_SYM = F32(1e-6)    # utils_quant.py:71-72 (python float -> fp32 opmath scalar)
EPS_ASYM = F32(1e-8)   # utils_quant.py:144,147


# --------------------------------------------------------------------------
# dtype emulation: values are always carried as float32 arrays; in "bf16" mode
# every array holds bf16-representable values and fl() re-rounds after each op.
# --------------------------------------------------------------------------
def bf16_round(a: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bf16 -> float32 (NaN stays NaN, inf stays inf)."""
    a = np.ascontiguousarray(a, dtype=F32)
    u = a.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) & 0xFFFF0000).astype(np.uint32)
    out = r.view(F32).copy()
    nan = np.isnan(a)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(a.shape)


def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """bf16-representable float32 array -> raw uint16 bit pattern."""
    a = np.ascontiguousarray(a, dtype=F32)
    return (a.view(np.uint32) >> 16).astype(np.uint16).reshape(a.shape)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint16)
    return (b.astype(np.uint32) << 16).view(F32).reshape(b.shape)


def _fl(dtype: str):
    if dtype == "fp32":
        return lambda a: np.asarray(a, dtype=F32)
    if dtype == "bf16":
        return bf16_round
    raise ValueError(f"dtype must be 'fp32' or 'bf16', got {dtype!r}")


# --------------------------------------------------------------------------
# reduction set: how the reference picks what one scale covers
# --------------------------------------------------------------------------
def as_rows(x: np.ndarray, layerwise: bool) -> np.ndarray:
    """View ``x`` as [rows, cols] where one row == one reduction set.

    utils_quant.py:50-70 / 110-143: layerwise -> the whole tensor; ndim <= 3
    -> the last dim; ndim == 4 -> the last two dims flattened per (d0, d1);
    ndim >= 5 -> ValueError.
    """
    if layerwise:
        return x.reshape(1, -1)
    if x.ndim <= 3:
        return x.reshape(-1, x.shape[-1]) if x.ndim >= 1 else x.reshape(1, 1)
    if x.ndim == 4:
        return x.reshape(x.shape[0] * x.shape[1], -1)
    raise ValueError("fake-quant input must have at most 4 dimensions")


def _nanmax(a: np.ndarray) -> np.ndarray:
    # np.max propagates NaN like torch.max
    return np.max(a, axis=1, keepdims=True)


# --------------------------------------------------------------------------
# SymQuantizer.forward  (utils_quant.py:37-74)
# --------------------------------------------------------------------------
def sym_forward(x: np.ndarray, num_bits: int, layerwise: bool = False, dtype: str = "fp32"):
    """Returns dict(y, codes, s, e) — y has x's shape; s, e are per-row [rows].

    dtype "bf16_amp": a bf16 tensor inside torch.autocast (what HF's Trainer wraps the step in,
    kd_trainer.py:106).  Autocast runs `reciprocal` in fp32, so `max + 1e-6` is still a bf16 op
    and everything after it — reciprocal, * Q, x * s (bf16 x fp32 promotes), round, s + 1e-6,
    the division — is fp32; y is a float32 tensor.  Pinned on the GPU against the restated
    chain run under the same autocast context (tests/test_gpu_parity.py)."""
    amp = dtype == "bf16_amp"
    fl = _fl("fp32" if amp else dtype)
    x = np.asarray(x, dtype=F32)
    x2 = as_rows(x, layerwise)
    Q = F32(2 ** (num_bits - 1) - 1)
    with np.errstate(all="ignore"):
        m = _nanmax(np.abs(x2))              # :51 / :56 / :63  (exact)
        d = fl(m + EPS_SYM)                  # :71  max_input + 1e-6
        if amp:
            d = bf16_round(d)
        r = fl(F32(1.0) / d)                 # :71  Q / d  ==  d.reciprocal() * Q
        s = fl(r * Q)
        p = fl(x2 * s)                       # :72  input * s
        q = np.rint(p).astype(F32)           # :72  torch.round == half-to-even
        e = fl(s + EPS_SYM)                  # :72  s + 1e-6
        y = fl(q / e)                        # :72  .div() — IEEE division
    return {
        "y": y.reshape(x.shape),
        "codes": q.reshape(x.shape),         # float-held integers (may be NaN)
        "s": s.reshape(-1),
        "e": e.reshape(-1),
    }


# --------------------------------------------------------------------------
# AsymQuantizer.forward  (utils_quant.py:96-149)
# --------------------------------------------------------------------------
def asym_forward(x: np.ndarray, num_bits: int, layerwise: bool = False, dtype: str = "fp32", div: str = "cpu"):
    """Returns dict(y, codes, a, beta) — a = alpha + 1e-8 per row, beta = row min.
    ``div``: how `.div(s)` (:146, s a Python int) is evaluated — "cpu": IEEE division (torch's CPU
    kernel; the pinned default), "cuda": multiply by fl32(1/S), ATen's CUDA kernel for a CPU-scalar
    divisor (pinned on the GPU box against the reference chain run eagerly on CUDA)."""
    fl = _fl(dtype)
    x = np.asarray(x, dtype=F32)
    x2 = as_rows(x, layerwise)
    S = F32(2 ** num_bits - 1)
    with np.errstate(all="ignore"):
        mx = np.max(x2, axis=1, keepdims=True)   # :111 / :118 / :128
        mn = np.min(x2, axis=1, keepdims=True)   # :112 / :124 / :137
        alpha = fl(mx - mn)
        beta = mn
        a = fl(alpha + EPS_ASYM)                 # :144
        n = fl(fl(x2 - beta) / a)                # :144
        q = np.rint(fl(n * S)).astype(F32)       # :146
        if div == "cuda":
            u = fl(q * (F32(1.0) / S))           # :146  .div(s) — ATen CUDA: a * (1 / b)
        else:
            u = fl(q / S)                        # :146  .div(s) — true division
        y = fl(fl(u * a) + beta)                 # :147  two roundings, no FMA
    return {
        "y": y.reshape(x.shape),
        "codes": q.reshape(x.shape),
        "a": a.reshape(-1),
        "beta": beta.reshape(-1),
    }


# --------------------------------------------------------------------------
# backward of both quantizers  (utils_quant.py:77-87, 152-162)
# --------------------------------------------------------------------------
def ste_backward(g: np.ndarray, x: np.ndarray, lo: float, hi: float, dtype: str = "fp32"):
    """gx = g where lo < x < hi else 0.  NaN x passes the gradient (both
    compares are false), +-inf is masked.  The clip bounds are compared in x's
    dtype (the 0-dim fp32 clip tensor does not promote a bf16 input)."""
    fl = _fl(dtype)
    g = np.asarray(g, dtype=F32)
    x = np.asarray(x, dtype=F32)
    lo_t = fl(np.asarray([lo], dtype=F32))[0]
    hi_t = fl(np.asarray([hi], dtype=F32))[0]
    with np.errstate(invalid="ignore"):
        clipped = (x >= hi_t) | (x <= lo_t)
    gx = np.where(clipped, F32(0.0), g).astype(F32)
    return {"gx": gx, "mask": ~clipped}


def pack_mask(mask: np.ndarray) -> np.ndarray:
    """Row-major bit packing used by the kernels' optional mask output: element
    i of the flattened tensor is bit (i % 8) of byte i // 8 (1 = grad passes)."""
    return np.packbits(np.asarray(mask, dtype=bool).reshape(-1), bitorder="little")


# --------------------------------------------------------------------------
# QuantizeLinear low-bit weight path  (utils_quant.py:202-242)
# --------------------------------------------------------------------------
LAYERWISE_SPLIT_NUMEL = 32768   # at::internal::GRAIN_SIZE: full reductions from this size on are thread-split


def _cascade(load, n: int):
    """ATen's ``multi_row_sum`` (aten/src/ATen/native/cpu/SumKernel.cpp): n steps over independent chains, level 0
    folded into level 1 after every 16 steps, 1 into 2 after every 256, 2 into 3 after every 4096."""
    ceil_log2 = 0 if n <= 1 else int(n - 1).bit_length()
    lp = max(4, ceil_log2 // 4)
    step, mask = 1 << lp, (1 << lp) - 1
    acc = [np.zeros_like(load(0)) for _ in range(4)]
    i = 0
    while i + step <= n:
        for _ in range(step):
            acc[0] = acc[0] + load(i)
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = np.zeros_like(acc[j - 1])
            if i & (mask << (j * lp)):
                break
    while i < n:
        acc[0] = acc[0] + load(i)
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    return acc[0]


def torch_cpu_row_sum(a: np.ndarray) -> np.ndarray:
    """``torch.sum(a, dim=-1)`` of a contiguous float32 [rows, cols] tensor on the CPU, bit for bit, as the AVX2
    build of ATen computes it (``cascade_sum`` -> ``vectorized_inner_sum``: vectors of 8 lanes, 4 vectors per step,
    a 4-level cascade per chain, left-over vectors, the fold over the 4 vectors, then tail and lanes in one scalar;
    rows shorter than one vector take the same scheme on scalars).  The kernel-side twin is
    llm-qat_b200/csrc/torch_sum_order.cuh; tests/test_torch_sum_order.py pins both to ``torch.sum`` itself.
    A row is never split over threads as long as there is more than one row (or fewer than 32768 elements)."""
    a = np.ascontiguousarray(a, dtype=F32)
    rows, cols = a.shape
    lanes = 8 if cols >= 8 else 1
    nvec = cols // lanes
    n = nvec // 4
    with np.errstate(all="ignore"):
        if n > 0:
            p = _cascade(lambda i: a[:, i * 4 * lanes:(i + 1) * 4 * lanes], n).reshape(rows, 4, lanes)
        else:
            p = np.zeros((rows, 4, lanes), F32)
        p0 = p[:, 0].copy()
        for v in range(n * 4, nvec):
            p0 = p0 + a[:, v * lanes:(v + 1) * lanes]
        for k in range(1, 4):
            p0 = p0 + p[:, k]
        fin = np.zeros(rows, F32)
        for e in range(nvec * lanes, cols):
            fin = fin + a[:, e]
        for lane in range(lanes):
            fin = fin + p0[:, lane]
    return fin


def _mean_rows(a: np.ndarray, layerwise: bool, dtype: str) -> np.ndarray:
    """``a.mean(dim=-1, keepdim=True)`` (:205-210, :219-224) as torch's CPU kernels compute it: the fp32 row sum in
    ATen's order (``torch_cpu_row_sum``), one true division by the column count, and for bf16 tensors a cast to
    fp32 before and one rounding to bf16 after (ReduceOps.cpp ``mean_out``) — bit-exact in both dtypes.
    Layerwise (``a.mean()`` of the whole tensor, unused by the model): below 32768 elements (at::internal::GRAIN_SIZE)
    the same cascade runs over the flattened tensor — exact; from there on torch splits the reduction over its
    intra-op threads, so the reference's own bits depend on the machine's thread count
    (tests/test_torch_sum_order.py shows it): that statistic is accumulated in float64 here and compared at a
    few-ulp tolerance in fp32."""
    fl = _fl(dtype)
    if layerwise and a.size >= LAYERWISE_SPLIT_NUMEL:
        return fl(np.asarray(np.mean(a.astype(np.float64)), dtype=F32).reshape(1, 1))
    if layerwise:          # one serial cascade over the flattened tensor: exact
        a = a.reshape(1, -1)
    with np.errstate(all="ignore"):
        return fl((torch_cpu_row_sum(a) / F32(a.shape[1])).astype(F32).reshape(-1, 1))


def lowbit_weight(w: np.ndarray, w_bits: int, layerwise: bool = False, dtype: str = "fp32"):
    """Effective forward weight for w_bits in {1, 2}: (q - w) + w  (:240-242)."""
    assert w_bits in (1, 2) and w.ndim == 2
    fl = _fl(dtype)
    w = np.asarray(w, dtype=F32)
    with np.errstate(all="ignore"):
        mean_abs = _mean_rows(np.abs(w), layerwise, dtype)
        if w_bits == 1:
            sf = mean_abs                                       # :205-210
            r = fl(w / sf)
            sgn = (r > 0).astype(F32) - (r < 0).astype(F32)     # torch.sign: NaN -> 0
            q = fl(sf * sgn)                                    # :211-213
        else:
            levels = F32(2 ** (w_bits - 1))                     # :217  "num_bits"
            clip = F32(1 - 1e-2)                                # :218
            sf = fl(F32(2.0) * mean_abs)                        # :219-224
            t = np.clip(fl(w / sf), -clip, clip)                # :229-231
            t = fl(fl(t * levels) - F32(0.5))                   # :232-233
            t = fl(np.rint(t).astype(F32) + F32(0.5))           # :228,235
            q = fl(fl(sf * t) / levels)                         # :226-237
        eff = fl(fl(q - w) + w)                                 # :240-242
    return {"w_eff": eff, "q": q, "sf": sf.reshape(-1)}


# --------------------------------------------------------------------------
# QuantizeLinear.forward  (utils_quant.py:190-254)
# --------------------------------------------------------------------------
def qlinear_forward(x, w, w_bits, a_bits, act_layerwise=False, weight_layerwise=False,
                    symmetric=True, dtype="fp32"):
    """out = F.linear(fake_quant(x), fake_quant(w)).  The contraction is done in
    float64 over the dequantized operands and rounded once to ``dtype`` — the
    reference's cuBLAS/MKL accumulation order is not part of the contract; the
    GEMM is compared at the tolerance north_star states (<= 1e-2 relative)."""
    fl = _fl(dtype)
    x = np.asarray(x, dtype=F32)
    w = np.asarray(w, dtype=F32)
    assert w.ndim == 2                                           # :192
    if w_bits >= 32:
        wq = w                                                   # :195-196
    elif w_bits >= 3:
        wq = sym_forward(w, w_bits, weight_layerwise, dtype)["y"]  # :197-201
    else:
        wq = lowbit_weight(w, w_bits, weight_layerwise, dtype)["w_eff"]
    if 2 < a_bits < 32:                                          # :244
        fwd = sym_forward if symmetric else asym_forward
        xq = fwd(x, a_bits, act_layerwise, dtype)["y"]
    else:
        xq = x
    out = xq.reshape(-1, x.shape[-1]).astype(np.float64) @ wq.astype(np.float64).T
    out = fl(out.astype(F32)).reshape(*x.shape[:-1], w.shape[0])
    return {"out": out, "xq": xq, "wq": wq}


# --------------------------------------------------------------------------
# K/V fake-quant call site  (modeling_llama_quant.py:320-327)
# --------------------------------------------------------------------------
def kv_fake_quant(k: np.ndarray, v: np.ndarray, kv_bits: int, dtype: str = "bf16"):
    """[bsz, q_len, hidden] K and V, one scale per token over all heads'
    channels, applied before the head split and RoPE; identity at kv_bits>=32."""
    if kv_bits >= 32:
        return k, v
    return (sym_forward(k, kv_bits, False, dtype)["y"],
            sym_forward(v, kv_bits, False, dtype)["y"])


# --------------------------------------------------------------------------
# bit-level comparison helpers shared by the tests
# --------------------------------------------------------------------------
def same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-for-bit equality of two float32 arrays, NaN == NaN for any payload."""
    return count_mismatch(a, b) == 0


def count_mismatch(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=F32)
    b = np.ascontiguousarray(b, dtype=F32)
    if a.shape != b.shape:
        return max(a.size, b.size)
    na, nb = np.isnan(a), np.isnan(b)
    neq = a.view(np.uint32) != b.view(np.uint32)
    return int(np.count_nonzero((neq & ~(na & nb)) | (na != nb)))
