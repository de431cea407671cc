"""Host-buffer front end: fake-quant forward + STE backward on tensors that live
in (pinned) host memory — the end-to-end shape of the reference's call when its
inputs are CPU tensors (utils_quant.py:37-87).  One call to the C ABI's
``qat_{sym,asym}_fwd_bwd_host`` pipelines row chunks through
H2D | kernels | D2H on three streams; results land in pinned host tensors.
"""
from __future__ import annotations

import torch

from . import _lib
from .utils_quant import _DTYPES, _clip_bounds, _reduction_view, _stream_ptr

_scratch = {}


def _device_scratch(nbytes: int, dev: torch.device, stream: int) -> torch.Tensor:
    """Staging memory of the pipeline, kept per (device, stream): calls on one stream are ordered, so they
    can share it; a call on another stream (or another device — "cuda" means the CURRENT one) gets its own."""
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (index, stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", index))
        _scratch[key] = buf
    return buf


def fake_quant_fwd_bwd_host(x: torch.Tensor, g: torch.Tensor | None, clip_val, num_bits: int, *,
                            symmetric: bool = True, device="cuda", y: torch.Tensor | None = None,
                            gx: torch.Tensor | None = None):
    """``x`` (and optionally ``g``): contiguous CPU tensors, fp32 or bf16, per-row
    (non-layerwise) reduction as in ``Quantizer.apply(x, clip, bits, False)``.
    Returns ``(y, gx)`` in pinned host memory; asynchronous on the current CUDA
    stream of ``device`` — synchronize (or wait on the stream) before reading."""
    if x.is_cuda or (g is not None and g.is_cuda):
        raise RuntimeError("fake_quant_fwd_bwd_host takes host tensors; use SymQuantizer.apply for CUDA tensors")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: llm-qat_b200 has no CPU fallback")
    if x.dtype not in _DTYPES:
        raise TypeError(f"float32 or bfloat16 expected, got {x.dtype}")
    if g is not None and (g.dtype != x.dtype or g.shape != x.shape):
        raise RuntimeError("grad must match the input's dtype and shape")
    x = x.contiguous()
    g = g.contiguous() if g is not None else None
    rows, cols = _reduction_view(x, False)
    lo, hi = _clip_bounds(clip_val)
    dt = _DTYPES[x.dtype]
    if y is None:
        y = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    if g is not None and gx is None:
        gx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    L = _lib.lib()
    dev = torch.device(device)
    nbytes = int(L.qat_host_scratch_bytes(rows, cols, dt, 1 if g is not None else 0))
    stream = _stream_ptr(dev)
    scratch = _device_scratch(nbytes, dev, stream)
    fn = L.qat_sym_fwd_bwd_host if symmetric else L.qat_asym_fwd_bwd_host
    with torch.cuda.device(scratch.device):
        rc = fn(x.data_ptr(), g.data_ptr() if g is not None else 0, y.data_ptr(),
                gx.data_ptr() if gx is not None else 0, lo, hi, rows, cols, dt, int(num_bits),
                scratch.data_ptr(), scratch.numel(), stream)
    _lib.check(rc, "qat_fwd_bwd_host")
    return y, gx
