"""Shared helpers for the parity tests: golden fixtures as raw bits, bit-level compares."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16}
_cache = {}


def golden(dtype: str):
    if dtype not in _cache:
        _cache[dtype] = np.load(os.path.join(GOLD, f"quant_{dtype}.npz"))
    return _cache[dtype]


def bits_to_f32(bits: np.ndarray, dtype: str) -> np.ndarray:
    """Stored raw bits -> float32 values (exact for both dtypes)."""
    if dtype == "bf16":
        return (bits.astype(np.uint32) << 16).view(np.float32)
    return bits.view(np.float32)


def bits_to_tensor(bits: np.ndarray, dtype: str, device="cpu") -> torch.Tensor:
    if dtype == "bf16":
        t = torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)
    else:
        t = torch.from_numpy(bits.view(np.int32).copy()).view(torch.float32)
    return t.to(device)


def tensor_to_f32(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def tensor_bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.view(torch.int32).numpy().view(np.uint32)


def mismatches(a_bits: np.ndarray, b_bits: np.ndarray, dtype: str) -> int:
    """Elements whose bit patterns differ, treating any NaN == any NaN."""
    if a_bits.shape != b_bits.shape:
        return max(a_bits.size, b_bits.size)
    a, b = bits_to_f32(a_bits, dtype), bits_to_f32(b_bits, dtype)
    na, nb = np.isnan(a), np.isnan(b)
    return int(np.count_nonzero(((a_bits != b_bits) & ~(na & nb)) | (na != nb)))


def f32_mismatches(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return mismatches(a.view(np.uint32), b.view(np.uint32), "fp32")


def quant_cases(g):
    """[(quantizer, input key, bits, layerwise)] present in a golden file."""
    out = []
    for k in g.files:
        if k.startswith("y/"):
            _, q, key, b, mode = k.split("/")
            out.append((q, key, int(b[1:]), mode == "lw"))
    return out


def clip_cases(g):
    out = []
    for k in g.files:
        if k.startswith("gx/"):
            _, q, key, mode, clip = k.split("/")
            lo, hi = clip[4:].split("_")
            out.append((q, key, mode == "lw", float(lo), float(hi), k))
    return out
