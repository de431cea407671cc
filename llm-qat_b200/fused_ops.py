"""Kernels for the callers either side of the fake-quant path, each an
autograd Function over the C ABI of libqat_b200.so:

* ``causal_attention(q, k, v)``     — reference models/modeling_llama_quant.py:352-377 (eager
  QK^T / sqrt(d) + causal mask + fp32 softmax + PV) and its backward, fused on tcgen05;
* ``kd_loss(student, teacher)``     — reference utils/kd_trainer.py:42-48 (KL batchmean of
  log_softmax(student) against softmax(teacher) over the vocabulary), one pass per direction.

CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import check


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _CausalAttention(torch.autograd.Function):
    """q, k, v: bf16 [B, S, H, 128] (contiguous).  Returns bf16 [B, S, H, 128]."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal):
        B, S, H, D = q.shape
        q, k, v = (t if t.is_contiguous() else t.contiguous() for t in (q.detach(), k.detach(), v.detach()))
        o = torch.empty_like(q)
        lse = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            check(_lib.lib().qat_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                          B, S, H, D, float(scale), int(causal), _stream(q.device)), "qat_attn_fwd")
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale, ctx.causal = float(scale), int(causal)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse = ctx.saved_tensors
        B, S, H, D = q.shape
        d_o = d_o.to(torch.bfloat16)
        if not d_o.is_contiguous():
            d_o = d_o.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty_like(lse)
        with torch.cuda.device(q.device):
            check(_lib.lib().qat_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr(),
                                          lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                          dv.data_ptr(), B, S, H, D, ctx.scale, ctx.causal, _stream(q.device)),
                  "qat_attn_bwd")
        return dq, dk, dv, None, None


def attention_supported(q: torch.Tensor) -> bool:
    return q.is_cuda and q.dtype == torch.bfloat16 and q.dim() == 4 and q.shape[-1] == 128


def causal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float | None = None,
                     causal: bool = True) -> torch.Tensor:
    """softmax(q k^T * scale + causal mask) v for bf16 [B, S, H, 128] tensors, fp32 softmax,
    P rounded to bf16 before P.v — modeling_llama_quant.py:352-377 without the [B, H, S, S] tensors."""
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise RuntimeError("causal_attention: CUDA tensors required; llm-qat_b200 has no CPU fallback")
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise TypeError("causal_attention takes bfloat16 tensors")
    if q.dim() != 4 or q.shape[-1] != 128 or k.shape != q.shape or v.shape != q.shape:
        raise RuntimeError(f"causal_attention: expected equal [B, S, H, 128] shapes, got {tuple(q.shape)}, "
                           f"{tuple(k.shape)}, {tuple(v.shape)}")
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    return _CausalAttention.apply(q, k, v, scale, causal)
