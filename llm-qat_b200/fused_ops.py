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
from .utils_quant import _on     # device guard that is free when the device is already current


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
        with _on(q.device):
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
        with _on(q.device):
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


class _KDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher):
        if student.shape != teacher.shape or student.dim() != 3:
            raise RuntimeError(f"kd_loss: expected equal [B, S, V] logits, got {tuple(student.shape)} and "
                               f"{tuple(teacher.shape)}")
        B, S, V = student.shape
        dt = _DT[student.dtype]
        s = student.detach()
        s = s if s.is_contiguous() else s.contiguous()
        t = teacher.detach().to(student.dtype)
        t = t if t.is_contiguous() else t.contiguous()
        dev = s.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        row_kl = torch.empty(B * S, dtype=torch.float32, device=dev)
        row_stat = torch.empty(B * S * 4, dtype=torch.float32, device=dev)
        with _on(dev):
            check(_lib.lib().qat_kd_loss_fwd(s.data_ptr(), t.data_ptr(), loss.data_ptr(), row_kl.data_ptr(),
                                             row_stat.data_ptr(), B * S, V, B, dt, _stream(dev)), "qat_kd_loss_fwd")
        ctx.save_for_backward(s, t, row_stat)
        ctx.dims = (B, S, V, dt)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        s, t, row_stat = ctx.saved_tensors
        B, S, V, dt = ctx.dims
        g = grad_loss.detach().to(device=s.device, dtype=torch.float32).contiguous()
        gs = torch.empty_like(s)
        with _on(s.device):
            check(_lib.lib().qat_kd_loss_bwd(s.data_ptr(), t.data_ptr(), row_stat.data_ptr(), g.data_ptr(),
                                             gs.data_ptr(), B * S, V, B, dt, _stream(s.device)), "qat_kd_loss_bwd")
        return gs, None


_DT = {torch.float32: _lib.QAT_F32, torch.bfloat16: _lib.QAT_BF16}


def kd_loss(student_logits: torch.Tensor, teacher_logits: torch.Tensor) -> torch.Tensor:
    """KL(batchmean) of log_softmax(student) against softmax(teacher) over dim 2 — kd_trainer.py:42-48
    (``KDTrainer.ce_loss``) — as one kernel per direction.  fp32 scalar loss; gradient flows to the student
    logits only (the teacher runs under no_grad, kd_trainer.py:55-60)."""
    if not (student_logits.is_cuda and teacher_logits.is_cuda):
        raise RuntimeError("kd_loss: CUDA tensors required; llm-qat_b200 has no CPU fallback")
    if student_logits.dtype not in _DT:
        raise TypeError(f"kd_loss takes float32 or bfloat16 logits, got {student_logits.dtype}")
    return _KDLoss.apply(student_logits, teacher_logits)


# ------------------------------------------------------------------ producers fused with the feed
def _amp_dtype(t: torch.Tensor) -> int:
    """QAT_BF16_AMP inside torch.autocast (the recipe's context), else QAT_BF16 — the same rule
    QuantizeLinear applies to its input (utils_quant._sym_amp)."""
    return _lib.QAT_BF16_AMP if (t.dtype == torch.bfloat16 and torch.is_autocast_enabled("cuda")) else _lib.QAT_BF16


def _feed_ptrs(blob, rows, cols):
    from .utils_quant import _feed_layout

    off_e, off_m, _ = _feed_layout(rows, cols)
    b = blob.data_ptr()
    return b, b + off_e, b + off_m


def _new_blob(rows, cols, dev):
    from .utils_quant import _feed_layout

    return torch.empty(_feed_layout(rows, cols)[2], dtype=torch.uint8, device=dev)


class _RMSNormFeed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps, bits, dt):
        C = x.shape[-1]
        xc = x.detach()
        xc = xc if xc.is_contiguous() else xc.contiguous()
        w = weight.detach()
        T = xc.numel() // C
        dev = xc.device
        y = torch.empty_like(xc)
        rstd = torch.empty(T, dtype=torch.float32, device=dev)
        blob = _new_blob(T, C, dev) if bits else torch.empty(0, dtype=torch.uint8, device=dev)
        c, e, m = _feed_ptrs(blob, T, C) if bits else (0, 0, 0)
        with _on(dev):
            check(_lib.lib().qat_rmsnorm_feed_fwd(xc.data_ptr(), w.data_ptr(), y.data_ptr(), rstd.data_ptr(), c, e, m,
                                                  -2.0, 2.0, T, C, float(eps), dt, int(bits) if bits else 8,
                                                  _stream(dev)), "qat_rmsnorm_feed_fwd")
        ctx.save_for_backward(xc, w, rstd)
        ctx.mark_non_differentiable(blob)
        return y, blob

    @staticmethod
    def backward(ctx, gy, _gblob):
        x, w, rstd = ctx.saved_tensors
        C = x.shape[-1]
        T = x.numel() // C
        g = gy.to(torch.bfloat16)
        g = g if g.is_contiguous() else g.contiguous()
        gx = torch.empty_like(x)
        gw = torch.empty_like(w)
        L = _lib.lib()
        nb = int(L.qat_rmsnorm_bwd_workspace_bytes(T, C))
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        with _on(x.device):
            check(L.qat_rmsnorm_bwd(g.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), gx.data_ptr(),
                                    gw.data_ptr(), ws.data_ptr(), nb, T, C, _stream(x.device)), "qat_rmsnorm_bwd")
        return gx, gw, None, None, None


def _register_feed(y: torch.Tensor, blob: torch.Tensor, bits: int, dt: int) -> None:
    """Hand the codes of ``y`` to the QuantizeLinear(s) that will consume this very tensor object."""
    from . import utils_quant as uq

    if uq._cache_mode() == 0:
        return
    K = y.shape[-1]
    uq.register_activation_feed(y, blob, y.numel() // K, K, dt, bits)


def rmsnorm_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.is_cuda
            and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192 and x.numel() > 0 and x.data_ptr() % 16 == 0)


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, feed_bits: int = 0) -> torch.Tensor:
    """LlamaRMSNorm.forward (modeling_llama_quant.py:121-129) in one kernel; with ``feed_bits`` the same
    launch also emits the int8 codes / divisors / STE mask of the result, which the QuantizeLinear layers
    fed by this tensor pick up instead of quantizing it again."""
    if not rmsnorm_supported(x, weight):
        raise RuntimeError("rmsnorm: needs CUDA bfloat16 tensors with hidden % 8 == 0 and <= 8192")
    dt = _amp_dtype(x)
    y, blob = _RMSNormFeed.apply(x, weight, eps, int(feed_bits), dt)
    if feed_bits:
        _register_feed(y, blob, int(feed_bits), dt)
    return y


class _SwiGLUFeed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up, bits, dt):
        C = gate.shape[-1]
        g = gate.detach()
        g = g if g.is_contiguous() else g.contiguous()
        u = up.detach()
        u = u if u.is_contiguous() else u.contiguous()
        T = g.numel() // C
        dev = g.device
        act = torch.empty_like(g)
        blob = _new_blob(T, C, dev) if bits else torch.empty(0, dtype=torch.uint8, device=dev)
        c, e, m = _feed_ptrs(blob, T, C) if bits else (0, 0, 0)
        with _on(dev):
            check(_lib.lib().qat_swiglu_feed_fwd(g.data_ptr(), u.data_ptr(), act.data_ptr(), c, e, m, -2.0, 2.0, T, C,
                                                 dt, int(bits) if bits else 8, _stream(dev)), "qat_swiglu_feed_fwd")
        ctx.save_for_backward(g, u)
        ctx.mark_non_differentiable(blob)
        return act, blob

    @staticmethod
    def backward(ctx, gact, _gblob):
        g, u = ctx.saved_tensors
        ga = gact.to(torch.bfloat16)
        ga = ga if ga.is_contiguous() else ga.contiguous()
        dg, du = torch.empty_like(g), torch.empty_like(u)
        with _on(g.device):
            check(_lib.lib().qat_swiglu_bwd(ga.data_ptr(), g.data_ptr(), u.data_ptr(), dg.data_ptr(), du.data_ptr(),
                                            g.numel(), _stream(g.device)), "qat_swiglu_bwd")
        return dg, du, None, None


def swiglu_supported(gate: torch.Tensor, up: torch.Tensor) -> bool:
    return (gate.is_cuda and gate.dtype == torch.bfloat16 and up.dtype == torch.bfloat16 and gate.shape == up.shape
            and gate.shape[-1] % 8 == 0 and gate.shape[-1] <= 16384 and gate.numel() > 0)


def swiglu(gate: torch.Tensor, up: torch.Tensor, feed_bits: int = 0) -> torch.Tensor:
    """``silu(gate) * up`` (modeling_llama_quant.py:235) in one kernel, optionally with the feed of the
    result for down_proj."""
    if not swiglu_supported(gate, up):
        raise RuntimeError("swiglu: needs equal-shape CUDA bfloat16 tensors with cols % 8 == 0 and <= 16384")
    dt = _amp_dtype(gate)
    act, blob = _SwiGLUFeed.apply(gate, up, int(feed_bits), dt)
    if feed_bits:
        _register_feed(act, blob, int(feed_bits), dt)
    return act


class _QKVPrep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cos, sin, pos, heads, kv_bits, lo, hi, dt):
        q, k, v = (t.detach() if t.is_contiguous() else t.detach().contiguous() for t in (q, k, v))
        hidden = q.shape[-1]
        T = q.numel() // hidden
        dev = q.device
        qo, ko, vo = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        quant = kv_bits < 32
        km = torch.empty(T * hidden // 8 if quant else 0, dtype=torch.uint8, device=dev)
        vm = torch.empty_like(km)
        pos = pos.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        with _on(dev):
            check(_lib.lib().qat_qkv_prep_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), qo.data_ptr(), ko.data_ptr(),
                                              vo.data_ptr(), km.data_ptr() if quant else 0,
                                              vm.data_ptr() if quant else 0, cos.data_ptr(), sin.data_ptr(),
                                              pos.data_ptr(), cos.shape[0], T, heads, hidden // heads, int(kv_bits),
                                              lo, hi, dt, _stream(dev)), "qat_qkv_prep_fwd")
        ctx.save_for_backward(km, vm, cos, sin, pos)
        ctx.meta = (T, heads, hidden // heads, quant)
        return qo, ko, vo

    @staticmethod
    def backward(ctx, dq_rot, dk_rot, dv_q):
        km, vm, cos, sin, pos = ctx.saved_tensors
        T, heads, hd, quant = ctx.meta
        a, b, c = (t.to(torch.bfloat16) for t in (dq_rot, dk_rot, dv_q))
        a, b, c = (t if t.is_contiguous() else t.contiguous() for t in (a, b, c))
        dq, dk, dv = torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)
        with _on(a.device):
            check(_lib.lib().qat_qkv_prep_bwd(a.data_ptr(), b.data_ptr(), c.data_ptr(), km.data_ptr() if quant else 0,
                                              vm.data_ptr() if quant else 0, cos.data_ptr(), sin.data_ptr(),
                                              pos.data_ptr(), cos.shape[0], dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                              T, heads, hd, _stream(a.device)), "qat_qkv_prep_bwd")
        return dq, dk, dv, None, None, None, None, None, None, None, None


def qkv_prep(q, k, v, cos_table, sin_table, position_ids, heads: int, kv_bits: int, clip=(-2.0, 2.0)):
    """K/V per-token fake-quant (SymQuantizer.apply(..., kv_bits, False), modeling_llama_quant.py:320-327)
    + rotary embedding of q and k (:334-341) in one launch.  q, k, v: bf16 [B, S, heads * 128];
    cos_table / sin_table: fp32 [max_pos, 128]; returns (q_rot, k_rot, v_q) bf16 in the same layout."""
    if not (q.is_cuda and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16):
        raise RuntimeError("qkv_prep: CUDA bfloat16 tensors required")
    if q.shape[-1] != heads * 128:
        raise RuntimeError("qkv_prep: head_dim must be 128")
    if k.shape != q.shape or v.shape != q.shape:
        raise RuntimeError(f"qkv_prep: q, k, v must have one shape, got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
    for name, tab in (("cos_table", cos_table), ("sin_table", sin_table)):
        if not (tab.is_cuda and tab.dtype == torch.float32 and tab.dim() == 2 and tab.shape[1] == 128 and tab.is_contiguous()):
            raise RuntimeError(f"qkv_prep: {name} must be a contiguous CUDA float32 [max_pos, 128] tensor")
    if cos_table.shape != sin_table.shape or cos_table.shape[0] == 0:
        raise RuntimeError("qkv_prep: cos_table and sin_table must have one non-empty shape")
    if position_ids.numel() != q.numel() // q.shape[-1]:
        raise RuntimeError("qkv_prep: one position id per token expected")
    dt = _amp_dtype(k)
    return _QKVPrep.apply(q, k, v, cos_table, sin_table, position_ids, int(heads), int(kv_bits), float(clip[0]),
                          float(clip[1]), dt)
