"""A minimal LLaMA decoder stack whose Linear layers and K/V fake-quant come from
an injected quantization module — ``llm_qat_b200.utils_quant`` (the product) or
``oracle.ref_module`` (the reference's eager path) — so that both run under the
same caller.  Restates the structure of the reference model file
(/root/reference/models/modeling_llama_quant.py): RMSNorm :112-129, rotary
:132-196, MLP :199-235, eager attention with fp32 softmax and pre-RoPE per-token
K/V fake-quant :238-393, decoder layer :396-467, model with per-layer gradient
checkpointing :724-747, untied lm_head :793; and the KD step of
/root/reference/utils/kd_trainer.py:42-81 (KL batchmean of log_softmax(student)
vs softmax(teacher)).  /root/reference cannot travel to the GPU box, hence this
harness; tests/test_harness.py checks it against the real reference layer here.
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint


@dataclass
class QatConfig:
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_attention_heads: int = 32
    num_hidden_layers: int = 32
    vocab_size: int = 32000
    max_position_embeddings: int = 2048
    rms_norm_eps: float = 1e-6
    initializer_range: float = 0.02
    w_bits: int = 4
    a_bits: int = 8
    kv_bits: int = 4

    @staticmethod
    def llama_7b(**kw):
        return QatConfig(**kw)

    @staticmethod
    def llama_13b(**kw):
        return QatConfig(hidden_size=5120, intermediate_size=13824, num_attention_heads=40,
                         num_hidden_layers=40, **kw)

    @staticmethod
    def tiny(**kw):
        base = dict(hidden_size=64, intermediate_size=176, num_attention_heads=4, num_hidden_layers=2,
                    vocab_size=128, max_position_embeddings=64)
        base.update(kw)
        return QatConfig(**base)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, h):
        var = h.to(torch.float32).pow(2).mean(-1, keepdim=True)
        h = h * torch.rsqrt(var + self.eps)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            h = h.to(self.weight.dtype)
        return self.weight * h


def _rope_tables(head_dim, n_pos, device, base=10000.0):
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2, device=device).float() / head_dim))
    freqs = torch.outer(torch.arange(n_pos, device=device, dtype=inv.dtype), inv)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


class Attention(nn.Module):
    def __init__(self, cfg: QatConfig, quant):
        super().__init__()
        H = cfg.hidden_size
        self.n_heads, self.head_dim, self.kv_bits = cfg.num_attention_heads, H // cfg.num_attention_heads, cfg.kv_bits
        mk = lambda: quant.QuantizeLinear(H, H, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)  # noqa: E731
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = mk(), mk(), mk(), mk()
        self.kv_quantizer = quant.SymQuantizer
        self.clip_k = torch.tensor([-2.0, 2.0])
        self.clip_v = torch.tensor([-2.0, 2.0])
        self.max_pos = cfg.max_position_embeddings

    def forward(self, h, mask, position_ids):
        b, s, H = h.shape
        q = self.q_proj(h).view(b, s, self.n_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(h)
        v = self.v_proj(h)
        if self.kv_bits < 32:   # per-token over all heads' channels, before the head split and RoPE
            k = self.kv_quantizer.apply(k, self.clip_k, self.kv_bits, False)
            v = self.kv_quantizer.apply(v, self.clip_v, self.kv_bits, False)
        k = k.view(b, s, self.n_heads, self.head_dim).transpose(1, 2)
        v = v.view(b, s, self.n_heads, self.head_dim).transpose(1, 2)
        cos, sin = _rope_tables(self.head_dim, max(self.max_pos, s), h.device)
        # rotary_emb(value_states, ...) returns tables in V's dtype (:334) — fp32 under autocast, where
        # the K/V fake-quant returns float32
        cos = cos.to(v.dtype)[position_ids].unsqueeze(1)
        sin = sin.to(v.dtype)[position_ids].unsqueeze(1)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(self.head_dim)
        if mask is not None:
            w = w + mask
            w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, device=w.device))
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(w, v).transpose(1, 2).reshape(b, s, H)
        return self.o_proj(o)


class MLP(nn.Module):
    def __init__(self, cfg: QatConfig, quant):
        super().__init__()
        H, I = cfg.hidden_size, cfg.intermediate_size
        self.gate_proj = quant.QuantizeLinear(H, I, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)
        self.down_proj = quant.QuantizeLinear(I, H, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)
        self.up_proj = quant.QuantizeLinear(H, I, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class DecoderLayer(nn.Module):
    def __init__(self, cfg: QatConfig, quant):
        super().__init__()
        self.self_attn = Attention(cfg, quant)
        self.mlp = MLP(cfg, quant)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, h, mask=None, position_ids=None):
        h = h + self.self_attn(self.input_layernorm(h), mask, position_ids)
        return h + self.mlp(self.post_attention_layernorm(h))


def causal_mask(b, s, dtype, device):
    m = torch.full((s, s), torch.finfo(dtype).min, device=device, dtype=dtype)
    m = torch.triu(m, diagonal=1)
    return m[None, None].expand(b, 1, s, s)


class CausalLM(nn.Module):
    """embed -> N decoder layers (checkpointed when training) -> norm -> lm_head;
    embed and lm_head are plain (unquantized), as in the reference (:581-583, :793)."""

    def __init__(self, cfg: QatConfig, quant, gradient_checkpointing=True):
        super().__init__()
        self.cfg = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([DecoderLayer(cfg, quant) for _ in range(cfg.num_hidden_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self.gradient_checkpointing = gradient_checkpointing
        self.apply(self._init)

    def _init(self, m):
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, self.cfg.initializer_range)
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(0.0, self.cfg.initializer_range)

    def forward(self, input_ids):
        b, s = input_ids.shape
        h = self.embed_tokens(input_ids)
        mask = causal_mask(b, s, h.dtype, h.device)
        pos = torch.arange(s, device=h.device)[None].expand(b, s)
        for layer in self.layers:
            if self.gradient_checkpointing and self.training and torch.is_grad_enabled():
                h = checkpoint(layer, h, mask, pos, use_reentrant=False)
            else:
                h = layer(h, mask, pos)
        return self.lm_head(self.norm(h))


class _PlainQuant:
    """Unquantized 'quant module' for the FP teacher (w_bits = a_bits = 32 never
    touches a quantizer; kv_bits = 32 skips the K/V call)."""

    class SymQuantizer:  # never applied
        @staticmethod
        def apply(x, *a):
            return x

    @staticmethod
    def QuantizeLinear(i, o, bias=False, w_bits=32, a_bits=32):
        return nn.Linear(i, o, bias=False)


def build_teacher(cfg: QatConfig):
    fp = QatConfig(**{**cfg.__dict__, "w_bits": 32, "a_bits": 32, "kv_bits": 32})
    t = CausalLM(fp, _PlainQuant, gradient_checkpointing=False)
    for p in t.parameters():
        p.requires_grad_(False)
    return t.eval()


def kd_loss(student_logits, teacher_logits):
    """kd_trainer.py:42-48: KL(batchmean) of log_softmax(student) vs softmax(teacher) over the vocab dim."""
    return F.kl_div(F.log_softmax(student_logits, dim=2), F.softmax(teacher_logits, dim=2), reduction="batchmean")


def qat_step(student, teacher, input_ids, optimizer, kd_loss_scale=1.0, autocast=False):
    """One training step of compute_loss_train + backward + optimizer (kd_trainer.py:53-127).
    ``autocast=True`` mirrors the recipe: HF's Trainer (run_train.sh `--bf16 True`) computes the
    loss — teacher and student forward — inside torch.autocast(bfloat16) (kd_trainer.py:106)."""
    ctx = torch.autocast(input_ids.device.type, dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with ctx:
        with torch.no_grad():
            t_logits = teacher(input_ids)
        s_logits = student(input_ids)
        loss = kd_loss_scale * kd_loss(s_logits, t_logits)
    del t_logits, s_logits
    loss.backward()
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss.detach()
