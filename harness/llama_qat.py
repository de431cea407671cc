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
    """LlamaRMSNorm (:112-129), same attribute names."""

    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.variance_epsilon = eps

    def forward(self, h):
        var = h.to(torch.float32).pow(2).mean(-1, keepdim=True)
        h = h * torch.rsqrt(var + self.variance_epsilon)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            h = h.to(self.weight.dtype)
        return self.weight * h


class RotaryEmbedding(nn.Module):
    """LlamaRotaryEmbedding (:132-171): fp32 cos / sin caches [1, 1, max_pos, dim], returned in x's dtype."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, device=None):
        super().__init__()
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float().to(device) / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        t = torch.arange(max_position_embeddings, device=inv_freq.device, dtype=inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        self.register_buffer("cos_cached", emb.cos()[None, None, :, :], persistent=False)
        self.register_buffer("sin_cached", emb.sin()[None, None, :, :], persistent=False)

    def forward(self, x, seq_len=None):
        return (self.cos_cached[:, :, :seq_len, ...].to(dtype=x.dtype),
                self.sin_cached[:, :, :seq_len, ...].to(dtype=x.dtype))


def _rotate_half(x):
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def _apply_rotary_pos_emb(q, k, cos, sin, position_ids):   # :181-196
    cos = cos.squeeze(1).squeeze(0)[position_ids].unsqueeze(1)
    sin = sin.squeeze(1).squeeze(0)[position_ids].unsqueeze(1)
    return (q * cos) + (_rotate_half(q) * sin), (k * cos) + (_rotate_half(k) * sin)


class Attention(nn.Module):
    """LlamaAttention (:238-393) with the reference's attribute names and call signature, so that
    llm_qat_b200.fuse_model treats it exactly like the reference's module."""

    def __init__(self, cfg: QatConfig, quant):
        super().__init__()
        H = cfg.hidden_size
        self.hidden_size, self.num_heads = H, cfg.num_attention_heads
        self.head_dim, self.kv_bits = H // cfg.num_attention_heads, cfg.kv_bits
        self.max_position_embeddings = cfg.max_position_embeddings
        mk = lambda: quant.QuantizeLinear(H, H, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)  # noqa: E731
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = mk(), mk(), mk(), mk()
        self.act_quantizer_k = self.act_quantizer_v = quant.SymQuantizer
        self.act_clip_val_k = torch.tensor([-2.0, 2.0])
        self.act_clip_val_v = torch.tensor([-2.0, 2.0])
        self.rotary_emb = RotaryEmbedding(self.head_dim, max_position_embeddings=cfg.max_position_embeddings)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False):
        b, s, H = hidden_states.shape
        q = self.q_proj(hidden_states).view(b, s, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(hidden_states)
        v = self.v_proj(hidden_states)
        if self.kv_bits < 32:   # per-token over all heads' channels, before the head split and RoPE
            k = self.act_quantizer_k.apply(k, self.act_clip_val_k, self.kv_bits, False)
            v = self.act_quantizer_v.apply(v, self.act_clip_val_v, self.kv_bits, False)
        k = k.view(b, s, self.num_heads, self.head_dim).transpose(1, 2)
        v = v.view(b, s, self.num_heads, self.head_dim).transpose(1, 2)
        # rotary_emb(value_states, ...) returns tables in V's dtype (:334) — fp32 under autocast, where
        # the K/V fake-quant returns float32
        cos, sin = self.rotary_emb(v, seq_len=s)
        q, k = _apply_rotary_pos_emb(q, k, cos, sin, position_ids)
        w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(self.head_dim)
        if attention_mask is not None:
            w = w + attention_mask
            w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, device=w.device))
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(w, v).transpose(1, 2).reshape(b, s, H)
        return self.o_proj(o), None, None


class MLP(nn.Module):
    def __init__(self, cfg: QatConfig, quant):
        super().__init__()
        H, I = cfg.hidden_size, cfg.intermediate_size
        self.gate_proj = quant.QuantizeLinear(H, I, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)
        self.down_proj = quant.QuantizeLinear(I, H, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)
        self.up_proj = quant.QuantizeLinear(H, I, bias=False, w_bits=cfg.w_bits, a_bits=cfg.a_bits)
        self.act_fn = nn.SiLU()

    def forward(self, x):
        return self.down_proj(self.act_fn(self.gate_proj(x)) * self.up_proj(x))


class DecoderLayer(nn.Module):
    def __init__(self, cfg: QatConfig, quant):
        super().__init__()
        self.self_attn = Attention(cfg, quant)
        self.mlp = MLP(cfg, quant)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, h, mask=None, position_ids=None):
        a, _, _ = self.self_attn(hidden_states=self.input_layernorm(h), attention_mask=mask, position_ids=position_ids)
        h = h + a
        return h + self.mlp(self.post_attention_layernorm(h))


def causal_mask(b, s, dtype, device):
    m = torch.full((s, s), torch.finfo(dtype).min, device=device, dtype=dtype)
    m = torch.triu(m, diagonal=1)
    return m[None, None].expand(b, 1, s, s)


class CausalLM(nn.Module):
    """embed -> N decoder layers (checkpointed when training) -> norm -> lm_head;
    embed and lm_head are plain (unquantized), as in the reference (:581-583, :793).
    ``fused=True`` applies llm_qat_b200.fuse_model (attention / MLP / RMSNorm kernels, default off)."""

    def __init__(self, cfg: QatConfig, quant, gradient_checkpointing=True, fused=False):
        super().__init__()
        self.cfg = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([DecoderLayer(cfg, quant) for _ in range(cfg.num_hidden_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self.gradient_checkpointing = gradient_checkpointing
        self.apply(self._init)
        self.fused = fused
        if fused:
            import llm_qat_b200

            llm_qat_b200.fuse_model(self)

    def _init(self, m):
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, self.cfg.initializer_range)
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(0.0, self.cfg.initializer_range)

    def forward(self, input_ids):
        b, s = input_ids.shape
        h = self.embed_tokens(input_ids)
        mask = causal_mask(b, s, h.dtype, h.device)
        if self.fused:   # what the patched _prepare_decoder_attention_mask does for the reference's LlamaModel
            import llm_qat_b200

            llm_qat_b200.mark_causal_mask(mask)
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


def build_teacher(cfg: QatConfig, fused=False):
    fp = QatConfig(**{**cfg.__dict__, "w_bits": 32, "a_bits": 32, "kv_bits": 32})
    t = CausalLM(fp, _PlainQuant, gradient_checkpointing=False, fused=fused)
    for p in t.parameters():
        p.requires_grad_(False)
    return t.eval()


def kd_loss(student_logits, teacher_logits):
    """kd_trainer.py:42-48: KL(batchmean) of log_softmax(student) vs softmax(teacher) over the vocab dim."""
    return F.kl_div(F.log_softmax(student_logits, dim=2), F.softmax(teacher_logits, dim=2), reduction="batchmean")


def qat_step(student, teacher, input_ids, optimizer, kd_loss_scale=1.0, autocast=False, loss_fn=None):
    """One training step of compute_loss_train + backward + optimizer (kd_trainer.py:53-127).
    ``autocast=True`` mirrors the recipe: HF's Trainer (run_train.sh `--bf16 True`) computes the
    loss — teacher and student forward — inside torch.autocast(bfloat16) (kd_trainer.py:106)."""
    ctx = torch.autocast(input_ids.device.type, dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with ctx:
        with torch.no_grad():
            t_logits = teacher(input_ids)
        s_logits = student(input_ids)
        loss = kd_loss_scale * (loss_fn or kd_loss)(s_logits, t_logits)
    del t_logits, s_logits
    loss.backward()
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss.detach()
