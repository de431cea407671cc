#!/usr/bin/env python
"""Generate tests/golden/layer_bf16.npz from the LIVE reference model file.

Build container only (``/root/reference`` does not exist on the GPU box):  python oracle/gen_golden_layer.py

Runs the UNMODIFIED /root/reference/models/modeling_llama_quant.py::LlamaDecoderLayer (on the reference's own
models/utils_quant.py) on the CPU in bfloat16 — W4A8KV8, hidden 256, 2 heads of 128, seq 64, the causal mask
its LlamaModel builds (_make_causal_mask, :60-92) — forward and backward, and stores inputs, the state dict,
the output and every gradient as raw bit patterns.  The GPU suite loads the same weights into the harness
layer on the product with llm_qat_b200.fuse_model (tcgen05 attention, K/V-quant + RoPE, RMSNorm / SwiGLU
producers, own backward GEMMs) and compares: that pins SURVEY.md 8(f)-1/3/4 to the reference model file itself,
not to a restatement.  Test infrastructure only.
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


def bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.int16).numpy().copy()


def main() -> int:
    from models.configuration_llama import LlamaConfig
    from models import modeling_llama_quant as M

    assert M.__file__.startswith(REF)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=688, num_attention_heads=2, num_hidden_layers=1, vocab_size=128,
                      max_position_embeddings=128, w_bits=4, a_bits=8, kv_bits=8)
    cfg.kv_bits = 8
    torch.manual_seed(0)
    layer = M.LlamaDecoderLayer(cfg)
    with torch.no_grad():
        for n, p in layer.named_parameters():
            if p.dim() == 2:
                p.normal_(0.0, 0.05)
            else:
                p.uniform_(0.5, 1.5)
    layer = layer.bfloat16()
    g = torch.Generator().manual_seed(7)
    b, s = 2, 64
    x = torch.randn(b, s, 256, generator=g).bfloat16().requires_grad_(True)
    go = torch.randn(b, s, 256, generator=g).bfloat16()
    mask = M._make_causal_mask(torch.Size((b, s)), torch.bfloat16, device=torch.device("cpu"))
    pos = torch.arange(s)[None].expand(b, s)
    y = layer(x, attention_mask=mask, position_ids=pos)[0]
    y.backward(go)
    store = {"x": bits(x), "go": bits(go), "y": bits(y), "gx": bits(x.grad), "meta": np.array([b, s, 256, 688, 2, 4, 8, 8])}
    for n, p in layer.named_parameters():
        store["w/" + n] = bits(p)
        store["g/" + n] = bits(p.grad)
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "layer_bf16.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, |y| = {float(y.float().norm()):.3f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
