#!/usr/bin/env python
"""GPU-box tool (torchrun, >= 2 ranks): the reference's own parallel recipe — FSDP full_shard,
auto-wrapped per decoder layer, bf16 (run_train.sh:42-43, kd_trainer.py:172-255) — on top of
this package.  FSDP re-materialises every layer's weights from shards on each forward, so this
is the case where a stale weight-code cache would silently corrupt training: the loss
trajectory with the caches on must equal the trajectory with them off, step for step, and the
unfused path (the reference's structure) must track it."""
import functools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.distributed.fsdp import FullyShardedDataParallel as FSDP, MixedPrecision
from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy

from harness import llama_qat as H
import llm_qat_b200


def run(mode_env, steps=5, orig_params=False, fused_model=False):
    for k, v in mode_env.items():
        os.environ[k] = v
    rank = dist.get_rank()
    # head_dim 128 so that fuse_model's attention kernel is eligible
    cfg = H.QatConfig(hidden_size=256, intermediate_size=688, num_attention_heads=2, num_hidden_layers=3,
                      vocab_size=512, max_position_embeddings=128, w_bits=4, a_bits=8, kv_bits=4)
    torch.manual_seed(0)
    model = H.CausalLM(cfg, llm_qat_b200.utils_quant, fused=fused_model).bfloat16().cuda()
    teacher = H.build_teacher(cfg, fused=fused_model).bfloat16().cuda()
    teacher.load_state_dict(model.state_dict())
    policy = functools.partial(transformer_auto_wrap_policy, transformer_layer_cls={H.DecoderLayer})
    model = FSDP(model, auto_wrap_policy=policy, device_id=torch.cuda.current_device(),
                 mixed_precision=MixedPrecision(param_dtype=torch.bfloat16, reduce_dtype=torch.bfloat16,
                                                buffer_dtype=torch.bfloat16),
                 limit_all_gathers=True, use_orig_params=orig_params)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(100 + rank)
    losses = []
    for step in range(steps):
        ids = torch.randint(0, cfg.vocab_size, (2, 64), generator=g).cuda()
        model.train()
        losses.append(float(H.qat_step(model, teacher, ids, opt, autocast=fused_model,
                                       loss_fn=llm_qat_b200.fused_ops.kd_loss if fused_model else None)))
        if step == 2:   # an eval-style forward between steps: weights unchanged, caches may hit
            model.eval()
            with torch.no_grad():
                model(ids)
    t = torch.tensor(losses, device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    return (t / dist.get_world_size()).tolist()


def main():
    dist.init_process_group("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    ok = True
    for orig in (False, True):
        on = run({"QAT_B200_CACHE": "1", "QAT_B200_FUSED_LINEAR": "1"}, orig_params=orig)
        off = run({"QAT_B200_CACHE": "0", "QAT_B200_FUSED_LINEAR": "1"}, orig_params=orig)
        unf = run({"QAT_B200_CACHE": "0", "QAT_B200_FUSED_LINEAR": "0"}, orig_params=orig)
        if dist.get_rank() == 0:
            same = on == off
            rel = max(abs(a - b) / max(abs(b), 1e-9) for a, b in zip(on, unf))
            print(f"use_orig_params={orig}: caches on  {['%.6f' % v for v in on]}")
            print(f"use_orig_params={orig}: caches off {['%.6f' % v for v in off]}  identical={same}")
            print(f"use_orig_params={orig}: unfused    {['%.6f' % v for v in unf]}  max rel diff vs fused {rel:.3e}")
            ok = ok and same and all(v == v for v in on)
        # llm_qat_b200.fuse_model (attention / MLP / RMSNorm kernels, producer-emitted codes) under the same wrapper
        f_on = run({"QAT_B200_CACHE": "1", "QAT_B200_FUSED_LINEAR": "1"}, orig_params=orig, fused_model=True)
        f_off = run({"QAT_B200_CACHE": "0", "QAT_B200_FUSED_LINEAR": "1"}, orig_params=orig, fused_model=True)
        if dist.get_rank() == 0:
            same = f_on == f_off
            rel = max(abs(a - b) / max(abs(b), 1e-9) for a, b in zip(f_on, on))
            print(f"use_orig_params={orig}: fuse_model, caches on  {['%.6f' % v for v in f_on]}")
            print(f"use_orig_params={orig}: fuse_model, caches off {['%.6f' % v for v in f_off]}  identical={same}  "
                  f"max rel diff vs quant-path-only {rel:.3e}")
            ok = ok and same and all(v == v for v in f_on)   # (this 5-step run is chaotic: even unfused vs fused differ by 10 %)
    if dist.get_rank() == 0:
        print("FSDP CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
