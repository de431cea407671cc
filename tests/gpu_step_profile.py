#!/usr/bin/env python
"""GPU-box tool: where does the full LLaMA-7B QAT step (BASELINE configs[3], fuse_model on) spend its
device time?  torch.profiler kernel table of 2 steps, grouped by kernel name."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import llm_qat_b200  # noqa: E402
from harness import llama_qat as H  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = H.QatConfig.llama_7b(w_bits=4, a_bits=8, kv_bits=4, num_hidden_layers=layers)
torch.manual_seed(0)
with torch.device("cuda"):
    student = H.CausalLM(cfg, llm_qat_b200.utils_quant, fused=True).bfloat16()
    teacher = H.build_teacher(cfg, fused=True).bfloat16()
teacher.load_state_dict(student.state_dict())
student.train()
opt = torch.optim.AdamW(student.parameters(), lr=2e-5)
ids = torch.randint(0, cfg.vocab_size, (1, 2048)).cuda()


def step():
    H.qat_step(student, teacher, ids, opt, autocast=True, loss_fn=llm_qat_b200.fused_ops.kd_loss)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
e1.synchronize()
print(f"step {e0.elapsed_time(e1)/3:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ka = [e for e in prof.key_averages() if getattr(e, "self_device_time_total", 0) > 0]
tot = sum(e.self_device_time_total for e in ka)
ours = sum(e.self_device_time_total for e in ka if "qat::" in e.key)
print(f"device kernel time per step {tot/2/1e3:.2f} ms, libqat_b200 {ours/2/1e3:.2f} ms, {sum(e.count for e in ka)//2} launches")
for e in sorted(ka, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"{e.self_device_time_total/2/1e3:9.3f} ms  x{e.count//2:5d}  {e.key[:120]}")
