#!/usr/bin/env python
"""GPU-box tool: HBM throughput of torch's own kernels by read:write mix — the ceiling that
write-dominated kernels (dequant 1:2, the autocast variant's fp32 y 1:2) should be judged by."""
import torch
def t(fn, nbytes, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return nbytes / (e0.elapsed_time(e1) / reps) / 1e6
n = 1 << 29
a16 = torch.randn(n, device="cuda", dtype=torch.bfloat16); b16 = torch.empty_like(a16)
a32 = torch.empty(n, device="cuda", dtype=torch.float32); c8 = torch.randint(-7, 8, (n,), device="cuda", dtype=torch.int8)
print(f"write only   (fill_ fp32, {n*4>>20} MiB):            {t(lambda: a32.fill_(1.0), n*4):8.0f} GB/s")
print(f"read only    (sum bf16):                          {t(lambda: a16.sum(), n*2):8.0f} GB/s")
print(f"1:1 copy     (bf16 -> bf16):                      {t(lambda: b16.copy_(a16), n*4):8.0f} GB/s")
print(f"1:2 convert  (bf16 -> fp32 copy_):                {t(lambda: a32.copy_(a16), n*6):8.0f} GB/s")
print(f"1:2 convert  (int8 -> bf16 copy_):                {t(lambda: b16.copy_(c8), n*3):8.0f} GB/s")
print(f"2:1          (bf16 add into bf16: 2 reads 1 write): {t(lambda: torch.add(a16, b16, out=b16), n*6):8.0f} GB/s")
