import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from llm_qat_b200._lib import CODES_I8, CODES_I16
from llm_qat_b200.utils_quant import fake_quant_forward, qlinear_i8, QuantizeLinear
torch.manual_seed(0)
for (T, K, N) in [(256, 512, 512), (256, 512, 1376), (256, 1376, 512), (256, 1280, 512), (256, 1408, 512), (200, 1040, 264), (256, 1376, 256), (128, 1376, 256)]:
    for scale in (1.0, 30.0):
        x = (torch.randn(T, K) * scale).bfloat16().cuda()
        w = (torch.randn(N, K) * 0.5).bfloat16().cuda()
        _, qx, _, ex, mx = fake_quant_forward(x, 8, False, True, want_y=False, codes_kind=CODES_I8, want_scales=True, mask_clip=(-2.0, 2.0))
        _, qw, _, ew, mw = fake_quant_forward(w, 4, False, True, want_y=False, codes_kind=CODES_I8, want_scales=True, mask_clip=(-2.0, 2.0))
        _, qx16, _, ex2, _ = fake_quant_forward(x, 8, False, True, want_y=False, codes_kind=CODES_I16, want_scales=True)
        codes_ok = bool((qx.short() == qx16.clamp(-127, 127)).all()) and bool((ex == ex2).all())
        out = qlinear_i8(qx, qw, ex, ew, torch.float32)
        ref = (qx.double() @ qw.double().t()) / (ex.double()[:, None] * ew.double()[None, :])
        rel = ((out.double() - ref).norm() / ref.norm()).item()
        bad_rows = ((out.double() - ref).abs().amax(dim=1) > 1e-3 * ref.abs().amax()).nonzero().flatten()[:8].tolist()
        bad_cols = ((out.double() - ref).abs().amax(dim=0) > 1e-3 * ref.abs().amax()).nonzero().flatten()[:8].tolist()
        print(f"T={T} K={K} N={N} scale={scale}: feed codes ok={codes_ok} gemm rel={rel:.2e} bad_rows={bad_rows} bad_cols={bad_cols}", flush=True)
