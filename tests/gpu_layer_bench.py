#!/usr/bin/env python
"""GPU-box tool: BASELINE configs 3 and 4 on ONE GPU, product vs the reference's
eager GPU path (oracle/ref_module.py), same harness.  Writes
gpurun_out/layer_bench.json.  (Test/bench infrastructure: may import oracle/.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from harness import llama_qat as H  # noqa: E402
from harness import qat_bench as B  # noqa: E402
from oracle import ref_module as R  # noqa: E402


def main():
    import llm_qat_b200

    layers = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
    cfg = H.QatConfig.llama_7b(w_bits=4, a_bits=8, kv_bits=4)
    out = {"config3_layer": {}, "config4_step": {}, "config3_layer_autocast": {}, "config4_step_autocast": {}}
    if "--autocast-only" in sys.argv:
        plain = ()
    else:
        plain = (("reference_eager_gpu", R, None), ("b200_unfused", llm_qat_b200.utils_quant, "0"),
                 ("b200_fused", llm_qat_b200.utils_quant, "1"))
    for name, quant, env in plain:
        if env is not None:
            os.environ["QAT_B200_FUSED_LINEAR"] = env
        out["config3_layer"][name] = B.time_layer(quant, cfg)
        print("config3", name, out["config3_layer"][name], flush=True)
    cfg4 = H.QatConfig.llama_7b(w_bits=4, a_bits=8, kv_bits=4, num_hidden_layers=layers)
    for name, quant, env in plain:
        if env is not None:
            os.environ["QAT_B200_FUSED_LINEAR"] = env
        torch.cuda.reset_peak_memory_stats()
        out["config4_step"][name] = B.time_qat_step(quant, cfg4)
        print("config4", name, out["config4_step"][name], flush=True)
    # the recipe's context: HF's Trainer runs the step inside torch.autocast(bf16) (kd_trainer.py:106)
    only = [a[7:] for a in sys.argv if a.startswith("--only=")]
    for name, quant, env, fm in (("reference_eager_gpu", R, None, False), ("b200_fused", llm_qat_b200.utils_quant, "1", False),
                                 ("b200_fused_model", llm_qat_b200.utils_quant, "1", True)):
        if only and name not in only:
            continue
        if env is not None:
            os.environ["QAT_B200_FUSED_LINEAR"] = env
        out["config3_layer_autocast"][name] = B.time_layer(quant, cfg, autocast=True, fused=fm)
        print("config3 autocast", name, out["config3_layer_autocast"][name], flush=True)
        torch.cuda.reset_peak_memory_stats()
        out["config4_step_autocast"][name] = B.time_qat_step(quant, cfg4, autocast=True, fused=fm)
        print("config4 autocast", name, out["config4_step_autocast"][name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "layer_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
