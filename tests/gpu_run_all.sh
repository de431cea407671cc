#!/bin/bash
# GPU-box driver used during development: parity tests, eager comparison, bench,
# then the ncu passes (launch list + one full capture of the dominant kernels).
mkdir -p gpurun_out; nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/tests.log
if [ "$1" != "quick" ]; then
  timeout 300 python tests/gpu_eager_compare.py > gpurun_out/eager.log 2>&1; echo "eager rc=$?"
  timeout 900 python tests/gpu_layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer bench rc=$?"; tail -8 gpurun_out/layer_bench.log
fi
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-qat-step"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  $CMD > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'rowquant_vec_kernel|ste_bwd_kernel|qlinear_i8_kernel' -s 16 -c 9 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  # K4 alone at config 2 (CTA-pair plan), same call: one more capture
  python tests/gpu_gemm_only.py 5 2 > gpurun_out/gemm_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:qlinear_i8_kernel -s 4 -c 1 -f -o gpurun_out/prof_gemm python tests/gpu_gemm_only.py 5 2 > gpurun_out/ncu_gemm.log 2>&1
  echo "ncu gemm rc=$?"
fi
