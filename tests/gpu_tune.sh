mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|fastdiv|FAILED|Error" > gpurun_out/tests.log; cat gpurun_out/tests.log
for it in 8 4 2; do
  QAT_B200_MAX_ITERS=$it timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_it$it.log 2>gpurun_out/bench_it$it.err
  python - $it <<'PY'
import json,sys
it=sys.argv[1]
l=json.loads(open(f'gpurun_out/bench_it{it}.log').read().strip().splitlines()[-1])
print('MAX_ITERS',it,'value',l['value'],'ms',l['ms_per_step'], {k:(v['us'],v['frac']) for k,v in l['roofline']['all_kernels'].items()}, 'fp32', {k:v['GBps'] for k,v in l['config1_fp32'].items()}, 'qlinear fwd ms', l['qlinear']['forward']['ms'])
PY
done
