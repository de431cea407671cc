#!/bin/bash
# GPU-box tool: A/B of programmatic dependent launch on the bench step and the layer.
for pdl in 1 0; do
  QAT_B200_PDL=$pdl python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-qat-step 2>gpurun_out/ab_pdl$pdl.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PDL', $pdl, 'value', d['value'], 'graph_step_us', d['roofline']['graph_step_us'], 'launch', d['config']['launch'], {k:(v['us'],v['frac']) for k,v in d['roofline']['all_kernels'].items()}, 'qlinear fwd ms', d['qlinear']['forward']['ms'], 'gemm', d['qlinear']['gemm']['ms'], 'e2e', d['e2e']['value'])"
  tail -2 gpurun_out/ab_pdl$pdl.err
done
