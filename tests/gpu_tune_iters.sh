#!/bin/bash
# GPU-box tool: K1/K2 vectors-per-thread cap against shape (QAT_B200_MAX_ITERS), from bench.py --shape-sweep.
for it in 2 4 8; do
  QAT_B200_MAX_ITERS=$it python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-qat-step --shape-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('MAX_ITERS', $it)
for k,v in d['shape_sweep'].items():
    if 'sym8' in k: print('  ', k, 'fwd', v['fwd']['us'], v['fwd']['frac'], 'fwd_bwd', v['fwd_bwd']['us'], v['fwd_bwd']['frac'])"
done
