#!/bin/bash
# GPU-box tool: e2e (host buffers) throughput against the pipeline's chunk size.
for mb in 2 4 8 16; do
  QAT_B200_HOST_CHUNK_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qat-step 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_mb', $mb, 'e2e', l['e2e']['value'], 'GB/s', l['e2e']['ms_per_step'], 'ms', 'qlinear', l['qlinear']['gemm'], l['qlinear']['forward'])"
done
