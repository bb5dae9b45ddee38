#!/bin/bash
# Residency experiment: pad the SQP kernel's dynamic shared memory to lower the resident CTAs per SM (5 -> 4 -> 3 -> 2).
for pad in 0 8000 28000 60000; do
  MPCC_SQPW_SMEM_PAD=$pad timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pad', $pad, 'ms/step', round(d['ms_per_step'],2), 'k_sqp_warp', d['kernels_ms']['k_sqp_warp'], d['latency_ms'])"
done
