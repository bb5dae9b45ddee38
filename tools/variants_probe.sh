#!/bin/bash
# A/B of prebuilt library variants (libmpcc_b200_<name>.so next to the default) on the bench workload.
cd "$(dirname "$0")/../mpcc_manipulator_b200"
cp libmpcc_b200.so /tmp/lib_default.so
for v in default "$@"; do
  if [ $v != default ]; then cp libmpcc_b200_$v.so libmpcc_b200.so; fi
  (cd ..; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],2), d['kernels_ms']['k_sqp_warp'], d['latency_ms'], d['last_step_stats'])")
  cp /tmp/lib_default.so libmpcc_b200.so
done
