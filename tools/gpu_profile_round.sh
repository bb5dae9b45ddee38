#!/bin/bash
# One GPU call: bench (plain), ncu launch list, ncu full captures of the two main kernels.  Outputs under gpurun_out/.
set -x
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err || exit 1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mlp|k_sqp_warp" -s 6 -c 3 -o gpurun_out/prof_r1_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
