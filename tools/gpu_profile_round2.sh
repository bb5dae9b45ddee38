#!/bin/bash
# Round 2, one GPU call: plain run, ncu launch list, ncu --set full captures of STEADY-STATE launches of the two main kernels
# (30 warm-up cycles first: every instance one SQP iteration, no start-up stragglers), the latency kernel, and the list of FP64 /
# DMMA pipe metrics this ncu knows.  Outputs under gpurun_out/ (scratch); tools/summarise_profiles_r2.py writes profiles/.
set -x
CMD="python bench.py --steps 3 --warmup 30 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/r2_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
ncu --query-metrics 2>/dev/null | grep -i -E "fp64|dmma" > gpurun_out/r2_metrics_fp64.txt
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mlp|k_sqp_warp" -s 93 -c 3 \
    --metrics sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fmaheavy.sum \
    -o gpurun_out/r2_prof_steady $CMD > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu2.log
python bench.py --config c5 --steps 60 --warmup 20 --no-cpu-baseline > gpurun_out/r2_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_sqp_cta" -s 60 -c 1 -o gpurun_out/r2_prof_cta python bench.py --config c5 --steps 60 --warmup 20 --no-cpu-baseline > gpurun_out/r2_ncu3.log 2>&1
tail -3 gpurun_out/r2_ncu3.log
