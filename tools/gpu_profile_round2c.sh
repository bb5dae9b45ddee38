#!/bin/bash
# Round 2, final build of the last session, one GPU call: plain run, ncu launch list of steady-state cycles, ncu --set full captures of one steady-state cycle's
# k_mlp_oz (int8-split MLP kernel on tcgen05) and k_sqp_warp launches (settle phase of 40 cycles first: every instance one SQP iteration).
# Outputs under gpurun_out/ (scratch); tools/summarise_profiles_r2b.py writes the summaries under profiles/.
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/r2c_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 60 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
$CMD > gpurun_out/r2c_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mlp_oz|k_sqp_warp" -s 130 -c 3 \
    -o gpurun_out/r2c_prof_steady $CMD > gpurun_out/r2c_ncu2.log 2>&1
tail -3 gpurun_out/r2c_ncu2.log
