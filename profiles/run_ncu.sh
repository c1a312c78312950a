#!/bin/bash
# Runs on the GPU box under gpurun: launch list + full captures of the hot kernels (B200_PROFILING.md recipe).
# usage: profiles/run_ncu.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
# 1) every launch of one capped solve with its device time
$NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $OUT/${TAG}_launches.csv python tools/profile_driver.py "C3_pcg_256^3" 1 > $OUT/${TAG}_launches.log 2>&1
# 2) full captures: level-0 GS colour pass, level-0 apply/residual, level-1 stencil GS pass (launch 96.. of the first FMG cycle)
$NCU --set full --import-source on -k regex:k_gs_l0 -s 2 -c 2 -o $OUT/${TAG}_gs_l0 -f python tools/profile_driver.py "C3_pcg_256^3" 1 > $OUT/${TAG}_gs_l0.log 2>&1
$NCU --set full --import-source on -k regex:k_apply_l0 -s 1 -c 2 -o $OUT/${TAG}_apply_l0 -f python tools/profile_driver.py "C3_pcg_256^3" 1 > $OUT/${TAG}_apply_l0.log 2>&1
$NCU --set full --import-source on -k regex:k_gs_stencil -s 96 -c 2 -o $OUT/${TAG}_gs_stencil_l1 -f python tools/profile_driver.py "C3_pcg_256^3" 1 > $OUT/${TAG}_gs_stencil.log 2>&1
ls -la $OUT
