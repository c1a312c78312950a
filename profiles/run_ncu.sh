#!/bin/bash
# Runs on the GPU box under gpurun: launch list + full captures of the hot kernels (B200_PROFILING.md recipe).
# usage: profiles/run_ncu.sh <tag> [what...]   what in: launches gs_l0 apply_l0 stencil
TAG=${1:-r01}; shift
WHAT=${@:-launches gs_l0 apply_l0 stencil}   # also: stencil_res potrf
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
DRV="python tools/profile_driver.py C3_pcg_256^3 1"
for w in $WHAT; do case $w in
  launches) # every launch of one capped solve with its device time
    $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $OUT/${TAG}_launches.csv $DRV > $OUT/${TAG}_launches.log 2>&1 ;;
  gs_l0)    $NCU --set full --import-source on -k regex:k_gs -s 2 -c 2 -o $OUT/${TAG}_gs_l0 -f $DRV > $OUT/${TAG}_gs_l0.log 2>&1 ;;
  apply_l0) $NCU --set full --import-source on -k regex:k_apply -s 1 -c 2 -o $OUT/${TAG}_apply_l0 -f $DRV > $OUT/${TAG}_apply_l0.log 2>&1 ;;
  stencil)  # level-1 colour passes (launches 102..109 of k_stencil_tile* in the first FMG cycle) and the level-1 residual (110)
    $NCU --set full --import-source on -k regex:k_stencil_tile -s 102 -c 9 -o $OUT/${TAG}_stencil_l1 -f $DRV > $OUT/${TAG}_stencil_l1.log 2>&1 ;;
  stencil_res) # the 8 residual-emitting level-1 colour passes of the first FMG cycle (launches 96..103 of k_stencil_tile*; r06b)
    $NCU --set full --import-source on -k regex:k_stencil_tile -s 96 -c 9 -o $OUT/${TAG}_stencil_l1_res -f $DRV > $OUT/${TAG}_stencil_l1_res.log 2>&1 ;;
  potrf)    # one 64 x 64 diagonal-block factorization of a C2 hierarchy rebuild with the densest warp sampling (a 25 us single-block kernel; r06j)
    $NCU --set full --warp-sampling-interval 0 --import-source on -k regex:k_potrf_inv_small -s 70 -c 1 -o $OUT/${TAG}_potrf -f python tools/coarse_factor_driver.py > $OUT/${TAG}_potrf.log 2>&1 ;;
esac; done
ls -la $OUT
