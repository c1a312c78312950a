#!/bin/bash
# ncu --set full of the 8 residual-emitting level-1 colour passes (first FMG cycle of a 256^3 solve capped at 1 iteration)
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06b}; O=gpurun_out; mkdir -p $O
timeout 900 ncu --clock-control none --set full --import-source on -k regex:k_stencil_tile -s 96 -c 9 -o $O/${T}_stencil_l1_res -f python tools/profile_driver.py "C3_pcg_256^3" 1 > $O/${T}_stencil_l1_res.log 2>&1
tail -2 $O/${T}_stencil_l1_res.log
