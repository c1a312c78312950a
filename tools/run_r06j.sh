#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06j}; O=gpurun_out; mkdir -p $O
timeout 400 ncu --clock-control none --set full --warp-sampling-interval 0 --import-source on -k regex:k_potrf_inv_small -s 70 -c 1 -o $O/${T}_potrf -f python tools/coarse_factor_driver.py > $O/${T}_potrf.log 2>&1; tail -n 2 $O/${T}_potrf.log
