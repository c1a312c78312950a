#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06i}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "residual_emitting or hierarchy or vcycle" > $O/${T}_pytest_parity.log 2>&1; tail -n 3 $O/${T}_pytest_parity.log
timeout 300 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -k "separate_residual or position_table or persistent" > $O/${T}_pytest_variants.log 2>&1; tail -n 3 $O/${T}_pytest_variants.log
timeout 400 ncu --clock-control none --set full --import-source on -k regex:k_potrf_inv_small -s 70 -c 1 -o $O/${T}_potrf -f python tools/coarse_factor_driver.py > $O/${T}_potrf.log 2>&1; tail -n 2 $O/${T}_potrf.log
