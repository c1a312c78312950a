#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06k}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${T}_pytest_parity.log 2>&1; tail -n 3 $O/${T}_pytest_parity.log
timeout 300 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "dense_coarse or C2 or C1" > $O/${T}_pytest_variants.log 2>&1; tail -n 3 $O/${T}_pytest_variants.log
timeout 120 python tools/coarse_factor_driver.py > $O/${T}_coarse.log 2>&1; tail -3 $O/${T}_coarse.log
timeout 300 python tools/bench_topopt.py > $O/${T}_topopt_C2.log 2>&1; tail -1 $O/${T}_topopt_C2.log | cut -c1-200
