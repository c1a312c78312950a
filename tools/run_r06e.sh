#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06e}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${T}_pytest_parity.log 2>&1; tail -n 3 $O/${T}_pytest_parity.log
echo early=1; timeout 300 python tools/time_ops.py > $O/${T}_time_ops_early1.log 2>&1; grep -E "level [12] |vcycle from level 1|FMG" $O/${T}_time_ops_early1.log
echo early=0; VF_ST_EARLY_TMA=0 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_early0.log 2>&1; grep -E "level [12] |vcycle from level 1|FMG" $O/${T}_time_ops_early0.log
