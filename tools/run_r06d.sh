#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06d}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -x -q > $O/${T}_pytest_parity.log 2>&1; tail -n 3 $O/${T}_pytest_parity.log
timeout 300 python tools/time_ops.py > $O/${T}_time_ops_fused.log 2>&1; grep -E "level [123] |vcycle|FMG" $O/${T}_time_ops_fused.log
VF_ST_POSTAB=0 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_nopostab.log 2>&1; grep -E "level 1|FMG" $O/${T}_time_ops_nopostab.log
VF_GS_RESIDUAL=0 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_separate.log 2>&1; grep -E "FMG" $O/${T}_time_ops_separate.log
