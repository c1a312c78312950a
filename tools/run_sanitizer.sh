#!/bin/bash
# compute-sanitizer memcheck + racecheck on small instances of the hot path: V-cycle / PCG parity cases incl. the
# TMA-staged level-0 smoother (forced on short rows), the Q2 operator, a small layer-by-layer run.  usage (under gpurun):
#   bash tools/run_sanitizer.sh <tag>
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r05}
O=gpurun_out
mkdir -p $O
SEL="test_vcycle_matches_oracle or test_residual_emitting_sweep or test_rebuild_every_solve_mode or test_pcg_parity or test_masked_pcg or test_direct_solve_single_level"
VF_GS_ROWS=2 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_q2.py -m gpu -x -q -k "$SEL or q2_apply or q2_cantilever" > $O/${T}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/${T}_memcheck.log | tail -3
VF_GS_ROWS=2 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_vcycle_matches_oracle or test_residual_emitting_sweep or test_direct_solve_single_level" > $O/${T}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/${T}_racecheck.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_mma_lbl.py tests/test_gpu_filters.py -m gpu -x -q > $O/${T}_memcheck_lbl_filters.log 2>&1
echo "memcheck (lbl, mma, filters) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/${T}_memcheck_lbl_filters.log | tail -3
