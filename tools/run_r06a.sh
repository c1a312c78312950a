#!/bin/bash
# round-2 session-3 check: residual-emitting pre-smoothing sweep (A/B), coarse factorization with the shared pivot reciprocal
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06a}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${T}_pytest_parity.log 2>&1; tail -3 $O/${T}_pytest_parity.log
timeout 300 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -k "separate_residual or dense_coarse" > $O/${T}_pytest_variants.log 2>&1; tail -3 $O/${T}_pytest_variants.log
timeout 300 python tools/time_ops.py > $O/${T}_time_ops_fused.log 2>&1; grep -E "level [012] |vcycle|FMG" $O/${T}_time_ops_fused.log
VF_GS_RESIDUAL=0 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_separate.log 2>&1; grep -E "vcycle|FMG" $O/${T}_time_ops_separate.log
VF_GS_RESIDUAL_MIN_NODES=1000000 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_fused_l1only.log 2>&1; grep -E "vcycle|FMG" $O/${T}_time_ops_fused_l1only.log
timeout 120 python tools/coarse_factor_driver.py > $O/${T}_coarse.log 2>&1; tail -3 $O/${T}_coarse.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.log 2> $O/${T}_bench.err; tail -1 $O/${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e'])
print(json.dumps(d['config']['extra'])[:600])
"
