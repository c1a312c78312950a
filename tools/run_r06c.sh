#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06c}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -x -q > $O/${T}_pytest_parity.log 2>&1; tail -n 3 $O/${T}_pytest_parity.log
timeout 300 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -k "separate_residual or position_table" > $O/${T}_pytest_variants.log 2>&1; tail -n 3 $O/${T}_pytest_variants.log
timeout 300 python tools/time_ops.py > $O/${T}_time_ops_fused.log 2>&1; grep -E "level [12] |vcycle|FMG" $O/${T}_time_ops_fused.log
VF_GS_RESIDUAL=0 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_separate.log 2>&1; grep -E "vcycle|FMG" $O/${T}_time_ops_separate.log
VF_ST_POSTAB=0 timeout 300 python tools/time_ops.py > $O/${T}_time_ops_nopostab.log 2>&1; grep -E "level 1|FMG" $O/${T}_time_ops_nopostab.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.log 2> $O/${T}_bench.err; tail -1 $O/${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e'])
"
