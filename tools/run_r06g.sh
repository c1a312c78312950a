#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06g}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; tail -n 3 $O/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.log 2> $O/${T}_bench.err; tail -1 $O/${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e'])
r=d['roofline']
print(d['ms_per_step'], {k:(round(v['ms_per_launch'],4), v['launches']) for k,v in r['families'].items()})
print(json.dumps(d['config']['extra'])[:900])
"
timeout 300 python tools/time_ops.py > $O/${T}_time_ops.log 2>&1; grep -E "level [012] |vcycle|FMG" $O/${T}_time_ops.log
