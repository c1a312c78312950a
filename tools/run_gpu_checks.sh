#!/bin/bash
# One-GPU round check under gpurun: full GPU test suite, headline bench (both arms), per-operation timings, the C2 / C5 drivers.
# usage: gpurun --timeout 2400 -- bash tools/run_gpu_checks.sh <tag>
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r05}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; tail -3 $O/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.log 2> $O/${T}_bench.err; tail -1 $O/${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['cpu_baseline'])
r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','traffic','share_of_step')}, r.get('fp64'))
for k,v in r['north_star_kernels'].items(): print(k, round(v['frac'],3), v.get('fp64',{}).get('frac'), v['avg_launch_us'])
print(json.dumps(d['config']['extra'])[:1500])
"
tail -3 $O/${T}_bench.err
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > $O/${T}_bench_ref.log 2>&1; tail -5 $O/${T}_bench_ref.log | cut -c1-700
timeout 300 python tools/time_ops.py > $O/${T}_time_ops.log 2>&1; grep -E "level [01] |vcycle|FMG" $O/${T}_time_ops.log
timeout 300 python tools/bench_topopt.py > $O/${T}_topopt_C2.log 2>&1; tail -1 $O/${T}_topopt_C2.log | cut -c1-300
timeout 300 python tools/bench_lbl.py > $O/${T}_lbl_C5.log 2>&1; tail -1 $O/${T}_lbl_C5.log | cut -c1-400
# ncu launch list of one solve capped at 1 PCG iteration (per-launch device times: shares of the step, not bench values)
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/${T}_launches.csv python tools/profile_driver.py "C3_pcg_256^3" 1 > $O/${T}_launches.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv > $O/${T}_launches_summary.txt 2>&1; head -14 $O/${T}_launches_summary.txt
