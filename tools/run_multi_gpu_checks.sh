#!/bin/bash
# N-GPU round check under gpurun --gpus N: NCCL parity (tools/dist_check.py, tests/test_gpu_nccl.py) and bench.py --gpus N (weak-scaling
# solve + the C4 topology optimization).  usage: gpurun --gpus N --timeout 600 -- bash tools/run_multi_gpu_checks.sh <tag> N
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r05}; N=${2:-2}
O=gpurun_out
mkdir -p $O
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > $O/${T}_dist_check_${N}gpu.log 2>&1; echo "dist_check rc=$?"
grep -c -- "-> OK" $O/${T}_dist_check_${N}gpu.log
if [ "$N" = "2" ]; then timeout 200 python -m pytest tests/test_gpu_nccl.py -q -s > $O/${T}_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; tail -1 $O/${T}_pytest_nccl.log; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > $O/${T}_bench_${N}gpu.log 2> $O/${T}_bench_${N}gpu.err; echo "bench rc=$?"
tail -1 $O/${T}_bench_${N}gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'iters', d['config']['pcg_iterations_per_solve'])
c=d['config'].get('extra',{}).get('C4',{})
print({k:c.get(k) for k in ('workload','ms_per_iteration','topopt_iterations_per_s','pcg_iterations','compliance')})
"
