"""Multi-GPU parity under pytest: the NCCL slab-partitioned solve (one process per GPU, ncclSend/Recv ghost planes, ncclAllReduce'd
PCG scalars) against the oracle's undivided solve, and the slab-partitioned layer-by-layer evaluator against the oracle's.  Needs two visible devices; on a single-GPU box the same control flow is covered
by the local groups of tests/test_gpu_slabs.py and tests/test_gpu_baseline_configs.py::test_slab_partitioned_solve_matches_oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_solve_matches_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two CUDA devices (found %d)" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "dist_oracle_check.py")]
    env = dict(os.environ); env["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 2) // 2))
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and r.stdout.count("-> OK") == 6        # 2 solves + 1 layer-by-layer run, on 2 ranks
