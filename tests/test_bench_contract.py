"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the reference algorithm timed on the host
cores) prints one JSON line with the keys the driver reads, uses the thread count it reports even when the launcher exported
OMP_NUM_THREADS=1 (torchrun does), and non-zero ranks of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ); env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C3_pcg_64^3", "--steps", "1", "--warmup", "0", *args],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_json_line():
    r = _run({"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "DOF*iters/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("MG-PCG DOF*iterations per second")
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] == os.cpu_count() and "PCG capped" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "DOF*iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "C3_pcg_64^3" and line["vs_baseline"] is None and line["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_silently():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
