"""Kernel-variant A/B timings: smoothing sweep / residual / apply on levels 0-2 and the FMG cycle, tagged with the VF_* environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from voxelfem_b200 import capi
workload = sys.argv[1] if len(sys.argv) > 1 else "C3_pcg_256^3"
s, mg = bench.setup(capi.Sim, capi.MG, workload, capi.DATA_DIR)
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("VF_")) or "default"
row = []
for l in (0, 1, 2):
    for op in ("smooth", "residual", "apply"):
        row.append("L%d %s %.4f" % (l, op, mg.time_op(op, l, reps=5 if l == 0 else 20)))
row.append("fmg %.4f" % mg.time_op("fmg", 0, reps=5))
import time
mg.update_stiffness(); mg.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    mg.update_stiffness()
mg.synchronize()
row.append("update_stiffness %.3f" % (1e3 * (time.perf_counter() - t0) / 3))
print("%-28s %s" % (tag, " | ".join(row)))
