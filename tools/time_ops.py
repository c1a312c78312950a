"""Per-operation device timings of the multigrid hierarchy (vf_mg_time_op) with roofline fractions."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voxelfem_b200 import capi  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "C3_pcg_256^3"
s, mg = bench.setup(capi.Sim, capi.MG, workload, capi.DATA_DIR)
levels = bench.WORKLOADS[workload][3]
peak = 6455.3
rows = []
for l in range(levels):
    nn = mg.nn(l)
    for op, bytes_per_node in (("smooth", 80.0 if l == 0 else 80.0 + 1944), ("residual", 80.0 if l == 0 else 80 + 1944), ("apply", 56.0 if l == 0 else 56 + 1944)):
        ms = mg.time_op(op, l, reps=5 if l == 0 else 20)
        gbs = bytes_per_node * nn / ms / 1e6
        rows.append((l, op, nn, ms, gbs, gbs / peak))
        print("level %d %-9s nodes %9d  %8.4f ms  %8.1f GB/s algorithmic  %5.1f%% of HBM peak" % (l, op, nn, ms, gbs, 100 * gbs / peak))
    if l > 0:
        try:
            ms = mg.time_op("smooth_residual", l, reps=20)
            print("level %d %-9s nodes %9d  %8.4f ms  (forward sweep that also emits the residual)" % (l, "smooth+r", nn, ms))
        except RuntimeError as e:
            print("level %d smooth+r unavailable: %s" % (l, e))
    if l < levels:
        for op in ("restrict", "prolong"):
            print("level %d %-9s %8.4f ms" % (l, op, mg.time_op(op, l, reps=20)))
print("coarse solve %.4f ms" % mg.time_op("coarse_solve", 0, reps=50))
for l in range(levels - 1, -1, -1):
    print("vcycle from level %d: %.4f ms" % (l, mg.time_op("vcycle", l, reps=5)))
print("FMG cycle: %.4f ms" % mg.time_op("fmg", 0, reps=5))
