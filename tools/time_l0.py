"""Level-0 kernel timings only (smoothing sweep, residual, apply) for quick A/B runs of kernel variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from voxelfem_b200 import capi
workload = sys.argv[1] if len(sys.argv) > 1 else "C3_pcg_256^3"
s, mg = bench.setup(capi.Sim, capi.MG, workload, capi.DATA_DIR)
tag = " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("VF_"))
for op in ("smooth", "residual", "apply"):
    print("%-40s level 0 %-9s %8.4f ms" % (tag, op, mg.time_op(op, 0, reps=5)))
