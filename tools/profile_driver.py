"""Minimal driver for ncu captures: builds the C3 (256^3) problem and runs a PCG solve capped at a few iterations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voxelfem_b200 import capi  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "C3_pcg_256^3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s, mg = bench.setup(capi.Sim, capi.MG, workload, capi.DATA_DIR)
ndof = s.N * s.num_nodes
x, b = capi.DeviceArray(ndof), capi.DeviceArray(ndof)
capi._check(capi.lib().vf_sim_build_load_vector_dev(s.h, b.ptr))
pcg = dict(bench.PCG); pcg["max_iter"] = iters
it, res = mg.pcg_dev(x, b, **pcg)
print("iterations", it, "residuals", res)
