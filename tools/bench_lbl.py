"""Config 5 (BASELINE.json configs[4]): 3D layer-by-layer simulation with warm-started multigrid (LayerByLayer.hh:223-296),
LayerByLayerObjective defaults (python/LayerByLayerObjective.py:19-20, 87-89): RAMP q = 3, PCG maxIter 50 / tol 1e-5 /
1 V-cycle / 1 smoothing sweep / no FMG, initial guesses from the N = 3 subspace method, gravity (0, -1, 0), build platform
y = 0 fully clamped, rho = 0.6.  Layers run along reference axis 1, so "256 layers on a 256x128x128 grid" is the grid
[128, 256, 128] here (SURVEY.md section 8, C5 caveat).  Prints layers/s and the per-layer PCG iteration counts."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelfem_b200 import capi  # noqa: E402


def run(ne=(128, 256, 128), levels=4, rho=0.6, method="N=3", increment=1, profile=False):
    ne = np.array(ne)
    s = capi.Sim(ne, np.zeros(3), ne.astype(float) / ne.max())
    s.set_isotropic(1.0, 0.3)
    s.set_interp(1, 1.0, 1e-4, 3.0, 3.0)
    s.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7)
    s.set_gravity([0.0, -1.0, 0.0])
    s.set_uniform_density(rho)
    mg = capi.MG(s, levels)
    ev = capi.LBL(mg)
    ev.select_init_method(method)
    capi.lib().vf_reset_kernel_launch_count()
    mg.prof_reset(); mg.prof_enable(profile)
    t0 = time.perf_counter()
    its, comps = ev.run(True, increment, 50, 1e-5, 1, 1, False)
    mg.synchronize()
    dt = time.perf_counter() - t0
    mg.prof_enable(False)
    prof = mg.prof_report()
    return {"workload": "C5_lbl_%dx%dx%d" % tuple(ne), "layers": int(len(its)), "seconds": dt, "layers_per_s": len(its) / dt,
            "pcg_iterations_total": int(its.sum()), "pcg_iterations_per_layer": [int(i) for i in its], "objective": ev.objective(),
            "dof_iterations_per_s": float(3 * s.num_nodes * its.sum() / dt), "gpu_launches": int(capi.lib().vf_kernel_launch_count()),
            "instrumented": profile,
            "device_ms_by_family": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])} if profile else None}


if __name__ == "__main__":
    small = "small" in sys.argv[1:]
    prof = "profile" in sys.argv[1:]    # CUDA events around every launch, no graph replay: per-family device time, slower wall clock
    print(json.dumps(run(ne=(32, 64, 32), levels=2, profile=prof) if small else run(profile=prof)))
