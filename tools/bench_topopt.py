"""Config 2 (BASELINE.json configs[1]): 3D cantilever 128x64x64 SIMP compliance topology optimization, smoothing filter r = 3
(linear) + projection beta = 1, OC updates, MultigridComplianceObjective defaults (cgIter 100, tol 1e-5, 1 V-cycle with 2
smoothing sweeps, FMG, warm start), 3 coarsening levels -- python/3DTopoptDemo.ipynb cells 1, 5 with the cantilever BC.
Prints topopt iterations/s (setup excluded) and the per-kernel-family device-time breakdown."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelfem_b200 import capi  # noqa: E402


def run(ne=(128, 64, 64), dom=(2.0, 1.0, 1.0), levels=3, iters=50, V=0.3, profile=True):
    s = capi.Sim(np.array(ne), np.zeros(3), np.array(dom))
    s.set_isotropic(1.0, 0.3)
    s.set_interp(0, 1.0, 1e-4, 3.0, 3.0)
    s.apply_bc_file(os.path.join(capi.DATA_DIR, "bcs", "3D", "cantilever_flexion_E.bc"))
    s.set_uniform_density(1.0)
    mg = capi.MG(s, levels)
    p = capi.Problem(mg, [("smooth", 3, 1), ("project", 1.0)], V)
    p.set_solver(100, 1e-5, 1, 2, True, False)
    x0 = np.full(int(np.prod(ne)), 0.5 + np.arctanh((2 * V - 1) * np.tanh(0.5)) / 1.0)   # ProjectionFilter(1).invert(V)
    p.set_vars(x0)
    p.oc_step()                     # warm-up iteration (hierarchy build, workspace allocation)
    mg.synchronize()
    mg.prof_reset(); mg.prof_enable(True)
    capi.lib().vf_reset_kernel_launch_count()
    t0 = time.perf_counter()
    pcg_its, evals = [], []
    for _ in range(iters):
        evals.append(p.oc_step())
        pcg_its.append(p.last_pcg_iters())
    mg.synchronize()
    dt = time.perf_counter() - t0
    mg.prof_enable(False)
    out = {"workload": "C2_topopt_%dx%dx%d" % tuple(ne), "iterations": iters, "seconds": dt, "topopt_iterations_per_s": iters / dt,
           "ms_per_iteration": 1e3 * dt / iters, "pcg_iterations": pcg_its, "oc_constraint_evals": evals,
           "compliance": p.compliance(), "volume_constraint": p.constraint(), "gpu_launches": int(capi.lib().vf_kernel_launch_count())}
    t0 = time.perf_counter()
    for _ in range(5):
        mg.update_stiffness()
    out["update_stiffness_ms"] = 1e3 * (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(5):
        p.objective_gradient()
    out["objective_gradient_ms_incl_d2h"] = 1e3 * (time.perf_counter() - t0) / 5
    if profile:
        prof = mg.prof_report()
        out["device_ms_by_family"] = {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    return out


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    print(json.dumps(run(iters=iters)))
