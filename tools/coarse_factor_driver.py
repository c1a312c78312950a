"""Builds the C2 (128x64x64, 3 levels) hierarchy once more after warm-up: target of the ncu launch list of the coarse factorization."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelfem_b200 import capi
ne = np.array([128, 64, 64])
s = capi.Sim(ne, np.zeros(3), np.array([2.0, 1.0, 1.0]))
s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0)
s.apply_bc_file(os.path.join(capi.DATA_DIR, "bcs", "3D/cantilever_flexion_E.bc")); s.set_uniform_density(0.5)
mg = capi.MG(s, 3)
for rep in range(3):
    s.set_uniform_density(0.5 + 0.01 * rep)
    t0 = time.perf_counter(); mg.update_stiffness(); mg.synchronize(); print("update_stiffness %.3f ms" % (1e3 * (time.perf_counter() - t0)))
