"""pyOptimizer -- the reference's optimizer module (python_bindings/Optimizer.cc:11-23) over libvoxelfem_b200:
MMA(numVars, numConstr, xmin, xmax, f, df_dx) with setInitialVar / step / enableGCMMA (MethodOfMovingAsymptotes.hh:28-470).
All O(n) work of a step runs in CUDA kernels (vf_mma_*); f and df_dx are the caller's Python callbacks, as in the reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
from voxelfem_b200.capi import MMA  # noqa: E402,F401
