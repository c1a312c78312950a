"""Import shim next to the pybind11 modules: re-exports voxelfem_b200/compat/mesh.py (see there)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
from voxelfem_b200.compat.mesh import *  # noqa: F401,F403,E402
