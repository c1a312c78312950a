"""voxelfem_b200 -- B200-native implementation of VoxelFEM's MG-PCG / topology-optimization hot path.

The compute path lives in libvoxelfem_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/voxelfem_b200.h).  `capi` is the ctypes binding; the drop-in `pyVoxelFEM` / `pyOptimizer`
pybind11 modules are built from voxelfem_b200/host/.
"""
__version__ = "0.1.0"
