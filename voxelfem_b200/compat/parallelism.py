"""Import shim for MeshFEM's `parallelism` module (TBB thread control, python/LayerByLayerOptimization.py:8).  The GPU
path has no host thread pool to size; the setters are accepted and ignored."""


def set_max_num_tbb_threads(n): return None
def set_hessian_assembly_num_threads(n): return None
def set_gradient_assembly_num_threads(n): return None
