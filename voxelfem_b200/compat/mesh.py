"""Import shim for MeshFEM's `mesh` module (imported, not used, by python/CoarseningLevelBenchmark.py:8)."""
