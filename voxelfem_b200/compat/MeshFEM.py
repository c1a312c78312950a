"""Import shim: the reference's drivers `import MeshFEM` only to make its bundled Python helpers importable
(python/CoarseningLevelBenchmark.py:8).  Nothing of MeshFEM's simplicial FEM is on the B200 hot path."""
