"""MeshFEM's `mesh` python module as far as the VoxelFEM drivers reach it (python/CoarseningLevelBenchmark.py:8 imports it; the
visualisation helpers use its .msh field reader / writer): MSHFieldParser, MSHFieldWriter (3rdParty/MeshFEM/src/python_bindings/
MSHFieldParser_bindings.cc, MSHFieldWriter_bindings.cc) on top of compat/msh.py.  MeshFEM's simplicial meshes themselves are not on
the B200 path."""
from voxelfem_b200.compat.msh import DomainType, MSHFieldParser, MSHFieldWriter  # noqa: F401
