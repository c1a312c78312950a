"""Golden values of the reference's density inputs (examples/densities/*.msh, binary Gmsh 2.2 with a per-element "density" field)
as read by voxelfem_b200/compat/msh.py and mapped onto the simulator grid by centroid (elementIndexFromMeshIO,
TensorProductSimulator.hh:729-744).  Run in the build container (needs /root/reference); writes tests/golden/msh_densities.json.
Only digests and a few sample values are stored -- the reference's files are not copied."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from voxelfem_b200.compat import msh  # noqa: E402

REF = "/root/reference/examples/densities"


class Grid:
    def __init__(self, ne, dmax): self.NbElementsPerDimension = np.array(ne); self.domain = (np.zeros(len(ne)), np.array(dmax, dtype=float))


out = {}
for name in sorted(os.listdir(REF)):
    m = msh.read_msh(os.path.join(REF, name))
    V, F = m["vertices"], m["elements"]
    dim = 2 if m["element_type"] == 3 else 3
    ext = V.max(axis=0) - V.min(axis=0)
    h = np.abs(V[F[0]] - V[F[0]][0]).max(axis=0)[:dim]                      # edge lengths of the first element
    ne = np.rint(ext[:dim] / h).astype(int)
    rho = msh.densities_from_msh(Grid(ne, ext[:dim]), os.path.join(REF, name))
    out[name] = dict(binary=bool(m["binary"]), num_vertices=int(V.shape[0]), num_elements=int(F.shape[0]), element_type=int(m["element_type"]),
                     fields={k: [d, list(a.shape)] for k, (d, a) in m["fields"].items()}, grid=[int(v) for v in ne],
                     density_sum=float(rho.sum()), density_min=float(rho.min()), density_max=float(rho.max()),
                     density_sha256=hashlib.sha256(np.ascontiguousarray(rho).tobytes()).hexdigest(),
                     sample_indices=[0, 1, int(rho.size // 3), int(rho.size // 2), int(rho.size - 1)],
                     sample_values=[float(rho[i]) for i in (0, 1, rho.size // 3, rho.size // 2, rho.size - 1)])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "msh_densities.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
