"""Gmsh 2.2 .msh field I/O (voxelfem_b200/compat/msh.py): round trips in both encodings, the
MSHFieldParser / MSHFieldWriter surface, the centroid mapping onto the simulator grid, and -- where the reference tree is present
(the build container) -- the reference's own density files against the committed digests (tests/golden/msh_densities.json,
made by tests/golden/make_msh_fixture.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from voxelfem_b200.compat import msh, tps_extras

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/examples/densities"


class Grid:
    """The part of the simulator API the host-side helpers need (no device)."""

    def __init__(self, ne, dmin, dmax):
        self.NbElementsPerDimension = np.array(ne); self.domain = (np.array(dmin, dtype=float), np.array(dmax, dtype=float))
    def getMesh(self): return tps_extras.getMesh(self)
    def numNodes(self): return int(np.prod(self.NbElementsPerDimension + 1))


@pytest.mark.parametrize("ne,dmax", [((5, 3), (2.5, 1.5)), ((4, 3, 2), (2.0, 1.5, 1.0))])
@pytest.mark.parametrize("binary", [True, False])
def test_round_trip_through_simulator_ordering(tmp_path, ne, dmax, binary):
    g = Grid(ne, np.zeros(len(ne)), dmax)
    rng = np.random.default_rng(3)
    rho = rng.uniform(0, 1, int(np.prod(ne)))
    u = rng.normal(size=(g.numNodes(), len(ne)))
    path = str(tmp_path / "f.msh")
    msh.write_fields(g, path, element_fields={"density": rho}, node_fields={"u": u}, binary=binary)
    m = msh.read_msh(path)
    assert m["binary"] == binary and m["element_type"] == (3 if len(ne) == 2 else 5)
    V, F = g.getMesh()
    assert np.array_equal(m["elements"], F) and np.allclose(m["vertices"], V, rtol=0, atol=0)
    assert np.array_equal(msh.densities_from_msh(g, path), rho)                       # bit-exact in both encodings (%.17g)
    p = msh.MSHFieldParser(path)
    assert p.scalarFieldNames() == ["density"] and p.vectorFieldNames() == ["u"]
    assert p.meshDimension() == len(ne) and p.meshDegree() == 1 and p.numElements() == rho.size
    assert np.array_equal(p.vectorField("u"), u) and np.array_equal(p.scalarField("density", msh.DomainType.PER_ELEMENT), rho)
    with pytest.raises(RuntimeError): p.scalarField("density", msh.DomainType.PER_NODE)


def test_centroid_mapping_ignores_the_files_numbering(tmp_path):
    """A file whose nodes run x-fastest and whose elements are shuffled (as the reference's files are numbered differently from
    the simulator) lands on the same grid cells."""
    ne = (6, 4)
    g = Grid(ne, (0, 0), (6.0, 4.0))
    xs, ys = np.meshgrid(np.arange(ne[0] + 1.0), np.arange(ne[1] + 1.0), indexing="xy")   # x fastest
    V = np.stack([xs.ravel(), ys.ravel()], axis=1)
    nid = lambda i, j: j * (ne[0] + 1) + i
    cells = [(i, j) for j in range(ne[1]) for i in range(ne[0])]
    perm = np.random.default_rng(0).permutation(len(cells))
    F = np.array([[nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)] for (i, j) in [cells[k] for k in perm]])
    val = np.array([10.0 * i + j for (i, j) in [cells[k] for k in perm]])
    path = str(tmp_path / "g.msh")
    w = msh.MSHFieldWriter(path, V, F, binary=True); w.addField("density", val); w.close()
    rho = msh.densities_from_msh(g, path).reshape(ne)
    assert np.array_equal(rho, 10.0 * np.arange(ne[0])[:, None] + np.arange(ne[1])[None, :])


def test_get_mesh_ordering():
    """getMesh (TensorProductSimulator.hh:747-777): Gmsh ordering = local vertex pairs with every odd pair swapped."""
    V, F = tps_extras.getMesh(Grid((2, 1), (0, 0), (2.0, 1.0)))
    assert V.shape == (6, 3) and np.all(V[:, 2] == 0)
    # nodes are numbered y-fastest: (0,0)=0 (0,1)=1 (1,0)=2 (1,1)=3 ...; element 0 has local nodes [0, 1, 2, 3] -> [0, 1, 3, 2]
    assert F.tolist() == [[0, 1, 3, 2], [2, 3, 5, 4]]
    V3, F3 = tps_extras.getMesh(Grid((1, 1, 1), (0, 0, 0), (1.0, 1.0, 1.0)))
    assert F3.tolist() == [[0, 1, 3, 2, 4, 5, 7, 6]]
    # counter-clockwise quads / positively oriented hexahedra
    a, b, c = V[F[0][0]], V[F[0][1]], V[F[0][3]]
    assert np.cross(b - a, c - a)[2] != 0


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_reference_density_files_match_committed_digests():
    gold = json.load(open(os.path.join(HERE, "golden", "msh_densities.json")))
    assert sorted(gold) == sorted(os.listdir(REF))
    for name, g in gold.items():
        m = msh.read_msh(os.path.join(REF, name))
        assert (bool(m["binary"]), m["vertices"].shape[0], m["elements"].shape[0], m["element_type"]) == (g["binary"], g["num_vertices"], g["num_elements"], g["element_type"])
        ext = m["vertices"].max(axis=0) - m["vertices"].min(axis=0)
        rho = msh.densities_from_msh(Grid(g["grid"], np.zeros(len(g["grid"])), ext[:len(g["grid"])]), os.path.join(REF, name))
        assert hashlib.sha256(np.ascontiguousarray(rho).tobytes()).hexdigest() == g["density_sha256"]
        assert [float(rho[i]) for i in g["sample_indices"]] == g["sample_values"]
        assert rho.min() == g["density_min"] and rho.max() == g["density_max"] and abs(rho.sum() - g["density_sum"]) < 1e-9
