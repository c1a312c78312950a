"""Degree-2 (Q2) elements on the GPU (vf_q2_*, csrc/vf_q2.cu) against the numpy restatement of the reference's generic element path
(oracle/q2ref.py).  Tolerances as for the Q1 operators: element matrix and a single operator
application to rounding, a converged solve to 1e-6 relative L2 (north_star)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import q2ref  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from voxelfem_b200 import capi as c
    assert c.device_count() > 0
    return c


def rel(a, b): return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def pair(capi, ne, dmax, nu=0.3, emin=1e-4, law=0):
    ne = np.array(ne); N = len(ne)
    g = capi.SimQ2(ne, np.zeros(N), np.array(dmax, dtype=float)); o = q2ref.Q2Sim(ne, np.zeros(N), np.array(dmax, dtype=float))
    for s in (g, o):
        s.set_isotropic(1.7, nu); s.set_interp(law, 1.0, emin, 3.0, 3.0)
    return g, o


@pytest.mark.parametrize("ne,dmax,nu", [((2, 2), (2.0, 2.0), 0.3), ((3, 2), (1.5, 0.5), 0.0), ((2, 2, 2), (2.0, 2.0, 2.0), 0.3), ((2, 1, 3), (0.5, 0.5, 0.375), 0.2)])
def test_q2_element_matrix(capi, ne, dmax, nu):
    g, o = pair(capi, ne, dmax, nu)
    assert g.num_nodes == o.num_nodes == int(np.prod(2 * np.array(ne) + 1)) and g.num_elements == o.num_elements
    K = g.K0()
    assert rel(K, o.K0) < 1e-13 and np.abs(K - K.T).max() == 0


@pytest.mark.parametrize("ne,dmax,law", [((5, 4), (2.5, 2.0), 0), ((1, 1), (1.0, 1.0), 0), ((7, 3), (7.0, 1.5), 1), ((4, 3, 2), (2.0, 1.5, 1.0), 0), ((1, 2, 1), (1.0, 2.0, 1.0), 1),
                                         ((5, 5, 6), (1.0, 1.0, 1.2), 0)])
def test_q2_apply_K(capi, ne, dmax, law):
    g, o = pair(capi, ne, dmax, law=law)
    rng = np.random.default_rng(7)
    rho = rng.uniform(0.0, 1.0, o.num_elements)
    g.set_densities(rho); o.set_densities(rho)
    assert rel(g.E(), o.E()) < 1e-15
    u = rng.normal(size=(o.num_nodes, o.N))
    f = o.apply_K(u)
    assert rel(g.apply_K(u), f) < 1e-12
    b = rng.normal(size=u.shape)
    assert rel(g.apply_K(u, out=b, zero_init=False, negate=True), b - f) < 1e-12      # applyK<ZeroInit = false, Negate = true>
    assert rel(g.apply_K(u, out=b, zero_init=False, negate=False), b + f) < 1e-12
    assert rel(g.element_energies(u), o.element_energies(u)) < 1e-12
    # symmetry and the rigid translations of the assembled operator
    v = rng.normal(size=u.shape)
    assert abs((v * g.apply_K(u)).sum() - (u * g.apply_K(v)).sum()) < 1e-10 * abs((v * f).sum())
    assert np.abs(g.apply_K(np.ones_like(u))).max() < 1e-11 * np.abs(f).max()


@pytest.mark.parametrize("ne,dmax", [((8, 4), (2.0, 1.0)), ((6, 3, 3), (2.0, 1.0, 1.0))])
def test_q2_cantilever_solve(capi, ne, dmax):
    """Clamped at x = 0, unit downward load on the x = max face: Jacobi-PCG on the device against the oracle's sparse direct solve."""
    g, o = pair(capi, ne, dmax, emin=1e-3)
    rng = np.random.default_rng(1)
    rho = rng.uniform(0.3, 1.0, o.num_elements)
    g.set_densities(rho); o.set_densities(rho)
    X = o.node_positions()
    fixed = np.repeat((X[:, 0] == 0)[:, None], o.N, axis=1)
    f = np.zeros((o.num_nodes, o.N))
    face = X[:, 0] == dmax[0]
    f[face, o.N - 1] = -1.0 / face.sum()
    uo = o.solve(f, fixed)
    ug, it, rr = g.pcg(np.zeros_like(f), f, fixed, max_iter=5000, tol=1e-11)
    assert rr <= 1e-11 and it < 5000
    assert np.linalg.norm(ug - uo) < 1e-6 * np.linalg.norm(uo) and np.all(ug[fixed] == 0)
    assert abs((f * ug).sum() - (f * uo).sum()) < 1e-8 * abs((f * uo).sum())         # compliance
    print("Q2 %s: Jacobi-PCG iterations %d, rel residual %.2e, rel-L2(u) vs direct solve %.2e" % (ne, it, rr, np.linalg.norm(ug - uo) / np.linalg.norm(uo)))
