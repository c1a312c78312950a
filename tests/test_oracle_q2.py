"""Pins the numpy restatement of the reference's generic (degree-2) element path, oracle/q2ref.py: the reference instantiates Q2
nowhere and ships no outputs (parity unpinned), so the checks are invariants and closed forms --
  symmetry, positive semi-definiteness and the rigid-body null space of K0;
  exact strain energies of linear AND quadratic displacement fields (a Q2 element reproduces both);
  agreement of the same construction at degree 1 with the Q1 oracle's K0;
  the assembled operator against the matrix-free element scatter; a manufactured solve."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import q2ref  # noqa: E402
from oracle import OracleSim  # noqa: E402


@pytest.mark.parametrize("N,h,nu", [(2, (1.0, 1.0), 0.3), (2, (0.5, 0.25), 0.0), (3, (1.0, 1.0, 1.0), 0.3), (3, (0.25, 0.5, 0.125), 0.2)])
def test_degree_one_matches_q1_oracle(N, h, nu):
    ne = np.array([2] * N)
    o = OracleSim(ne, np.zeros(N), ne * np.array(h)); o.set_isotropic(1.3, nu)
    K = q2ref.element_stiffness(N, 1, np.array(h), q2ref.isotropic_tensor(N, 1.3, nu))
    assert np.abs(K - o.K0()).max() < 1e-14 * np.abs(K).max()


@pytest.mark.parametrize("N,h", [(2, (1.0, 0.5)), (3, (0.5, 1.0, 0.25))])
def test_q2_element_matrix_invariants(N, h):
    D = q2ref.isotropic_tensor(N, 2.0, 0.3)
    K = q2ref.element_stiffness(N, 2, np.array(h), D)
    n = N * 3 ** N
    assert K.shape == (n, n) and np.abs(K - K.T).max() < 1e-14 * np.abs(K).max()
    w = np.linalg.eigvalsh(K)
    nrigid = 3 if N == 2 else 6
    assert np.all(w[:nrigid] < 1e-12 * w[-1]) and w[nrigid] > 1e-6 * w[-1]              # exactly the rigid-body modes are free
    s = q2ref.Q2Sim([1] * N, np.zeros(N), np.array(h)); s.set_isotropic(2.0, 0.3)
    X = s.node_positions()
    rng = np.random.default_rng(0)
    # linear field u = A x + b: energy = vol * eps : C : eps
    A = rng.normal(size=(N, N)); b = rng.normal(size=N)
    u = X @ A.T + b
    eps = 0.5 * (A + A.T)
    flat = np.array([eps[i, i] for i in range(N)] + ([eps[0, 1]] if N == 2 else [eps[1, 2], eps[0, 2], eps[0, 1]]))
    shear = np.array([1.0] * N + [2.0] * (len(flat) - N))
    exact = np.prod(h) * (flat * shear) @ D @ (flat * shear)
    assert abs(u.ravel() @ K @ u.ravel() - exact) < 1e-12 * abs(exact)
    # quadratic field u_c = x^T Q_c x: strain is linear in x, energy integrated exactly by a 2-point Gauss rule per axis
    Q = rng.normal(size=(N, N, N)); Q = 0.5 * (Q + Q.transpose(0, 2, 1))
    uq = np.einsum("ni,cij,nj->nc", X, Q, X)
    gp, gw = q2ref.gauss01(2)
    e_ref = 0.0
    import itertools
    for q in itertools.product(range(2), repeat=N):
        x = np.array([gp[q[d]] * h[d] for d in range(N)])
        G = 2.0 * np.einsum("cij,j->ci", Q, x)                                         # grad u (c, i)
        e = 0.5 * (G + G.T)
        fl = np.array([e[i, i] for i in range(N)] + ([e[0, 1]] if N == 2 else [e[1, 2], e[0, 2], e[0, 1]]))
        e_ref += np.prod([gw[q[d]] for d in range(N)]) * np.prod(h) * (fl * shear) @ D @ (fl * shear)
    assert abs(uq.ravel() @ K @ uq.ravel() - e_ref) < 1e-11 * abs(e_ref)


def test_q2_operator_assembly_and_solve():
    s = q2ref.Q2Sim([3, 2, 2], np.zeros(3), np.array([1.5, 1.0, 1.0])); s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-3, 3.0, 3.0)
    rng = np.random.default_rng(2)
    s.set_densities(rng.uniform(0.1, 1.0, s.num_elements))
    assert s.num_nodes == 7 * 5 * 5 and s.enodes.shape == (12, 27)
    u = rng.normal(size=(s.num_nodes, 3))
    K = s.assemble()
    assert np.abs((K @ u.ravel()).reshape(-1, 3) - s.apply_K(u)).max() < 1e-12
    assert abs(0.5 * u.ravel() @ (K @ u.ravel()) - 0.5 * (s.E() * s.element_energies(u)).sum()) < 1e-10
    # manufactured solve: clamp the plane x = 0, recover a random admissible field from its own load
    X = s.node_positions()
    fixed = np.repeat((X[:, 0] == 0)[:, None], 3, axis=1)
    ut = rng.normal(size=(s.num_nodes, 3)); ut[fixed] = 0
    f = s.apply_K(ut)
    us = s.solve(f, fixed)
    assert np.abs(us - ut).max() < 1e-8 * np.abs(ut).max()
