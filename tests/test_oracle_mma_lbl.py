"""CPU tests of the oracle's MMA (MethodOfMovingAsymptotes.hh) and layer-by-layer evaluator (LayerByLayer.hh)."""
import os

import numpy as np
import pytest

import oracle
from mma_problems import compliance_like, svanberg_toy
from oracle import OracleLBL, OracleMG, OracleMMA, OracleSim


@pytest.mark.parametrize("gcmma", [False, True])
def test_mma_reproduces_svanberg_toy_optimum(gcmma):
    n, m, lo, hi, f, df, x0, xstar = svanberg_toy()
    opt = OracleMMA(n, m, lo, hi, f, df)
    opt.enableGCMMA(gcmma)
    opt.setInitialVar(x0)
    for _ in range(12):
        opt.step()
    x = opt.getOptimalVar()
    assert np.abs(x - xstar).max() < (2e-4 if gcmma else 2e-6)
    assert np.abs(x - np.array([2.0175, 1.7800, 1.2375])).max() < 1e-4      # published 4-digit optimum
    if not gcmma:
        from scipy.optimize import minimize
        r = minimize(lambda z: f(z)[0], x0, jac=lambda z: df(z)[0], bounds=[(0, 5)] * 3, method="SLSQP", options={"ftol": 1e-15},
                     constraints=[{"type": "ineq", "fun": lambda z: -f(z)[1:], "jac": lambda z: -df(z)[1:]}])
        assert np.abs(x - r.x).max() < 1e-6
    assert abs(f(x)[0] - 8.7702) < 2e-4
    assert f(x)[1:].max() < 1e-6


def test_mma_matches_scipy_on_volume_constrained_problem():
    from scipy.optimize import minimize
    n, m, lo, hi, f, df, x0, _ = compliance_like(40)
    opt = OracleMMA(n, m, lo, hi, f, df)
    opt.setInitialVar(x0)
    for _ in range(60):
        opt.step()
    x = opt.getOptimalVar()
    r = minimize(lambda z: f(z)[0], x0, jac=lambda z: df(z)[0], bounds=list(zip(lo, hi)), method="SLSQP",
                 constraints=[{"type": "ineq", "fun": lambda z: -f(z)[1], "jac": lambda z: -df(z)[1]}], options={"ftol": 1e-15, "maxiter": 500})
    assert abs(f(x)[0] - r.fun) < 1e-6 * abs(r.fun)
    assert f(x)[1] < 1e-7


def _lbl_setup(cls_sim, cls_mg, ne=(8, 8, 4), levels=1, rho=None):
    ne = np.array(ne)
    s = cls_sim(ne, np.zeros(len(ne)), ne.astype(float) / ne[0])
    s.set_isotropic(1.0, 0.3)
    s.set_interp(1, 1.0, 1e-4, 3.0, 3.0)       # RAMP, q = 3 (python/LayerByLayerObjective.py:19-20)
    s.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7)   # clamp the build plate y = 0
    g = np.zeros(len(ne)); g[1] = -1.0
    s.set_gravity(g)
    s.set_densities(rho if rho is not None else np.random.default_rng(3).uniform(0.3, 1.0, int(np.prod(ne))))
    return s, cls_mg(s, levels)


def test_lbl_subspace_recurrences_match_brute_force():
    """InitGenSubspace's recurrences for A = U^T K U, b = U^T f (LayerByLayer.hh:149-202) versus the brute-force values the
    reference documents them to equal: the subspace guess must make the PCG start closer than the zero guess, and the whole
    run must agree with a run that uses zero guesses (same per-layer solutions up to the solver tolerance)."""
    s, mg = _lbl_setup(OracleSim, OracleMG)
    runs = {}
    for method in ("zero", "N=3", "fd", "constant"):
        ev = OracleLBL(mg)
        ev.select_init_method(method)
        it, comp = ev.run(True, 1, 200, 1e-10, 1, 1, False)
        runs[method] = (it, comp, ev.objective(), ev.gradient())
    it0, c0, o0, g0 = runs["zero"]
    assert len(it0) == 8
    for method in ("N=3", "fd", "constant"):
        it, c, o, g = runs[method]
        assert np.allclose(c, c0, rtol=1e-7)
        assert abs(o - o0) < 1e-7 * abs(o0)
        assert np.abs(g - g0).max() < 1e-6 * np.abs(g0).max()
    assert runs["N=3"][0][1:].sum() < it0[1:].sum()      # warm starts pay off


def test_lbl_gradient_finite_difference():
    """fd_validation protocol (3rdParty/MeshFEM/python/fd_validation.py:34-57) on the layer-by-layer objective."""
    rho = np.random.default_rng(5).uniform(0.4, 0.9, 8 * 4 * 4)
    s, mg = _lbl_setup(OracleSim, OracleMG, ne=(8, 4, 4), rho=rho)
    ev = OracleLBL(mg)
    ev.run(True, 1, 300, 1e-12, 1, 1, False)
    g = ev.gradient()
    d = np.random.default_rng(6).normal(size=rho.size)
    eps = 1e-5
    vals = []
    for sgn in (+1, -1):
        s.set_densities(rho + sgn * eps * d)
        ev.run(True, 1, 300, 1e-12, 1, 1, False)
        vals.append(ev.objective())
    fd = (vals[0] - vals[1]) / (2 * eps)
    assert abs(fd - g @ d) < 1e-5 * abs(fd)


def test_lbl_layer_increment_and_errors():
    s, mg = _lbl_setup(OracleSim, OracleMG)
    ev = OracleLBL(mg)
    it, comp = ev.run(True, 2, 100, 1e-8, 1, 1, False)
    assert len(it) == 4
    with pytest.raises(RuntimeError, match="Unrecognized"):
        ev.select_init_method("bogus")
    s.set_gravity([0.0, 0.0, 0.0])
    with pytest.raises(RuntimeError, match="gravity"):
        ev.run(True, 1, 10, 1e-5, 1, 1, False)
