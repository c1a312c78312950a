"""CPU pinning of the oracle's topology-optimization layer against independent numpy
restatements of the published formulas and against finite differences -- the reference ships no golden vectors for it."""
import itertools
import os

import numpy as np
import pytest

import oracle
from oracle import OracleMG, OracleProblem, OracleSim


def smoothing_reference(x, shape, r, linear):
    """SmoothingFilter (TopologyOptimizationFilter.hh:312-386): offsets [-r, r]^N, w = (r + 1) - |offset|_2 (dropped if <= 0) or 1,
    neighbour indices reflected (i < 0 -> -i - 1, i >= n -> 2n - i - 1), output = sum w x / sum w."""
    x = np.asarray(x, dtype=float).reshape(shape)
    out = np.zeros_like(x)
    offs = [(o, (r + 1) - np.linalg.norm(o) if linear else 1.0) for o in itertools.product(range(-r, r + 1), repeat=len(shape))]
    offs = [(o, w) for o, w in offs if w > 0]
    wsum = sum(w for _, w in offs)
    for idx in np.ndindex(*shape):
        acc = 0.0
        for o, w in offs:
            j = []
            for i, d, n in zip(idx, o, shape):
                k = i + d
                k = -k - 1 if k < 0 else (2 * n - k - 1 if k >= n else k)
                j.append(k)
            acc += w * x[tuple(j)]
        out[idx] = acc / wsum
    return out.ravel()


@pytest.mark.parametrize("shape,r,linear", [((7, 5), 1, False), ((6, 9), 2, True), ((5, 4, 6), 1, True), ((4, 5, 3), 2, False), ((8, 3, 4), 3, True)])
def test_smoothing_filter_matches_formula(shape, r, linear):
    rng = np.random.default_rng(0)
    x, y = rng.uniform(size=int(np.prod(shape))), rng.uniform(size=int(np.prod(shape)))
    fx = oracle.smoothing_filter(x, shape, r, int(linear))
    assert np.abs(fx - smoothing_reference(x, shape, r, linear)).max() < 1e-14
    assert abs(fx @ y - x @ oracle.smoothing_filter(y, shape, r, int(linear))) < 1e-13      # apply == backprop (:297-310)
    assert np.abs(oracle.smoothing_filter(np.ones_like(x), shape, r, int(linear)) - 1).max() < 1e-15   # partition of unity


@pytest.mark.parametrize("beta", [0.5, 1.0, 4.0, 16.0])
def test_projection_filter_closed_form_and_derivative(beta):
    x = np.random.default_rng(1).uniform(size=200)
    p = oracle.projection_apply(x, beta)
    ref = (np.tanh(beta / 2) + np.tanh(beta * (x - 0.5))) / (2 * np.tanh(beta / 2))          # TopologyOptimizationFilter.hh:199-232
    assert np.abs(p - ref).max() < 1e-15
    g = np.random.default_rng(2).normal(size=x.size)
    eps = 1e-6
    fd = (oracle.projection_apply(x + eps, beta) - oracle.projection_apply(x - eps, beta)) / (2 * eps)
    assert np.abs(oracle.projection_backprop(g, x, beta) - g * fd).max() < 1e-7 * max(1.0, beta)
    # invert: atanh((2 y - 1) tanh(beta / 2)) / beta + 1 / 2 maps the projected value back
    assert np.abs(np.arctanh((2 * p - 1) * np.tanh(beta / 2)) / beta + 0.5 - x).max() < 1e-9


def _small_problem(data_dir, ne=(8, 4, 4), levels=1, filters=(("smooth", 1, 1), ("project", 2.0))):
    s = OracleSim(np.array(ne), np.zeros(3), np.array([2.0, 1.0, 1.0]))
    s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-3, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc"))
    s.set_uniform_density(1.0)
    p = OracleProblem(OracleMG(s, levels), list(filters), 0.4)
    p.set_solver(400, 1e-13, 1, 2, True, True)
    return s, p


def test_compliance_sensitivity_finite_difference(data_dir):
    """dJ/d rho_e = -1/2 gamma rho^(gamma-1) (E0 - Emin) u_e^T K0 u_e (TensorProductSimulator.hh:972-1005) against central
    differences of J = 1/2 f.u with u from the direct solver."""
    ne = (6, 4, 4)
    s = OracleSim(np.array(ne), np.zeros(3), np.array([1.5, 1.0, 1.0]))
    s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-3, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc"))
    rho = np.random.default_rng(3).uniform(0.3, 0.9, int(np.prod(ne)))
    s.set_densities(rho)
    f = s.build_load()
    g = s.compliance_gradient(s.solve(f))
    d = np.random.default_rng(4).normal(size=rho.size)
    eps = 1e-6
    J = []
    for sgn in (1, -1):
        s.set_densities(rho + sgn * eps * d)
        J.append(0.5 * float((f * s.solve(f)).sum()))
    assert abs((J[0] - J[1]) / (2 * eps) - g @ d) < 1e-6 * abs(g @ d)
    assert (g < 0).all()                                                                   # stiffer is always better for compliance


def test_problem_gradients_through_the_filter_chain(data_dir):
    """evaluateObjectiveGradient / evaluateConstraintsJacobian (TopologyOptimizationProblem.hh:77-121) = chain rule through
    smoothing + projection: central differences on the design variables."""
    s, p = _small_problem(data_dir)
    n = s.num_elements
    x = np.random.default_rng(5).uniform(0.3, 0.7, n)
    p.set_vars(x)
    gJ, gc = p.objective_gradient(), p.constraint_jacobian()
    d = np.random.default_rng(6).normal(size=n)
    eps = 1e-6
    vals = []
    for sgn in (1, -1):
        p.set_vars(x + sgn * eps * d)
        vals.append((p.compliance(), p.constraint()))
    assert abs((vals[0][0] - vals[1][0]) / (2 * eps) - gJ @ d) < 2e-6 * abs(gJ @ d)
    assert abs((vals[0][1] - vals[1][1]) / (2 * eps) - gc @ d) < 1e-7 * abs(gc @ d)
    # constraint value: 1 - mean(rho_phys) / V (TopologyOptimizationConstraint.hh:30-32)
    p.set_vars(x)
    assert abs(p.constraint() - (1 - p.physical_vars().mean() / 0.4)) < 1e-14


def test_oc_step_matches_independent_bisection(data_dir):
    """OCOptimizer::step (OptimalityCriterion.hh:51-134) restated in numpy on top of the oracle's filters: bracket dilation 32
    around the midpoint (floor 0.01), halve / double up to 100 times, bisect until |c| <= ctol; update
    clamp(x (dJ / (dc lambda))^p, x -+ m, [0, 1])."""
    s, p = _small_problem(data_dir)
    n, shape, V = s.num_elements, (8, 4, 4), 0.4
    x = np.full(n, 0.45) * (1 + 0.1 * np.sin(np.arange(n)))
    p.set_vars(x)
    dJ, dc = p.objective_gradient(), p.constraint_jacobian()
    m, pw, ctol = 0.2, 0.5, 1e-6

    def stepped(lam):
        with np.errstate(invalid="ignore", divide="ignore"):
            cand = x * np.power(dJ / (dc * lam), pw)
        cand = np.where(np.isfinite(cand), cand, x)
        return np.clip(np.clip(cand, x - m, x + m), 0.0, 1.0)

    def ceval(lam):
        phys = oracle.projection_apply(oracle.smoothing_filter(stepped(lam), shape, 1, 1), 2.0)
        return 1 - phys.mean() / V
    lo, hi = 1.0, 2.0
    mid = 0.5 * (lo + hi)
    hi = 32 * hi + (1 - 32) * mid
    lo = max(32 * lo + (1 - 32) * mid, 0.01)
    nit = 0
    while nit < 100 and not ceval(lo) < 0:
        hi, lo, nit = lo, lo / 2, nit + 1
    if nit == 0:
        while nit < 100 and not ceval(hi) > 0:
            lo, hi, nit = hi, hi * 2, nit + 1
    while True:
        mid = 0.5 * (lo + hi)
        v = ceval(mid)
        if abs(v) <= ctol:
            break
        if v < 0: lo = mid
        if v > 0: hi = mid
    p.oc_step(m, pw, ctol)
    assert np.abs(p.design_vars() - stepped(mid)).max() < 1e-12
    assert abs(p.constraint()) <= ctol * (1 + 1e-9)
