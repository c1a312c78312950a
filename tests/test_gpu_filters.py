"""GPU parity of UpsampleFilter / VertexToCellFilter / LangelaarFilter / PythonFilter (SURVEY.md 8f rank 2;
TopologyOptimizationFilter.hh:247-275, 418-712) against the oracle restatements, stand-alone and inside a TopologyOptimizationProblem."""
import os

import numpy as np
import pytest

import oracle
from oracle import OracleMG, OracleSim

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(11)


@pytest.fixture(scope="module")
def capi():
    from voxelfem_b200 import capi as c
    assert c.device_count() > 0
    return c


def relmax(a, b): return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("shape,f", [((3, 4), 2), ((5, 3), 4), ((9, 5), 2), ((3, 4, 2), 2), ((2, 3, 3), 3), ((17, 9, 9), 2)])
def test_upsample(capi, shape, f):
    x = RNG.normal(size=shape)
    assert relmax(capi.upsample_filter(x, shape, f), oracle.upsample(x, shape, f)) < 1e-14
    g = RNG.normal(size=tuple((s - 1) * f + 1 for s in shape))
    assert relmax(capi.upsample_filter_backprop(g, shape, f), oracle.upsample_backprop(g, shape, f)) < 1e-14


@pytest.mark.parametrize("shape", [(3, 4), (33, 17), (4, 3, 5), (17, 9, 9)])
def test_vertex_to_cell(capi, shape):
    x = RNG.normal(size=shape)
    assert relmax(capi.vertex_to_cell_filter(x, shape), oracle.vertex_to_cell(x, shape)) < 1e-15
    g = RNG.normal(size=tuple(s - 1 for s in shape))
    assert relmax(capi.vertex_to_cell_filter_backprop(g, shape), oracle.vertex_to_cell_backprop(g, shape)) < 1e-15


@pytest.mark.parametrize("shape", [(7, 6), (32, 16), (3, 4, 3), (8, 12, 6)])
def test_langelaar(capi, shape):
    x = RNG.uniform(0.02, 1.0, shape); prev = RNG.uniform(0.0, 1.0, shape)
    yg, sg = capi.langelaar_filter(x, shape, out_prev=prev)
    yo, so = oracle.langelaar(x, shape, out_prev=prev)
    assert relmax(yg, yo) < 1e-13 and relmax(sg, so) < 1e-13
    w = RNG.normal(size=shape)
    assert relmax(capi.langelaar_filter_backprop(w, x, yo, so, shape), oracle.langelaar_backprop(w, x, yo, so, shape)) < 1e-12


def _problem(capi, ne, dom, bc, filters, vol, data_dir, rho0=0.5):
    s = capi.Sim(np.array(ne), np.zeros(len(ne)), np.array(dom, dtype=float))
    s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", bc)); s.set_uniform_density(rho0)
    mg = capi.MG(s, 2)
    p = capi.Problem(mg, filters, vol)
    p.set_solver(100, 1e-10, 1, 2, True, False)
    return s, mg, p


def test_chain_with_dimension_changing_filters(capi, data_dir):
    """design variables on a coarse VERTEX grid -> Upsample(2) -> VertexToCell -> Smoothing -> Projection -> element densities
    (the chain of the reference's multi-resolution demos): sizes, forward values and back-propagated gradients."""
    ne = (16, 8)
    filters = [("upsample", 2), ("vertex_to_cell",), ("smooth", 1, 1), ("project", 2.0)]
    s, mg, p = _problem(capi, ne, (2.0, 1.0), "mbb_N.bc", filters, 0.5, data_dir)
    assert list(p.grid_dims()) == [9, 5] and list(p.grid_dims(True)) == [16, 8] and p.nv == 45      # (17, 9) vertices <- (9, 5)
    x = RNG.uniform(0.2, 0.9, 45)
    p.set_vars(x)
    v1 = oracle.upsample(x, (9, 5), 2); v2 = oracle.vertex_to_cell(v1, (17, 9)); v3 = oracle.smoothing_filter(v2, ne, 1, 1); v4 = oracle.projection_apply(v3, 2.0)
    assert relmax(p.physical_vars(), v4) < 1e-13
    # gradient: oracle compliance gradient at the GPU's displacement, pulled back through the oracle's filters
    os_ = OracleSim(np.array(ne), np.zeros(2), np.array([2.0, 1.0])); os_.set_isotropic(1.0, 0.3); os_.set_interp(0, 1.0, 1e-4, 3.0, 3.0); os_.set_densities(v4)
    g = os_.compliance_gradient(p.u())
    g = oracle.projection_backprop(g, v3, 2.0); g = oracle.smoothing_filter(g, ne, 1, 1); g = oracle.vertex_to_cell_backprop(g, (17, 9)); g = oracle.upsample_backprop(g, (9, 5), 2)
    assert relmax(p.objective_gradient(), g) < 1e-10
    dc = np.full(int(np.prod(ne)), -1.0 / (0.5 * np.prod(ne)))
    dc = oracle.projection_backprop(dc, v3, 2.0); dc = oracle.smoothing_filter(dc, ne, 1, 1); dc = oracle.vertex_to_cell_backprop(dc, (17, 9)); dc = oracle.upsample_backprop(dc, (9, 5), 2)
    assert relmax(p.constraint_jacobian(), dc) < 1e-12
    assert abs(p.constraint() - (1 - v4.mean() / 0.5)) < 1e-13
    n = p.oc_step()                                                          # the OC update runs on the 45 design variables
    assert n > 0 and p.design_vars().shape == (45,) and abs(p.constraint()) <= 1e-6


@pytest.mark.parametrize("ne,dom,bc", [((16, 8), (2.0, 1.0), "mbb_N.bc"), ((8, 8, 4), (2.0, 2.0, 1.0), "3D/cantilever_flexion_E.bc")])
def test_chain_with_langelaar_filter(capi, data_dir, ne, dom, bc):
    filters = [("smooth", 1, 1), ("langelaar",)]
    s, mg, p = _problem(capi, ne, dom, bc, filters, 0.6, data_dir)
    n = int(np.prod(ne))
    prev = np.zeros(n)
    for trial in range(2):          # the second application sees the first one's output as the array's previous content (3D)
        x = RNG.uniform(0.2, 1.0, n)
        p.set_vars(x)
        v1 = oracle.smoothing_filter(x, ne, 1, 1)
        v2, sm = oracle.langelaar(v1, ne, out_prev=prev)
        assert relmax(p.physical_vars(), v2) < 1e-12
        os_ = OracleSim(np.array(ne), np.zeros(len(ne)), np.array(dom, dtype=float)); os_.set_isotropic(1.0, 0.3); os_.set_interp(0, 1.0, 1e-4, 3.0, 3.0); os_.set_densities(v2)
        g = os_.compliance_gradient(p.u())
        g = oracle.langelaar_backprop(g, v1, v2, sm, ne); g = oracle.smoothing_filter(g, ne, 1, 1)
        assert relmax(p.objective_gradient(), g) < 1e-9
        prev = v2


def test_python_filter_in_chain(capi, data_dir):
    ne = (16, 8)
    calls = []
    def ap(x): calls.append("a"); return x ** 2
    def bp(g, v): calls.append("b"); return g * 2 * v
    s, mg, p = _problem(capi, ne, (2.0, 1.0), "mbb_N.bc", [("smooth", 1, 0), ("python", ap, bp)], 0.5, data_dir)
    x = RNG.uniform(0.3, 1.0, int(np.prod(ne)))
    p.set_vars(x)
    v1 = oracle.smoothing_filter(x, ne, 1, 0)
    assert relmax(p.physical_vars(), v1 ** 2) < 1e-14
    dc = oracle.smoothing_filter(np.full(x.size, -1.0 / (0.5 * x.size)) * 2 * v1, ne, 1, 0)
    assert relmax(p.constraint_jacobian(), dc) < 1e-13 and "a" in calls and "b" in calls
    def bad(x): raise ValueError("boom")
    s2, mg2, p2 = _problem(capi, ne, (2.0, 1.0), "mbb_N.bc", [("python", bad, bp)], 0.5, data_dir)
    with pytest.raises(ValueError, match="boom"):
        p2.set_vars(x)
