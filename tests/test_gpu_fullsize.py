"""Size-independent properties at BASELINE.json's full single-GPU size (3D 256^3, configs[2]) where the CPU oracle is too slow
to run side by side: symmetry, linearity and null space of the level-0 operator, agreement of the two level-0 kernels
(symmetry-adapted vs dense), adjointness of the grid transfers, and a full MG-PCG solve on a heterogeneous density field
whose residual is re-evaluated independently."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NE = (256, 256, 256)


@pytest.fixture(scope="module")
def capi():
    from voxelfem_b200 import capi as c
    assert c.device_count() > 0
    return c


@pytest.fixture(scope="module")
def problem(capi, data_dir):
    ne = np.array(NE)
    s = capi.Sim(ne, np.zeros(3), np.ones(3))
    s.set_isotropic(1.0, 0.3)
    s.set_interp(0, 1.0, 1e-5, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc"))
    # heterogeneous field of SURVEY.md section 8d: one pass of the linear r = 2 filter over uniform noise
    rho = np.clip(capi.smoothing_filter(np.random.default_rng(0).uniform(0, 1, int(np.prod(ne))), ne, 2, 1), 0, 1)
    s.set_densities(rho)
    return s


def test_level0_operator_properties_at_full_size(capi, problem):
    s = problem
    rng = np.random.default_rng(1)
    nn = s.num_nodes
    u, v = rng.standard_normal((nn, 3)), rng.standard_normal((nn, 3))
    Ku, Kv = s.apply_K(u), s.apply_K(v)
    scale = np.abs(Ku).max()
    assert abs((u * Kv).sum() - (v * Ku).sum()) < 1e-11 * abs((u * Kv).sum())          # symmetry
    a, b = 0.7, -1.3
    assert np.abs(s.apply_K(a * u + b * v) - (a * Ku + b * Kv)).max() < 1e-12 * scale   # linearity
    t = np.tile(np.array([1.0, -2.0, 0.5]), (nn, 1))
    assert np.abs(s.apply_K(t)).max() < 1e-12 * scale                                   # rigid translations are in the null space
    assert (u * Ku).sum() > 0                                                           # positive semi-definite
    # out (+=, -=) K u variants against the plain product
    base = rng.standard_normal((nn, 3))
    assert np.abs(s.apply_K(u, out=base, zero_init=False, negate=True) - (base - Ku)).max() < 1e-12 * scale
    # the dense kernel (any material) against the symmetry-adapted one (VF_L0_DENSE is read when the material is set)
    os.environ["VF_L0_DENSE"] = "1"
    try:
        s.set_isotropic(1.0, 0.3)
        Ku_dense = s.apply_K(u)
    finally:
        del os.environ["VF_L0_DENSE"]
        s.set_isotropic(1.0, 0.3)
    assert np.abs(Ku_dense - Ku).max() < 1e-13 * scale
    assert np.abs(s.apply_K(u) - Ku).max() == 0                                         # deterministic


def test_full_size_solve_and_transfers(capi, problem):
    s = problem
    mg = capi.MG(s, 5)
    f = s.build_load()
    u, it, res = mg.pcg(np.zeros_like(f), f, 100, 1e-10, 1, 1, True)
    assert 5 <= it <= 40, it
    assert np.all(np.diff(np.log(res)) < 0.5)                                           # no blow-up along the way
    free = ~np.stack([(s.dirichlet_mask() >> c) & 1 for c in range(3)], axis=1).astype(bool)
    r = (f - s.apply_K(u)) * free                                                       # residual re-evaluated outside the solver
    assert np.linalg.norm(r) <= 1.05e-10 * np.linalg.norm(f)
    assert np.abs(u[~free]).max() == 0                                                  # Dirichlet values imposed exactly
    assert 0.5 * (f * u).sum() > 0
    # restriction is the transpose of interpolation on every level pair (MultigridSolver.hh:216-275 vs :130-212)
    rng = np.random.default_rng(2)
    for l in range(2):
        xc = rng.standard_normal((mg.nn(l + 1), 3)); yf = rng.standard_normal((mg.nn(l), 3))
        lhs, rhs = (mg.interpolate(l, xc) * yf).sum(), (xc * mg.restrict(l, yf)).sum()
        assert abs(lhs - rhs) < 1e-11 * abs(lhs)
    # Galerkin identity on the first coarse level: A_1 x = R A_0 P x   (MultigridSolver.hh:711-819)
    xc = rng.standard_normal((mg.nn(1), 3))
    a1 = mg.apply_K(1, xc)
    ref = mg.restrict(0, s.apply_K(mg.interpolate(0, xc)))
    assert np.abs(a1 - ref).max() < 1e-11 * np.abs(ref).max()
    # The residual-emitting sweep at level 1 of the full-size hierarchy (4.2 GB stencil, the level that streams from HBM): its iterate
    # equals the plain sweep's (same arithmetic) and its residual equals computeResidual of that iterate (MultigridSolver.hh:452-458, 527-541),
    # for both sweep directions; Dirichlet components are exactly zero.
    dm1 = mg.get_sim(1).dirichlet_mask()
    bits = (dm1[:, None] >> np.arange(3)[None, :]) & 1
    u1 = rng.standard_normal((mg.nn(1), 3)); u1[bits == 1] = 0
    b1 = rng.standard_normal(u1.shape)
    scale = np.abs(mg.residual(1, u1, b1)).max()
    for fwd in (True, False):
        us, rs = mg.smooth_residual(1, u1, b1, fwd)
        assert np.abs(us - mg.smooth(1, u1, b1, fwd)).max() <= 1e-13 * np.abs(us).max()
        assert np.isfinite(rs).all() and np.all(rs[bits == 1] == 0.0)
        assert np.abs(rs - mg.residual(1, us, b1)).max() < 1e-11 * scale
