"""Slab-partitioned MG-PCG: the grid cut into slabs along axis 0 must reproduce the undivided solve.

Local groups run all parts in one process on one GPU (exchanges are device copies), exercising exactly the control flow,
windows, ghost-plane exchanges, stencil completion and replicated coarse levels that the NCCL ranks run."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from voxelfem_b200 import capi as c
    return c


def build_reference(capi, ne, dom, bc, levels, rho, emin):
    s = capi.Sim(np.array(ne), np.zeros(3), np.array(dom))
    s.set_isotropic(1.0, 0.3)
    s.set_interp(0, 1.0, emin, 3.0, 3.0)
    s.apply_bc_file(bc)
    s.set_densities(rho)
    mg = capi.MG(s, levels)
    f = s.build_load()
    return s, mg, f


def solve_slabs(capi, ne, dom, bc, levels, first_rep, nparts, rho, emin, f_global, pcg):
    ranges = capi.slab_ranges(ne[0], nparts, 2 ** first_rep)
    sims, mgs = [], []
    for (a, b) in ranges:
        s = capi.SlabSim(np.array(ne), np.zeros(3), np.array(dom), a, b, share_stream_with=sims[0] if sims else None)
        s.set_isotropic(1.0, 0.3)
        s.set_interp(0, 1.0, emin, 3.0, 3.0)
        s.apply_bc_file(bc)
        s.set_densities(s.window_of_elements(rho))
        sims.append(s)
        mgs.append(capi.SlabMG(s, levels, first_rep))
    # the windows of the load vector built per part must equal the slices of the global one (forces split over the GLOBAL box)
    for s in sims:
        np.testing.assert_array_equal(s.build_load(), s.window_of_nodal(f_global))
    grp = capi.SlabGroup(mgs)
    xs, bs = [], []
    for s in sims:
        n = s.num_nodes * 3
        x, b = capi.DeviceArray(n), capi.DeviceArray(n)
        b.upload(capi.to_soa(s.window_of_nodal(f_global)))
        xs.append(x); bs.append(b)
    it, res = grp.pcg_dev(xs, bs, **pcg)
    nn = np.array(ne) + 1
    u = np.full(tuple(nn) + (3,), np.nan)
    for s, x in zip(sims, xs):
        w = capi.from_soa(x.download(), 3).reshape((s.plane_hi - s.plane_lo + 1,) + tuple(nn[1:]) + (3,))
        u[s.own_lo:s.own_hi + 1] = w[s.own_lo - s.plane_lo:s.own_hi - s.plane_lo + 1]
    grp.close()
    return u.reshape(-1, 3), it, res


@pytest.mark.parametrize("ne,levels,first_rep,nparts", [
    ((16, 8, 8), 2, 1, 2),      # level 0 windowed; levels 1, 2 replicated
    ((16, 8, 8), 2, 2, 2),      # levels 0, 1 windowed; coarsest replicated
    ((32, 8, 8), 2, 2, 4),      # four slabs (interior parts have two neighbours)
    ((48, 16, 8), 3, 2, 3),     # uneven widths, a replicated level below the first replicated one
    ((32, 16, 16), 3, 3, 2),    # three windowed levels
])
@pytest.mark.parametrize("fmg", [True, False])
def test_slab_pcg_matches_undivided_solve(capi, data_dir, ne, levels, first_rep, nparts, fmg):
    dom = (ne[0] / 8.0, ne[1] / 8.0, ne[2] / 8.0)
    bc = os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc")
    rho = np.random.default_rng(3).uniform(0.2, 1.0, int(np.prod(ne)))
    pcg = dict(max_iter=60, tol=1e-10, mg_iterations=1, mg_smoothing=1, fmg=fmg)
    s, mg, f = build_reference(capi, ne, dom, bc, levels, rho, 1e-4)
    u_ref, it_ref, res_ref = mg.pcg(np.zeros_like(f), f, pcg["max_iter"], pcg["tol"], 1, 1, fmg)
    u, it, res = solve_slabs(capi, ne, dom, bc, levels, first_rep, nparts, rho, 1e-4, f, pcg)
    assert not np.isnan(u).any()
    # same algorithm, same colour order: only the summation order of the dot products / stencil completion differs, and the
    # residual-emitting sweep adds its <= 26 contributions per node in no fixed order (red.add): the residual norms agree to
    # rounding, which a history falling by ten orders of magnitude amplifies to ~1e-6 relative in its last entries
    assert it == it_ref
    np.testing.assert_allclose(res, res_ref, rtol=1e-4)
    assert np.linalg.norm(u - u_ref) <= 1e-9 * np.linalg.norm(u_ref)


def test_slab_mbb_partial_dirichlet(capi, data_dir):
    """MBB rollers (dirichlety / dirichletyz) exercise point Gauss-Seidel on partially constrained nodes and mask coarsening."""
    ne, dom, levels = (32, 16, 8), (4.0, 2.0, 1.0), 2
    bc = os.path.join(data_dir, "bcs", "3D", "mbb_N.bc")
    rho = np.full(int(np.prod(ne)), 0.5)
    pcg = dict(max_iter=80, tol=1e-9, mg_iterations=1, mg_smoothing=2, fmg=True)
    s, mg, f = build_reference(capi, ne, dom, bc, levels, rho, 1e-5)
    u_ref, it_ref, _ = mg.pcg(np.zeros_like(f), f, pcg["max_iter"], pcg["tol"], 1, 2, True)
    u, it, _ = solve_slabs(capi, ne, dom, bc, levels, 2, 2, rho, 1e-5, f, pcg)
    assert it == it_ref
    assert np.linalg.norm(u - u_ref) <= 1e-9 * np.linalg.norm(u_ref)


@pytest.mark.parametrize("ne,levels,first_rep,nparts,radius", [((32, 8, 8), 2, 2, 2, 2), ((48, 16, 8), 2, 1, 3, 3), ((32, 16, 16), 3, 2, 4, 1)])
def test_slab_topopt_matches_undivided_problem(capi, data_dir, ne, levels, first_rep, nparts, radius):
    """Config-4 style loop (filters + partitioned MG-PCG compliance + volume constraint + OC) on a local slab group against the
    undivided device problem: same compliance, sensitivities, bisection path and design variables, iteration by iteration."""
    dom = (ne[0] / 8.0, ne[1] / 8.0, ne[2] / 8.0)
    bc = os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc")
    V = 0.3
    filters = [("smooth", radius, 1), ("project", 1.0)]
    x0 = np.full(int(np.prod(ne)), 0.5 + np.arctanh((2 * V - 1) * np.tanh(0.5)))
    x0 = x0 * (1 + 0.05 * np.sin(np.arange(x0.size)))                    # break the uniformity so that halos matter
    s, mg, _ = build_reference(capi, ne, dom, bc, levels, np.ones(int(np.prod(ne))), 1e-4)
    ref = capi.Problem(mg, filters, V)
    ref.set_solver(200, 1e-12, 1, 2, True, False)      # tight solves: what remains is the difference of the two code paths
    ref.set_vars(x0)

    sims, mgs = [], []
    for (a, b) in capi.slab_ranges(ne[0], nparts, 2 ** first_rep):
        ps = capi.SlabSim(np.array(ne), np.zeros(3), np.array(dom), a, b, share_stream_with=sims[0] if sims else None)
        ps.set_isotropic(1.0, 0.3); ps.set_interp(0, 1.0, 1e-4, 3.0, 3.0); ps.apply_bc_file(bc)
        ps.set_densities(ps.window_of_elements(np.ones(int(np.prod(ne)))))
        sims.append(ps); mgs.append(capi.SlabMG(ps, levels, first_rep))
    grp = capi.SlabGroup(mgs)
    top = capi.SlabProblem(list(zip(sims, mgs)), grp, filters, V)
    top.set_solver(200, 1e-12, 1, 2, True, False)
    top.set_vars(x0)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    for it in range(3):
        assert rel(top.physical_vars(), ref.physical_vars()) < (1e-13 if it == 0 else 1e-6)
        assert abs(top.compliance() - ref.compliance()) < 1e-8 * abs(ref.compliance()), it
        assert abs(top.constraint() - ref.constraint()) < 1e-12
        assert rel(top.objective_gradient(), ref.objective_gradient()) < 1e-7
        assert rel(top.constraint_jacobian(), ref.constraint_jacobian()) < 1e-12
        na, nb = top.oc_step(), ref.oc_step()
        assert na == nb and abs(top.last_pcg_iters - ref.last_pcg_iters()) <= 2
        assert rel(top.design_vars(), ref.design_vars()) < 1e-6
    grp.close()


@pytest.mark.parametrize("ne,levels,first_rep,nparts,method,inc", [((16, 8, 4), 2, 1, 2, "N=3", 1), ((24, 8, 8), 2, 1, 3, "fd", 1), ((16, 8, 4), 2, 2, 2, "N=2", 2), ((16, 8, 4), 2, 1, 4, "zero", 1)])
def test_slab_layer_by_layer_matches_oracle(capi, ne, levels, first_rep, nparts, method, inc):
    """The layer-by-layer evaluator on a local slab group (vf_group_lbl_*: every slab holds a piece of every layer) against the
    oracle's undivided LayerByLayerEvaluator (LayerByLayer.hh:223-296): layer schedule bit-exact, per-layer compliances, objective and
    gradient within 1e-8 with solves converged to 1e-11, iteration counts side by side."""
    from oracle import OracleLBL, OracleMG, OracleSim
    ne = np.array(ne)
    dom = ne.astype(float) / ne[0]
    rho = np.random.default_rng(3).uniform(0.3, 1.0, int(np.prod(ne)))
    g = np.array([0.0, -1.0, 0.0])

    def prep(s, r):
        s.set_isotropic(1.0, 0.3); s.set_interp(1, 1.0, 1e-4, 3.0, 3.0)
        s.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7)          # clamp the build plate y = 0
        s.set_gravity(g); s.set_densities(r)
    o = OracleSim(ne, np.zeros(3), dom); prep(o, rho)
    oe = OracleLBL(OracleMG(o, levels)); oe.select_init_method(method)
    oi, oc = oe.run(True, inc, 200, 1e-11, 1, 1, False)
    sims, mgs = [], []
    for (a, b) in capi.slab_ranges(int(ne[0]), nparts, 2 ** first_rep):
        ps = capi.SlabSim(ne, np.zeros(3), dom, a, b, share_stream_with=sims[0] if sims else None)
        prep(ps, ps.window_of_elements(rho))
        sims.append(ps); mgs.append(capi.SlabMG(ps, levels, first_rep))
    grp = capi.SlabGroup(mgs)
    ge = capi.SlabLBL(grp, ne); ge.select_init_method(method)
    layers = []
    gi, gc = ge.run(True, inc, 200, 1e-11, 1, 1, False, callback=lambda l, c, it: layers.append(l))
    assert layers == list(range(int(ne[1]), 0, -inc))
    assert len(gi) == len(oi) and np.abs(gi.astype(int) - oi.astype(int)).max() <= 1, (gi, oi)
    assert np.abs(gc - oc).max() < 1e-8 * np.abs(oc).max()
    assert abs(ge.objective() - oe.objective()) < 1e-8 * abs(oe.objective())
    assert np.abs(ge.gradient() - oe.gradient()).max() < 1e-8 * np.abs(oe.gradient()).max()
    gi2, gc2 = ge.run(False, inc, 200, 1e-11, 1, 1, False)                         # warm start from the full-shape solution of the first run
    assert np.abs(gc2 - oc).max() < 1e-8 * np.abs(oc).max() and gi2[0] <= 1
    ge.close(); grp.close()
