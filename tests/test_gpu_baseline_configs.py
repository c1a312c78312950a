"""GPU-vs-oracle parity ON BASELINE.json's OWN CONFIGURATIONS (C1, C2, C3, C5 and a slab-partitioned solve).

Both sides run side by side on the named grids, boundary conditions, densities and solver parameters (SURVEY.md 8d "Synthetic
inputs"): the CUDA path through the C ABI, the CPU oracle (oracle/vfo.cpp) on the host cores.  Gates (BASELINE.json north_star):
converged displacement within 1e-6 relative L2, compliance and sensitivities within 1e-8 relative (both sides solved to the
tight tolerance for that check), PCG iteration counts printed side by side and within +-1.

The oracle costs seconds (C1, C5), tens of seconds (C2) and about a minute (C3, 256^3) on the GPU box's host cores.
"""
import os
import time

import numpy as np
import pytest

from oracle import OracleLBL, OracleMG, OracleProblem, OracleSim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from voxelfem_b200 import capi as c
    assert c.device_count() > 0
    return c


def rel_l2(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def build(Sim, MG, ne, dmax, bc, levels, data_dir, rho, emin, law=0):
    ne = np.array(ne)
    s = Sim(ne, np.zeros(len(ne)), np.array(dmax, dtype=float))
    s.set_isotropic(1.0, 0.3)                               # examples/materials/B9Creator.material
    s.set_interp(law, 1.0, emin, 3.0, 3.0)
    if bc is not None:
        s.apply_bc_file(os.path.join(data_dir, "bcs", bc))
    if np.isscalar(rho):
        s.set_uniform_density(rho)
    else:
        s.set_densities(rho)
    return s, (MG(s, levels) if levels else None)


def solve_both(capi, ne, dmax, bc, levels, data_dir, rho, emin, pcg):
    out = []
    for Sim, MG in ((capi.Sim, capi.MG), (OracleSim, OracleMG)):
        s, mg = build(Sim, MG, ne, dmax, bc, levels, data_dir, rho, emin)
        f = s.build_load()
        t0 = time.time()
        u, it, res = mg.pcg(np.zeros_like(f), f, **pcg)
        out.append(dict(u=u, it=it, res=np.asarray(res), f=f, seconds=time.time() - t0, grad=s.compliance_gradient(u)))
    return out


def check_solve(tag, g, o, tol):
    bn = np.linalg.norm(o["f"])
    cg, co = 0.5 * float((g["f"] * g["u"]).sum()), 0.5 * float((o["f"] * o["u"]).sum())
    print("\n%s: PCG iterations gpu=%d oracle=%d | final relative residual gpu=%.6e oracle=%.6e | rel-L2(u)=%.3e | compliance gpu=%.12e oracle=%.12e (rel %.2e) | "
          "sensitivities rel %.2e | seconds gpu=%.2f oracle=%.2f"
          % (tag, g["it"], o["it"], g["res"][-1] / bn, o["res"][-1] / bn, rel_l2(g["u"], o["u"]), cg, co, rel(cg, co), relmax(g["grad"], o["grad"]), g["seconds"], o["seconds"]))
    assert abs(g["it"] - o["it"]) <= 1                                           # iteration counts side by side
    assert g["res"][-1] <= tol * bn and o["res"][-1] <= tol * bn                 # both converged to the requested tolerance
    n = min(len(g["res"]), len(o["res"]))
    assert np.allclose(g["res"][:n - 1], o["res"][:n - 1], rtol=1e-4)            # the same residual history
    assert np.array_equal(g["f"], o["f"])
    assert rel_l2(g["u"], o["u"]) <= 1e-6                                        # converged displacement
    assert rel(cg, co) <= 1e-8                                                   # compliance
    assert relmax(g["grad"], o["grad"]) <= 1e-8                                  # sensitivities


PCG_BENCH = dict(max_iter=100, tol=1e-10, mg_iterations=1, mg_smoothing=1, fmg=True)   # python/CoarseningLevelBenchmark.py:22-29


def test_C1_mbb_256x128(capi, data_dir):
    """configs[0]: 2D MBB beam 256 x 128 Q1, one MG-PCG solve exactly as python/CoarseningLevelBenchmark.py:76-100 sets it up."""
    g, o = solve_both(capi, (256, 128), (2.0, 1.0), "mbb_N.bc", 3, data_dir, 0.5, 1e-5, PCG_BENCH)
    check_solve("C1 2D MBB 256x128, 3 levels", g, o, 1e-10)


def test_C3_cantilever_256_cubed(capi, data_dir):
    """configs[2]: one MG-PCG solve at 256^3 (50.9 M DOF) -- the bench.py workload -- against the oracle's solve of the same system."""
    g, o = solve_both(capi, (256, 256, 256), (1.0, 1.0, 1.0), "3D/cantilever_flexion_E.bc", 5, data_dir, 0.5, 1e-5, PCG_BENCH)
    check_solve("C3 3D cantilever 256^3, 5 levels", g, o, 1e-10)


def test_C3_heterogeneous_128_cubed(capi, data_dir):
    """SURVEY.md 8d's heterogeneous field rho = clip(smooth(rng(0).uniform)) on the C3 problem at 128^3 (the oracle needs seconds)."""
    ne = (128, 128, 128)
    raw = np.random.default_rng(0).uniform(0.0, 1.0, int(np.prod(ne)))
    rho = np.clip(capi.smoothing_filter(raw, ne, 2, 1), 0.0, 1.0)
    g, o = solve_both(capi, ne, (1.0, 1.0, 1.0), "3D/cantilever_flexion_E.bc", 4, data_dir, rho, 1e-5, PCG_BENCH)
    check_solve("C3 heterogeneous 128^3, 4 levels", g, o, 1e-10)


def test_C2_topopt_128x64x64_first_iterations(capi, data_dir):
    """configs[1]: 3D cantilever 128 x 64 x 64 SIMP compliance topopt, SmoothingFilter(3, Linear) + ProjectionFilter(1), volume
    fraction 0.3, OC -- the first 5 OC iterations side by side.  The solve tolerance is tightened from the objective's default
    1e-5 to 1e-10 so that compliance and sensitivities are comparable at 1e-8."""
    ne, vol = (128, 64, 64), 0.3
    filters = [("smooth", 3, 1), ("project", 1.0)]
    import math
    x0 = math.atanh((2 * vol - 1) * math.tanh(0.5)) + 0.5                          # ProjectionFilter(1).invert(vol)
    probs = []
    for Sim, MG, Problem in ((capi.Sim, capi.MG, capi.Problem), (OracleSim, OracleMG, OracleProblem)):
        s, mg = build(Sim, MG, ne, (2.0, 1.0, 1.0), "3D/cantilever_flexion_E.bc", 2, data_dir, vol, 1e-4)
        p = Problem(mg, filters, vol)
        p.set_solver(100, 1e-10, 1, 2, True, False)
        p.set_vars(np.full(int(np.prod(ne)), x0))
        probs.append(p)
    gp, op = probs
    for it in range(6):
        cg, co = gp.compliance(), op.compliance()
        dg, do = gp.objective_gradient(), op.objective_gradient()
        print("\nC2 OC iteration %d: compliance gpu=%.12e oracle=%.12e (rel %.2e) | sensitivities rel %.2e | PCG iterations gpu=%d oracle=%d | volume constraint gpu=%.3e oracle=%.3e"
              % (it, cg, co, rel(cg, co), relmax(dg, do), gp.last_pcg_iters(), op.last_pcg_iters(), gp.constraint(), op.constraint()))
        assert rel(cg, co) <= 1e-8 and relmax(dg, do) <= 1e-8
        assert abs(gp.last_pcg_iters() - op.last_pcg_iters()) <= 1
        assert rel_l2(gp.u(), op.u()) <= 1e-6
        assert np.abs(gp.physical_vars() - op.physical_vars()).max() <= 1e-9
        if it == 5:
            break
        ng, no = gp.oc_step(), op.oc_step()
        assert ng == no                                                        # the same bracket / bisection path


def test_C5_layer_by_layer_64_layers(capi):
    """configs[4] at 64 layers (grid [32, 64, 32], layers along axis 1): LayerByLayerObjective defaults (python/LayerByLayerObjective.py:19-20:
    RAMP q = 3, maxIter 50, tol 1e-5, one V-cycle, one smoothing step, no FMG, init N = 3), gravity (0, -1, 0), build plate clamped."""
    ne = (32, 64, 32)
    rho = np.clip(0.6 + 0.3 * np.random.default_rng(0).standard_normal(int(np.prod(ne))), 0.05, 1.0)
    runs = []
    for Sim, MG, LBL in ((capi.Sim, capi.MG, capi.LBL), (OracleSim, OracleMG, OracleLBL)):
        s = Sim(np.array(ne), np.zeros(3), np.array([1.0, 2.0, 1.0]))
        s.set_isotropic(1.0, 0.3); s.set_interp(1, 1.0, 1e-4, 3.0, 3.0)
        s.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7); s.set_gravity(np.array([0, -1.0, 0])); s.set_densities(rho)
        ev = LBL(MG(s, 3)); ev.select_init_method("N=3")
        its, comps = ev.run(True, 1, 50, 1e-5, 1, 1, False)
        runs.append((np.asarray(its), np.asarray(comps), ev.objective(), ev.gradient()))
    (ig, cg, og, gg), (io, co, oo, go) = runs
    print("\nC5 64 layers: total PCG iterations gpu=%d oracle=%d (per layer max difference %d) | objective gpu=%.12e oracle=%.12e (rel %.2e) | gradient rel %.2e"
          % (ig.sum(), io.sum(), np.abs(ig - io).max(), og, oo, rel(og, oo), relmax(gg, go)))
    assert len(ig) == len(io) == 64                                              # the layer schedule
    assert np.abs(ig - io).max() <= 1
    # every layer is solved to tol 1e-5 only (the schedule's own setting), so compliances agree to that, not to 1e-8
    assert np.allclose(cg, co, rtol=2e-5) and rel(og, oo) <= 2e-5 and relmax(gg, go) <= 1e-4


def test_C5_layer_by_layer_tight_tolerance(capi):
    """The same schedule solved to 1e-10 per layer: objective and gradient at the north_star tolerances."""
    ne = (16, 32, 16)
    rho = np.clip(0.6 + 0.3 * np.random.default_rng(1).standard_normal(int(np.prod(ne))), 0.05, 1.0)
    runs = []
    for Sim, MG, LBL in ((capi.Sim, capi.MG, capi.LBL), (OracleSim, OracleMG, OracleLBL)):
        s = Sim(np.array(ne), np.zeros(3), np.array([1.0, 2.0, 1.0]))
        s.set_isotropic(1.0, 0.3); s.set_interp(1, 1.0, 1e-4, 3.0, 3.0)
        s.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7); s.set_gravity(np.array([0, -1.0, 0])); s.set_densities(rho)
        ev = LBL(MG(s, 2)); ev.select_init_method("N=3")
        its, comps = ev.run(True, 1, 100, 1e-10, 1, 1, False)
        runs.append((np.asarray(its), np.asarray(comps), ev.objective(), ev.gradient()))
    (ig, cg, og, gg), (io, co, oo, go) = runs
    print("\nC5 tight: iterations gpu=%d oracle=%d | objective rel %.2e | gradient rel %.2e" % (ig.sum(), io.sum(), rel(og, oo), relmax(gg, go)))
    assert np.abs(ig - io).max() <= 1 and np.allclose(cg, co, rtol=1e-8) and rel(og, oo) <= 1e-8 and relmax(gg, go) <= 1e-8


def test_slab_partitioned_solve_matches_oracle(capi, data_dir):
    """the solve partitioned into slabs along axis 0 (windows, ghost-plane exchanges, sub-assembled Galerkin stencils,
    replicated coarse levels -- the control flow the NCCL ranks run, here as a local group on one device) against the ORACLE's
    undivided solve.  The NCCL transport itself is covered by tests/test_gpu_nccl.py when two devices are visible."""
    from test_gpu_slabs import solve_slabs
    ne, dom, levels, first_rep = (64, 32, 32), (2.0, 1.0, 1.0), 3, 2
    rho = np.random.default_rng(5).uniform(0.2, 1.0, int(np.prod(ne)))
    bc = os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc")
    so, mo = build(OracleSim, OracleMG, ne, dom, "3D/cantilever_flexion_E.bc", levels, data_dir, rho, 1e-4)
    f = so.build_load()
    uo, ito, reso = mo.pcg(np.zeros_like(f), f, **PCG_BENCH)
    for nparts in (2, 4):
        u, it, res = solve_slabs(capi, ne, dom, bc, levels, first_rep, nparts, rho, 1e-4, f, PCG_BENCH)
        co, cg = 0.5 * float((f * uo).sum()), 0.5 * float((f * u.reshape(uo.shape)).sum())
        print("\nslabs=%d: PCG iterations gpu=%d oracle=%d | rel-L2(u)=%.3e | compliance rel %.2e" % (nparts, it, ito, rel_l2(u.reshape(uo.shape), uo), rel(cg, co)))
        assert abs(it - ito) <= 1 and rel_l2(u.reshape(uo.shape), uo) <= 1e-6 and rel(cg, co) <= 1e-8
