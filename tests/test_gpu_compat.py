"""The reference's Python workflows, written against the reference's own names (pyVoxelFEM / pyOptimizer), run on the
GPU through voxelfem_b200/compat and are checked against the CPU oracle:
  * python/CoarseningLevelBenchmark.py:76-100 (single MG-PCG solve with it_callback),
  * python/3DTopoptDemo.ipynb cells 1,5 (filters + MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer),
  * python/LayerByLayerObjective.py:75-124 (getIntermediateFabricationShape + LayerByLayerEvaluator)."""
import os
import sys

import numpy as np
import pytest

from oracle import OracleLBL, OracleMG, OracleProblem, OracleSim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER_MODULES = ("pyVoxelFEM", "pyOptimizer", "benchmark", "parallelism", "MeshFEM", "mesh")


@pytest.fixture(scope="module", params=["pybind11", "python-shim"])
def vf(request):
    """pyVoxelFEM in both flavours: the pybind11 extension built from host/VoxelFEM.hh (voxelfem_b200/pybind) and the pure-Python
    module over the ctypes layer (voxelfem_b200/compat).  Whichever directory is on sys.path provides `pyVoxelFEM`, `pyOptimizer`,
    `benchmark`, ... under the names the reference's drivers import."""
    d = os.path.join(ROOT, "voxelfem_b200", "pybind" if request.param == "pybind11" else "compat")
    saved = {k: sys.modules.pop(k) for k in DRIVER_MODULES if k in sys.modules}
    sys.path.insert(0, d)
    import importlib
    m = importlib.import_module("pyVoxelFEM")
    assert os.path.dirname(os.path.abspath(m.__file__)) == d
    yield m
    sys.path.remove(d)
    for k in DRIVER_MODULES:
        sys.modules.pop(k, None)
    sys.modules.update(saved)


def rel_l2(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("dim,grid,levels", [(2, [64, 32], 2), (3, [32, 16, 16], 2)])
def test_coarsening_level_benchmark_flow(vf, data_dir, dim, grid, levels):
    bc = os.path.join(data_dir, "bcs", "3D" if dim == 3 else "", "cantilever_flexion_E.bc")
    corners = [[0, 0], [2, 1]] if dim == 2 else [[0, 0, 0], [2, 1, 1]]
    tps = vf.TensorProductSimulator([1] * dim, corners, grid)
    tps.readMaterial(os.path.join(data_dir, "materials", "B9Creator.material"))
    tps.applyDisplacementsAndLoadsFromFile(bc)
    tps.E_min = 1e-5
    tps.setUniformDensities(0.5)
    assert list(tps.NbElementsPerDimension) == grid and tps.numNodes() == int(np.prod(np.array(grid) + 1))
    f = tps.buildLoadVector()
    mg = tps.multigridSolver(levels)
    seen = []

    def it_callback(it, u_curr, residual):
        assert u_curr.shape == f.shape and residual.shape == f.shape
        seen.append((it, np.linalg.norm(residual), u_curr.copy()))
    u = mg.preconditionedConjugateGradient(np.zeros_like(f), f, maxIter=100, tol=1e-10, it_callback=it_callback, mgIterations=1,
                                           mgSmoothingIterations=1, fullMultigrid=True)
    o = OracleSim(np.array(grid), np.zeros(dim), np.array(corners[1], dtype=float))
    o.set_isotropic(1.0, 0.3); o.set_interp(0, 1.0, 1e-5, 3.0, 3.0); o.apply_bc_file(bc); o.set_uniform_density(0.5)
    uo, ito, reso = OracleMG(o, levels).pcg(np.zeros_like(f), f, 100, 1e-10, 1, 1, True)
    assert abs(len(seen) - ito) <= 1 and [s[0] for s in seen] == list(range(1, len(seen) + 1))
    assert rel_l2(u, uo) < 1e-6
    k = min(len(seen), ito)
    assert np.allclose([s[1] for s in seen][:k], reso[:k], rtol=1e-4)
    assert rel_l2(seen[-1][2], u) < 1e-14                         # the callback's iterate of the last iteration is the result
    r = mg.computeResidual(0, u, f)
    assert np.linalg.norm(r) <= 1e-10 * np.linalg.norm(f) * 1.01
    assert np.all(mg.zeroOutDirichletComponents(0, np.ones_like(f))[tps.getDirichletMask()] == 0)
    assert rel_l2(tps.applyK(u), mg.applyK(0, u)) < 1e-13


def test_topopt_demo_flow(vf, data_dir):
    grid, V = [16, 8, 8], 0.3
    bc = os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc")
    tps = vf.TensorProductSimulator([1, 1, 1], [[0, 0, 0], [2, 1, 1]], grid)
    tps.readMaterial(os.path.join(data_dir, "materials", "B9Creator.material"))
    tps.applyDisplacementsAndLoadsFromFile(bc)
    pf = vf.ProjectionFilter(); pf.beta = 1
    filters = [vf.SmoothingFilter(2, vf.SmoothingFilter.Type.Linear), pf]
    objective = vf.MultigridComplianceObjective(tps.multigridSolver(2))
    objective.tol = 1e-9
    top = vf.TopologyOptimizationProblem(tps, objective, [vf.TotalVolumeConstraint(V)], filters)
    x0 = pf.invert(V) * np.ones(tps.numElements())
    top.setVars(x0)
    oc = vf.OCOptimizer(top)

    o = OracleSim(np.array(grid), np.zeros(3), np.array([2.0, 1.0, 1.0]))
    o.set_isotropic(1.0, 0.3); o.set_interp(0, 1.0, 1e-4, 3.0, 3.0); o.apply_bc_file(bc); o.set_uniform_density(1.0)
    op = OracleProblem(OracleMG(o, 2), [("smooth", 2, 1), ("project", 1.0)], V)
    op.set_solver(100, 1e-9, 1, 2, True, False); op.set_vars(x0)
    for it in range(3):
        assert abs(top.evaluateObjective() - op.compliance()) < 1e-8 * abs(op.compliance())
        assert abs(top.evaluateConstraints()[0] - op.constraint()) < 1e-9      # after an OC step both sit within ctol = 1e-6 of zero
        g = top.evaluateObjectiveGradient()
        assert np.abs(g - op.objective_gradient()).max() < 1e-7 * np.abs(g).max()
        assert top.evaluateConstraintsJacobian().shape == (1, top.numVars())
        oc.step(); op.oc_step()
        assert np.abs(top.getVars() - op.design_vars()).max() < 1e-6
    assert np.abs(top.getDensities() - op.physical_vars()).max() < 1e-6
    assert np.abs(tps.getDensities() - top.getDensities()).max() == 0           # the simulator carries the physical densities
    # filterChain view: backprop of a physical-space gradient equals the problem's own chain rule
    dJ = objective.gradient()
    assert np.abs(top.filterChain.backprop(dJ) - top.evaluateObjectiveGradient()).max() < 1e-12 * np.abs(dJ).max()
    assert abs(objective.compliance() - 0.5 * float((objective.f() * objective.u()).sum())) < 1e-12 * abs(objective.compliance())


def test_layer_by_layer_flow(vf):
    ne = [8, 8, 4]
    rho = np.random.default_rng(3).uniform(0.3, 1.0, int(np.prod(ne)))
    tps = vf.TensorProductSimulator([1, 1, 1], [[0, 0, 0], [1.0, 1.0, 0.5]], ne)
    tps.setDensities(rho)
    lbl_sim = tps.getIntermediateFabricationShape(1, False, vf.InterpolationLaw.RAMP)
    lbl_sim.q = 3
    assert np.allclose(lbl_sim.gravity, [0, -1, 0]) and lbl_sim.interpolationLaw == vf.InterpolationLaw.RAMP
    mask = lbl_sim.getDirichletMask().reshape(9, 9, 5, 3)
    assert mask[:, 0].all() and not mask[:, 1:].any()                           # build platform clamped, nothing else
    mg = lbl_sim.multigridSolver(1)
    ev = vf.LayerByLayerEvaluator(lbl_sim)
    ev.selectInitMethod("N=3")
    layers, pcg_its = [], []
    # the reference's callback shapes: lblCallback(l, compliance, grad_compliance, u) (LayerByLayer.hh:222, 277-279) and the PCG
    # it_callback(it, x, r) handed to every layer's solve (:265)
    def lbl_cb(layer, c, grad, u):
        assert grad.shape == (lbl_sim.numElements(),) and u.shape == (lbl_sim.numNodes(), 3)
        assert np.abs(grad - lbl_sim.complianceGradient(u)).max() <= 1e-12 * np.abs(grad).max()
        layers.append((layer, c, 0.5 * float((u * 0).sum())))
    def it_cb(it, x, r):
        assert x.shape == r.shape == (lbl_sim.numNodes(), 3)
        pcg_its.append(it)
    assert ev.run(mg, True, 1, maxIter=50, tol=1e-8, it_callback=it_cb, mgIterations=1, mgSmoothingIterations=1, fullMultigrid=False,
                  lblCallback=lbl_cb) is None
    assert len(pcg_its) >= 8 and pcg_its[0] == 1
    o = OracleSim(np.array(ne), np.zeros(3), np.array([1.0, 1.0, 0.5]))
    o.set_isotropic(1.0, 0.0); o.set_interp(1, 1.0, 1e-4, 3.0, 3.0)    # default material ETensor(1, 0)
    o.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7); o.set_gravity(np.array([0, -1.0, 0])); o.set_densities(rho)
    oe = OracleLBL(OracleMG(o, 1)); oe.select_init_method("N=3")
    its, comps = oe.run(True, 1, 50, 1e-8, 1, 1, False)
    assert len(layers) == len(its) == 8
    assert np.allclose([c for _, c, _ in layers], comps, rtol=1e-6)
    assert abs(ev.objective() - oe.objective()) < 1e-6 * abs(oe.objective())
    assert np.abs(ev.gradient() - oe.gradient()).max() < 1e-5 * np.abs(oe.gradient()).max()
    # downsampling helpers (TensorProductSimulator.hh:1926-1992)
    coarse = tps.downsample(1)
    tps.downsampleDensityFieldTo(rho, coarse)
    assert np.allclose(coarse.getDensities().reshape(4, 4, 2), rho.reshape(4, 2, 4, 2, 2, 2).mean(axis=(1, 3, 5)))
    gc = np.arange(coarse.numElements(), dtype=float)
    assert abs(tps.upsampleDensityGradientFrom(coarse, gc).sum() - gc.sum()) < 1e-12 * gc.sum()


def test_mma_module(vf):
    import pyOptimizer
    assert os.path.dirname(os.path.abspath(pyOptimizer.__file__)) == os.path.dirname(os.path.abspath(vf.__file__))
    from mma_problems import svanberg_toy
    n, m, lo, hi, f, df, x0, xstar = svanberg_toy()
    opt = pyOptimizer.MMA(n, m, lo, hi, f, df)
    opt.setInitialVar(x0)
    for _ in range(12):
        opt.step()
    assert np.abs(opt.getOptimalVar() - xstar).max() < 2e-6


def _lbl_optimization_problem_class(vf, benchmark, tps):
    """python/LayerByLayerOptimization.py:54-79 (getClass) restated: a python subclass of the problem class named by getClassName."""
    baseClass = eval(vf.getClassName(tps, "TopologyOptimizationProblem"), {"pyVoxelFEM": vf})

    class LayerByLayerOptimizationProblem(baseClass):
        def __init__(self, tps, optObj, constraints, filters, layObj, weight):
            super().__init__(tps, optObj, constraints, filters)
            self.optObj, self.layObj, self.weight, self.skipLayers, self.opt_obj, self.lbl_obj = optObj, layObj, weight, 1, 0, 0
        def setVars(self, x, verbose=False):
            benchmark.start_timer_section('Optimization setVars')
            super().setVars(x)
            benchmark.stop_timer_section('Optimization setVars')
            self.layObj.setVars(self.getDensities(), self.skipLayers, verbose)
        def evaluateObjective(self):
            self.opt_obj = super().evaluateObjective()
            self.lbl_obj = self.layObj.energy()
            return self.opt_obj + self.weight * self.lbl_obj
        def evaluateObjectiveGradient(self):
            return super().evaluateObjectiveGradient() + self.weight * self.filterChain.backprop(self.layObj.gradient())
    return LayerByLayerOptimizationProblem


class _LayObj:
    """python/LayerByLayerObjective.py:75-139 restated (no downsampling, no symmetry): the layer-by-layer objective of a design."""
    def __init__(self, vf, tps, levels, pcgVars):
        self.vf, self.tps, self.pcgVars = vf, tps, pcgVars
        self.sim = tps.getIntermediateFabricationShape(1, False, vf.InterpolationLaw.RAMP)
        self.sim.q = 3
        self.mg = self.sim.multigridSolver(levels)
        self.ev = None
    def setVars(self, x, skipLayers=1, verbose=False):
        self.tps.downsampleDensityFieldTo(x, self.sim)
        if self.ev is None: self.ev = self.vf.LayerByLayerEvaluator(self.sim)
        self.ev.selectInitMethod("N=3")
        self.ev.run(self.mg, True, skipLayers, **self.pcgVars, verbose=verbose, lblCallback=None)
    def energy(self): return self.ev.objective()
    def gradient(self): return self.ev.gradient()


def test_layer_by_layer_optimization_flow(vf, data_dir):
    """python/LayerByLayerOptimization.py:56-79, 180-200 + its MMA / OC drivers (:13-38): a python-SUBCLASSED problem
    (trampoline) whose objective adds the layer-by-layer energy, driven by pyOptimizer.MMA and by OCOptimizer.step."""
    import benchmark
    import pyOptimizer
    benchmark.reset()
    grid, maxVolume, weight = [16, 8], 0.6, 0.1
    filters = [vf.SmoothingFilter(radius=2, type=vf.SmoothingFilter.Type.Linear), vf.ProjectionFilter(beta=5)]
    constraints = [vf.TotalVolumeConstraint(maxVolume)]
    uniformDensity = filters[-1].invert(maxVolume)
    tps = vf.TensorProductSimulator([1, 1], [[0, 0], [2, 1]], grid)
    tps.setDensities(np.ones(int(np.prod(grid))) * uniformDensity)
    tps.readMaterial(os.path.join(data_dir, "materials", "B9Creator.material"))
    tps.applyDisplacementsAndLoadsFromFile(os.path.join(data_dir, "bcs", "mbb_N.bc"))
    optObj = vf.MultigridComplianceObjective(tps.multigridSolver(2))
    optObj.mgSmoothingIterations = 1
    optObj.tol = 1e-9
    cg = {"opt": 0, "lay": 0}
    optObj.residual_cb = lambda i, r: cg.__setitem__("opt", cg["opt"] + 1)
    pcgVars = {'maxIter': 50, 'tol': 1e-9, 'mgIterations': 1, 'mgSmoothingIterations': 1, 'fullMultigrid': False,
               'it_callback': lambda i, _, r: cg.__setitem__("lay", cg["lay"] + 1)}
    layObj = _LayObj(vf, tps, 2, pcgVars)
    Problem = _lbl_optimization_problem_class(vf, benchmark, tps)
    top = Problem(tps, optObj, constraints, filters, layObj, weight)
    x0 = tps.getDensities()
    top.setVars(x0)
    assert cg["opt"] > 0 and cg["lay"] > 0
    J = top.evaluateObjective()
    assert abs(J - (top.opt_obj + weight * top.lbl_obj)) < 1e-14 * abs(J) and top.lbl_obj > 0
    # the subclass's gradient: finite differences of the combined objective along a random direction
    g = top.evaluateObjectiveGradient()
    d = np.random.default_rng(0).normal(size=x0.size); h = 1e-5
    top.setVars(x0 + h * d); Jp = top.evaluateObjective()
    top.setVars(x0 - h * d); Jm = top.evaluateObjective()
    assert abs((Jp - Jm) / (2 * h) - float(g @ d)) < 2e-5 * abs(float(g @ d))
    top.setVars(x0)
    # MMA exactly as python/LayerByLayerOptimization.py:13-38 wires it
    n = top.numVars()
    def gradients(x): return np.stack([top.evaluateObjectiveGradient(), -top.evaluateConstraintsJacobian()[0]])
    def objAndConstr(x):
        top.setVars(x)
        return np.stack([top.evaluateObjective(), -top.evaluateConstraints()[0]])
    mma = pyOptimizer.MMA(n, 1, np.zeros(n), np.ones(n), objAndConstr, gradients)
    mma.setInitialVar(x0)
    J0 = objAndConstr(x0)[0]
    for _ in range(4): mma.step()
    assert top.evaluateObjective() < J0 and -top.evaluateConstraints()[0] <= 1e-6       # descends, volume constraint respected
    # OC on the subclassed problem: gradient from the override, search on the device, result through the override's setVars
    top.setVars(x0)
    before = dict(cg)
    oc = vf.OCOptimizer(top)
    oc.step()
    assert cg["lay"] > before["lay"]                                    # the override's setVars ran the layer-by-layer simulation
    assert abs(top.evaluateConstraints()[0]) <= 1e-6 and top.evaluateObjective() < J0
    oc.step(inplace=False)
    assert abs(top.evaluateConstraints()[0]) <= 1e-6
    rep = benchmark.report
    rep()


@pytest.mark.parametrize("dim,grid,corner", [(2, [8, 6], [4.0, 3.0]), (3, [6, 4, 4], [3.0, 2.0, 2.0])])
def test_export_and_postprocessing_methods(vf, data_dir, dim, grid, corner, tmp_path):
    """getK, getMesh, sampleNodalField, getDirichletVarsAndValues, getForceMask, getBCIndicatorField,
    constantStrainLoad, solveWithImposedLoads, debugMulticolorElementVisit, the intermediate-shape transfers and .msh output --
    checked against the operators of the solve path and against closed forms."""
    bc = os.path.join(data_dir, "bcs", "3D" if dim == 3 else "", "cantilever_flexion_E.bc")
    tps = vf.TensorProductSimulator([1] * dim, [[0] * dim, corner], grid)
    tps.readMaterial(os.path.join(data_dir, "materials", "B9Creator.material"))
    tps.applyDisplacementsAndLoadsFromFile(bc)
    tps.E_min = 0.0; tps.gamma = 1.0                                      # E(rho) = rho: constantStrainLoad scales by the DENSITY
    rng = np.random.default_rng(11)
    rho = rng.uniform(0.2, 1.0, tps.numElements())
    tps.setDensities(rho)
    nn = tps.numNodes()
    # getK: the upper triangle of the assembled matrix reproduces the matrix-free operator
    K = tps.getK()
    Kfull = K + K.T - __import__("scipy.sparse", fromlist=["diags"]).diags(K.diagonal())
    u = rng.normal(size=(nn, dim))
    assert rel_l2((Kfull @ u.ravel()).reshape(nn, dim), tps.applyK(u)) < 1e-13
    assert (K - __import__("scipy.sparse", fromlist=["triu"]).triu(K)).nnz == 0
    # Dirichlet / force bookkeeping
    mask = tps.getDirichletMask()
    dvars, dvals = tps.getDirichletVarsAndValues()
    assert sorted(dvars) == sorted(np.flatnonzero(mask.ravel()).tolist()) and all(v == 0 for v in dvals)
    f = tps.buildLoadVector()
    assert np.array_equal(np.asarray(tps.getForceMask()), f != 0)
    ind = np.asarray(tps.getBCIndicatorField())
    bits = (mask * (1 << np.arange(dim))).sum(axis=1) + (1 << dim) * ((f != 0) * (1 << np.arange(dim))).sum(axis=1)
    assert np.array_equal(ind, bits.astype(float))
    # mesh + sampling: nodal values at the vertices, the mean of the corner values at element centroids, clamping outside
    V, F = tps.getMesh()
    assert V.shape == (nn, 3) and F.shape == (tps.numElements(), 2 ** dim)
    assert np.allclose(V[:, :dim], np.array([tps.nodePosition(i) for i in range(nn)]))
    assert rel_l2(np.asarray(tps.sampleNodalField(u, V[:, :dim])), u) < 1e-13
    cent = V[F].mean(axis=1)[:, :dim]
    assert rel_l2(np.asarray(tps.sampleNodalField(u, cent)), u[F].mean(axis=1)) < 1e-13
    far = np.array([[-5.0] * dim, [1e3] * dim])
    assert np.allclose(np.asarray(tps.sampleNodalField(u, far)), u[[0, nn - 1]])
    # constant-strain load: multilinear elements reproduce u(x) = eps x, so the load is K u_lin when E(rho) = rho
    eps = rng.normal(size=(dim, dim)); eps = 0.5 * (eps + eps.T)
    assert rel_l2(np.asarray(tps.constantStrainLoad(eps)), tps.applyK(V[:, :dim] @ eps.T)) < 1e-12
    flat = [eps[i, i] for i in range(dim)] + ([eps[0, 1]] if dim == 2 else [eps[1, 2], eps[0, 2], eps[0, 1]])
    assert rel_l2(np.asarray(tps.constantStrainLoad(flat)), np.asarray(tps.constantStrainLoad(eps))) < 1e-15
    # direct solve with the imposed loads
    us = np.asarray(tps.solveWithImposedLoads())
    assert rel_l2(us, tps.solve(f)) < 1e-12 and np.all(us[mask] == 0)
    # multicoloured element visit: a permutation, colour by colour (2^N parity classes, axis 0 outermost)
    rank = np.asarray(tps.debugMulticolorElementVisit()).astype(int).reshape(grid)
    assert sorted(rank.ravel().tolist()) == list(range(tps.numElements()))
    par = lambda r: tuple(int(v) % 2 for v in np.unravel_index(int(np.flatnonzero(rank.ravel() == r)[0]), grid))
    cols = [par(r) for r in range(tps.numElements())]
    assert cols == sorted(cols)
    # intermediate fabrication shape transfers
    tps2 = vf.TensorProductSimulator([1] * dim, [[0] * dim, corner], grid)
    tps2.readMaterial(os.path.join(data_dir, "materials", "B9Creator.material")); tps2.setDensities(rho)
    g = np.zeros(dim); g[1] = -1.0; tps2.gravity = g
    lo = [-1e-9] * dim; hi = [c + 1e-9 for c in corner]; hi[1] = 1e-9
    tps2.addDirichletCondition([0.0] * dim, lo, hi, "xyz"[:dim])
    inter = tps2.getIntermediateFabricationShape(0.5)
    ui = np.asarray(tps2.transferVFieldToIntermediateFabricationShape(inter, u))
    shp = tuple(np.array(grid) + 1)
    assert np.array_equal(ui, u.reshape(shp + (dim,))[:, :grid[1] // 2 + 1].reshape(-1, dim))
    acc = np.ones(tps2.numElements())
    gi = rng.normal(size=inter.numElements())
    tps2.accumElementScalarFieldFromIntermediateFabricationShape(inter, gi, acc)
    exp = np.ones(grid); exp[:, :grid[1] // 2] += gi.reshape(tuple(inter.NbElementsPerDimension))
    assert np.array_equal(acc, exp.ravel())
    # .msh output of the simulator's fields and reading it back
    sys.path.insert(0, ROOT)
    from voxelfem_b200.compat import msh
    path = str(tmp_path / "out.msh")
    msh.write_fields(tps, path, element_fields={"density": rho}, node_fields={"u": u})
    assert np.array_equal(msh.densities_from_msh(tps, path), rho)
    assert np.array_equal(msh.MSHFieldParser(path).vectorField("u"), u)
