"""pyVoxelFEM -- the reference's Python module surface (python_bindings/VoxelFEM.cc:400-432) over libvoxelfem_b200.

Put this directory on PYTHONPATH and the reference's drivers (python/CoarseningLevelBenchmark.py,
python/LayerByLayerObjective.py, the 3D topopt notebook) import it under the name they expect.  Same class, method,
argument and default names as the pybind11 module; array conventions as there: nodal fields are (numNodes, N)
float64 copies in and out, densities flat over (ex, ey[, ez]) row-major (VoxelFEM.cc:76-195).

Everything here is host glue: each call lands in one C-ABI entry point of include/voxelfem_b200.h, which launches CUDA
kernels.  There is no CPU path; without a usable GPU the first constructor raises.  Methods of the reference that lie
outside the MG-PCG / topopt hot path ("out of scope") raise NotImplementedError by name.
"""
import ctypes as C
import enum
import json
import math

import os
import sys
import weakref

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
from voxelfem_b200 import capi  # noqa: E402
from voxelfem_b200.compat import tps_extras  # noqa: E402


class InterpolationLaw(enum.IntEnum):          # VoxelFEM.cc:407-410
    SIMP = 0
    RAMP = 1


SIMP, RAMP = InterpolationLaw.SIMP, InterpolationLaw.RAMP   # export_values()


class NumberType(enum.IntEnum):                # VoxelFEM.cc:415-418
    DOUBLE = 0
    FLOAT = 1


def _unsupported(name):
    def f(self, *a, **k):
        raise NotImplementedError("pyVoxelFEM.%s.%s is outside the B200 hot path (SURVEY.md section 8: out of scope)" % (type(self).__name__, name))
    f.__name__ = name
    return f


class _ETensor:
    """Minimal stand-in for MeshFEM's ElasticityTensor (ElasticityTensor.hh:100-131): the flattened 6x6 / 3x3 matrix."""

    def __init__(self, N, D=None):
        self.N = N
        self.D = np.zeros((6, 6)) if D is None else np.array(D, dtype=np.float64)
        self.isotropic = None

    def setIsotropic(self, E, nu):
        self.isotropic = (float(E), float(nu))


class _TPS:
    """detail.TensorProductSimulator1_1[_1] (VoxelFEM.cc:76-154)."""
    BUILD_DIRECTION = 1

    def __init__(self, domainBBox, elementsPerDimension):
        ne = np.asarray(elementsPerDimension, dtype=np.int64)
        self._dmin = np.asarray(domainBBox[0], dtype=np.float64).copy()
        self._dmax = np.asarray(domainBBox[1], dtype=np.float64).copy()
        self._s = capi.Sim(ne, self._dmin, self._dmax)
        self._N = len(ne)
        self._interp = dict(law=int(SIMP), E0=1.0, Emin=1e-4, gamma=3.0, q=3.0)   # TensorProductSimulator.hh:2160-2166 defaults
        self._gravity = np.zeros(self._N)
        self._et = _ETensor(self._N)
        self._et.setIsotropic(1.0, 0.0)       # ETensor(1, 0), TensorProductSimulator.hh:2114
        self._material_set = False

    # ---- sizes / indexing (TensorProductSimulator.hh:209-279, 1532-1651) ----
    def numNodes(self): return int(self._s.num_nodes)
    def numElements(self): return int(self._s.num_elements)
    @property
    def NbElementsPerDimension(self): return self._s.ne.copy()
    gridShape = NbElementsPerDimension
    @property
    def NbNodesPerDimension(self): return self._s.ne + 1
    @property
    def domain(self): return (self._dmin.copy(), self._dmax.copy())
    bbox = domain
    @property
    def dx(self): return (self._dmax - self._dmin) / self._s.ne
    @property
    def elementVolume(self): return float(np.prod(self.dx))
    def nodePosition(self, ni): return self._dmin + np.array(np.unravel_index(int(ni), tuple(self._s.ne + 1))) * self.dx
    def elementIndexForGridCell(self, cellIdxs): return int(np.ravel_multi_index(tuple(int(i) for i in cellIdxs), tuple(self._s.ne)))
    def elemNodeGlobalIndex(self, ei, n):
        e = np.array(np.unravel_index(int(ei), tuple(self._s.ne)))
        off = np.array([(int(n) >> (self._N - 1 - d)) & 1 for d in range(self._N)])
        return int(np.ravel_multi_index(tuple(e + off), tuple(self._s.ne + 1)))
    def elementNodes(self, ei): return [self.elemNodeGlobalIndex(ei, n) for n in range(2 ** self._N)]

    # ---- material / interpolation law (:2055-2102) ----
    def readMaterial(self, materialPath):
        m = json.load(open(materialPath))
        if m.get("type") != "isotropic_material":
            raise NotImplementedError("only isotropic_material files are supported (Materials.cc:291-311)")
        self._et.setIsotropic(m["young"], m["poisson"])
        self._s.set_isotropic(float(m["young"]), float(m["poisson"]))
        self._material_set = True

    def _get_et(self): return self._et
    def _set_et(self, et):
        self._et = et
        if getattr(et, "isotropic", None): self._s.set_isotropic(*et.isotropic)
        else: self._s.set_elasticity_tensor(np.asarray(et.D, dtype=np.float64))
        self._material_set = True
    ETensor = property(_get_et, _set_et)

    def _push_interp(self): self._s.set_interp(**self._interp)
    def _interp_prop(key, cast=float):
        def g(self): return cast(self._interp[key])
        def s(self, v):
            self._interp[key] = cast(v); self._push_interp()
        return property(g, s)
    interpolationLaw = _interp_prop("law", lambda v: InterpolationLaw(int(v)))
    E_0 = _interp_prop("E0")
    E_min = _interp_prop("Emin")
    gamma = _interp_prop("gamma")
    q = _interp_prop("q")
    del _interp_prop

    def _get_gravity(self): return self._gravity.copy()
    def _set_gravity(self, g):
        self._gravity = np.asarray(g, dtype=np.float64).copy(); self._s.set_gravity(self._gravity)
    gravity = property(_get_gravity, _set_gravity)

    # ---- densities (:290-331, 2088-2102) ----
    def getDensities(self): return self._s.densities()
    def setDensities(self, rho): self._s.set_densities(rho)
    def setUniformDensities(self, density): self._s.set_uniform_density(float(density))
    def setDensity(self, ei, value):
        rho = self._s.densities(); rho[int(ei)] = value; self._s.set_densities(rho)
    def setDensitiesFromCoarseGrid(self, upscalingFactor, rho):
        c = np.asarray(rho, dtype=np.float64).reshape(tuple(self._s.ne // int(upscalingFactor)))
        for d in range(self._N): c = np.repeat(c, int(upscalingFactor), axis=d)
        self._s.set_densities(c.ravel())
    def elementDensity(self, ei): return float(self._s.densities()[int(ei)])
    def getYoungModulusScaleFactor(self): return self._s.E()
    def elementYoungModulusScaleFactor(self, ei): return float(self._s.E()[int(ei)])
    def setFabricationMaskHeightByLayer(self, l): self._s.set_mask_layer(int(l))
    def getFabricationMaskHeight(self):
        a, b, h = C.c_int64(), C.c_int64(), C.c_double()
        capi._check(self._s.L.vf_sim_get_mask_info(self._s.h, C.byref(a), C.byref(b), C.byref(h)))
        return h.value

    # ---- boundary conditions and loads (:464-652, 1269-1288) ----
    def applyDisplacementsAndLoadsFromFile(self, bcPath): self._s.apply_bc_file(bcPath)
    def addDirichletCondition(self, u, minCorner, maxCorner, componentMask="xyz"):
        cm = sum(1 << "xyz".index(ch) for ch in componentMask.lower())
        self._s.add_dirichlet(u, minCorner, maxCorner, cm)
    def applySymmetryConditions(self, symmetry_axes, minMaxFace=None):
        bits = lambda a: sum(1 << d for d, v in enumerate(a) if v)
        self._s.apply_symmetry_conditions(bits(symmetry_axes), bits(minMaxFace) if minMaxFace is not None else 0)
    def getDirichletMask(self):
        m = self._s.dirichlet_mask()
        return np.stack([(m >> c) & 1 for c in range(self._N)], axis=1).astype(bool)
    def buildLoadVector(self): return self._s.build_load()

    # ---- operators and solves ----
    def fullDensityElementStiffnessMatrix(self): return self._s.K0()
    def elementStiffnessMatrix(self, ei): return self._s.K0() * self.elementYoungModulusScaleFactor(ei)
    def applyK(self, u): return self._s.apply_K(u)
    def solve(self, f): return self._s.solve(f)
    def complianceGradient(self, u): return self._s.compliance_gradient(u)
    def elementEnergyDensity(self, u): return self._s.energy_density(u)
    def multigridSolver(self, numCoarseningLevels): return _MG(self, int(numCoarseningLevels))
    def mesh(self): return self
    def clearCachedElementStiffness(self): pass

    # ---- layer-by-layer helpers (:1852-1923) and downsampling (:1926-1992) ----
    def getIntermediateFabricationShape(self, yfrac, validateBoundaryConditions=True, law=SIMP):
        if not 0 <= yfrac <= 1: raise RuntimeError("hfrac is out of bounds")
        ne = self._s.ne.copy(); full = int(ne[1]); nh = int(round(yfrac * full))
        if abs(yfrac * full - nh) > 1e-10: raise RuntimeError("hfrac chops off a noninteger number of element layers")
        unexpected = RuntimeError("Original simulator has unexpected boundary conditions for layer-by-layer simulation")
        if validateBoundaryConditions:      # TensorProductSimulator.hh:1885-1907
            if not np.any(self._gravity): raise unexpected
            if self._s.num_force_nodes() != 0: raise unexpected
            m = self._s.dirichlet_mask().reshape(tuple(self._s.ne + 1))
            full = (1 << self._N) - 1
            base = np.take(m, 0, axis=1)
            if np.any(base != full) or np.any(np.delete(m, 0, axis=1) != 0) or self._s.has_nonzero_dirichlet_values(): raise unexpected
        ne[1] = nh
        dmax = self._dmax.copy(); dmax[1] = self._dmin[1] + yfrac * (self._dmax[1] - self._dmin[1])
        r = _TPS((self._dmin, dmax), ne)
        # law from the argument; E_min, E_0, gamma from the parent; the RAMP factor keeps its default (:1877-1883)
        r._interp = dict(r._interp, law=int(law), Emin=self._interp["Emin"], E0=self._interp["E0"], gamma=self._interp["gamma"]); r._push_interp()
        r.ETensor = self._et
        self.transferDensitiesToIntermediateFabricationShape(r)
        g = self._gravity.copy()
        if not np.any(g): g[1] = -1.0
        r.gravity = g
        eps = 1e-9 * float(np.max(dmax - self._dmin))
        lo, hi = self._dmin - eps, dmax + eps
        hi[1] = self._dmin[1] + eps
        r._s.add_dirichlet(np.zeros(self._N), lo, hi, (1 << self._N) - 1)      # build platform fully clamped (:1913-1920)
        return r
    def transferDensitiesToIntermediateFabricationShape(self, intermediateTPS):
        shp = tuple(self._s.ne); nh = int(intermediateTPS._s.ne[1])
        intermediateTPS.setDensities(self.getDensities().reshape(shp)[:, :nh].ravel())
    def downsample(self, downsamplingLevels):
        f = 2 ** int(downsamplingLevels)
        if np.any(self._s.ne % f): raise RuntimeError("Grid size must be divisible by 2^downsamplingLevels")
        r = _TPS((self._dmin, self._dmax), self._s.ne // f)
        r._interp = dict(self._interp); r._interp["law"] = int(SIMP); r._interp["q"] = 3.0; r._push_interp()
        r.ETensor = self._et
        return r
    def _factor(self, coarse):
        f = int(self._s.ne[0] // coarse._s.ne[0])
        if np.any(coarse._s.ne * f != self._s.ne): raise RuntimeError("Invalid downsampled simulator")
        return f
    def downsampleDensityFieldTo(self, densities, coarseTPS):
        f = self._factor(coarseTPS); d = np.asarray(densities, dtype=np.float64)
        if d.size != self.numElements(): raise RuntimeError("Invalid input densities size (%d vs %d)" % (d.size, self.numElements()))
        shp = []
        for n in coarseTPS._s.ne: shp += [int(n), f]
        c = d.reshape(shp).sum(axis=tuple(range(1, 2 * self._N, 2))) * (1.0 / f ** self._N)
        coarseTPS.setDensities(c.ravel())
    def upsampleDensityGradientFrom(self, coarseTPS, g_coarse):
        f = self._factor(coarseTPS); g = np.asarray(g_coarse, dtype=np.float64)
        if g.size != coarseTPS.numElements(): raise RuntimeError("Invalid coarse gradient size")
        g = g.reshape(tuple(coarseTPS._s.ne))
        for d in range(self._N): g = np.repeat(g, f, axis=d)
        return g.ravel() * (1.0 / f ** self._N)

    # ---- export / post-processing next to the solve path: compat/tps_extras.py ----
    def _dirichletConditions(self): return self._s.dirichlet_conditions()
    def _forceNodes(self): return self._s.force_nodes()
    def getK(self): return tps_extras.getK(self)
    def constantStrainLoad(self, eps): return tps_extras.constantStrainLoad(self, eps)
    def solveWithImposedLoads(self): return tps_extras.solveWithImposedLoads(self)
    def getDirichletVarsAndValues(self): return tps_extras.getDirichletVarsAndValues(self)
    def getForceMask(self): return tps_extras.getForceMask(self)
    def getBCIndicatorField(self): return tps_extras.getBCIndicatorField(self)
    def sampleNodalField(self, u, p): return tps_extras.sampleNodalField(self, u, p)
    def getMesh(self): return tps_extras.getMesh(self)
    def debugMulticolorElementVisit(self): return tps_extras.debugMulticolorElementVisit(self)
    def transferVFieldToIntermediateFabricationShape(self, intermediateTPS, u): return tps_extras.transferVFieldToIntermediateFabricationShape(self, intermediateTPS, u)
    def accumElementScalarFieldFromIntermediateFabricationShape(self, intermediateTPS, rho_in, rho_accum):
        tps_extras.accumElementScalarFieldFromIntermediateFabricationShape(self, intermediateTPS, rho_in, rho_accum)


class _LevelSim:
    """What MG.getSimulator(l) exposes of a coarse-level simulator."""
    def __init__(self, mg, l): self._mg, self._l = mg, l
    def numNodes(self): return int(self._mg._m.nn(self._l))
    def getDirichletMask(self):
        m = self._mg._m.get_sim(self._l).dirichlet_mask(); N = self._mg._m.N
        return np.stack([(m >> c) & 1 for c in range(N)], axis=1).astype(bool)


class _MG:
    """detail.MultigridSolver1_1[_1] (VoxelFEM.cc:155-195)."""

    def __init__(self, tps, levels):
        self._tps, self._m = tps, capi.MG(tps._s, levels)
        self.mg = self

    def getSimulator(self, l): return self._tps if l == 0 else _LevelSim(self, l)
    def computeResidual(self, l, u, b): return self._m.residual(l, u, b)
    def applyK(self, l, u): return self._m.apply_K(l, u)
    def zeroOutDirichletComponents(self, l, u):
        u = np.array(u, dtype=np.float64)
        u[self.getSimulator(l).getDirichletMask()] = 0.0
        return u
    def updateStiffnessMatrices(self): self._m.update_stiffness()
    def setSymmetricGaussSeidel(self, symmetric): self._m.set_symmetric_gs(symmetric)
    def setFabricationMaskHeightByLayer(self, h): self._m.set_mask_layer(int(h))
    def debug_get_x(self, l): return self._m.debug_get("x", l)
    def debug_get_b(self, l): return self._m.debug_get("b", l)
    def debugMulticolorVisit(self): return self._m.debug_multicolor_visit()

    def solve(self, u, f, numSteps, numSmoothingSteps, stiffnessUpdated=False, zeroDirichlet=False, it_callback=None, fullMultigrid=False):
        if it_callback is None:
            return self._m.solve(u, f, numSteps, numSmoothingSteps, stiffnessUpdated, zeroDirichlet, fullMultigrid)
        x = np.array(u, dtype=np.float64)      # callback after every cycle (MultigridSolver.hh:566-570): one cycle per call
        for i in range(int(numSteps)):
            x = self._m.solve(x, f, 1, numSmoothingSteps, stiffnessUpdated or i > 0, zeroDirichlet, fullMultigrid and i == 0)
            it_callback(i, x)
        return x

    def preconditionedConjugateGradient(self, u, b, maxIter, tol, it_callback=None, mgIterations=1, mgSmoothingIterations=1, fullMultigrid=False):
        m = self._m
        cb = None
        if it_callback is not None:
            n = m.nn(0) * m.N

            def cb(it, rnorm):     # the reference hands (it, x, r) to the callback (MultigridSolver.hh:1043-1045, 1146-1147)
                it_callback(it, m.pcg_iterate(), m.pcg_residual())
        x, _, _ = m.pcg(u, b, int(maxIter), float(tol), int(mgIterations), int(mgSmoothingIterations), bool(fullMultigrid), False, cb)
        return x


def TensorProductSimulator(degreesPerDimension, domainBBox, elementsPerDimension, numberType=NumberType.DOUBLE):
    """Factory (VoxelFEM.cc:422-430): only the instantiations the reference's bindings register, <double,1,1[,1]> (:301-308)."""
    if list(degreesPerDimension) not in ([1, 1], [1, 1, 1]) or NumberType(numberType) != NumberType.DOUBLE:
        raise RuntimeError("No template instantiation matching degreesPerDimension/number type!")
    return _TPS(domainBBox, elementsPerDimension)


def getClassName(simulator, name):
    # nameMangler (VoxelFEM.cc:33-35, 251-253): name + "1_1[_1]" (+ "" for double)
    return "pyVoxelFEM.detail." + name + "_".join(["1"] * simulator._N)


# ---------------------------------------------------------------------------------------------------------------------
# Filters and constraints (VoxelFEM.cc:311-397)
# ---------------------------------------------------------------------------------------------------------------------
class _Filter:
    def __init__(self): self._in = self._out = None
    def setInputDimensions(self, gridDims): self._in = np.asarray(gridDims, dtype=np.int64); self._out = self._in if self._out is None else self._out
    def setOutputDimensions(self, gridDims): self._out = np.asarray(gridDims, dtype=np.int64); self._in = self._out if self._in is None else self._in
    @property
    def inputDimensions(self): return self._in
    @property
    def outputDimensions(self): return self._out
    def _dims(self):
        if self._in is None: raise RuntimeError("Grid dimensions are not set")
        return self._in
    def apply(self, x): return self._apply(np.asarray(x, dtype=np.float64).ravel())
    def _backprop(self, g, vars_in): raise NotImplementedError


class SmoothingFilter(_Filter):          # TopologyOptimizationFilter.hh:283-400
    class Type(enum.IntEnum):
        Const = 0
        Linear = 1

    def __init__(self, radius=1, type=Type.Const):
        super().__init__(); self.radius, self.type = int(radius), SmoothingFilter.Type(type)
    def _apply(self, x): return capi.smoothing_filter(x, self._dims(), self.radius, int(self.type))
    def _backprop(self, g, vars_in): return capi.smoothing_filter(g, self._dims(), self.radius, int(self.type))   # symmetric (:297-310)
    def _spec(self): return ("smooth", self.radius, int(self.type))


class ProjectionFilter(_Filter):         # TopologyOptimizationFilter.hh:199-245
    def __init__(self, beta=1.0):
        super().__init__(); self.beta = beta
    def _get_beta(self): return self._beta
    def _set_beta(self, beta):
        if beta <= 0: raise RuntimeError("Beta parameter has to be positive (received beta = %f)" % beta)
        self._beta = float(beta)
    beta = property(_get_beta, _set_beta)
    def invert(self, filteredValue):
        if filteredValue > 1.0 or filteredValue < 0.0:
            raise RuntimeError("ProjectionFilter::invert domain error: target density for inversion is outside [0, 1].")
        return math.atanh((2 * filteredValue - 1) * math.tanh(0.5 * self._beta)) / self._beta + 0.5
    def _apply(self, x): return capi.projection_apply(x, self._beta)
    def _backprop(self, g, vars_in): return capi.projection_backprop(g, vars_in, self._beta)
    def _spec(self): return ("project", self._beta)


class UpsampleFilter(_Filter):           # TopologyOptimizationFilter.hh:418-523
    def __init__(self, factor=2):
        super().__init__(); self._factor = int(factor)
    def setInputDimensions(self, gridDims):
        d = np.asarray(gridDims, dtype=np.int64)
        if np.any(d < 2): raise RuntimeError("Interpolation can only be applied to a 2^d grid or larger")
        self._in = d; self._out = (d - 1) * self._factor + 1
    def setOutputDimensions(self, gridDims):
        d = np.asarray(gridDims, dtype=np.int64)
        if np.any(d < 2): raise RuntimeError("Interpolation can only be applied to a 2^d grid or larger")
        self._out = d; self._in = (d - 1) // self._factor + 1
        if np.any((self._in - 1) * self._factor + 1 != self._out): raise RuntimeError("Output size is not divisible by factor")
    def _apply(self, x): return capi.upsample_filter(x, self._dims(), self._factor)
    def _backprop(self, g, vars_in): return capi.upsample_filter_backprop(g, self._dims(), self._factor)
    def _spec(self): return ("upsample", self._factor)


class VertexToCellFilter(_Filter):       # TopologyOptimizationFilter.hh:528-598
    def setInputDimensions(self, gridDims):
        d = np.asarray(gridDims, dtype=np.int64)
        if np.any(d < 2): raise RuntimeError("Input grid must be 2^d or larger.")
        self._in = d; self._out = d - 1
    def setOutputDimensions(self, gridDims):
        self._out = np.asarray(gridDims, dtype=np.int64); self._in = self._out + 1
    def _apply(self, x): return capi.vertex_to_cell_filter(x, self._dims())
    def _backprop(self, g, vars_in): return capi.vertex_to_cell_filter_backprop(g, self._dims())
    def _spec(self): return ("vertex_to_cell",)


class LangelaarFilter(_Filter):          # TopologyOptimizationFilter.hh:601-712
    def __init__(self):
        super().__init__(); self._filtered = self._smax = None
    def _apply(self, x):
        prev = self._filtered if self._filtered is not None and self._filtered.size == x.size else None
        self._filtered, self._smax = capi.langelaar_filter(x, self._dims(), out_prev=prev)   # m_cachedFiltered, m_cachedSmax (:622, 697-701)
        return self._filtered.copy()
    def _backprop(self, g, vars_in):
        if self._filtered is None: raise RuntimeError("LangelaarFilter.backprop before apply")
        return capi.langelaar_filter_backprop(g, vars_in, self._filtered, self._smax, self._dims())
    def _spec(self): return ("langelaar",)


class PythonFilter(_Filter):             # TopologyOptimizationFilter.hh:247-275
    def __init__(self):
        super().__init__(); self.apply_cb = None; self.backprop_cb = None
    def _apply(self, x):
        if self.apply_cb is None: raise RuntimeError("Apply callback must be configured")
        out = np.zeros(int(np.prod(self._out))); self.apply_cb(x, out); return out          # callbacks write into `out` (Eigen::Ref)
    def _backprop(self, g, vars_in):
        if self.backprop_cb is None: raise RuntimeError("Backprop callback must be configured")
        out = np.zeros(int(np.prod(self._in))); self.backprop_cb(g, vars_in, out); return out
    def _spec(self):
        return ("python", lambda x: self._apply(x), lambda g, v: self._backprop(g, v))


class FilterChain:                       # TopologyOptimizationFilter.hh:90-187
    def __init__(self, filters, outGridDimensions):
        self._filters = list(filters)
        dims = np.asarray(outGridDimensions, dtype=np.int64)
        self._out_dims = dims.copy()
        for f in reversed(self._filters):            # setOutputDimensions (:117-132): from the physical grid backwards
            f.setOutputDimensions(dims); dims = np.asarray(f.inputDimensions, dtype=np.int64)
        self._dims = dims
        self._vars = [np.zeros(int(np.prod(self._dims)))] + [np.zeros(int(np.prod(f.outputDimensions))) for f in self._filters]
    @property
    def filters(self): return self._filters
    def numVars(self): return int(np.prod(self._dims))
    numPhysicalVars = numVars            # the reference binds numPhysicalVars to numVars (VoxelFEM.cc:331)
    def gridDims(self): return self._dims
    def physicalGridDims(self): return self._out_dims
    def setDesignVars(self, xDesign):
        x = np.asarray(xDesign, dtype=np.float64).ravel()
        if x.size != self.numVars(): raise RuntimeError("Variable size mismatch")
        v = [x.copy()]
        for f in self._filters: v.append(f._apply(v[-1]))
        self._vars = v
    def designVars(self): return self._vars[0].copy()
    def physicalVars(self): return self._vars[-1].copy()
    def backprop(self, g):
        g = np.asarray(g, dtype=np.float64).ravel()
        if g.size != self._vars[-1].size: raise RuntimeError("Size mismatch")
        for i in range(len(self._filters) - 1, -1, -1): g = self._filters[i]._backprop(g, self._vars[i])
        return g


class TotalVolumeConstraint:             # TopologyOptimizationConstraint.hh:24-40
    def __init__(self, volumeFraction): self.volumeFraction = float(volumeFraction)


# ---------------------------------------------------------------------------------------------------------------------
# Objectives, problem, optimizers (VoxelFEM.cc:226-297)
# ---------------------------------------------------------------------------------------------------------------------
class _MGComplianceObjective:            # TopologyOptimizationObjective.hh:60-105
    def __init__(self, mg_solver):
        self.mg = mg_solver
        self.cgIter, self.tol, self.mgIterations, self.mgSmoothingIterations = 100, 1e-5, 1, 2
        self.fullMultigrid, self.zeroInit, self.residual_cb = True, False, None
        self._problem = None

    def _p(self):
        pr = self._problem() if self._problem is not None else None
        if pr is None: raise RuntimeError("objective is not attached to a TopologyOptimizationProblem yet")
        return pr._p
    def compliance(self): return self._p().compliance()
    def u(self): return self._p().u()
    def f(self): return self.mg._tps.buildLoadVector()
    def gradient(self): return self.mg._tps.complianceGradient(self.u())
    def updateCache(self, xPhys):
        raise NotImplementedError("updateCache is driven by TopologyOptimizationProblem.setVars on the device (vf_top_set_vars)")


def MultigridComplianceObjective(mg_solver): return _MGComplianceObjective(mg_solver)


def ComplianceObjective(simulator):
    raise NotImplementedError("ComplianceObjective (CHOLMOD direct solve of the fine system) is not on the B200 path; "
                              "use MultigridComplianceObjective(tps.multigridSolver(levels))")


class _TOProblem:
    """detail.TopologyOptimizationProblem (TopologyOptimizationProblem.hh:17-155), device-resident through vf_top_*."""

    def __init__(self, simulator, objective, constraints, filters):
        if not isinstance(objective, _MGComplianceObjective): raise NotImplementedError("objective must be a MultigridComplianceObjective")
        if len(constraints) != 1 or not isinstance(constraints[0], TotalVolumeConstraint):
            raise NotImplementedError("constraints must be [TotalVolumeConstraint] (the OC optimizer requires exactly that, OptimalityCriterion.hh:43-45)")
        self._sim, self._obj, self._constraints, self._filters = simulator, objective, list(constraints), list(filters)
        dims = np.asarray(simulator.NbElementsPerDimension, dtype=np.int64)
        for f in reversed(self._filters): f.setOutputDimensions(dims); dims = np.asarray(f.inputDimensions, dtype=np.int64)
        self._p = capi.Problem(objective.mg._m, [f._spec() for f in self._filters], self._constraints[0].volumeFraction)
        self._chain = _DeviceChainView(self)
        self._res_cb_set = None
        objective._problem = weakref.ref(self)

    def _sync_solver(self):
        o = self._obj
        if o.residual_cb is not self._res_cb_set:       # residual_cb(it, r) (TopologyOptimizationObjective.hh:93, 104)
            cb = o.residual_cb
            self._p.set_residual_callback(None if cb is None else (lambda it, rn: cb(it, o.mg._m.pcg_residual())))
            self._res_cb_set = cb
        self._p.set_solver(int(o.cgIter), float(o.tol), int(o.mgIterations), int(o.mgSmoothingIterations), bool(o.fullMultigrid), bool(o.zeroInit))
    def numVars(self): return self._p.nv
    def setVars(self, x, forceUpdate=False):
        self._sync_solver(); self._p.set_vars(x); return True
    def getVars(self): return self._p.design_vars()
    def getDensities(self): return self._p.physical_vars()
    def evaluateObjective(self): return self._p.compliance()
    def evaluateObjectiveGradient(self): return self._p.objective_gradient()
    def evaluateConstraints(self): return np.array([self._p.constraint()])
    def evaluateConstraintsJacobian(self): return self._p.constraint_jacobian().reshape(1, -1)
    @property
    def objective(self): return self._obj
    @property
    def filters(self): return self._filters
    @property
    def filterChain(self): return self._chain
    @property
    def constraints(self): return self._constraints


class _DeviceChainView:
    """problem.filterChain: reads the device-resident chain of the problem; backprop runs the stand-alone filter kernels."""
    def __init__(self, problem): self._pr = problem
    @property
    def filters(self): return self._pr._filters
    def numVars(self): return self._pr.numVars()
    numPhysicalVars = numVars
    def gridDims(self): return self._pr._p.grid_dims(False)
    def physicalGridDims(self): return self._pr._p.grid_dims(True)
    def designVars(self): return self._pr._p.design_vars()
    def physicalVars(self): return self._pr._p.physical_vars()
    def setDesignVars(self, xDesign): self._pr.setVars(xDesign)
    def backprop(self, g):
        # back-propagation of an arbitrary physical-space gradient through the problem's filters at the current design variables
        fc = FilterChain(self._pr._filters, self.physicalGridDims()); fc.setDesignVars(self.designVars())
        return fc.backprop(g)


def TopologyOptimizationProblem(simulator, objective, constraints, filters): return _TOProblem(simulator, objective, constraints, filters)


class _OCOptimizer:                      # OptimalityCriterion.hh:38-149
    def __init__(self, problem): self._pr = problem
    def step(self, m=0.2, p=0.5, ctol=1e-6, inplace=True):
        pr = self._pr
        pr._sync_solver()
        overridden = any(getattr(type(pr), n) is not getattr(_TOProblem, n) for n in ("setVars", "evaluateObjective", "evaluateObjectiveGradient"))
        if inplace and not overridden:       # everything stays on the device (vf_top_oc_step)
            pr._p.oc_step(m, p, ctol); return
        # A Python subclass overrides the problem's virtual methods (trampoline, VoxelFEM.cc:58-66) or the caller asked for the
        # out-of-place variant (:57-60): the objective gradient comes from the (possibly overridden) method, the multiplier search
        # runs on the device against the problem's own chain and constraint, and the result goes through setVars (:133).
        stepped, _ = pr._p.oc_search(pr.evaluateObjectiveGradient(), m, p, ctol)
        pr.setVars(stepped)


def OCOptimizer(problem): return _OCOptimizer(problem)


class _LBL:                              # LayerByLayer.hh:25-309
    def __init__(self, lblSim): self._sim, self._l, self._method = lblSim, None, "N=3"
    def selectInitMethod(self, method):
        self._method = method
        if self._l is not None: self._l.select_init_method(method)
    def run(self, solver, zeroInit, layerIncrement, maxIter, tol, it_callback=None, mgIterations=1, mgSmoothingIterations=1,
            fullMultigrid=False, verbose=False, lblCallback=None):
        if self._l is None or self._l.mg is not solver._m:
            self._l = capi.LBL(solver._m); self._l.select_init_method(self._method)
        lbl, m = self._l, solver._m
        cb = pcb = None
        if lblCallback is not None or verbose:
            def cb(layer, compliance, iters):
                if verbose: print("Layer %d: %s" % (layer, repr(compliance)))        # LayerByLayer.hh:275-276
                if lblCallback is not None: lblCallback(layer, compliance, lbl.layer_gradient(), lbl.layer_u())   # cb(l, compliance, grad_compliance, u) (:222, 277-279)
        if it_callback is not None:
            def pcb(it, rnorm): it_callback(it, m.pcg_iterate(), m.pcg_residual())      # (it, x, r), MultigridSolver.hh:1043-1045
        self._its, self._compliances = lbl.run(bool(zeroInit), int(layerIncrement), int(maxIter), float(tol), int(mgIterations), int(mgSmoothingIterations),
                                               bool(fullMultigrid), cb, pcb)             # the reference's run returns nothing
    def objective(self): return self._l.objective()
    def gradient(self): return self._l.gradient()


def LayerByLayerEvaluator(lblSim): return _LBL(lblSim)


class detail:
    """Namespace mirroring the pybind11 `detail` submodule's mangled class names (VoxelFEM.cc:19-32)."""
    TensorProductSimulator1_1 = TensorProductSimulator1_1_1 = _TPS
    MultigridSolver1_1 = MultigridSolver1_1_1 = _MG
    TopologyOptimizationProblem1_1 = TopologyOptimizationProblem1_1_1 = _TOProblem
    MultigridComplianceObjective1_1 = MultigridComplianceObjective1_1_1 = _MGComplianceObjective
    OCOptimizer1_1 = OCOptimizer1_1_1 = _OCOptimizer
    LayerByLayerEvaluator1_1 = LayerByLayerEvaluator1_1_1 = _LBL
