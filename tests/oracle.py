"""ctypes wrapper around the CPU oracle (oracle/libvfo.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the product
package voxelfem_b200.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "libvfo.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    src = os.path.join(_ROOT, "oracle", "vfo.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


MMA_F_CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
MMA_DF_CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build_oracle()
    L = C.CDLL(_LIB_PATH)
    vp, ci, cd, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    sig = {
        "vfo_last_error": (C.c_char_p, []),
        "vfo_num_threads": (ci, []),
        "vfo_set_num_threads": (None, [ci]),
        "vfo_sim_create": (vp, [ci, _ip, _dp, _dp]),
        "vfo_sim_destroy": (None, [vp]),
        "vfo_sim_num_nodes": (i64, [vp]),
        "vfo_sim_num_elements": (i64, [vp]),
        "vfo_sim_set_isotropic": (None, [vp, cd, cd]),
        "vfo_sim_set_D": (None, [vp, _dp]),
        "vfo_sim_get_K0": (None, [vp, _dp]),
        "vfo_sim_set_interp": (None, [vp, ci, cd, cd, cd, cd]),
        "vfo_sim_set_gravity": (None, [vp, _dp]),
        "vfo_sim_set_densities": (ci, [vp, _dp]),
        "vfo_sim_set_uniform_density": (ci, [vp, cd]),
        "vfo_sim_get_E": (None, [vp, _dp]),
        "vfo_sim_get_densities": (None, [vp, _dp]),
        "vfo_sim_apply_bcs": (ci, [vp, ci, _i32p, _i32p, _dp, _dp, _dp]),
        "vfo_sim_add_dirichlet": (ci, [vp, _dp, _dp, _dp, ci]),
        "vfo_sim_get_dirichlet_mask": (None, [vp, _u8p]),
        "vfo_sim_num_force_nodes": (i64, [vp]),
        "vfo_sim_build_load": (None, [vp, _dp]),
        "vfo_sim_apply_K": (None, [vp, _dp, _dp, ci, ci]),
        "vfo_sim_set_mask_layer": (ci, [vp, i64]),
        "vfo_sim_mask_info": (None, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "vfo_sim_compliance_gradient": (None, [vp, _dp, _dp, ci]),
        "vfo_sim_energy_density": (None, [vp, _dp, _dp]),
        "vfo_sim_solve": (ci, [vp, _dp, _dp]),
        "vfo_sim_zero_dirichlet": (None, [vp, _dp]),
        "vfo_sim_masked_dot": (cd, [vp, _dp, _dp]),
        "vfo_mg_create": (vp, [vp, ci]),
        "vfo_mg_destroy": (None, [vp]),
        "vfo_mg_num_levels": (ci, [vp]),
        "vfo_mg_get_sim": (vp, [vp, ci]),
        "vfo_mg_get_coarsened_fine_K0": (None, [vp, ci, _dp]),
        "vfo_mg_update_stiffness": (ci, [vp]),
        "vfo_mg_apply_K": (ci, [vp, ci, _dp, _dp]),
        "vfo_mg_residual": (ci, [vp, ci, _dp, _dp, _dp]),
        "vfo_mg_smooth": (ci, [vp, ci, _dp, _dp, ci]),
        "vfo_mg_restrict": (None, [vp, ci, _dp, _dp]),
        "vfo_mg_interpolate": (None, [vp, ci, _dp, _dp, ci]),
        "vfo_mg_get_stencil": (ci, [vp, ci, _dp]),
        "vfo_mg_coarse_solve": (ci, [vp, _dp, _dp]),
        "vfo_mg_solve": (ci, [vp, _dp, _dp, ci, ci, ci, ci, ci, _dp]),
        "vfo_mg_pcg": (ci, [vp, _dp, _dp, ci, cd, ci, ci, ci, ci, C.POINTER(ci), _dp]),
        "vfo_mg_get_pcg_residual": (None, [vp, _dp]),
        "vfo_mg_set_symmetric_gs": (None, [vp, ci]),
        "vfo_mg_set_mask_layer": (ci, [vp, i64]),
        "vfo_mg_decrement_mask": (ci, [vp, ci]),
        "vfo_mg_debug_get": (None, [vp, ci, ci, _dp]),
        "vfo_mg_debug_multicolor_visit": (None, [vp, _i32p]),
        "vfo_smoothing_filter": (None, [ci, _ip, ci, ci, _dp, _dp]),
        "vfo_projection_apply": (None, [i64, cd, _dp, _dp]),
        "vfo_projection_backprop": (None, [i64, cd, _dp, _dp, _dp]),
        "vfo_filter_upsample": (None, [ci, _ip, ci, _dp, _dp]),
        "vfo_filter_upsample_backprop": (None, [ci, _ip, ci, _dp, _dp]),
        "vfo_filter_v2c": (None, [ci, _ip, _dp, _dp]),
        "vfo_filter_v2c_backprop": (None, [ci, _ip, _dp, _dp]),
        "vfo_filter_langelaar": (None, [ci, _ip, _dp, _dp, _dp]),
        "vfo_filter_langelaar_backprop": (None, [ci, _ip, _dp, _dp, _dp, _dp, _dp]),
        "vfo_problem_create": (vp, [vp, ci, _dp, cd]),
        "vfo_problem_destroy": (None, [vp]),
        "vfo_problem_set_solver": (None, [vp, ci, cd, ci, ci, ci, ci]),
        "vfo_problem_set_vars": (ci, [vp, _dp]),
        "vfo_problem_get_vars": (None, [vp, ci, _dp]),
        "vfo_problem_compliance": (cd, [vp]),
        "vfo_problem_constraint": (cd, [vp]),
        "vfo_problem_objective_gradient": (None, [vp, _dp]),
        "vfo_problem_constraint_jacobian": (None, [vp, _dp]),
        "vfo_problem_get_u": (None, [vp, _dp]),
        "vfo_problem_last_pcg_iters": (ci, [vp]),
        "vfo_problem_oc_step": (ci, [vp, cd, cd, cd, C.POINTER(ci)]),
        "vfo_problem_get_lambda": (None, [vp, C.POINTER(cd), C.POINTER(cd)]),
        "vfo_lbl_create": (vp, [vp]),
        "vfo_lbl_destroy": (None, [vp]),
        "vfo_lbl_select_init_method": (ci, [vp, C.c_char_p]),
        "vfo_lbl_run": (ci, [vp, ci, i64, ci, cd, ci, ci, ci]),
        "vfo_lbl_objective": (cd, [vp]),
        "vfo_lbl_gradient": (None, [vp, _dp]),
        "vfo_lbl_num_layers_run": (ci, [vp]),
        "vfo_lbl_layer_info": (None, [vp, _i32p, _dp]),
        "vfo_mma_create": (vp, [ci, ci, _dp, _dp, MMA_F_CB, MMA_DF_CB, vp]),
        "vfo_mma_destroy": (None, [vp]),
        "vfo_mma_enable_gcmma": (None, [vp, ci]),
        "vfo_mma_set_initial_var": (None, [vp, _dp]),
        "vfo_mma_step": (ci, [vp]),
        "vfo_mma_get_optimal_var": (None, [vp, _dp]),
        "vfo_mma_newton_iterations": (i64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().vfo_last_error().decode())


# ---------------------------------------------------------------------------
# Boundary-condition files (same JSON schema as the reference's examples/bcs/*.bc;
# parsing rules: 3rdParty/MeshFEM/src/lib/MeshFEM/BoundaryConditions.cc:219-380)
# ---------------------------------------------------------------------------
def parse_bc(path_or_dict, dmin, dmax):
    """Return (kind, cmask, values, bmin, bmax) arrays for the axis-aligned box regions."""
    cfg = path_or_dict
    if not isinstance(cfg, dict):
        with open(path_or_dict) as f:
            cfg = json.load(f)
    dmin = np.asarray(dmin, dtype=float)
    dmax = np.asarray(dmax, dtype=float)
    N = len(dmin)
    kinds, masks, vals, los, his = [], [], [], [], []
    for reg in cfg["regions"]:
        t = reg["type"]
        cm = 7
        if t.startswith("dirichlet"):
            rest = t[9:]
            comp = ""
            for ch in rest:
                if ch < "x" or ch > "z":
                    break
                comp += ch
            if len(comp) > 3:
                raise RuntimeError("invalid mask")
            if comp:
                cm = sum(1 << "xyz".index(c) for c in set(comp))
            if rest[len(comp):] != "":
                raise RuntimeError("Invalid type '%s'" % t)
            kind = 0
        elif t == "force":
            kind = 1
        else:
            raise RuntimeError("Illegal constraint type, only \"dirichlet\" and \"force\" accepted")

        def pad(v):
            v = [float(x) for x in v][:3]
            return v + [0.0] * (3 - len(v))
        if "box%" in reg:
            lo = np.array(pad(reg["box%"]["minCorner"]))
            hi = np.array(pad(reg["box%"]["maxCorner"]))
            lo[:N] = dmin + lo[:N] * (dmax - dmin)  # BBox::interpolatePoint (Geometry.hh:259-262)
            hi[:N] = dmin + hi[:N] * (dmax - dmin)
        else:
            lo = np.array(pad(reg["box"]["minCorner"]))
            hi = np.array(pad(reg["box"]["maxCorner"]))
        kinds.append(kind)
        masks.append(cm)
        vals.append(pad(reg["value"]))
        los.append(lo)
        his.append(hi)
    return (np.array(kinds, dtype=np.int32), np.array(masks, dtype=np.int32),
            np.ascontiguousarray(vals, dtype=np.float64), np.ascontiguousarray(los, dtype=np.float64),
            np.ascontiguousarray(his, dtype=np.float64))


def to_soa(u):
    """(numNodes, N) array -> flat SoA (component-major) as stored by VField (ColMajor)."""
    return np.ascontiguousarray(np.asarray(u, dtype=np.float64).T).ravel()


def from_soa(flat, N):
    return np.ascontiguousarray(flat.reshape(N, -1).T)


class OracleSim:
    def __init__(self, ne, dmin=None, dmax=None, _handle=None):
        self.L = lib()
        if _handle is not None:
            self.h = _handle
            self.N = None
        else:
            ne = np.ascontiguousarray(ne, dtype=np.int64)
            self.N = len(ne)
            if dmin is None:
                dmin = np.zeros(self.N)
                dmax = ne.astype(float)
            self.dmin = np.ascontiguousarray(dmin, dtype=np.float64)
            self.dmax = np.ascontiguousarray(dmax, dtype=np.float64)
            self.ne = ne
            self.h = self.L.vfo_sim_create(self.N, ne, self.dmin, self.dmax)
        self.owned = True

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vfo_sim_destroy(self.h)
            self.h = None

    @property
    def num_nodes(self): return self.L.vfo_sim_num_nodes(self.h)
    @property
    def num_elements(self): return self.L.vfo_sim_num_elements(self.h)
    @property
    def nn(self): return self.ne + 1

    def set_isotropic(self, E, nu): self.L.vfo_sim_set_isotropic(self.h, E, nu)
    def set_elasticity_tensor(self, D): self.L.vfo_sim_set_D(self.h, np.ascontiguousarray(D, dtype=np.float64))   # setETensor (TensorProductSimulator.hh:343-347)

    def K0(self):
        ke = self.N * 2 ** self.N
        out = np.zeros((ke, ke))
        self.L.vfo_sim_get_K0(self.h, out)
        return out

    def set_interp(self, law=0, E0=1.0, Emin=1e-4, gamma=3.0, q=3.0): self.L.vfo_sim_set_interp(self.h, law, E0, Emin, gamma, q)
    def set_gravity(self, g): self.L.vfo_sim_set_gravity(self.h, np.ascontiguousarray(g, dtype=np.float64))
    def set_densities(self, rho): _check(self.L.vfo_sim_set_densities(self.h, np.ascontiguousarray(rho, dtype=np.float64).ravel()))
    def set_uniform_density(self, v): _check(self.L.vfo_sim_set_uniform_density(self.h, v))

    def E(self):
        out = np.zeros(self.num_elements)
        self.L.vfo_sim_get_E(self.h, out)
        return out

    def apply_bc_file(self, path):
        k, m, v, lo, hi = parse_bc(path, self.dmin, self.dmax)
        _check(self.L.vfo_sim_apply_bcs(self.h, len(k), k, m, v, lo, hi))

    def add_dirichlet(self, u, lo, hi, cmask=7):
        pad = lambda a: np.ascontiguousarray(list(a) + [0.0] * (3 - len(a)), dtype=np.float64)
        _check(self.L.vfo_sim_add_dirichlet(self.h, pad(u), pad(lo), pad(hi), cmask))

    def dirichlet_mask(self):
        out = np.zeros(self.num_nodes, dtype=np.uint8)
        self.L.vfo_sim_get_dirichlet_mask(self.h, out)
        return out

    def build_load(self):
        f = np.zeros(self.num_nodes * self.N)
        self.L.vfo_sim_build_load(self.h, f)
        return from_soa(f, self.N)

    def apply_K(self, u, out=None, zero_init=True, negate=False):
        o = np.zeros(self.num_nodes * self.N) if out is None else to_soa(out)
        self.L.vfo_sim_apply_K(self.h, to_soa(u), o, int(zero_init), int(negate))
        return from_soa(o, self.N)

    def set_mask_layer(self, l): _check(self.L.vfo_sim_set_mask_layer(self.h, l))

    def mask_info(self):
        a, b = C.c_int64(), C.c_int64()
        self.L.vfo_sim_mask_info(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def compliance_gradient(self, u, g=None):
        out = np.zeros(self.num_elements) if g is None else np.ascontiguousarray(g, dtype=np.float64).copy()
        self.L.vfo_sim_compliance_gradient(self.h, to_soa(u), out, int(g is not None))
        return out

    def energy_density(self, u):
        out = np.zeros(self.num_elements)
        self.L.vfo_sim_energy_density(self.h, to_soa(u), out)
        return out

    def solve(self, f):
        u = np.zeros(self.num_nodes * self.N)
        _check(self.L.vfo_sim_solve(self.h, to_soa(f), u))
        return from_soa(u, self.N)


class OracleMG:
    def __init__(self, sim, levels):
        self.L = lib()
        self.sim = sim
        self.N = sim.N
        self.h = self.L.vfo_mg_create(sim.h, levels)
        if not self.h:
            raise RuntimeError(self.L.vfo_last_error().decode())
        self.levels = levels
        self._sims = []
        for l in range(levels + 1):
            s = OracleSim(None, _handle=self.L.vfo_mg_get_sim(self.h, l))
            s.N = sim.N
            s.ne = sim.ne // (2 ** l)
            s.dmin, s.dmax = sim.dmin, sim.dmax
            self._sims.append(s)

    def __del__(self):
        if getattr(self, "h", None):
            self._sims = []
            self.L.vfo_mg_destroy(self.h)
            self.h = None

    def get_sim(self, l): return self._sims[l]
    def nn(self, l): return self._sims[l].num_nodes

    def coarsened_fine_K0(self, fi):
        ke = self.N * 2 ** self.N
        out = np.zeros((ke, ke))
        self.L.vfo_mg_get_coarsened_fine_K0(self.h, fi, out)
        return out

    def update_stiffness(self): _check(self.L.vfo_mg_update_stiffness(self.h))

    def apply_K(self, l, u):
        out = np.zeros(self.nn(l) * self.N)
        _check(self.L.vfo_mg_apply_K(self.h, l, to_soa(u), out))
        return from_soa(out, self.N)

    def residual(self, l, u, b):
        out = np.zeros(self.nn(l) * self.N)
        _check(self.L.vfo_mg_residual(self.h, l, to_soa(u), to_soa(b), out))
        return from_soa(out, self.N)

    def smooth(self, l, u, b, forward=True):
        uu = to_soa(u)
        _check(self.L.vfo_mg_smooth(self.h, l, uu, to_soa(b), int(forward)))
        return from_soa(uu, self.N)

    def restrict(self, lf, fine):
        out = np.zeros(self.nn(lf + 1) * self.N)
        self.L.vfo_mg_restrict(self.h, lf, to_soa(fine), out)
        return from_soa(out, self.N)

    def interpolate(self, lf, coarse, fine=None):
        out = np.zeros(self.nn(lf) * self.N) if fine is None else to_soa(fine)
        self.L.vfo_mg_interpolate(self.h, lf, to_soa(coarse), out, int(fine is not None))
        return from_soa(out, self.N)

    def stencil(self, l):
        ns = 3 ** self.N
        out = np.zeros(self.nn(l) * ns * self.N * self.N)
        _check(self.L.vfo_mg_get_stencil(self.h, l, out))
        return out.reshape(self.nn(l), ns, self.N, self.N)

    def coarse_solve(self, f):
        l = self.levels
        out = np.zeros(self.nn(l) * self.N)
        _check(self.L.vfo_mg_coarse_solve(self.h, to_soa(f), out))
        return from_soa(out, self.N)

    def solve(self, u, f, num_steps, num_smooth, stiffness_updated=False, zero_dirichlet=False, fmg=False):
        out = np.zeros(self.nn(0) * self.N)
        _check(self.L.vfo_mg_solve(self.h, to_soa(u), to_soa(f), num_steps, num_smooth, int(stiffness_updated), int(zero_dirichlet), int(fmg), out))
        return from_soa(out, self.N)

    def pcg(self, u, b, max_iter, tol, mg_iterations=1, mg_smoothing=1, fmg=False, dirichlet_ok=False):
        x = to_soa(u)
        it = C.c_int(0)
        res = np.zeros(max(max_iter, 1) + 1)
        _check(self.L.vfo_mg_pcg(self.h, x, to_soa(b), max_iter, tol, mg_iterations, mg_smoothing, int(fmg), int(dirichlet_ok), C.byref(it), res))
        return from_soa(x, self.N), it.value, res[:it.value]

    def set_stiffness_prebuilt(self, on):
        """bench infrastructure: skip the per-call hierarchy rebuild inside pcg (the caller has called update_stiffness)."""
        self.L.vfo_mg_set_stiffness_prebuilt.argtypes = [C.c_void_p, C.c_int]; self.L.vfo_mg_set_stiffness_prebuilt.restype = None
        self.L.vfo_mg_set_stiffness_prebuilt(self.h, int(on))

    def pcg_residual(self):
        out = np.zeros(self.nn(0) * self.N)
        self.L.vfo_mg_get_pcg_residual(self.h, out)
        return from_soa(out, self.N)

    def set_symmetric_gs(self, s): self.L.vfo_mg_set_symmetric_gs(self.h, int(s))
    def set_mask_layer(self, l): _check(self.L.vfo_mg_set_mask_layer(self.h, l))
    def decrement_mask(self, inc): _check(self.L.vfo_mg_decrement_mask(self.h, inc))

    def debug_get(self, which, l):
        out = np.zeros(self.nn(l) * self.N)
        self.L.vfo_mg_debug_get(self.h, {"x": 0, "b": 1, "r": 2}[which], l, out)
        return from_soa(out, self.N)

    def debug_multicolor_visit(self):
        out = np.zeros(self.nn(0), dtype=np.int32)
        self.L.vfo_mg_debug_multicolor_visit(self.h, out)
        return out


class OracleProblem:
    """TopologyOptimizationProblem + MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer."""

    def __init__(self, mg, filters, vol_frac):
        # filters: list of ("smooth", radius, type) / ("project", beta); type 0 = Const, 1 = Linear
        self.L = lib()
        self.mg = mg
        spec = []
        for f in filters:
            if f[0] == "smooth":
                spec += [0, f[1], f[2], 0.0]
            else:
                spec += [1, 0, 0, f[1]]
        spec = np.ascontiguousarray(spec if spec else [0.0], dtype=np.float64)
        self.h = self.L.vfo_problem_create(mg.h, len(filters), spec, vol_frac)
        if not self.h:
            raise RuntimeError(self.L.vfo_last_error().decode())
        self.ne = mg.sim.num_elements
        self.N = mg.N

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vfo_problem_destroy(self.h)
            self.h = None

    def set_solver(self, cg_iter=100, tol=1e-5, mg_it=1, mg_smooth=2, fmg=True, zero_init=False):
        self.L.vfo_problem_set_solver(self.h, cg_iter, tol, mg_it, mg_smooth, int(fmg), int(zero_init))

    def set_vars(self, x): _check(self.L.vfo_problem_set_vars(self.h, np.ascontiguousarray(x, dtype=np.float64)))

    def design_vars(self):
        o = np.zeros(self.ne); self.L.vfo_problem_get_vars(self.h, 0, o); return o

    def physical_vars(self):
        o = np.zeros(self.ne); self.L.vfo_problem_get_vars(self.h, 1, o); return o

    def compliance(self): return self.L.vfo_problem_compliance(self.h)
    def constraint(self): return self.L.vfo_problem_constraint(self.h)

    def objective_gradient(self):
        o = np.zeros(self.ne); self.L.vfo_problem_objective_gradient(self.h, o); return o

    def constraint_jacobian(self):
        o = np.zeros(self.ne); self.L.vfo_problem_constraint_jacobian(self.h, o); return o

    def u(self):
        o = np.zeros(self.mg.nn(0) * self.N); self.L.vfo_problem_get_u(self.h, o); return from_soa(o, self.N)

    def last_pcg_iters(self): return self.L.vfo_problem_last_pcg_iters(self.h)

    def oc_step(self, m=0.2, p=0.5, ctol=1e-6):
        n = C.c_int(0)
        _check(self.L.vfo_problem_oc_step(self.h, m, p, ctol, C.byref(n)))
        return n.value


def smoothing_filter(x, shape, radius, ftype):
    shape = np.ascontiguousarray(shape, dtype=np.int64)
    out = np.zeros(int(np.prod(shape)))
    lib().vfo_smoothing_filter(len(shape), shape, radius, ftype, np.ascontiguousarray(x, dtype=np.float64).ravel(), out)
    return out


def projection_apply(x, beta):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel(); out = np.zeros_like(x)
    lib().vfo_projection_apply(len(x), beta, x, out); return out


def projection_backprop(g, vars_, beta):
    g = np.ascontiguousarray(g, dtype=np.float64).ravel(); out = np.zeros_like(g)
    lib().vfo_projection_backprop(len(g), beta, g, np.ascontiguousarray(vars_, dtype=np.float64).ravel(), out); return out


def _f(a): return np.ascontiguousarray(a, dtype=np.float64).ravel()
def _s(shape): return np.ascontiguousarray(shape, dtype=np.int64)


def upsample(x, coarse_shape, factor):
    cs = _s(coarse_shape); out = np.zeros(int(np.prod((cs - 1) * factor + 1)))
    lib().vfo_filter_upsample(len(cs), cs, factor, _f(x), out); return out


def upsample_backprop(g, coarse_shape, factor):
    cs = _s(coarse_shape); out = np.zeros(int(np.prod(cs)))
    lib().vfo_filter_upsample_backprop(len(cs), cs, factor, _f(g), out); return out


def vertex_to_cell(x, vertex_shape):
    vs = _s(vertex_shape); out = np.zeros(int(np.prod(vs - 1)))
    lib().vfo_filter_v2c(len(vs), vs, _f(x), out); return out


def vertex_to_cell_backprop(g, vertex_shape):
    vs = _s(vertex_shape); out = np.zeros(int(np.prod(vs)))
    lib().vfo_filter_v2c_backprop(len(vs), vs, _f(g), out); return out


def langelaar(x, shape, out_prev=None):
    """returns (filtered, smax); out_prev = previous content of the output array (zeros on a fresh filter)."""
    sz = _s(shape); n = int(np.prod(sz))
    out = np.zeros(n) if out_prev is None else _f(out_prev).copy(); smax = np.zeros(n)
    lib().vfo_filter_langelaar(len(sz), sz, _f(x), out, smax); return out, smax


def langelaar_backprop(g, vars_, filtered, smax, shape):
    sz = _s(shape); out = np.zeros(int(np.prod(sz)))
    lib().vfo_filter_langelaar_backprop(len(sz), sz, _f(g), _f(vars_), _f(filtered), _f(smax), out); return out


class OracleLBL:
    """LayerByLayerEvaluator (LayerByLayer.hh:25-309)."""

    def __init__(self, mg):
        self.L = lib(); self.mg = mg
        self.h = self.L.vfo_lbl_create(mg.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vfo_lbl_destroy(self.h); self.h = None

    def select_init_method(self, m): _check(self.L.vfo_lbl_select_init_method(self.h, m.encode()))

    def run(self, zero_init=True, layer_increment=1, max_iter=50, tol=1e-5, mg_iterations=1, mg_smoothing=1, fmg=False):
        _check(self.L.vfo_lbl_run(self.h, int(zero_init), layer_increment, max_iter, tol, mg_iterations, mg_smoothing, int(fmg)))
        n = self.L.vfo_lbl_num_layers_run(self.h)
        it = np.zeros(n, dtype=np.int32); c = np.zeros(n)
        self.L.vfo_lbl_layer_info(self.h, it, c)
        return it, c

    def objective(self): return self.L.vfo_lbl_objective(self.h)

    def gradient(self):
        g = np.zeros(self.mg.sim.num_elements); self.L.vfo_lbl_gradient(self.h, g); return g


class OracleMMA:
    """MMA (MethodOfMovingAsymptotes.hh:28-469); f(x) -> (m+1,), df_dx(x) -> (m+1, n)."""

    def __init__(self, n, m, xmin, xmax, f, df_dx):
        self.L = lib(); self.n, self.m = n, m

        def _f(xp, out, _):
            x = np.ctypeslib.as_array(xp, shape=(n,))
            np.ctypeslib.as_array(out, shape=(m + 1,))[:] = np.asarray(f(x.copy()), dtype=np.float64).ravel()

        def _df(xp, out, _):
            x = np.ctypeslib.as_array(xp, shape=(n,))
            np.ctypeslib.as_array(out, shape=(m + 1, n))[:] = np.asarray(df_dx(x.copy()), dtype=np.float64).reshape(m + 1, n)
        self._cbs = (MMA_F_CB(_f), MMA_DF_CB(_df))
        self.h = self.L.vfo_mma_create(n, m, np.ascontiguousarray(xmin, dtype=np.float64), np.ascontiguousarray(xmax, dtype=np.float64), self._cbs[0], self._cbs[1], None)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vfo_mma_destroy(self.h); self.h = None

    def enableGCMMA(self, e): self.L.vfo_mma_enable_gcmma(self.h, int(e))
    def setInitialVar(self, x): self.L.vfo_mma_set_initial_var(self.h, np.ascontiguousarray(x, dtype=np.float64))
    def step(self): _check(self.L.vfo_mma_step(self.h))

    def getOptimalVar(self):
        x = np.zeros(self.n); self.L.vfo_mma_get_optimal_var(self.h, x); return x

    def newton_iterations(self): return self.L.vfo_mma_newton_iterations(self.h)
