"""ctypes binding of the C ABI in include/voxelfem_b200.h (libvoxelfem_b200.so).

This is plumbing for tests/ and bench.py; the reference-facing Python module is the
pybind11 `pyVoxelFEM` built from voxelfem_b200/host/.  Nothing here computes on the CPU:
every numeric call goes to the CUDA library and raises if it (or a GPU) is missing.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvoxelfem_b200.so")
DATA_DIR = os.path.join(_HERE, "data")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
PCG_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_void_p)
LBL_CALLBACK = C.CFUNCTYPE(None, C.c_int64, C.c_double, C.c_int, C.c_void_p)
FILTER_APPLY_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64, C.c_void_p)
FILTER_BACKPROP_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_void_p)
MMA_F_CALLBACK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
MMA_DF_CALLBACK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

_lib = None


class VoxelFEMError(RuntimeError):
    pass


def _guarded(ctype, fn):
    """ctypes prints and swallows exceptions raised inside callbacks: catch the first one here so that the caller can re-raise it
    once the C call has returned (and skip the remaining invocations)."""
    failed = []
    if fn is None:
        return ctype(), failed

    def wrapper(*a):
        if failed:
            return
        try:
            fn(*a)
        except BaseException as e:   # noqa: BLE001 -- re-raised by _reraise
            failed.append(e)
    return ctype(wrapper), failed


def _reraise(failed):
    if failed:
        raise failed[0]


def lib():
    """Load libvoxelfem_b200.so; fails loudly if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VoxelFEMError("libvoxelfem_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ci, cd, i64, sz = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_size_t
    pvp = C.POINTER(C.c_void_p)
    sig = {
        "vf_last_error": (C.c_char_p, []),
        "vf_device_count": (ci, [C.POINTER(ci)]),
        "vf_set_device": (ci, [ci]),
        "vf_version": (ci, []),
        "vf_kernel_launch_count": (i64, []),
        "vf_reset_kernel_launch_count": (None, []),
        "vf_measure_fp64_peak": (ci, [C.POINTER(cd)]),
        "vf_benchmark_enable": (None, [ci]),
        "vf_benchmark_enabled": (ci, []),
        "vf_benchmark_reset": (None, []),
        "vf_benchmark_start_timer_section": (None, [C.c_char_p]),
        "vf_benchmark_stop_timer_section": (None, [C.c_char_p]),
        "vf_benchmark_start_timer": (None, [C.c_char_p]),
        "vf_benchmark_stop_timer": (None, [C.c_char_p]),
        "vf_benchmark_add_message": (None, [C.c_char_p]),
        "vf_benchmark_report": (sz, [ci, C.c_char_p, sz]),
        "vf_sim_create": (ci, [ci, _ip, _dp, _dp, pvp]),
        "vf_sim_destroy": (ci, [vp]),
        "vf_sim_num_nodes": (i64, [vp]),
        "vf_sim_num_elements": (i64, [vp]),
        "vf_sim_set_elasticity_tensor": (ci, [vp, _dp]),
        "vf_sim_set_isotropic": (ci, [vp, cd, cd]),
        "vf_sim_get_K0": (ci, [vp, _dp]),
        "vf_sim_set_interpolation": (ci, [vp, ci, cd, cd, cd, cd]),
        "vf_sim_set_gravity": (ci, [vp, _dp]),
        "vf_sim_set_densities": (ci, [vp, _dp]),
        "vf_sim_set_uniform_density": (ci, [vp, cd]),
        "vf_sim_get_densities": (ci, [vp, _dp]),
        "vf_sim_get_young_moduli": (ci, [vp, _dp]),
        "vf_sim_apply_bc_regions": (ci, [vp, ci, _i32p, _i32p, _dp, _dp, _dp]),
        "vf_sim_add_dirichlet_box": (ci, [vp, _dp, _dp, _dp, ci]),
        "vf_sim_apply_symmetry_conditions": (ci, [vp, ci, ci]),
        "vf_sim_get_dirichlet_mask": (ci, [vp, _u8p]),
        "vf_sim_num_force_nodes": (i64, [vp]),
        "vf_sim_num_dirichlet_nodes": (i64, [vp]),
        "vf_sim_get_dirichlet_conditions": (ci, [vp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), _dp]),
        "vf_sim_get_force_nodes": (ci, [vp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), _dp]),
        "vf_sim_num_nonzero_dirichlet_values": (i64, [vp]),
        "vf_sim_build_load_vector": (ci, [vp, _dp]),
        "vf_sim_build_load_vector_dev": (ci, [vp, vp]),
        "vf_sim_apply_K": (ci, [vp, _dp, _dp, ci, ci]),
        "vf_sim_set_mask_layer": (ci, [vp, i64]),
        "vf_sim_get_mask_info": (ci, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(cd)]),
        "vf_sim_compliance_gradient": (ci, [vp, _dp, _dp, ci]),
        "vf_sim_element_energy_density": (ci, [vp, _dp, _dp]),
        "vf_sim_solve": (ci, [vp, _dp, _dp]),
        "vf_mg_create": (ci, [vp, ci, pvp]),
        "vf_mg_destroy": (ci, [vp]),
        "vf_mg_num_levels": (ci, [vp]),
        "vf_mg_level_num_nodes": (i64, [vp, ci]),
        "vf_mg_level_grid": (ci, [vp, ci, _ip]),
        "vf_mg_level_dirichlet_mask": (ci, [vp, ci, _u8p]),
        "vf_mg_get_coarsened_fine_K0": (ci, [vp, ci, _dp]),
        "vf_mg_update_stiffness_matrices": (ci, [vp]),
        "vf_mg_apply_K": (ci, [vp, ci, _dp, _dp]),
        "vf_mg_compute_residual": (ci, [vp, ci, _dp, _dp, _dp]),
        "vf_mg_smooth": (ci, [vp, ci, _dp, _dp, ci]),
        "vf_mg_smooth_residual": (ci, [vp, ci, _dp, _dp, ci, _dp]),
        "vf_mg_restrict": (ci, [vp, ci, _dp, _dp]),
        "vf_mg_interpolate": (ci, [vp, ci, _dp, _dp, ci]),
        "vf_mg_get_stencil": (ci, [vp, ci, _dp]),
        "vf_mg_coarse_solve": (ci, [vp, _dp, _dp]),
        "vf_mg_solve": (ci, [vp, _dp, _dp, ci, ci, ci, ci, ci, _dp]),
        "vf_mg_pcg": (ci, [vp, _dp, _dp, ci, cd, ci, ci, ci, ci, C.POINTER(ci), _dp, PCG_CALLBACK, vp]),
        "vf_mg_pcg_io": (ci, [vp, _dp, _dp, _dp, ci, cd, ci, ci, ci, ci, C.POINTER(ci), _dp, PCG_CALLBACK, vp]),
        "vf_mg_get_pcg_iterate": (ci, [vp, _dp]),
        "vf_dev_filter_smooth": (ci, [vp, ci, _ip, ci, ci, vp, vp]),
        "vf_dev_filter_project": (ci, [vp, i64, cd, vp, vp]),
        "vf_dev_filter_project_backprop": (ci, [vp, i64, cd, vp, vp, vp]),
        "vf_dev_oc_update": (ci, [vp, i64, vp, vp, vp, cd, cd, cd, vp]),
        "vf_dev_sum": (ci, [vp, i64, vp, C.POINTER(cd)]),
        "vf_sim_set_densities_dev": (ci, [vp, vp]),
        "vf_sim_compliance_gradient_dev": (ci, [vp, vp, vp, ci]),
        "vf_sim_stream": (vp, [vp]),
        "vf_sim_synchronize": (ci, [vp]),
        "vf_mg_pcg_dev": (ci, [vp, vp, vp, ci, cd, ci, ci, ci, ci, C.POINTER(ci), _dp, PCG_CALLBACK, vp]),
        "vf_mg_get_pcg_residual": (ci, [vp, _dp]),
        "vf_mg_set_symmetric_gauss_seidel": (ci, [vp, ci]),
        "vf_mg_set_rebuild_every_solve": (ci, [vp, ci]),
        "vf_mg_set_mask_layer": (ci, [vp, i64]),
        "vf_mg_decrement_mask": (ci, [vp, ci]),
        "vf_mg_debug_get": (ci, [vp, ci, ci, _dp]),
        "vf_mg_debug_multicolor_visit": (ci, [vp, _i32p]),
        "vf_dev_alloc": (ci, [sz, pvp]),
        "vf_dev_free": (ci, [vp]),
        "vf_dev_upload": (ci, [vp, _dp, sz]),
        "vf_dev_download": (ci, [_dp, vp, sz]),
        "vf_dev_memset_zero": (ci, [vp, sz]),
        "vf_mg_stream": (vp, [vp]),
        "vf_mg_synchronize": (ci, [vp]),
        "vf_prof_enable": (ci, [vp, ci]),
        "vf_prof_reset": (ci, [vp]),
        "vf_prof_num_categories": (ci, []),
        "vf_prof_name": (C.c_char_p, [ci]),
        "vf_prof_get": (ci, [vp, ci, C.POINTER(i64), C.POINTER(cd), C.POINTER(cd)]),
        "vf_mg_time_op": (ci, [vp, ci, ci, ci, ci, C.POINTER(cd)]),
        "vf_filter_smooth": (ci, [ci, _ip, ci, ci, _dp, _dp]),
        "vf_filter_project": (ci, [i64, cd, _dp, _dp]),
        "vf_filter_project_backprop": (ci, [i64, cd, _dp, _dp, _dp]),
        "vf_top_create": (ci, [vp, ci, _dp, cd, pvp]),
        "vf_top_num_vars": (i64, [vp]),
        "vf_top_num_physical_vars": (i64, [vp]),
        "vf_top_get_grid_dims": (ci, [vp, ci, _ip]),
        "vf_top_set_python_filter": (ci, [vp, ci, FILTER_APPLY_CB, FILTER_BACKPROP_CB, vp]),
        "vf_top_set_residual_callback": (ci, [vp, PCG_CALLBACK, vp]),
        "vf_filter_upsample": (ci, [ci, _ip, ci, _dp, _dp]),
        "vf_filter_upsample_backprop": (ci, [ci, _ip, ci, _dp, _dp]),
        "vf_filter_vertex_to_cell": (ci, [ci, _ip, _dp, _dp]),
        "vf_filter_vertex_to_cell_backprop": (ci, [ci, _ip, _dp, _dp]),
        "vf_filter_langelaar": (ci, [ci, _ip, _dp, _dp, _dp]),
        "vf_filter_langelaar_backprop": (ci, [ci, _ip, _dp, _dp, _dp, _dp, _dp]),
        "vf_top_destroy": (ci, [vp]),
        "vf_top_set_solver": (ci, [vp, ci, cd, ci, ci, ci, ci]),
        "vf_top_set_vars": (ci, [vp, _dp]),
        "vf_top_get_vars": (ci, [vp, ci, _dp]),
        "vf_top_compliance": (ci, [vp, C.POINTER(cd)]),
        "vf_top_constraint": (ci, [vp, C.POINTER(cd)]),
        "vf_top_objective_gradient": (ci, [vp, _dp]),
        "vf_top_constraint_jacobian": (ci, [vp, _dp]),
        "vf_top_get_u": (ci, [vp, _dp]),
        "vf_top_last_pcg_iterations": (ci, [vp]),
        "vf_top_oc_step": (ci, [vp, cd, cd, cd, C.POINTER(ci)]),
        "vf_top_get_lambda_bracket": (ci, [vp, C.POINTER(cd), C.POINTER(cd)]),
        "vf_sim_create_slab": (ci, [ci, _ip, _dp, _dp, i64, i64, vp, pvp]),
        "vf_sim_window": (ci, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        "vf_mg_create_slab": (ci, [vp, ci, ci, pvp]),
        "vf_group_create_local": (ci, [ci, pvp, pvp]),
        "vf_nccl_unique_id": (ci, [vp]),
        "vf_group_create_nccl": (ci, [vp, ci, ci, vp, pvp]),
        "vf_group_destroy": (ci, [vp]),
        "vf_group_pcg_dev": (ci, [vp, pvp, pvp, ci, cd, ci, ci, ci, ci, C.POINTER(ci), _dp, PCG_CALLBACK, vp]),
        "vf_group_top_create": (ci, [vp, ci, _dp, cd, pvp]),
        "vf_group_top_destroy": (ci, [vp]),
        "vf_group_top_halo_layers": (i64, [vp]),
        "vf_group_top_set_solver": (ci, [vp, ci, cd, ci, ci, ci, ci]),
        "vf_group_top_set_vars": (ci, [vp, _dp]),
        "vf_group_top_get_vars": (ci, [vp, ci, _dp]),
        "vf_group_top_compliance": (ci, [vp, C.POINTER(cd)]),
        "vf_group_top_constraint": (ci, [vp, C.POINTER(cd)]),
        "vf_group_top_objective_gradient": (ci, [vp, _dp]),
        "vf_group_top_constraint_jacobian": (ci, [vp, _dp]),
        "vf_group_top_last_pcg_iterations": (ci, [vp]),
        "vf_group_top_get_u": (ci, [vp, ci, _dp]),
        "vf_group_top_oc_step": (ci, [vp, cd, cd, cd, C.POINTER(ci)]),
        "vf_q2_create": (ci, [ci, _ip, _dp, _dp, pvp]),
        "vf_q2_destroy": (ci, [vp]),
        "vf_q2_num_nodes": (i64, [vp]),
        "vf_q2_num_elements": (i64, [vp]),
        "vf_q2_set_isotropic": (ci, [vp, cd, cd]),
        "vf_q2_set_elasticity_tensor": (ci, [vp, _dp]),
        "vf_q2_get_K0": (ci, [vp, _dp]),
        "vf_q2_set_interpolation": (ci, [vp, ci, cd, cd, cd, cd]),
        "vf_q2_set_densities": (ci, [vp, _dp]),
        "vf_q2_get_young_moduli": (ci, [vp, _dp]),
        "vf_q2_apply_K": (ci, [vp, _dp, _dp, ci, ci]),
        "vf_q2_element_energies": (ci, [vp, _dp, _dp]),
        "vf_q2_pcg": (ci, [vp, _dp, _dp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), ci, cd, C.POINTER(ci), C.POINTER(cd)]),
        "vf_group_lbl_create": (ci, [vp, pvp]),
        "vf_group_lbl_destroy": (ci, [vp]),
        "vf_group_lbl_select_init_method": (ci, [vp, C.c_char_p]),
        "vf_group_lbl_run": (ci, [vp, ci, i64, ci, cd, ci, ci, ci, LBL_CALLBACK, vp]),
        "vf_group_lbl_objective": (ci, [vp, C.POINTER(cd)]),
        "vf_group_lbl_gradient": (ci, [vp, _dp]),
        "vf_group_lbl_num_layer_iterations": (ci, [vp, C.POINTER(ci)]),
        "vf_lbl_create": (ci, [vp, pvp]),
        "vf_lbl_destroy": (ci, [vp]),
        "vf_lbl_select_init_method": (ci, [vp, C.c_char_p]),
        "vf_lbl_run": (ci, [vp, ci, i64, ci, cd, ci, ci, ci, LBL_CALLBACK, vp, PCG_CALLBACK, vp]),
        "vf_lbl_get_layer_u": (ci, [vp, _dp]),
        "vf_lbl_get_layer_gradient": (ci, [vp, _dp]),
        "vf_top_oc_search": (ci, [vp, vp, cd, cd, cd, _dp, C.POINTER(ci)]),
        "vf_lbl_objective": (ci, [vp, C.POINTER(cd)]),
        "vf_lbl_gradient": (ci, [vp, _dp]),
        "vf_mma_create": (ci, [i64, ci, _dp, _dp, pvp]),
        "vf_mma_destroy": (ci, [vp]),
        "vf_mma_enable_gcmma": (ci, [vp, ci]),
        "vf_mma_set_initial_var": (ci, [vp, _dp]),
        "vf_mma_step": (ci, [vp, MMA_F_CALLBACK, MMA_DF_CALLBACK, vp, ci]),
        "vf_mma_get_optimal_var": (ci, [vp, _dp]),
        "vf_mma_get_optimal_var_dev": (ci, [vp, pvp]),
        "vf_mma_newton_iterations": (i64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def _check(rc):
    if rc != 0:
        msg = lib().vf_last_error().decode()
        if rc == 2:
            raise VoxelFEMError(msg)  # std::logic_error in the reference
        raise VoxelFEMError(msg)


def device_count():
    n = C.c_int(0)
    _check(lib().vf_device_count(C.byref(n)))
    return n.value


def parse_bc(path_or_dict, dmin, dmax):
    """Parse a VoxelFEM/MeshFEM .bc JSON file into box regions in absolute coordinates
    (rules of MeshFEM BoundaryConditions.cc:219-380 restricted to the `dirichlet[xyz]*` / `force`
    box regions the voxel simulator accepts, TensorProductSimulator.hh:600-652)."""
    cfg = path_or_dict
    if not isinstance(cfg, dict):
        with open(path_or_dict) as f:
            cfg = json.load(f)
    dmin = np.asarray(dmin, dtype=float)
    dmax = np.asarray(dmax, dtype=float)
    N = len(dmin)
    kinds, masks, vals, los, his = [], [], [], [], []

    def pad(v):
        v = [float(x) for x in v][:3]
        return v + [0.0] * (3 - len(v))

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
                raise VoxelFEMError("invalid mask")
            if comp:
                cm = sum(1 << "xyz".index(c) for c in set(comp))
            if rest[len(comp):] != "":
                raise VoxelFEMError("Invalid type '%s'" % t)
            kind = 0
        elif t == "force":
            kind = 1
        else:
            raise VoxelFEMError("Illegal constraint type, only \"dirichlet\" and \"force\" accepted")
        if "box%" in reg:
            lo = np.array(pad(reg["box%"]["minCorner"]))
            hi = np.array(pad(reg["box%"]["maxCorner"]))
            lo[:N] = dmin + lo[:N] * (dmax - dmin)
            hi[:N] = dmin + hi[:N] * (dmax - dmin)
        elif "box" in reg:
            lo = np.array(pad(reg["box"]["minCorner"]))
            hi = np.array(pad(reg["box"]["maxCorner"]))
        else:
            raise VoxelFEMError("only box / box% regions are supported")
        kinds.append(kind)
        masks.append(cm)
        vals.append(pad(reg["value"]))
        los.append(lo)
        his.append(hi)
    return (np.array(kinds, dtype=np.int32), np.array(masks, dtype=np.int32),
            np.ascontiguousarray(vals, dtype=np.float64), np.ascontiguousarray(los, dtype=np.float64),
            np.ascontiguousarray(his, dtype=np.float64))


def to_soa(u):
    """(numNodes, N) -> flat component-major VField storage."""
    return np.ascontiguousarray(np.asarray(u, dtype=np.float64).T).ravel()


def from_soa(flat, N):
    return np.ascontiguousarray(flat.reshape(N, -1).T)


class _Handle:
    """Owns one C-ABI handle and, strongly, the handles of the device objects built on top of it (a solver holds a raw
    pointer to its simulator, a problem to its solver).  close() destroys the dependents first.  Python's cycle collector
    finalises unreachable objects in arbitrary order and clears weak references before it does, so the order cannot be
    left to __del__: whichever wrapper dies first closes the whole subtree, the others find their handle already closed."""

    def __init__(self, destroy, h):
        self.destroy, self.h, self.children = destroy, h, []

    def adopt(self, child):
        self.children = [c for c in self.children if c.h] + [child]

    def close(self):
        for c in self.children:
            c.close()
        self.children = []
        if self.h:
            self.destroy(self.h)
            self.h = None


class _Owned:
    def _own(self, h, destroy, *parents):
        self._box = _Handle(destroy, h)
        for p in parents:
            p._box.adopt(self._box)

    @property
    def h(self):
        b = getattr(self, "_box", None)
        return b.h if b is not None else None

    def close(self):
        b = getattr(self, "_box", None)
        if b is not None:
            b.close()

    def __del__(self):
        self.close()


class DeviceArray:
    """A device-resident array of doubles (vf_dev_alloc)."""

    def __init__(self, n):
        self.L = lib()
        p = C.c_void_p()
        _check(self.L.vf_dev_alloc(int(n), C.byref(p)))
        self.ptr, self.n = p, int(n)

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64).ravel()
        assert host.size == self.n
        _check(self.L.vf_dev_upload(self.ptr, host, self.n))

    def download(self):
        out = np.empty(self.n)
        _check(self.L.vf_dev_download(out, self.ptr, self.n))
        return out

    def zero(self):
        _check(self.L.vf_dev_memset_zero(self.ptr, self.n))

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.vf_dev_free(self.ptr)
            self.ptr = None


class Sim(_Owned):
    """TensorProductSimulator<double,1,1[,1]> device state (vf_sim)."""

    def __init__(self, ne, dmin=None, dmax=None):
        self.L = lib()
        ne = np.ascontiguousarray(ne, dtype=np.int64)
        self.N = len(ne)
        if dmin is None:
            dmin = np.zeros(self.N)
            dmax = ne.astype(float)
        self.ne = ne
        self.dmin = np.ascontiguousarray(dmin, dtype=np.float64)
        self.dmax = np.ascontiguousarray(dmax, dtype=np.float64)
        h = C.c_void_p()
        _check(self.L.vf_sim_create(self.N, ne, self.dmin, self.dmax, C.byref(h)))
        self._own(h, self.L.vf_sim_destroy)

    @property
    def num_nodes(self): return self.L.vf_sim_num_nodes(self.h)
    @property
    def num_elements(self): return self.L.vf_sim_num_elements(self.h)
    @property
    def nn(self): return self.ne + 1

    def set_isotropic(self, E, nu): _check(self.L.vf_sim_set_isotropic(self.h, E, nu))
    def set_elasticity_tensor(self, D): _check(self.L.vf_sim_set_elasticity_tensor(self.h, np.ascontiguousarray(D, dtype=np.float64)))

    def K0(self):
        ke = self.N * 2 ** self.N
        out = np.zeros((ke, ke))
        _check(self.L.vf_sim_get_K0(self.h, out))
        return out

    def set_interp(self, law=0, E0=1.0, Emin=1e-4, gamma=3.0, q=3.0): _check(self.L.vf_sim_set_interpolation(self.h, law, E0, Emin, gamma, q))

    def set_gravity(self, g):
        g = list(g) + [0.0] * (3 - len(g))
        _check(self.L.vf_sim_set_gravity(self.h, np.ascontiguousarray(g, dtype=np.float64)))

    def set_densities(self, rho): _check(self.L.vf_sim_set_densities(self.h, np.ascontiguousarray(rho, dtype=np.float64).ravel()))
    def set_uniform_density(self, v): _check(self.L.vf_sim_set_uniform_density(self.h, v))

    def densities(self):
        out = np.zeros(self.num_elements); _check(self.L.vf_sim_get_densities(self.h, out)); return out

    def E(self):
        out = np.zeros(self.num_elements); _check(self.L.vf_sim_get_young_moduli(self.h, out)); return out

    def apply_bc_file(self, path):
        k, m, v, lo, hi = parse_bc(path, self.dmin, self.dmax)
        _check(self.L.vf_sim_apply_bc_regions(self.h, len(k), k, m, v, lo, hi))

    def add_dirichlet(self, u, lo, hi, cmask=7):
        pad = lambda a: np.ascontiguousarray(list(a) + [0.0] * (3 - len(a)), dtype=np.float64)
        _check(self.L.vf_sim_add_dirichlet_box(self.h, pad(u), pad(lo), pad(hi), cmask))

    def apply_symmetry_conditions(self, axes_mask, max_face_mask=0): _check(self.L.vf_sim_apply_symmetry_conditions(self.h, axes_mask, max_face_mask))

    def dirichlet_mask(self):
        out = np.zeros(self.num_nodes, dtype=np.uint8); _check(self.L.vf_sim_get_dirichlet_mask(self.h, out)); return out

    def num_force_nodes(self): return int(self.L.vf_sim_num_force_nodes(self.h))

    def dirichlet_conditions(self):
        """(nodes, component masks, values (n, N)) of the stored Dirichlet conditions, ascending node index."""
        n = int(self.L.vf_sim_num_dirichlet_nodes(self.h))
        nodes, masks, vals = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1), dtype=np.uint8), np.zeros(max(n, 1) * self.N)
        _check(self.L.vf_sim_get_dirichlet_conditions(self.h, nodes, masks, vals))
        return nodes[:n], masks[:n], vals[:n * self.N].reshape(n, self.N)

    def force_nodes(self):
        """(nodes, forces (n, N)) of the stored nodal forces."""
        n = self.num_force_nodes()
        nodes, f = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1) * self.N)
        _check(self.L.vf_sim_get_force_nodes(self.h, nodes, f))
        return nodes[:n], f[:n * self.N].reshape(n, self.N)
    def has_nonzero_dirichlet_values(self): return int(self.L.vf_sim_num_nonzero_dirichlet_values(self.h)) != 0

    def build_load(self):
        f = np.zeros(self.num_nodes * self.N); _check(self.L.vf_sim_build_load_vector(self.h, f)); return from_soa(f, self.N)

    def apply_K(self, u, out=None, zero_init=True, negate=False):
        o = np.zeros(self.num_nodes * self.N) if out is None else to_soa(out)
        _check(self.L.vf_sim_apply_K(self.h, to_soa(u), o, int(zero_init), int(negate)))
        return from_soa(o, self.N)

    def set_mask_layer(self, l): _check(self.L.vf_sim_set_mask_layer(self.h, l))

    def mask_info(self):
        a, b, h = C.c_int64(), C.c_int64(), C.c_double()
        _check(self.L.vf_sim_get_mask_info(self.h, C.byref(a), C.byref(b), C.byref(h)))
        return a.value, b.value

    def compliance_gradient(self, u, g=None):
        out = np.zeros(self.num_elements) if g is None else np.ascontiguousarray(g, dtype=np.float64).copy()
        _check(self.L.vf_sim_compliance_gradient(self.h, to_soa(u), out, int(g is not None)))
        return out

    def energy_density(self, u):
        out = np.zeros(self.num_elements); _check(self.L.vf_sim_element_energy_density(self.h, to_soa(u), out)); return out

    def solve(self, f):
        u = np.zeros(self.num_nodes * self.N); _check(self.L.vf_sim_solve(self.h, to_soa(f), u)); return from_soa(u, self.N)


class _LevelView:
    def __init__(self, mg, l):
        self.mg, self.l = mg, l

    @property
    def num_nodes(self): return self.mg.nn(self.l)

    def dirichlet_mask(self):
        out = np.zeros(self.num_nodes, dtype=np.uint8)
        _check(self.mg.L.vf_mg_level_dirichlet_mask(self.mg.h, self.l, out))
        return out


class MG(_Owned):
    """MultigridSolver (vf_mg)."""

    def __init__(self, sim, levels):
        self.L = lib()
        self.sim, self.N, self.levels = sim, sim.N, levels
        h = C.c_void_p()
        _check(self.L.vf_mg_create(sim.h, levels, C.byref(h)))
        self._own(h, self.L.vf_mg_destroy, sim)

    def get_sim(self, l): return _LevelView(self, l)
    def nn(self, l): return self.L.vf_mg_level_num_nodes(self.h, l)

    def coarsened_fine_K0(self, fi):
        ke = self.N * 2 ** self.N
        out = np.zeros((ke, ke)); _check(self.L.vf_mg_get_coarsened_fine_K0(self.h, fi, out)); return out

    def update_stiffness(self): _check(self.L.vf_mg_update_stiffness_matrices(self.h))

    def apply_K(self, l, u):
        out = np.zeros(self.nn(l) * self.N); _check(self.L.vf_mg_apply_K(self.h, l, to_soa(u), out)); return from_soa(out, self.N)

    def residual(self, l, u, b):
        out = np.zeros(self.nn(l) * self.N); _check(self.L.vf_mg_compute_residual(self.h, l, to_soa(u), to_soa(b), out)); return from_soa(out, self.N)

    def smooth(self, l, u, b, forward=True):
        uu = to_soa(u); _check(self.L.vf_mg_smooth(self.h, l, uu, to_soa(b), int(forward))); return from_soa(uu, self.N)

    def smooth_residual(self, l, u, b, forward=True):
        """One smoothing sweep that also returns computeResidual of its result (stored-stencil levels): (u, r)."""
        uu = to_soa(u); r = np.zeros_like(uu)
        _check(self.L.vf_mg_smooth_residual(self.h, l, uu, to_soa(b), int(forward), r))
        return from_soa(uu, self.N), from_soa(r, self.N)

    def restrict(self, lf, fine):
        out = np.zeros(self.nn(lf + 1) * self.N); _check(self.L.vf_mg_restrict(self.h, lf, to_soa(fine), out)); return from_soa(out, self.N)

    def interpolate(self, lf, coarse, fine=None):
        out = np.zeros(self.nn(lf) * self.N) if fine is None else to_soa(fine)
        _check(self.L.vf_mg_interpolate(self.h, lf, to_soa(coarse), out, int(fine is not None)))
        return from_soa(out, self.N)

    def stencil(self, l):
        ns = 3 ** self.N
        out = np.zeros(self.nn(l) * ns * self.N * self.N)
        _check(self.L.vf_mg_get_stencil(self.h, l, out))
        return out.reshape(self.nn(l), ns, self.N, self.N)

    def coarse_solve(self, f):
        out = np.zeros(self.nn(self.levels) * self.N); _check(self.L.vf_mg_coarse_solve(self.h, to_soa(f), out)); return from_soa(out, self.N)

    def solve(self, u, f, num_steps, num_smooth, stiffness_updated=False, zero_dirichlet=False, fmg=False):
        out = np.zeros(self.nn(0) * self.N)
        _check(self.L.vf_mg_solve(self.h, to_soa(u), to_soa(f), num_steps, num_smooth, int(stiffness_updated), int(zero_dirichlet), int(fmg), out))
        return from_soa(out, self.N)

    def pcg(self, u, b, max_iter, tol, mg_iterations=1, mg_smoothing=1, fmg=False, dirichlet_ok=False, callback=None):
        u0 = to_soa(u)
        x = np.empty_like(u0)
        it = C.c_int(0)
        res = np.zeros(max(max_iter, 1) + 1)
        cb, failed = _guarded(PCG_CALLBACK, (lambda i, r, _u: callback(i, r)) if callback else None)
        _check(self.L.vf_mg_pcg_io(self.h, u0, to_soa(b), x, max_iter, tol, mg_iterations, mg_smoothing, int(fmg), int(dirichlet_ok), C.byref(it), res, cb, None))
        _reraise(failed)
        return from_soa(x, self.N), it.value, res[:it.value]

    def pcg_dev(self, x_dev, b_dev, max_iter, tol, mg_iterations=1, mg_smoothing=1, fmg=False, dirichlet_ok=False):
        it = C.c_int(0)
        res = np.zeros(max(max_iter, 1) + 1)
        _check(self.L.vf_mg_pcg_dev(self.h, x_dev.ptr, b_dev.ptr, max_iter, tol, mg_iterations, mg_smoothing, int(fmg), int(dirichlet_ok), C.byref(it), res, PCG_CALLBACK(), None))
        return it.value, res[:it.value]

    def pcg_iterate(self):
        out = np.zeros(self.nn(0) * self.N); _check(self.L.vf_mg_get_pcg_iterate(self.h, out)); return from_soa(out, self.N)

    def pcg_residual(self):
        out = np.zeros(self.nn(0) * self.N); _check(self.L.vf_mg_get_pcg_residual(self.h, out)); return from_soa(out, self.N)

    def set_symmetric_gs(self, s): _check(self.L.vf_mg_set_symmetric_gauss_seidel(self.h, int(s)))
    def set_rebuild_every_solve(self, on=True): _check(self.L.vf_mg_set_rebuild_every_solve(self.h, int(on)))
    def set_mask_layer(self, l): _check(self.L.vf_mg_set_mask_layer(self.h, l))
    def decrement_mask(self, inc): _check(self.L.vf_mg_decrement_mask(self.h, inc))

    def debug_get(self, which, l):
        out = np.zeros(self.nn(l) * self.N)
        _check(self.L.vf_mg_debug_get(self.h, {"x": 0, "b": 1, "r": 2}[which], l, out))
        return from_soa(out, self.N)

    def debug_multicolor_visit(self):
        out = np.zeros(self.nn(0), dtype=np.int32); _check(self.L.vf_mg_debug_multicolor_visit(self.h, out)); return out

    def stream(self): return self.L.vf_mg_stream(self.h)
    def synchronize(self): _check(self.L.vf_mg_synchronize(self.h))

    OPS = {"smooth": 0, "residual": 1, "apply": 2, "restrict": 3, "prolong": 4, "coarse_solve": 5, "vcycle": 6, "fmg": 7, "smooth_residual": 8}

    def time_op(self, op, level=0, reps=10, nsmooth=1):
        ms = C.c_double()
        _check(self.L.vf_mg_time_op(self.h, self.OPS[op], level, reps, nsmooth, C.byref(ms)))
        return ms.value

    def prof_enable(self, on=True): _check(self.L.vf_prof_enable(self.h, int(on)))
    def prof_reset(self): _check(self.L.vf_prof_reset(self.h))

    def prof_report(self):
        out = {}
        for c in range(self.L.vf_prof_num_categories()):
            n, ms, un = C.c_int64(), C.c_double(), C.c_double()
            _check(self.L.vf_prof_get(self.h, c, C.byref(n), C.byref(ms), C.byref(un)))
            if n.value:
                out[self.L.vf_prof_name(c).decode()] = {"launches": n.value, "ms": ms.value, "units": un.value}
        return out


# ---------------------------------------------------------------------------------------------
# Slab-partitioned solver (multi-GPU)
# ---------------------------------------------------------------------------------------------
def slab_ranges(ne0, nparts, align):
    """Even split of ne0 element layers into nparts slabs whose boundaries are multiples of `align`."""
    units = ne0 // align
    assert units * align == ne0 and units >= nparts, "grid not divisible into %d slabs aligned to %d" % (nparts, align)
    cuts = [align * ((units * i) // nparts) for i in range(nparts + 1)]
    return [(cuts[i], cuts[i + 1]) for i in range(nparts)]


def slab_window(ne0, slab_begin, slab_end):
    """Node planes stored (plane_lo, plane_hi) and owned (own_lo, own_hi) by the part holding element layers
    [slab_begin, slab_end) of ne0 (mirrors vf_sim_create_slab): one ghost plane per neighbour; a plane shared by two
    slabs is owned by the lower one."""
    plane_lo, plane_hi = max(slab_begin - 1, 0), min(slab_end + 1, ne0)
    own_hi = slab_end if slab_end == ne0 else slab_end - 1
    return plane_lo, plane_hi, slab_begin, own_hi


class SlabSim(Sim):
    """One slab of a TensorProductSimulator: stores the window of node planes [plane_lo, plane_hi] of the global grid."""

    def __init__(self, ne_global, dmin, dmax, slab_begin, slab_end, share_stream_with=None):
        self.L = lib()
        self.ne_global = np.ascontiguousarray(ne_global, dtype=np.int64)
        self.N = len(self.ne_global)
        self.dmin = np.ascontiguousarray(dmin, dtype=np.float64)
        self.dmax = np.ascontiguousarray(dmax, dtype=np.float64)
        h = C.c_void_p()
        _check(self.L.vf_sim_create_slab(self.N, self.ne_global, self.dmin, self.dmax, int(slab_begin), int(slab_end),
                                         share_stream_with.h if share_stream_with is not None else None, C.byref(h)))
        self._own(h, self.L.vf_sim_destroy)
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(self.L.vf_sim_window(h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.plane_lo, self.plane_hi, self.own_lo, self.own_hi = a.value, b.value, c.value, d.value
        self.slab = (int(slab_begin), int(slab_end))
        self.ne = self.ne_global.copy()
        self.ne[0] = self.plane_hi - self.plane_lo

    def window_of_nodal(self, field):
        """Slice of a global (numNodes, N) nodal field stored by this part."""
        nn = self.ne_global + 1
        return np.ascontiguousarray(field.reshape(tuple(nn) + (self.N,))[self.plane_lo:self.plane_hi + 1]).reshape(-1, self.N)

    def window_of_elements(self, values):
        return np.ascontiguousarray(np.asarray(values).reshape(tuple(self.ne_global))[self.plane_lo:self.plane_hi]).ravel()


class SlabMG(MG):
    def __init__(self, sim, levels, first_replicated_level):
        self.L = lib()
        self.sim, self.N, self.levels = sim, sim.N, levels
        h = C.c_void_p()
        _check(self.L.vf_mg_create_slab(sim.h, levels, first_replicated_level, C.byref(h)))
        self._own(h, self.L.vf_mg_destroy, sim)


class SlabGroup(_Owned):
    """A MultigridSolver cut into slabs along axis 0: a local group (all parts in this process, one device) or one NCCL rank."""

    def __init__(self, parts, rank=None, world=None, unique_id=None):
        self.L = lib()
        self.parts = list(parts)
        h = C.c_void_p()
        if unique_id is None:
            arr = (C.c_void_p * len(self.parts))(*[p.h for p in self.parts])
            _check(self.L.vf_group_create_local(len(self.parts), arr, C.byref(h)))
        else:
            assert len(self.parts) == 1
            buf = C.create_string_buffer(bytes(unique_id), 128)
            _check(self.L.vf_group_create_nccl(self.parts[0].h, rank, world, buf, C.byref(h)))
        self._own(h, self.L.vf_group_destroy, *self.parts)

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        _check(lib().vf_nccl_unique_id(buf))
        return buf.raw

    def pcg_dev(self, xs, bs, max_iter, tol, mg_iterations=1, mg_smoothing=1, fmg=False, dirichlet_ok=False):
        it = C.c_int(0)
        res = np.zeros(max(max_iter, 1) + 1)
        xa = (C.c_void_p * len(xs))(*[x.ptr for x in xs])
        ba = (C.c_void_p * len(bs))(*[b.ptr for b in bs])
        _check(self.L.vf_group_pcg_dev(self.h, xa, ba, max_iter, tol, mg_iterations, mg_smoothing, int(fmg), int(dirichlet_ok), C.byref(it), res, PCG_CALLBACK(), None))
        return it.value, res[:it.value]

class Problem(_Owned):
    """TopologyOptimizationProblem + MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer (vf_top)."""

    KINDS = {"smooth": 0, "project": 1, "upsample": 2, "vertex_to_cell": 3, "langelaar": 4, "python": 5}

    def __init__(self, mg, filters, vol_frac):
        """filters: ("smooth", radius, type) | ("project", beta) | ("upsample", factor) | ("vertex_to_cell",) | ("langelaar",) |
        ("python", apply(in) -> out, backprop(d_dout, vars) -> d_din)"""
        self.L = lib()
        self.mg = mg
        spec, self._py = [], []
        for i, f in enumerate(filters):
            k = self.KINDS[f[0]]
            if k == 0:
                spec += [0, f[1], f[2], 0.0]
            elif k == 1:
                spec += [1, 0, 0, f[1]]
            elif k == 2:
                spec += [2, f[1], 0, 0.0]
            else:
                spec += [k, 0, 0, 0.0]
            if k == 5:
                self._py.append((i, f[1], f[2]))
        spec = np.ascontiguousarray(spec if spec else [0.0], dtype=np.float64)
        h = C.c_void_p()
        _check(self.L.vf_top_create(mg.h, len(filters), spec, vol_frac, C.byref(h)))
        self._own(h, self.L.vf_top_destroy, mg)
        self.ne = mg.sim.num_elements
        self.N = mg.N
        self.nv = int(self.L.vf_top_num_vars(self.h))
        self._cbs, self._failed = [], []
        for i, ap, bp in self._py:
            def a_cb(pin, nin, pout, nout, _u, ap=ap):
                try:
                    np.ctypeslib.as_array(pout, (nout,))[:] = np.asarray(ap(np.ctypeslib.as_array(pin, (nin,)).copy()), dtype=np.float64).ravel(); return 0
                except BaseException as e:   # noqa: BLE001
                    self._failed.append(e); return 1

            def b_cb(pg, nout, pv, nin, pout, _u, bp=bp):
                try:
                    np.ctypeslib.as_array(pout, (nin,))[:] = np.asarray(bp(np.ctypeslib.as_array(pg, (nout,)).copy(), np.ctypeslib.as_array(pv, (nin,)).copy()), dtype=np.float64).ravel(); return 0
                except BaseException as e:   # noqa: BLE001
                    self._failed.append(e); return 1
            ca, cb = FILTER_APPLY_CB(a_cb), FILTER_BACKPROP_CB(b_cb)
            self._cbs += [ca, cb]
            _check(self.L.vf_top_set_python_filter(self.h, i, ca, cb, None))

    def _call(self, rc):
        if self._failed:
            e = self._failed[0]; self._failed.clear(); raise e
        _check(rc)

    def set_residual_callback(self, fn):
        """fn(iteration, residual_norm) after every PCG iteration of the problem's solves; None clears it."""
        def guarded(i, r, _u):
            try:
                fn(i, r)
            except BaseException as e:   # noqa: BLE001
                self._failed.append(e)
        self._res_cb = PCG_CALLBACK(guarded) if fn is not None else PCG_CALLBACK()
        _check(self.L.vf_top_set_residual_callback(self.h, self._res_cb, None))

    def grid_dims(self, physical=False):
        d = np.zeros(self.N, dtype=np.int64); _check(self.L.vf_top_get_grid_dims(self.h, int(physical), d)); return d

    def set_solver(self, cg_iter=100, tol=1e-5, mg_it=1, mg_smooth=2, fmg=True, zero_init=False):
        _check(self.L.vf_top_set_solver(self.h, cg_iter, tol, mg_it, mg_smooth, int(fmg), int(zero_init)))

    def set_vars(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        if x.size != self.nv: raise VoxelFEMError("Variable size mismatch")
        self._call(self.L.vf_top_set_vars(self.h, x))

    def design_vars(self):
        o = np.zeros(self.nv); _check(self.L.vf_top_get_vars(self.h, 0, o)); return o

    def physical_vars(self):
        o = np.zeros(self.ne); _check(self.L.vf_top_get_vars(self.h, 1, o)); return o

    def compliance(self):
        v = C.c_double(); _check(self.L.vf_top_compliance(self.h, C.byref(v))); return v.value

    def constraint(self):
        v = C.c_double(); _check(self.L.vf_top_constraint(self.h, C.byref(v))); return v.value

    def objective_gradient(self):
        o = np.zeros(self.nv); self._call(self.L.vf_top_objective_gradient(self.h, o)); return o

    def constraint_jacobian(self):
        o = np.zeros(self.nv); self._call(self.L.vf_top_constraint_jacobian(self.h, o)); return o

    def u(self):
        o = np.zeros(self.mg.nn(0) * self.N); _check(self.L.vf_top_get_u(self.h, o)); return from_soa(o, self.N)

    def last_pcg_iters(self): return self.L.vf_top_last_pcg_iterations(self.h)

    def oc_step(self, m=0.2, p=0.5, ctol=1e-6):
        n = C.c_int(0); self._call(self.L.vf_top_oc_step(self.h, m, p, ctol, C.byref(n))); return n.value

    def oc_search(self, dJ=None, m=0.2, p=0.5, ctol=1e-6):
        """Bracket + bisection of OCOptimizer::step with a caller-supplied objective gradient; returns the stepped variables
        without setting them (vf_top_oc_search)."""
        n = C.c_int(0); out = np.zeros(self.nv)
        g = None if dJ is None else np.ascontiguousarray(dJ, dtype=np.float64).ravel()
        self._call(self.L.vf_top_oc_search(self.h, None if g is None else g.ctypes.data_as(C.c_void_p), m, p, ctol, out, C.byref(n)))
        return out, n.value

    def lambda_bracket(self):
        a, b = C.c_double(), C.c_double(); _check(self.L.vf_top_get_lambda_bracket(self.h, C.byref(a), C.byref(b))); return a.value, b.value


def smoothing_filter(x, shape, radius, ftype):
    shape = np.ascontiguousarray(shape, dtype=np.int64)
    out = np.zeros(int(np.prod(shape)))
    _check(lib().vf_filter_smooth(len(shape), shape, radius, ftype, np.ascontiguousarray(x, dtype=np.float64).ravel(), out))
    return out


def _fa(a): return np.ascontiguousarray(a, dtype=np.float64).ravel()
def _sa(shape): return np.ascontiguousarray(shape, dtype=np.int64)


def upsample_filter(x, coarse_shape, factor):                         # UpsampleFilter::apply
    cs = _sa(coarse_shape); out = np.zeros(int(np.prod((cs - 1) * factor + 1)))
    _check(lib().vf_filter_upsample(len(cs), cs, int(factor), _fa(x), out)); return out


def upsample_filter_backprop(g, coarse_shape, factor):
    cs = _sa(coarse_shape); out = np.zeros(int(np.prod(cs)))
    _check(lib().vf_filter_upsample_backprop(len(cs), cs, int(factor), _fa(g), out)); return out


def vertex_to_cell_filter(x, vertex_shape):                           # VertexToCellFilter::apply
    vs = _sa(vertex_shape); out = np.zeros(int(np.prod(vs - 1)))
    _check(lib().vf_filter_vertex_to_cell(len(vs), vs, _fa(x), out)); return out


def vertex_to_cell_filter_backprop(g, vertex_shape):
    vs = _sa(vertex_shape); out = np.zeros(int(np.prod(vs)))
    _check(lib().vf_filter_vertex_to_cell_backprop(len(vs), vs, _fa(g), out)); return out


def langelaar_filter(x, shape, out_prev=None):
    """LangelaarFilter::apply -> (filtered, smax).  out_prev: previous content of the output array (zeros for a fresh filter)."""
    sz = _sa(shape); n = int(np.prod(sz))
    out = np.zeros(n) if out_prev is None else _fa(out_prev).copy(); smax = np.zeros(n)
    _check(lib().vf_filter_langelaar(len(sz), sz, _fa(x), out, smax)); return out, smax


def langelaar_filter_backprop(g, vars_, filtered, smax, shape):
    sz = _sa(shape); out = np.zeros(int(np.prod(sz)))
    _check(lib().vf_filter_langelaar_backprop(len(sz), sz, _fa(g), _fa(vars_), _fa(filtered), _fa(smax), out)); return out


def projection_apply(x, beta):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel(); out = np.zeros_like(x)
    _check(lib().vf_filter_project(len(x), beta, x, out)); return out


def projection_backprop(g, vars_, beta):
    g = np.ascontiguousarray(g, dtype=np.float64).ravel(); out = np.zeros_like(g)
    _check(lib().vf_filter_project_backprop(len(g), beta, g, np.ascontiguousarray(vars_, dtype=np.float64).ravel(), out)); return out


class SimQ2(_Owned):
    """TensorProductSimulator<double, 2, 2[, 2]> on the reference's generic element path (vf_q2_*, csrc/vf_q2.cu): K0, applyK, element
    energies and a Jacobi-preconditioned CG.  Nodal fields are (numNodes, N) arrays over the (2 ne + 1)^N node grid."""

    def __init__(self, ne, dmin=None, dmax=None):
        self.L = lib()
        self.ne = np.ascontiguousarray(ne, dtype=np.int64); self.N = len(self.ne)
        dmin = np.zeros(self.N) if dmin is None else np.ascontiguousarray(dmin, dtype=np.float64)
        dmax = self.ne.astype(np.float64) if dmax is None else np.ascontiguousarray(dmax, dtype=np.float64)
        h = C.c_void_p()
        _check(self.L.vf_q2_create(self.N, self.ne, dmin, dmax, C.byref(h)))
        self._own(h, self.L.vf_q2_destroy)
        self.num_nodes, self.num_elements = int(self.L.vf_q2_num_nodes(self.h)), int(self.L.vf_q2_num_elements(self.h))

    def set_isotropic(self, E, nu): _check(self.L.vf_q2_set_isotropic(self.h, E, nu))
    def set_elasticity_tensor(self, D): _check(self.L.vf_q2_set_elasticity_tensor(self.h, np.ascontiguousarray(D, dtype=np.float64)))
    def set_interp(self, law=0, E0=1.0, Emin=1e-4, gamma=3.0, q=3.0): _check(self.L.vf_q2_set_interpolation(self.h, law, E0, Emin, gamma, q))
    def set_densities(self, rho): _check(self.L.vf_q2_set_densities(self.h, np.ascontiguousarray(rho, dtype=np.float64).ravel()))

    def K0(self):
        k = self.N * 3 ** self.N
        out = np.zeros((k, k)); _check(self.L.vf_q2_get_K0(self.h, out)); return out

    def E(self):
        out = np.zeros(self.num_elements); _check(self.L.vf_q2_get_young_moduli(self.h, out)); return out

    def apply_K(self, u, out=None, zero_init=True, negate=False):
        o = np.zeros(self.num_nodes * self.N) if out is None else to_soa(out)
        _check(self.L.vf_q2_apply_K(self.h, to_soa(u), o, int(zero_init), int(negate)))
        return from_soa(o, self.N)

    def element_energies(self, u):
        out = np.zeros(self.num_elements); _check(self.L.vf_q2_element_energies(self.h, to_soa(u), out)); return out

    def pcg(self, x0, b, fixed, max_iter=2000, tol=1e-10):
        """-> (x, iterations, relative residual); fixed: boolean (numNodes, N) mask of clamped components."""
        x = to_soa(x0).copy(); it, rr = C.c_int(0), C.c_double(0)
        fx = np.ascontiguousarray(np.asarray(fixed, dtype=np.uint8).T).ravel()
        _check(self.L.vf_q2_pcg(self.h, x, to_soa(b), fx, max_iter, tol, C.byref(it), C.byref(rr)))
        return from_soa(x, self.N), it.value, rr.value


class LBL(_Owned):
    """LayerByLayerEvaluator (LayerByLayer.hh:25-309) on the GPU."""

    def __init__(self, mg):
        self.L = lib(); self.mg = mg
        h = C.c_void_p()
        _check(self.L.vf_lbl_create(mg.h, C.byref(h)))
        self._own(h, self.L.vf_lbl_destroy, mg)

    def select_init_method(self, m): _check(self.L.vf_lbl_select_init_method(self.h, m.encode()))

    def run(self, zero_init=True, layer_increment=1, max_iter=50, tol=1e-5, mg_iterations=1, mg_smoothing=1, fmg=False, callback=None,
            pcg_callback=None):
        """callback(layer, compliance, pcg_iterations) per layer -- layer_u() / layer_gradient() are valid inside it;
        pcg_callback(it, residual_norm) per PCG iteration of every layer (mg.pcg_iterate() / mg.pcg_residual() valid inside)."""
        its, cs = [], []

        def _cb(layer, compliance, iters, _):
            its.append(iters); cs.append(compliance)
            if callback is not None:
                callback(layer, compliance, iters)
        cb, failed = _guarded(LBL_CALLBACK, _cb)
        pcb, pfailed = _guarded(PCG_CALLBACK, (lambda i, r, _u: pcg_callback(i, r)) if pcg_callback else None)
        _check(self.L.vf_lbl_run(self.h, int(zero_init), layer_increment, max_iter, tol, mg_iterations, mg_smoothing, int(fmg), cb, None, pcb, None))
        _reraise(failed); _reraise(pfailed)
        return np.array(its, dtype=np.int32), np.array(cs)

    def layer_u(self):
        n = self.mg.nn(0) * self.mg.N
        out = np.zeros(n); _check(self.L.vf_lbl_get_layer_u(self.h, out)); return from_soa(out, self.mg.N)

    def layer_gradient(self):
        g = np.zeros(self.mg.sim.num_elements); _check(self.L.vf_lbl_get_layer_gradient(self.h, g)); return g

    def objective(self):
        v = C.c_double(0); _check(self.L.vf_lbl_objective(self.h, C.byref(v))); return v.value

    def gradient(self):
        g = np.zeros(self.mg.sim.num_elements); _check(self.L.vf_lbl_gradient(self.h, g)); return g


class SlabLBL(_Owned):
    """LayerByLayerEvaluator (LayerByLayer.hh:25-309) on a slab group: vf_group_lbl_* (every slab holds a piece of every layer)."""

    def __init__(self, group, ne_global):
        self.L = lib(); self.group = group; self.ne_global = int(np.prod(ne_global))
        h = C.c_void_p()
        _check(self.L.vf_group_lbl_create(group.h, C.byref(h)))
        self._own(h, self.L.vf_group_lbl_destroy, group)

    def select_init_method(self, m): _check(self.L.vf_group_lbl_select_init_method(self.h, m.encode()))

    def run(self, zero_init=True, layer_increment=1, max_iter=50, tol=1e-5, mg_iterations=1, mg_smoothing=1, fmg=False, callback=None):
        """-> (PCG iterations per layer, compliance per layer); callback(layer, compliance, pcg_iterations)."""
        its, cs = [], []

        def _cb(layer, compliance, iters, _):
            its.append(iters); cs.append(compliance)
            if callback is not None: callback(layer, compliance, iters)
        cb, failed = _guarded(LBL_CALLBACK, _cb)
        _check(self.L.vf_group_lbl_run(self.h, int(zero_init), layer_increment, max_iter, tol, mg_iterations, mg_smoothing, int(fmg), cb, None))
        _reraise(failed)
        return np.array(its, dtype=np.int32), np.array(cs)

    def objective(self):
        v = C.c_double(0); _check(self.L.vf_group_lbl_objective(self.h, C.byref(v))); return v.value

    def gradient(self):
        g = np.zeros(self.ne_global); _check(self.L.vf_group_lbl_gradient(self.h, g)); return g


class MMA:
    """pyOptimizer.MMA (python_bindings/Optimizer.cc:11-23): MMA(numVars, numConstr, xmin, xmax, f, df_dx)."""

    def __init__(self, numVars, numConstr, xmin, xmax, f, df_dx):
        self.L = lib(); self.n, self.m = int(numVars), int(numConstr)
        n, m = self.n, self.m
        self._err = None

        def _f(xp, out, _):
            try:
                x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
                np.ctypeslib.as_array(out, shape=(m + 1,))[:] = np.asarray(f(x), dtype=np.float64).ravel()
                return 0
            except Exception as e:  # surfaced by step()
                self._err = e
                return 1

        def _df(xp, out, _):
            try:
                x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
                np.ctypeslib.as_array(out, shape=(m + 1, n))[:] = np.asarray(df_dx(x), dtype=np.float64).reshape(m + 1, n)
                return 0
            except Exception as e:
                self._err = e
                return 1
        self._cbs = (MMA_F_CALLBACK(_f), MMA_DF_CALLBACK(_df))
        h = C.c_void_p()
        _check(self.L.vf_mma_create(n, m, np.ascontiguousarray(xmin, dtype=np.float64), np.ascontiguousarray(xmax, dtype=np.float64), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vf_mma_destroy(self.h); self.h = None

    def enableGCMMA(self, enable): _check(self.L.vf_mma_enable_gcmma(self.h, int(enable)))
    def setInitialVar(self, x): _check(self.L.vf_mma_set_initial_var(self.h, np.ascontiguousarray(x, dtype=np.float64)))

    def step(self):
        self._err = None
        rc = self.L.vf_mma_step(self.h, self._cbs[0], self._cbs[1], None, 0)
        if self._err is not None:
            raise self._err
        _check(rc)

    def getOptimalVar(self):
        x = np.zeros(self.n); _check(self.L.vf_mma_get_optimal_var(self.h, x)); return x

    def newton_iterations(self): return self.L.vf_mma_newton_iterations(self.h)


def slab_halo_range(sb, se, ne0, R):
    """Element layers [elo, ehi) a slab owning [sb, se) holds once R halo layers per neighbour are attached (clipped at the grid)."""
    return max(0, sb - R), min(ne0, se + R)


class SlabProblem(_Owned):
    """Compliance topology optimization (TopologyOptimizationProblem + MultigridComplianceObjective + TotalVolumeConstraint +
    OCOptimizer, TopologyOptimizationProblem.hh:17-155, OptimalityCriterion.hh:38-149) on a grid partitioned into slabs along axis 0
    (BASELINE.json configs[3]): the vf_group_top_* entry points of the C ABI.  Every part owns the design variables of its element
    layers; filter halos, the volume / compliance all-reduces and the OC bracket / bisection all run inside the library (device
    copies in a local group, NCCL between ranks).  `parts`: list of (SlabSim, SlabMG) living in this process -- all parts of a local
    group, or this rank's single part of an NCCL group.  Whole-grid host arrays cross the boundary."""

    def __init__(self, parts, group, filters, vol_frac, dist=None):
        self.L = lib()
        self.parts, self.group = list(parts), group
        s0 = self.parts[0][0]
        self.N, self.gne = s0.N, [int(v) for v in s0.ne_global]
        self.ne_global = int(np.prod(self.gne))
        spec = []
        for f in filters:
            if f[0] == "smooth": spec += [0, f[1], f[2], 0.0]
            elif f[0] == "project": spec += [1, 0, 0, f[1]]
            else: raise VoxelFEMError("slab-partitioned problems support the Smoothing and Projection filters")
        spec = np.ascontiguousarray(spec if spec else [0.0], dtype=np.float64)
        h = C.c_void_p()
        _check(self.L.vf_group_top_create(group.h, len(filters), spec, float(vol_frac), C.byref(h)))
        self._own(h, self.L.vf_group_top_destroy, group)
        self.R = int(self.L.vf_group_top_halo_layers(self.h))
        self.stream_handle = self.L.vf_sim_stream(s0.h)

    @property
    def last_pcg_iters(self): return int(self.L.vf_group_top_last_pcg_iterations(self.h))

    def set_solver(self, cg_iter=100, tol=1e-5, mg_it=1, mg_smooth=2, fmg=True, zero_init=False):
        _check(self.L.vf_group_top_set_solver(self.h, cg_iter, tol, mg_it, mg_smooth, int(fmg), int(zero_init)))

    def set_vars(self, x_global):
        """Design variables of the whole grid (host array); every part keeps the layers it owns."""
        x = np.ascontiguousarray(x_global, dtype=np.float64).ravel()
        assert x.size == self.ne_global
        _check(self.L.vf_group_top_set_vars(self.h, x))

    def _get(self, fn, *a):
        out = np.zeros(self.ne_global); _check(fn(self.h, *a, out)); return out

    def design_vars(self): return self._get(self.L.vf_group_top_get_vars, 0)
    def physical_vars(self): return self._get(self.L.vf_group_top_get_vars, 1)
    def objective_gradient(self): return self._get(self.L.vf_group_top_objective_gradient)
    def constraint_jacobian(self): return self._get(self.L.vf_group_top_constraint_jacobian)

    def compliance(self):
        v = C.c_double(0); _check(self.L.vf_group_top_compliance(self.h, C.byref(v))); return v.value

    def constraint(self):
        v = C.c_double(0); _check(self.L.vf_group_top_constraint(self.h, C.byref(v))); return v.value

    def u_window(self, part=0):
        """Displacement window of local part `part`, (nodes of the window, 3)."""
        s = self.parts[part][0]
        out = np.zeros(int(np.prod([int(v) + 1 for v in s.ne])) * 3)
        _check(self.L.vf_group_top_get_u(self.h, part, out)); return from_soa(out, 3)

    def oc_step(self, m=0.2, p=0.5, ctol=1e-6):
        """OCOptimizer::step (OptimalityCriterion.hh:51-134); returns the number of constraint evaluations."""
        n = C.c_int(0); _check(self.L.vf_group_top_oc_step(self.h, m, p, ctol, C.byref(n))); return n.value
