"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/voxelfem_b200.h declares, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "voxelfem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", text))
    names -= {"vf_pcg_callback", "vf_lbl_callback", "vf_mma_f_callback", "vf_mma_df_callback"}
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from voxelfem_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) > 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_signatures_cover_header():
    from voxelfem_b200 import capi
    L = capi.lib()
    for n in declared_symbols():
        assert getattr(L, n).argtypes is not None or n in ("vf_last_error", "vf_version"), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from voxelfem_b200 import capi
    import numpy as np
    with pytest.raises(capi.VoxelFEMError, match="CUDA"):
        capi.Sim(np.array([4, 4]))


def test_product_does_not_reference_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "voxelfem_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hh", ".cc", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                if "oracle" in txt.lower() and fn != "__init__.py":
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_handle_tree_destroys_dependents_first():
    """The ctypes layer owns C-ABI handles in a tree (simulator -> solvers -> problems): whichever wrapper is finalised first,
    dependents are destroyed before what they point into, and every handle is destroyed exactly once."""
    from voxelfem_b200.capi import _Handle
    log = []
    sim = _Handle(lambda h: log.append(("sim", h)), 1)
    mg1, mg2 = _Handle(lambda h: log.append(("mg", h)), 2), _Handle(lambda h: log.append(("mg", h)), 3)
    top = _Handle(lambda h: log.append(("top", h)), 4)
    sim.adopt(mg1); sim.adopt(mg2); mg1.adopt(top)
    mg2.close()                                  # a solver dropped on its own
    sim.close()                                  # the simulator goes first (cycle collector order): subtree first
    top.close(); mg1.close(); sim.close()        # late finalisers find their handles closed
    assert log == [("mg", 3), ("top", 4), ("mg", 2), ("sim", 1)]
    assert sim.h is None and mg1.h is None and top.h is None
