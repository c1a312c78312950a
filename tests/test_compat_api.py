"""The pyVoxelFEM / pyOptimizer compatibility modules expose the API surface of the reference's pybind11 modules
(python_bindings/VoxelFEM.cc, Optimizer.cc).  The expected names are a committed fixture extracted from the reference
sources by tests/golden/make_api_fixture.py; methods outside the hot path must exist and raise NotImplementedError by name."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "voxelfem_b200", "compat"))
API = json.load(open(os.path.join(ROOT, "tests", "golden", "pyvoxelfem_api.json")))


def _classes():
    import pyVoxelFEM as m
    return {"TensorProductSimulator": m._TPS, "MultigridSolver": m._MG, "TopologyOptimizationProblem": m._TOProblem,
            "MultigridComplianceObjective": m._MGComplianceObjective, "LayerByLayerEvaluator": m._LBL, "OCOptimizer": m._OCOptimizer,
            "FilterChain": m.FilterChain, "ProjectionFilter": m.ProjectionFilter, "SmoothingFilter": m.SmoothingFilter,
            "Filter": m.SmoothingFilter, "TotalVolumeConstraint": m.TotalVolumeConstraint}


def test_module_level_names():
    import pyVoxelFEM as m
    missing = [n for n in API["module"] if not hasattr(m, n)]
    assert not missing, missing
    assert m.InterpolationLaw.SIMP == m.SIMP and m.InterpolationLaw.RAMP == m.RAMP
    assert {"DOUBLE", "FLOAT"} <= set(m.NumberType.__members__)
    assert {"Const", "Linear"} <= set(m.SmoothingFilter.Type.__members__)


@pytest.mark.parametrize("cls", sorted(set(API) - {"module", "pyOptimizer.MMA", "Constraint", "ComplianceObjective", "PythonFilter",
                                                    "UpsampleFilter", "VertexToCellFilter", "LangelaarFilter"}))
def test_class_surface(cls):
    import inspect
    c = _classes()[cls]
    # instance attributes set in __init__ (def_readwrite members) count as well
    src = inspect.getsource(c)
    missing = [n for n in API[cls] if not hasattr(c, n) and ("self.%s" % n) not in src]
    assert not missing, (cls, missing)


def test_mma_surface():
    import pyOptimizer
    missing = [n for n in API["pyOptimizer.MMA"] if not hasattr(pyOptimizer.MMA, n)]
    assert not missing, missing


def test_out_of_scope_names_fail_loudly():
    import pyVoxelFEM as m
    for name in ("PythonFilter", "UpsampleFilter", "VertexToCellFilter", "LangelaarFilter"):
        with pytest.raises(NotImplementedError, match=name):
            getattr(m, name)()
    with pytest.raises(NotImplementedError, match="MultigridComplianceObjective"):
        m.ComplianceObjective(None)
    with pytest.raises(NotImplementedError, match="getK"):
        m._TPS.getK(object.__new__(m._TPS))
    with pytest.raises(RuntimeError, match="No template instantiation"):
        m.TensorProductSimulator([2, 2], [[0, 0], [1, 1]], [4, 4])


def test_host_side_filter_helpers():
    import pyVoxelFEM as m
    pf = m.ProjectionFilter(4.0)
    import math
    for v in (0.1, 0.5, 0.9):   # invert(apply(x)) == x for the closed form (TopologyOptimizationFilter.hh:199-232)
        y = (math.tanh(2.0) + math.tanh(4.0 * (v - 0.5))) / (2 * math.tanh(2.0))
        assert abs(pf.invert(y) - v) < 1e-12
    with pytest.raises(RuntimeError, match="positive"):
        pf.beta = -1
    with pytest.raises(RuntimeError, match="domain error"):
        pf.invert(1.5)
