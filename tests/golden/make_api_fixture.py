"""Extracts the Python API surface the reference's pybind11 modules define (class -> method / property names) from
python_bindings/VoxelFEM.cc and Optimizer.cc into tests/golden/pyvoxelfem_api.json.  Run in the build container
(needs /root/reference); the fixture travels with the repo.      python tests/golden/make_api_fixture.py"""
import json
import os
import re

REF = "/root/reference/python_bindings"
out = {}
src = open(os.path.join(REF, "VoxelFEM.cc")).read()
src = re.sub(r"//[^\n]*", "", src)
# split the source at every py::class_< ... > / py::enum_ statement; names are taken from the nameMangler / literal argument
stmts = re.split(r"(?=py::class_<|py::enum_<)", src)
alias = {"TPS": "TensorProductSimulator", "MG": "MultigridSolver", "TOProblem": "TopologyOptimizationProblem", "CO": "ComplianceObjective",
         "MGCO": "MultigridComplianceObjective", "LBL": "LayerByLayerEvaluator", "OCO": "OCOptimizer", "Filter_": "Filter", "FC": "FilterChain",
         "PyF": "PythonFilter", "PF": "ProjectionFilter", "SF": "SmoothingFilter", "UF": "UpsampleFilter", "VCF": "VertexToCellFilter",
         "LF": "LangelaarFilter", "TVC": "TotalVolumeConstraint", "C": "Constraint"}
for st in stmts:
    m = re.match(r"py::class_<\s*(\w+)", st)
    if not m:
        continue
    body = re.split(r"\n\s*;\s*\n", st)[0]   # statement ends at a line holding only ";"
    name = alias.get(m.group(1))
    if name is None:
        continue
    names = re.findall(r"\.def(?:_property(?:_readonly)?|_readwrite|_readonly)?\(\s*\"(\w+)\"", body)
    out.setdefault(name, [])
    out[name] = sorted(set(out[name]) | set(names))
# SmoothingFilter's methods are attached to the pySF variable after the enum
m = re.search(r"pySF\s*\n\s*(\.def.*?);", src, re.S)
out["SmoothingFilter"] = sorted(set(out.get("SmoothingFilter", [])) | set(re.findall(r"\"(\w+)\"", m.group(1))) - {"Const", "Linear"})
out["module"] = sorted(set(re.findall(r"\bm\.def\(\s*\"(\w+)\"", src)) | {"InterpolationLaw", "NumberType", "FilterChain", "ProjectionFilter",
                       "SmoothingFilter", "PythonFilter", "UpsampleFilter", "VertexToCellFilter", "LangelaarFilter", "TotalVolumeConstraint", "detail"})
opt = open(os.path.join(REF, "Optimizer.cc")).read()
out["pyOptimizer.MMA"] = sorted(set(re.findall(r"\.def\(\s*\"(\w+)\"", opt)))
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyvoxelfem_api.json"), "w"), indent=1, sort_keys=True)
print({k: len(v) for k, v in out.items()})
