import os, sys, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "voxelfem_b200", "pybind"))
import numpy as np
import pyVoxelFEM as vf
def P(*a): print(*a, flush=True)
DATA = os.path.join(ROOT, "voxelfem_b200", "data")
for dim, grid, corners, bc in ((3, [16, 8, 8], [[0, 0, 0], [2, 1, 1]], "3D/cantilever_flexion_E.bc"), (2, [16, 8], [[0, 0], [2, 1]], "mbb_N.bc")):
    tps = vf.TensorProductSimulator([1] * dim, corners, grid)
    tps.readMaterial(os.path.join(DATA, "materials", "B9Creator.material"))
    tps.applyDisplacementsAndLoadsFromFile(os.path.join(DATA, "bcs", bc))
    pf = vf.ProjectionFilter(); pf.beta = 1
    filters = [vf.SmoothingFilter(2, vf.SmoothingFilter.Type.Linear), pf]
    objective = vf.MultigridComplianceObjective(tps.multigridSolver(2))
    cons = [vf.TotalVolumeConstraint(0.3)]
    top = vf.TopologyOptimizationProblem(tps, objective, cons, filters); P("factory ok", type(top))
    top.setVars(pf.invert(0.3) * np.ones(tps.numElements())); P("setVars", top.evaluateObjective())
    oc = vf.OCOptimizer(top); oc.step(); P("oc", top.evaluateObjective(), top.evaluateConstraints())
    P(top.filterChain.backprop(objective.gradient())[:3], objective.compliance())
