"""The C++ host classes (voxelfem_b200/host/VoxelFEM.hh: the reference's class and method names over the C ABI) driven by
host_smoke.cpp on the GPU; every printed number is compared with the CPU oracle on the same problem."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import OracleMG, OracleProblem, OracleSim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "voxelfem_b200", "host")


@pytest.fixture(scope="module")
def smoke_output(data_dir):
    exe = os.path.join(HOST, "host_smoke")
    if not os.path.exists(exe):
        subprocess.run(["make"], cwd=HOST, check=True)
    r = subprocess.run([exe, data_dir], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = {}
    for line in r.stdout.splitlines():
        k, _, v = line.partition(" ")
        out[k] = v
    return out


def _oracle_solve(ne, dom, bc, levels, data_dir):
    o = OracleSim(np.array(ne), np.zeros(len(ne)), np.array(dom, dtype=float))
    o.set_isotropic(1.0, 0.3); o.set_interp(0, 1.0, 1e-5, 3.0, 3.0); o.apply_bc_file(os.path.join(data_dir, "bcs", bc)); o.set_uniform_density(0.5)
    f = o.build_load()
    u, it, res = OracleMG(o, levels).pcg(np.zeros_like(f), f, 100, 1e-10, 1, 1, True)
    return it, 0.5 * float((f * u).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("tag,ne,dom,bc", [("solve2d", (64, 32), (2, 1), "cantilever_flexion_E.bc"), ("solve3d", (32, 16, 16), (2, 1, 1), "3D/cantilever_flexion_E.bc")])
def test_cpp_single_solve(smoke_output, data_dir, tag, ne, dom, bc):
    it, comp = _oracle_solve(ne, dom, bc, 2, data_dir)
    o = smoke_output
    assert abs(int(o[tag + "_iterations"]) - it) <= 1
    assert int(o[tag + "_callbacks"]) == int(o[tag + "_iterations"])
    assert abs(float(o[tag + "_compliance"]) - comp) < 1e-8 * abs(comp)
    assert float(o[tag + "_relres"]) < 1.01e-10 and abs(float(o[tag + "_cb_relres"]) - float(o[tag + "_relres"])) < 1e-12
    assert abs(float(o[tag + "_cb_xnorm_minus_xnorm"])) < 1e-12       # the last callback saw the final iterate
    assert float(o[tag + "_applyK_consistency"]) < 1e-12


@pytest.mark.gpu
def test_cpp_topopt_and_errors(smoke_output, data_dir):
    o = smoke_output
    V = 0.3
    s = OracleSim(np.array([16, 8, 8]), np.zeros(3), np.array([2.0, 1.0, 1.0]))
    s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0); s.apply_bc_file(os.path.join(data_dir, "bcs", "3D", "cantilever_flexion_E.bc")); s.set_uniform_density(1.0)
    op = OracleProblem(OracleMG(s, 2), [("smooth", 2, 1), ("project", 1.0)], V)
    op.set_solver(100, 1e-9, 1, 2, True, False)
    op.set_vars(np.full(16 * 8 * 8, 0.5 + np.arctanh((2 * V - 1) * np.tanh(0.5))))
    for it in range(3):
        assert abs(float(o["topopt_compliance_%d" % it]) - op.compliance()) < 1e-8 * abs(op.compliance())
        assert abs(float(o["topopt_constraint_%d" % it]) - op.constraint()) < 1e-9
        op.oc_step()
    assert abs(float(o["topopt_sum_vars"]) - op.design_vars().sum()) < 1e-6 * op.design_vars().sum()
    assert abs(float(o["topopt_sum_gradient"]) - op.objective_gradient().sum()) < 1e-6 * abs(op.objective_gradient().sum())
    assert abs(float(o["topopt_jacobian_entry"]) - op.constraint_jacobian()[0]) < 1e-12
    assert abs(float(o["mma_mean"]) - 0.2) < 1e-5                        # the volume constraint is active at the optimum
    assert o["odd_grid_error"].startswith("runtime_error") and "divisible" in o["odd_grid_error"]


@pytest.mark.gpu
def test_cpp_slab_problem_and_q2(smoke_output):
    """The slab-partitioned problem (SlabTopologyOptimizationProblem over vf_group_top_*, two local slabs) reproduces the undivided
    C++ problem iteration by iteration; the Q2 simulator reproduces the numpy restatement of the generic element path."""
    o = smoke_output
    for it in range(3):
        c, cu = float(o["slab_compliance_%d" % it]), float(o["topopt_compliance_%d" % it])
        assert abs(c - cu) < 1e-8 * abs(cu), it
        assert abs(float(o["slab_constraint_%d" % it]) - float(o["topopt_constraint_%d" % it])) < 1e-9
    assert abs(float(o["slab_sum_vars"]) - float(o["topopt_sum_vars"])) < 1e-7 * float(o["topopt_sum_vars"])
    assert abs(float(o["slab_sum_gradient"]) - float(o["topopt_sum_gradient"])) < 1e-6 * abs(float(o["topopt_sum_gradient"]))
    assert int(o["slab_halo_layers"]) == 4                                  # max(2 * 2, 2 + 1) for SmoothingFilter(2) + ProjectionFilter
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import q2ref
    s = q2ref.Q2Sim([3, 2, 2], np.zeros(3), np.array([1.5, 1.0, 1.0])); s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-3, 3.0, 3.0)
    X = s.node_positions()
    u = X * np.array([0.01, -0.003, -0.003])
    assert int(o["q2_nodes"]) == s.num_nodes
    assert float(o["q2_K0_asymmetry"]) == 0 and float(o["q2_K0_translation"]) < 1e-13
    assert abs(float(o["q2_K0_trace"]) - np.trace(s.K0)) < 1e-12 * np.trace(s.K0)
    e = s.element_energies(u).sum()
    assert abs(float(o["q2_linear_field_energy"]) - e) < 1e-12 * e
    assert abs(float(o["q2_uKu"]) - (u * s.apply_K(u)).sum()) < 1e-12 * e
    # closed form: full density (E = 1), uniaxial stress state eps = (0.01, -0.003, -0.003) -> sigma_xx = 0.01, energy = vol * sigma : eps
    assert abs(e - 1.5 * 0.01 * 0.01) < 1e-12


def test_cpp_host_header_compiles_and_fails_loudly_without_gpu():
    """CPU check: the header-only host API and its smoke program build against the C ABI; without a GPU the program must
    report the missing device instead of computing anything."""
    import torch
    subprocess.run(["make"], cwd=HOST, check=True, capture_output=True)
    assert os.path.exists(os.path.join(HOST, "host_smoke"))
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(HOST, "host_smoke")], capture_output=True, text=True)
    assert r.returncode != 0 and "no usable CUDA device" in r.stdout
