// host_smoke.cpp -- the reference's C++ workflow for the hot path, written against the reference's class names
// (VoxelFEM.hh), compiled by __graft_entry__.build() and run on the GPU by tests/test_gpu_host_api.py, which checks
// the printed numbers against the CPU restatement kept under tests/:
//   1. python/CoarseningLevelBenchmark.py:76-100 in C++: one MG-PCG solve with an iteration callback (2D and 3D),
//   2. python/3DTopoptDemo.ipynb cells 1, 5: filters + MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer,
//   3. error behaviour: the exception the reference throws for a grid that cannot be coarsened (MultigridSolver.hh:52).
#include <cstdio>
#include <string>

#include "VoxelFEM.hh"

using namespace voxelfem_b200;

template<typename TPS>
static void single_solve(const char *tag, const typename TPS::BBoxN &dom, const typename TPS::EigenNDIndex &ne, const std::string &data, const std::string &bc, size_t levels) {
    auto tps = std::make_shared<TPS>(dom, ne);
    tps->readMaterial(data + "/materials/B9Creator.material");
    tps->applyDisplacementsAndLoadsFromFile(data + "/bcs/" + bc);
    tps->setE_min(1e-5);
    tps->setUniformDensities(0.5);
    const VField f = tps->buildLoadVector();
    using MG = MultigridSolver<double, 1, 1>;
    (void)sizeof(MG);
    auto run = [&](auto &mg) {
        VField x(tps->numNodes(), TPS::N);
        size_t calls = 0; double lastRes = 0, lastDiff = -1;
        mg.preconditionedConjugateGradient(x, f, 100, 1e-10, [&](size_t it, const VField &xi, const VField &r) { ++calls; lastRes = r.norm(); (void)it; lastDiff = xi.norm(); }, 1, 1, true);
        VField r; mg.computeResidual(0, x, f, r);
        std::printf("%s_iterations %zu\n%s_callbacks %zu\n%s_compliance %.15e\n%s_relres %.6e\n%s_cb_relres %.6e\n%s_cb_xnorm_minus_xnorm %.3e\n", tag, mg.lastPCGIterations(), tag, calls, tag,
                    0.5 * f.dot(x), tag, r.norm() / f.norm(), tag, lastRes / f.norm(), tag, lastDiff - x.norm());
        const VField Ku = tps->applyK(x), Ku2 = mg.applyK(0, x);
        double d = 0; for (size_t i = 0; i < Ku.size(); ++i) d = std::max(d, std::abs(Ku.data()[i] - Ku2.data()[i]));
        std::printf("%s_applyK_consistency %.3e\n", tag, d);
    };
    if constexpr (TPS::N == 2) { MultigridSolver<double, 1, 1> mg(tps, levels); run(mg); }
    else { MultigridSolver<double, 1, 1, 1> mg(tps, levels); run(mg); }
}

int main(int argc, char **argv) {
    std::setvbuf(stdout, nullptr, _IOLBF, 0);      // line-buffered also into a pipe: a crash must not swallow what was printed
    const std::string data = argc > 1 ? argv[1] : "voxelfem_b200/data";
    try {
        using TPS2 = TensorProductSimulator<double, 1, 1>;
        using TPS3 = TensorProductSimulator<double, 1, 1, 1>;
        single_solve<TPS2>("solve2d", {{0, 0}, {2, 1}}, {64, 32}, data, "cantilever_flexion_E.bc", 2);
        single_solve<TPS3>("solve3d", {{0, 0, 0}, {2, 1, 1}}, {32, 16, 16}, data, "3D/cantilever_flexion_E.bc", 2);

        {   // topology optimization, 3 OC iterations
            auto tps = std::make_shared<TPS3>(TPS3::BBoxN{{0, 0, 0}, {2, 1, 1}}, TPS3::EigenNDIndex{16, 8, 8});
            tps->readMaterial(data + "/materials/B9Creator.material");
            tps->applyDisplacementsAndLoadsFromFile(data + "/bcs/3D/cantilever_flexion_E.bc");
            auto mg = std::make_shared<MultigridSolver<double, 1, 1, 1>>(tps, 2);
            auto objective = std::make_shared<MultigridComplianceObjective<TPS3>>(mg);
            objective->tol = 1e-9;
            auto pf = std::make_shared<ProjectionFilter<double>>(1.0);
            TopologyOptimizationProblem<TPS3> top(*tps, objective, {std::make_shared<TotalVolumeConstraint<double>>(0.3)},
                                                  {std::make_shared<SmoothingFilter<double>>(2, SmoothingFilter<double>::Type::Linear), pf});
            top.setVars(VXd(top.numVars(), pf->invert(0.3)));
            OCOptimizer<TopologyOptimizationProblem<TPS3>> oc(top);
            for (int it = 0; it < 3; ++it) {
                std::printf("topopt_compliance_%d %.15e\ntopopt_constraint_%d %.6e\n", it, top.evaluateObjective(), it, top.evaluateConstraints()[0]);
                oc.step();
            }
            double s = 0, g = 0; for (double v : top.getVars()) s += v; for (double v : top.evaluateObjectiveGradientAndReturn()) g += v;
            std::printf("topopt_sum_vars %.12e\ntopopt_sum_gradient %.12e\ntopopt_jacobian_entry %.12e\n", s, g, top.evaluateConstraintsJacobianAndReturn()[0]);
        }
        {   // MMA on a separable problem: min sum (x - 0.3)^2  s.t. mean(x) <= 0.2
            const int n = 50;
            MMA opt(n, 1, VXd(n, 0.0), VXd(n, 1.0),
                    [&](const VXd &x) { double f = 0, m = 0; for (double v : x) { f += (v - 0.3) * (v - 0.3); m += v; } return VXd{f, m / n - 0.2}; },
                    [&](const VXd &x) { VXd d(2 * n); for (int i = 0; i < n; ++i) { d[i] = 2 * (x[i] - 0.3); d[n + i] = 1.0 / n; } return d; });
            opt.setInitialVar(VXd(n, 0.5));
            for (int i = 0; i < 30; ++i) opt.step();
            double m = 0; for (double v : opt.getOptimalVar()) m += v;
            std::printf("mma_mean %.9e\n", m / n);
        }
        {   // the same problem cut into two slabs of a local group (vf_group_top_*): 2 OC iterations side by side with the undivided values above
            SlabTopologyOptimizationProblem::Setup st;
            st.domainMax = {{2, 1, 1}}; st.elements = {{16, 8, 8}}; st.bcPath = data + "/bcs/3D/cantilever_flexion_E.bc";
            st.young = 1.0; st.poisson = 0.3; st.numCoarseningLevels = 2; st.firstReplicatedLevel = 1;
            st.filters = {{0, 2, 1, 0.0}, {1, 0, 0, 1.0}}; st.volumeFraction = 0.3;
            SlabTopologyOptimizationProblem top(st, {{0, 8}, {8, 16}});
            top.setSolver(100, 1e-9, 1, 2, true, false);
            top.setVars(VXd(top.numVars(), ProjectionFilter<double>(1.0).invert(0.3)));
            for (int it = 0; it < 3; ++it) {
                std::printf("slab_compliance_%d %.15e\nslab_constraint_%d %.6e\n", it, top.evaluateObjective(), it, top.evaluateConstraint());
                top.ocStep();
            }
            double s = 0, g = 0; for (double v : top.getVars()) s += v; for (double v : top.evaluateObjectiveGradient()) g += v;
            std::printf("slab_sum_vars %.12e\nslab_sum_gradient %.12e\nslab_halo_layers %lld\n", s, g, (long long)top.filterHaloLayers());
        }
        {   // degree-2 elements (vf_q2_*): element matrix invariants, the operator on a linear field, a clamped solve
            using Q2 = TensorProductSimulatorQ2<3>;
            Q2 q2(Q2::BBoxN{{0, 0, 0}, {1.5, 1.0, 1.0}}, Q2::EigenNDIndex{3, 2, 2});
            q2.setIsotropicETensor(1.0, 0.3);
            q2.setInterpolation(InterpolationLaw::SIMP, 1.0, 1e-3, 3.0, 3.0);
            const auto K = q2.fullDensityElementStiffnessMatrix();
            double asym = 0, rowsum = 0, tr = 0;
            for (size_t i = 0; i < 81; ++i) { tr += K[i * 81 + i]; for (size_t j = 0; j < 81; ++j) asym = std::max(asym, std::abs(K[i * 81 + j] - K[j * 81 + i])); }
            for (size_t i = 0; i < 81; ++i) { double r = 0; for (size_t j = 0; j < 81; j += 3) r += K[i * 81 + j]; rowsum = std::max(rowsum, std::abs(r)); }   // K0 * (translation along x) = 0
            const auto nn = q2.NbNodesPerDimension();
            VField u(q2.numNodes(), 3), f;
            for (size_t a = 0, n = 0; a < nn[0]; ++a) for (size_t b = 0; b < nn[1]; ++b) for (size_t c = 0; c < nn[2]; ++c, ++n) {
                const double x = 0.25 * a, y = 0.25 * b, z = 0.25 * c;              // node spacing h / 2 = 0.25
                u(n, 0) = 0.01 * x; u(n, 1) = -0.003 * y; u(n, 2) = -0.003 * z;     // uniaxial strain state of a nu = 0.3 material
            }
            q2.applyK<true, false>(u, f);
            double energy = 0; for (double e : q2.elementEnergies(u)) energy += e;
            std::printf("q2_nodes %zu\nq2_K0_asymmetry %.3e\nq2_K0_translation %.3e\nq2_K0_trace %.12e\nq2_linear_field_energy %.12e\nq2_uKu %.12e\n", q2.numNodes(), asym, rowsum, tr, energy, u.dot(f));
        }
        try {
            auto odd = std::make_shared<TPS2>(TPS2::BBoxN{{0, 0}, {1, 1}}, TPS2::EigenNDIndex{6, 5});
            MultigridSolver<double, 1, 1> mg(odd, 1);
            std::printf("odd_grid_error none\n");
        } catch (const std::runtime_error &e) { std::printf("odd_grid_error runtime_error: %s\n", e.what()); }
    } catch (const std::exception &e) {
        std::printf("FAILED %s\n", e.what());
        return 1;
    }
    return 0;
}
