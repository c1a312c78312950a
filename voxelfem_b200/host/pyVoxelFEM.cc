// pyVoxelFEM.cc -- pybind11 module `pyVoxelFEM` over the host classes of VoxelFEM.hh (which sit on the C ABI of libvoxelfem_b200).
//
// Same module surface as the reference's binding (python_bindings/VoxelFEM.cc:76-432): factory TensorProductSimulator(degreesPerDimension,
// domainBBox, elementsPerDimension, numberType); detail.TensorProductSimulator1_1[_1], detail.MultigridSolver1_1[_1],
// detail.TopologyOptimizationProblem1_1[_1] (python-subclassable through a trampoline, :58-66), MultigridComplianceObjective,
// LayerByLayerEvaluator, OCOptimizer, the filters, FilterChain, TotalVolumeConstraint, InterpolationLaw, NumberType, getClassName.
// Array conventions as there: nodal fields are (numNodes, N) float64 numpy arrays copied in and out, densities flat over
// (ex, ey[, ez]) row-major.  Eigen is not available: the casters below convert numpy arrays to the header's VField / VXd.
// The export / post-processing methods next to the hot path are bound from voxelfem_b200/compat/tps_extras.py (shared with the ctypes flavour).
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "VoxelFEM.hh"

namespace py = pybind11;
using namespace voxelfem_b200;
using NpArr = py::array_t<double, py::array::c_style | py::array::forcecast>;

static VXd to_vxd(const NpArr &a) { const double *p = a.data(); return VXd(p, p + a.size()); }
static py::array_t<double> from_vxd(const VXd &v) { py::array_t<double> a((py::ssize_t)v.size()); std::copy(v.begin(), v.end(), a.mutable_data()); return a; }
static VField to_vfield(const NpArr &a, size_t N) {
    if (a.ndim() != 2 || (size_t)a.shape(1) != N) throw std::runtime_error("expected a (numNodes, " + std::to_string(N) + ") array");
    const size_t n = (size_t)a.shape(0); VField f(n, N); auto r = a.unchecked<2>();
    for (size_t i = 0; i < n; ++i) for (size_t c = 0; c < N; ++c) f(i, c) = r(i, c);
    return f;
}
static py::array_t<double> from_vfield(const VField &f) {
    py::array_t<double> a({(py::ssize_t)f.rows(), (py::ssize_t)f.cols()}); auto w = a.mutable_unchecked<2>();
    for (size_t i = 0; i < f.rows(); ++i) for (size_t c = 0; c < f.cols(); ++c) w(i, c) = f(i, c);
    return a;
}
template<size_t N> static std::array<double, N> to_vnd(const std::vector<double> &v) { if (v.size() != N) throw std::runtime_error("expected " + std::to_string(N) + " values"); std::array<double, N> r; std::copy(v.begin(), v.end(), r.begin()); return r; }
template<size_t N> static std::array<size_t, N> to_idx(const std::vector<size_t> &v) { if (v.size() != N) throw std::runtime_error("expected " + std::to_string(N) + " indices"); std::array<size_t, N> r; std::copy(v.begin(), v.end(), r.begin()); return r; }
template<class A> static py::array_t<double> np_of(const A &v) { py::array_t<double> a((py::ssize_t)v.size()); std::copy(v.begin(), v.end(), a.mutable_data()); return a; }
template<class A> static py::array_t<int64_t> npi_of(const A &v) { py::array_t<int64_t> a((py::ssize_t)v.size()); std::copy(v.begin(), v.end(), a.mutable_data()); return a; }

// MeshFEM's ElasticityTensor as far as the drivers use it (ElasticityTensor.hh:100-131): setIsotropic(E, nu)
struct PyETensor { bool isotropic = true; double E = 1, nu = 0; std::vector<double> D; void setIsotropic(double E_, double nu_) { isotropic = true; E = E_; nu = nu_; } };

template<size_t N> struct Names { static std::string mangle(const std::string &n) { std::string s = n; for (size_t d = 0; d < N; ++d) s += (d ? "_1" : "1"); return s; } };

// Trampoline: a python subclass may override setVars / evaluateObjective / evaluateObjectiveGradient (VoxelFEM.cc:58-66)
template<typename TPS>
class PyTopologyOptimizationProblem : public TopologyOptimizationProblem<TPS> {
public:
    using Base = TopologyOptimizationProblem<TPS>;
    using Base::Base;
    bool hasHostOverrides() const override {
        py::gil_scoped_acquire gil;
        return bool(py::get_override(static_cast<const Base *>(this), "setVars")) || bool(py::get_override(static_cast<const Base *>(this), "evaluateObjective")) ||
               bool(py::get_override(static_cast<const Base *>(this), "evaluateObjectiveGradient"));
    }
    bool setVars(const VXd &x, bool forceUpdate = false) override {
        py::gil_scoped_acquire gil;
        if (py::function f = py::get_override(static_cast<const Base *>(this), "setVars")) { py::object r = f(from_vxd(x)); return r.is_none() ? true : r.cast<bool>(); }
        return Base::setVars(x, forceUpdate);
    }
    double evaluateObjective() const override {
        py::gil_scoped_acquire gil;
        if (py::function f = py::get_override(static_cast<const Base *>(this), "evaluateObjective")) return f().cast<double>();
        return Base::evaluateObjective();
    }
    VXd evaluateObjectiveGradientAndReturn() const override {
        py::gil_scoped_acquire gil;
        if (py::function f = py::get_override(static_cast<const Base *>(this), "evaluateObjectiveGradient")) return to_vxd(NpArr::ensure(f()));
        return Base::evaluateObjectiveGradientAndReturn();
    }
};

// problem.filterChain: the device-resident chain of a problem (FilterChain interface, TopologyOptimizationFilter.hh:90-187)
template<typename TPS> struct ProblemChainView { TopologyOptimizationProblem<TPS> *p; };

// what MG.getSimulator(l > 0) exposes of a coarse-level simulator
template<typename MG> struct LevelSimView { std::shared_ptr<MG> mg; size_t l; };

template<size_t... Degrees>
void addTPSBindings(py::module &m, py::module &detail) {
    using TPS = TensorProductSimulator<double, Degrees...>;
    using MG = MultigridSolver<double, Degrees...>;
    constexpr size_t N = sizeof...(Degrees);
    using NM = Names<N>;
    auto mask_to_np = [](const std::vector<uint8_t> &mk) {
        py::array_t<bool> a({(py::ssize_t)mk.size(), (py::ssize_t)N}); auto w = a.template mutable_unchecked<2>();
        for (size_t i = 0; i < mk.size(); ++i) for (size_t c = 0; c < N; ++c) w(i, c) = (mk[i] >> c) & 1;
        return a;
    };
    py::class_<TPS, std::shared_ptr<TPS>> tps(detail, NM::mangle("TensorProductSimulator").c_str());
    tps.def("numNodes", &TPS::numNodes).def("numElements", &TPS::numElements)
       .def("getDensities", [](const TPS &s) { return from_vxd(s.getDensities()); })
       .def("setDensities", [](TPS &s, const NpArr &rho) { s.setDensities(to_vxd(rho)); }, py::arg("rho"))
       .def("readMaterial", &TPS::readMaterial, py::arg("materialPath"))
       .def("setDensity", &TPS::setDensity, py::arg("ei"), py::arg("value"))
       .def("setUniformDensities", &TPS::setUniformDensities, py::arg("density"))
       .def("setDensitiesFromCoarseGrid", [](TPS &s, size_t f, const NpArr &rho) { s.setDensitiesFromCoarseGrid(f, to_vxd(rho)); }, py::arg("upscalingFactor"), py::arg("rho"))
       .def("elementDensity", &TPS::elementDensity, py::arg("ei"))
       .def("elementYoungModulusScaleFactor", &TPS::elementYoungModulusScaleFactor, py::arg("ei"))
       .def("getYoungModulusScaleFactor", [](const TPS &s) { return from_vxd(s.getYoungModulusScaleFactor()); })
       .def("setFabricationMaskHeightByLayer", &TPS::setFabricationMaskHeightByLayer, py::arg("l"))
       .def("getFabricationMaskHeight", &TPS::getFabricationMaskHeight)
       .def("solve", [](TPS &s, const NpArr &f) { return from_vfield(s.solve(to_vfield(f, N))); }, py::arg("f"))
       .def("complianceGradient", [](const TPS &s, const NpArr &u) { return from_vxd(s.complianceGradientFlattened(to_vfield(u, N))); })
       .def("multigridSolver", [](std::shared_ptr<TPS> s, size_t levels) { return std::make_shared<MG>(s, levels); })
       .def("applyK", [](const TPS &s, const NpArr &u) { return from_vfield(s.applyK(to_vfield(u, N))); }, py::arg("u"))
       .def("buildLoadVector", [](const TPS &s) { return from_vfield(s.buildLoadVector()); })
       .def("nodePosition", [](const TPS &s, size_t ni) { return np_of(s.nodePosition(ni)); }, py::arg("ni"))
       .def("applyDisplacementsAndLoadsFromFile", &TPS::applyDisplacementsAndLoadsFromFile, py::arg("bcPath"))
       .def("elementIndexForGridCell", [](const TPS &s, const std::vector<size_t> &c) { return s.elementIndexForGridCell(to_idx<N>(c)); }, py::arg("cellIdxs"))
       .def("getDirichletMask", [mask_to_np](const TPS &s) { return mask_to_np(s.getDirichletMask()); })
       .def("elementStiffnessMatrix", [](const TPS &s, size_t ei) { const auto K = s.elementStiffnessMatrix(ei); const py::ssize_t k = N << N; py::array_t<double> a({k, k}); std::copy(K.begin(), K.end(), a.mutable_data()); return a; }, py::arg("ei"))
       .def("elementNodes", &TPS::elementNodes, py::arg("ei"))
       .def("elemNodeGlobalIndex", &TPS::elemNodeGlobalIndex, py::arg("ei"), py::arg("n"))
       .def("addDirichletCondition", [](TPS &s, const std::vector<double> &u, const std::vector<double> &lo, const std::vector<double> &hi, const std::string &cm) {
                s.addDirichletCondition(to_vnd<N>(u), to_vnd<N>(lo), to_vnd<N>(hi), cm); }, py::arg("u"), py::arg("minCorner"), py::arg("maxCorner"), py::arg("componentMask") = "xyz")
       .def("mesh", [](std::shared_ptr<TPS> s) { return s; }, "Hack to support MeshFEM's `simu_tils` helpers")
       .def_property_readonly("domain", [](const TPS &s) { return py::make_tuple(np_of(s.domain().minCorner), np_of(s.domain().maxCorner)); })
       .def_property_readonly("bbox", [](const TPS &s) { return py::make_tuple(np_of(s.domain().minCorner), np_of(s.domain().maxCorner)); })
       .def_property_readonly("gridShape", [](const TPS &s) { return npi_of(s.NbElementsPerDimension()); })
       .def_property_readonly("NbElementsPerDimension", [](const TPS &s) { return npi_of(s.NbElementsPerDimension()); })
       .def_property_readonly("NbNodesPerDimension", [](const TPS &s) { return npi_of(s.NbNodesPerDimension()); })
       .def_property("interpolationLaw", &TPS::interpolationLaw, &TPS::setInterpolationLaw)
       .def_property("E_0", &TPS::E_0, &TPS::setE_0).def_property("E_min", &TPS::E_min, &TPS::setE_min)
       .def_property("gamma", &TPS::SIMPExponent, &TPS::setSIMPExponent).def_property("q", &TPS::RAMPFactor, &TPS::setRAMPFactor)
       .def_property("gravity", [](const TPS &s) { return np_of(s.getGravity()); }, [](TPS &s, const std::vector<double> &g) { s.setGravity(to_vnd<N>(g)); })
       .def_property("ETensor", [](const TPS &s) { const auto &e = s.getETensor(); PyETensor r; r.isotropic = e.isotropic; r.E = e.E; r.nu = e.nu; r.D = e.D; return r; },
                     [](TPS &s, const PyETensor &e) { typename TPS::ETensorState st; st.isotropic = e.isotropic; st.E = e.E; st.nu = e.nu; st.D = e.D; s.setETensor(st); }, "Elasticty tensor")
       .def_property_readonly("dx", [](const TPS &s) { return np_of(s.getStretchings()); }, "Get the dimensions of a grid cell")
       .def_property_readonly("elementVolume", [](const TPS &s) { return s.elementVolume(0); })
       .def("fullDensityElementStiffnessMatrix", [](const TPS &s) { const auto K = s.fullDensityElementStiffnessMatrix(); const py::ssize_t k = N << N; py::array_t<double> a({k, k}); std::copy(K.begin(), K.end(), a.mutable_data()); return a; })
       .def("clearCachedElementStiffness", [](TPS &) {})
       .def("elementEnergyDensity", [](const TPS &s, const NpArr &u) { return from_vxd(s.elementEnergyDensity(to_vfield(u, N))); }, py::arg("u"))
       .def("getIntermediateFabricationShape", &TPS::getIntermediateFabricationShape, py::arg("yfrac"), py::arg("validateBoundaryConditions") = true, py::arg("law") = InterpolationLaw::SIMP)
       .def("transferDensitiesToIntermediateFabricationShape", &TPS::transferDensitiesToIntermediateFabricationShape, py::arg("intermediateTPS"))
       .def("applySymmetryConditions", [](TPS &s, const std::vector<bool> &axes, const std::vector<bool> &faces) {
                std::array<bool, N> a{}, f{}; for (size_t d = 0; d < N; ++d) { a[d] = d < axes.size() && axes[d]; f[d] = d < faces.size() && faces[d]; } s.applySymmetryConditions(a, f); },
            py::arg("symmetry_axes"), py::arg("minMaxFace") = std::vector<bool>(N, false))
       .def("downsample", &TPS::downsample, py::arg("downsamplingLevels"))
       .def("downsampleDensityFieldTo", [](const TPS &s, const NpArr &rho, TPS &coarse) { s.downsampleDensityFieldTo(to_vxd(rho), coarse); }, py::arg("densities"), py::arg("coarseTPS"))
       .def("upsampleDensityGradientFrom", [](const TPS &s, const TPS &coarse, const NpArr &g) { return from_vxd(s.upsampleDensityGradientFrom(coarse, to_vxd(g))); }, py::arg("coarseTPS"), py::arg("g_coarse"));
    // Export / post-processing methods next to the solve path: one host-side implementation written
    // against the public simulator API (voxelfem_b200/compat/tps_extras.py), shared with the ctypes flavour of this module.
    tps.def("_dirichletConditions", [](const TPS &s) {
            const size_t n = (size_t)vf_sim_num_dirichlet_nodes(s.handle());
            py::array_t<int64_t> nodes((py::ssize_t)n); py::array_t<uint8_t> masks((py::ssize_t)n); py::array_t<double> vals({(py::ssize_t)n, (py::ssize_t)N});
            std::vector<int64_t> hn(std::max<size_t>(n, 1)); std::vector<uint8_t> hm(std::max<size_t>(n, 1)); std::vector<double> hv(std::max<size_t>(n, 1) * N);
            if (vf_sim_get_dirichlet_conditions(s.handle(), hn.data(), hm.data(), hv.data()) != 0) throw std::runtime_error(vf_last_error());
            std::copy(hn.begin(), hn.begin() + n, nodes.mutable_data()); std::copy(hm.begin(), hm.begin() + n, masks.mutable_data()); std::copy(hv.begin(), hv.begin() + n * N, vals.mutable_data());
            return py::make_tuple(nodes, masks, vals); })
       .def("_forceNodes", [](const TPS &s) {
            const size_t n = (size_t)vf_sim_num_force_nodes(s.handle());
            py::array_t<int64_t> nodes((py::ssize_t)n); py::array_t<double> f({(py::ssize_t)n, (py::ssize_t)N});
            std::vector<int64_t> hn(std::max<size_t>(n, 1)); std::vector<double> hf(std::max<size_t>(n, 1) * N);
            if (vf_sim_get_force_nodes(s.handle(), hn.data(), hf.data()) != 0) throw std::runtime_error(vf_last_error());
            std::copy(hn.begin(), hn.begin() + n, nodes.mutable_data()); std::copy(hf.begin(), hf.begin() + n * N, f.mutable_data());
            return py::make_tuple(nodes, f); });
    for (const char *name : {"getK", "constantStrainLoad", "solveWithImposedLoads", "getDirichletVarsAndValues", "getForceMask", "getBCIndicatorField", "sampleNodalField", "getMesh",
                             "debugMulticolorElementVisit", "transferVFieldToIntermediateFabricationShape", "accumElementScalarFieldFromIntermediateFabricationShape"}) {
        const std::string n = name;
        tps.def(name, [n](py::object self, py::args a, py::kwargs kw) { return py::module_::import("voxelfem_b200.compat.tps_extras").attr(n.c_str())(self, *a, **kw); });
    }

    using LSV = LevelSimView<MG>;
    py::class_<LSV>(detail, NM::mangle("CoarseLevelSimulator").c_str())
        .def("numNodes", [](const LSV &v) { return v.mg->numNodes(v.l); })
        .def("getDirichletMask", [mask_to_np](const LSV &v) { return mask_to_np(v.mg->levelDirichletMask(v.l)); });

    py::class_<MG, std::shared_ptr<MG>>(detail, NM::mangle("MultigridSolver").c_str())
        .def("getSimulator", [](std::shared_ptr<MG> mg, size_t l) -> py::object { if (l == 0) return py::cast(mg->getSimulatorPtr()); return py::cast(LSV{mg, l}); }, py::arg("l"))
        .def("computeResidual", [](MG &mg, size_t l, const NpArr &u, const NpArr &b) { VField r; mg.computeResidual(l, to_vfield(u, N), to_vfield(b, N), r); return from_vfield(r); }, py::arg("l"), py::arg("u"), py::arg("b"))
        .def("applyK", [](MG &mg, size_t l, const NpArr &u) { return from_vfield(mg.applyK(l, to_vfield(u, N))); }, py::arg("l"), py::arg("u"))
        .def("zeroOutDirichletComponents", [](MG &mg, size_t l, const NpArr &u) { VField v = to_vfield(u, N); mg.zeroOutDirichletComponents(l, v); return from_vfield(v); }, py::arg("l"), py::arg("u"))
        .def("updateStiffnessMatrices", &MG::updateStiffnessMatrices)
        .def("setSymmetricGaussSeidel", &MG::setSymmetricGaussSeidel, py::arg("symmetric"))
        .def("solve", [](MG &mg, const NpArr &u, const NpArr &f, size_t numSteps, size_t numSmoothingSteps, bool stiffnessUpdated, bool zeroDirichlet, py::object it_callback, bool fmg) {
                typename MG::MGCallback cb = nullptr;
                if (!it_callback.is_none()) cb = [it_callback](size_t i, const VField &x) { it_callback(i, from_vfield(x)); };
                return from_vfield(mg.solve(to_vfield(u, N), to_vfield(f, N), numSteps, numSmoothingSteps, stiffnessUpdated, zeroDirichlet, cb, fmg)); },
             py::arg("u"), py::arg("f"), py::arg("numSteps"), py::arg("numSmoothingSteps"), py::arg("stiffnessUpdated") = false, py::arg("zeroDirichlet") = false,
             py::arg("it_callback") = py::none(), py::arg("fullMultigrid") = false)
        .def("setFabricationMaskHeightByLayer", &MG::setFabricationMaskHeightByLayer, py::arg("h"))
        .def("preconditionedConjugateGradient", [](MG &mg, const NpArr &u, const NpArr &b, size_t maxIter, double tol, py::object it_callback, size_t mgIterations, size_t mgSmoothingIterations, bool fmg) {
                typename MG::PCGCallback cb = nullptr; std::exception_ptr err;
                if (!it_callback.is_none()) cb = [it_callback, &err](size_t i, const VField &x, const VField &r) { if (err) return; try { it_callback(i, from_vfield(x), from_vfield(r)); } catch (...) { err = std::current_exception(); } };
                VField x = to_vfield(u, N);                  // the binding copies u and returns the new x (VoxelFEM.cc:174-186)
                mg.preconditionedConjugateGradient(x, to_vfield(b, N), maxIter, tol, cb, mgIterations, mgSmoothingIterations, fmg);
                if (err) std::rethrow_exception(err);
                return from_vfield(x); },
             py::arg("u"), py::arg("b"), py::arg("maxIter"), py::arg("tol"), py::arg("it_callback") = py::none(), py::arg("mgIterations") = 1, py::arg("mgSmoothingIterations") = 1, py::arg("fullMultigrid") = false)
        .def("debug_get_x", [](MG &mg, size_t l) { return from_vfield(mg.debug_get_x(l)); }, py::arg("l"))
        .def("debug_get_b", [](MG &mg, size_t l) { return from_vfield(mg.debug_get_b(l)); }, py::arg("l"))
        .def("debugMulticolorVisit", [](MG &mg) { const auto v = mg.debugMulticolorVisit(); py::array_t<int32_t> a((py::ssize_t)v.size()); std::copy(v.begin(), v.end(), a.mutable_data()); return a; });

    using TOProblem = TopologyOptimizationProblem<TPS>;
    using PyTOProblem = PyTopologyOptimizationProblem<TPS>;
    using MGCO = MultigridComplianceObjective<TPS>;
    using FiltersList = typename TOProblem::FiltersList;
    using ConstraintsList = typename TOProblem::ConstraintsList;
    using ObjectivePtr = typename TOProblem::ObjectivePtr;
    using PCV = ProblemChainView<TPS>;

    // "Constructor" function (VoxelFEM.cc:236-240): forwards to the class registered for this simulator type, so that the object is built
    // by the class's own py::init (one construction path for the factory, the mangled class and python subclasses of it)
    m.def("TopologyOptimizationProblem", [detail](std::shared_ptr<TPS> s, py::object o, py::object c, py::object f) {
              return detail.attr(NM::mangle("TopologyOptimizationProblem").c_str())(s, o, c, f); },
          py::arg("simulator"), py::arg("objective"), py::arg("constraints"), py::arg("filters"));
    m.def("MultigridComplianceObjective", [](std::shared_ptr<MG> mg) { return std::make_shared<MGCO>(mg); }, py::arg("mg_solver"));
    m.def("ComplianceObjective", [](TPS &) -> py::object { PyErr_SetString(PyExc_NotImplementedError, "ComplianceObjective (CHOLMOD direct solve of the fine system) is not on the B200 path; use MultigridComplianceObjective(tps.multigridSolver(levels))"); throw py::error_already_set(); }, py::arg("simulator"));

    py::class_<PCV>(detail, NM::mangle("ProblemFilterChain").c_str())
        .def("numVars", [](const PCV &v) { return v.p->numVars(); })
        .def("numPhysicalVars", [](const PCV &v) { return v.p->numVars(); })      // the reference binds numPhysicalVars to numVars (VoxelFEM.cc:331)
        .def("gridDims", [](const PCV &v) { return npi_of(v.p->gridDims(false)); })
        .def("physicalGridDims", [](const PCV &v) { return npi_of(v.p->gridDims(true)); })
        .def("designVars", [](const PCV &v) { return from_vxd(v.p->getVars()); })
        .def("physicalVars", [](const PCV &v) { return from_vxd(v.p->getDensities()); })
        .def("setDesignVars", [](PCV &v, const NpArr &x) { v.p->setVars(to_vxd(x)); }, py::arg("xDesign"))
        .def("backprop", [](const PCV &v, const NpArr &g) { return from_vxd(v.p->backpropThroughFilters(to_vxd(g))); }, py::arg("g"))
        .def_property_readonly("filters", [](const PCV &v) { return v.p->getFilters(); });

    py::class_<TOProblem, PyTOProblem>(detail, NM::mangle("TopologyOptimizationProblem").c_str())
        .def(py::init<TPS &, ObjectivePtr, ConstraintsList, FiltersList>(), py::keep_alive<1, 2>())
        .def("evaluateObjective", &TOProblem::evaluateObjective)
        .def("evaluateObjectiveGradient", [](const TOProblem &p) { return from_vxd(p.TOProblem::evaluateObjectiveGradientAndReturn()); })
        .def("evaluateConstraints", [](const TOProblem &p) { return from_vxd(p.evaluateConstraints()); })
        .def("evaluateConstraintsJacobian", [](const TOProblem &p) { const VXd g = p.evaluateConstraintsJacobianAndReturn(); py::array_t<double> a({(py::ssize_t)1, (py::ssize_t)g.size()}); std::copy(g.begin(), g.end(), a.mutable_data()); return a; })
        .def("numVars", &TOProblem::numVars)
        .def("getVars", [](const TOProblem &p) { return from_vxd(p.getVars()); })
        .def("setVars", [](TOProblem &p, const NpArr &x, bool force) { return p.TOProblem::setVars(to_vxd(x), force); }, py::arg("x"), py::arg("forceUpdate") = false)
        .def("getDensities", [](const TOProblem &p) { return from_vxd(p.getDensities()); })
        .def_property_readonly("objective", &TOProblem::getObjective)
        .def_property_readonly("filters", &TOProblem::getFilters)
        .def_property_readonly("filterChain", py::cpp_function([](TOProblem &p) { return PCV{&p}; }, py::keep_alive<0, 1>()))
        .def_property_readonly("constraints", &TOProblem::getConstraints);

    m.def("getClassName", [](TPS &, const std::string &name) { return std::string("pyVoxelFEM.detail.") + NM::mangle(name); }, py::arg("simulator"), py::arg("name"));

    py::class_<MGCO, std::shared_ptr<MGCO>>(detail, NM::mangle("MultigridComplianceObjective").c_str())
        .def_property_readonly("mg", [](const MGCO &o) { return std::static_pointer_cast<MG>(o.mgHolder()); })
        .def_readwrite("cgIter", &MGCO::cgIter).def_readwrite("tol", &MGCO::tol).def_readwrite("mgIterations", &MGCO::mgIterations)
        .def_readwrite("mgSmoothingIterations", &MGCO::mgSmoothingIterations).def_readwrite("fullMultigrid", &MGCO::fullMultigrid).def_readwrite("zeroInit", &MGCO::zeroInit)
        .def_property("residual_cb", [](const MGCO &) { return py::none(); }, [](MGCO &o, py::object cb) {
                if (cb.is_none()) o.residual_cb = nullptr; else o.residual_cb = [cb](size_t i, const VField &r) { cb(i, from_vfield(r)); }; })
        .def("compliance", &MGCO::compliance)
        .def("u", [](const MGCO &o) { return from_vfield(o.u()); })
        .def("f", [](const MGCO &o) { return from_vfield(o.f()); })
        .def("gradient", [](const MGCO &o) { return from_vxd(o.gradient()); })
        .def("updateCache", [](MGCO &, py::object) { PyErr_SetString(PyExc_NotImplementedError, "updateCache is driven by TopologyOptimizationProblem.setVars on the device (vf_top_set_vars)"); throw py::error_already_set(); }, py::arg("xPhys"));

    using LBL = LayerByLayerEvaluator<TPS>;
    py::class_<LBL>(detail, NM::mangle("LayerByLayerEvaluator").c_str())
        .def("selectInitMethod", &LBL::selectInitMethod, py::arg("method"), "Select method by name ['zero', 'fd', 'N=1', 'N=2', ...]")
        .def("run", [](LBL &e, MG &solver, bool zeroInit, size_t layerIncrement, size_t maxIter, double tol, py::object it_callback, size_t mgIterations, size_t mgSmoothingIterations,
                       bool fmg, bool verbose, py::object lblCallback) {
                typename MG::PCGCallback icb = nullptr; typename LBL::LBLCallback lcb = nullptr;
                if (!it_callback.is_none()) icb = [it_callback](size_t i, const VField &x, const VField &r) { it_callback(i, from_vfield(x), from_vfield(r)); };
                if (!lblCallback.is_none()) lcb = [lblCallback](size_t l, double c, const VXd &g, const VField &u) { lblCallback(l, c, from_vxd(g), from_vfield(u)); };
                e.run(solver, zeroInit, layerIncrement, maxIter, tol, icb, mgIterations, mgSmoothingIterations, fmg, verbose, lcb); },
             py::arg("solver"), py::arg("zeroInit"), py::arg("layerIncrement"), py::arg("maxIter"), py::arg("tol"), py::arg("it_callback") = py::none(), py::arg("mgIterations") = 1,
             py::arg("mgSmoothingIterations") = 1, py::arg("fullMultigrid") = false, py::arg("verbose") = false, py::arg("lblCallback") = py::none())
        .def("objective", &LBL::objective)
        .def("gradient", [](const LBL &e) { return from_vxd(e.gradient()); });
    m.def("LayerByLayerEvaluator", [](std::shared_ptr<TPS> s) { return std::make_unique<LBL>(s); }, py::arg("lblSim"));

    using OCO = OCOptimizer<TOProblem>;
    py::class_<OCO>(detail, NM::mangle("OCOptimizer").c_str())
        .def(py::init<TOProblem &>(), py::arg("problem"), py::keep_alive<1, 2>())
        .def("step", &OCO::step, py::arg("m") = 0.2, py::arg("p") = 0.5, py::arg("ctol") = 1e-6, py::arg("inplace") = true);
    m.def("OCOptimizer", [detail](py::object p) {
              if (!py::isinstance<TOProblem>(p)) throw py::reference_cast_error();   // not this simulator type: try the next overload
              return detail.attr(NM::mangle("OCOptimizer").c_str())(p); }, py::arg("problem"));
}

PYBIND11_MODULE(pyVoxelFEM, m) {
    m.doc() = "Voxel-based finite element codebase (B200-native hot path over libvoxelfem_b200)";
    py::module detail = m.def_submodule("detail");
    py::register_exception_translator([](std::exception_ptr p) {   // std::logic_error (the PCG's NaN guard) and runtime_error both surface as RuntimeError, as with pybind11's defaults
        try { if (p) std::rethrow_exception(p); } catch (const std::logic_error &e) { PyErr_SetString(PyExc_RuntimeError, e.what()); }
    });
    py::enum_<InterpolationLaw>(m, "InterpolationLaw").value("SIMP", InterpolationLaw::SIMP).value("RAMP", InterpolationLaw::RAMP).export_values();
    py::class_<PyETensor>(detail, "ElasticityTensor").def(py::init<>()).def("setIsotropic", &PyETensor::setIsotropic, py::arg("E"), py::arg("nu"))
        .def_readwrite("D", &PyETensor::D).def_readonly("isotropic", &PyETensor::isotropic);

    addTPSBindings<1, 1>(m, detail);
    addTPSBindings<1, 1, 1>(m, detail);

    using Filter_ = Filter<double>;
    py::class_<Filter_, std::shared_ptr<Filter_>>(detail, "Filter")
        .def("setInputDimensions", &Filter_::setInputDimensions, py::arg("gridDims"))
        .def("setOutputDimensions", &Filter_::setOutputDimensions, py::arg("gridDims"))
        .def_property_readonly("inputDimensions", [](const Filter_ &f) { return npi_of(f.inputDimensions()); })
        .def_property_readonly("outputDimensions", [](const Filter_ &f) { return npi_of(f.outputDimensions()); })
        .def("apply", [](Filter_ &f, const NpArr &x) { f.checkGridDimensionsAreSet(); return from_vxd(f.apply(to_vxd(x))); }, py::arg("x"));

    using FC = FilterChain<double>;
    py::class_<FC, std::shared_ptr<FC>>(m, "FilterChain")
        .def(py::init<typename FC::Filters, const GridDims &>(), py::arg("filters"), py::arg("outGridDimensions"))
        .def("numVars", &FC::numVars).def("numPhysicalVars", &FC::numVars)
        .def("gridDims", [](const FC &c) { return npi_of(c.gridDims()); }, "Input grid dimensions")
        .def("physicalGridDims", [](const FC &c) { return npi_of(c.physicalGridDims()); }, "Output grid dimensions")
        .def("setDesignVars", [](FC &c, const NpArr &x) { c.setDesignVars(to_vxd(x)); }, py::arg("xDesign"))
        .def("backprop", [](const FC &c, const NpArr &g) { return from_vxd(c.backprop(to_vxd(g))); }, py::arg("g"))
        .def("designVars", [](const FC &c) { return from_vxd(c.designVars()); })
        .def("physicalVars", [](const FC &c) { return from_vxd(c.physicalVars()); })
        .def_property_readonly("filters", &FC::filters);

    using PyF = PythonFilter<double>;
    // callbacks: apply_cb(in, out) / backprop_cb(in, vars, out) write into the numpy array `out` (Eigen::Ref in the reference, :253-254)
    py::class_<PyF, Filter_, std::shared_ptr<PyF>>(m, "PythonFilter")
        .def(py::init<>())
        .def_property("apply_cb", [](const PyF &) { return py::none(); }, [](PyF &f, py::object cb) {
                if (cb.is_none()) { f.apply_cb = nullptr; return; }
                f.apply_cb = [cb](const VXd &in, VXd &out) { py::gil_scoped_acquire gil; auto o = from_vxd(out); cb(from_vxd(in), o); std::copy(o.data(), o.data() + o.size(), out.begin()); }; })
        .def_property("backprop_cb", [](const PyF &) { return py::none(); }, [](PyF &f, py::object cb) {
                if (cb.is_none()) { f.backprop_cb = nullptr; return; }
                f.backprop_cb = [cb](const VXd &in, const VXd &vars, VXd &out) { py::gil_scoped_acquire gil; auto o = from_vxd(out); cb(from_vxd(in), from_vxd(vars), o); std::copy(o.data(), o.data() + o.size(), out.begin()); }; });

    using PF = ProjectionFilter<double>;
    py::class_<PF, Filter_, std::shared_ptr<PF>>(m, "ProjectionFilter")
        .def(py::init<double>(), py::arg("beta")).def(py::init<>())
        .def("invert", &PF::invert, py::arg("filteredValue"))
        .def_property("beta", &PF::getBeta, &PF::setBeta);

    using SF = SmoothingFilter<double>;
    py::class_<SF, Filter_, std::shared_ptr<SF>> pySF(m, "SmoothingFilter");
    py::enum_<SF::Type>(pySF, "Type").value("Const", SF::Type::Const).value("Linear", SF::Type::Linear);
    pySF.def(py::init<size_t, SF::Type>(), py::arg("radius") = 1, py::arg("type") = SF::Type::Const)
        .def_readwrite("radius", &SF::radius).def_readwrite("type", &SF::type);

    py::class_<UpsampleFilter<double>, Filter_, std::shared_ptr<UpsampleFilter<double>>>(m, "UpsampleFilter").def(py::init<size_t>(), py::arg("factor"));
    py::class_<VertexToCellFilter<double>, Filter_, std::shared_ptr<VertexToCellFilter<double>>>(m, "VertexToCellFilter").def(py::init<>());
    py::class_<LangelaarFilter<double>, Filter_, std::shared_ptr<LangelaarFilter<double>>>(m, "LangelaarFilter").def(py::init<>());

    using C_ = Constraint<double>;
    py::class_<C_, std::shared_ptr<C_>>(detail, "Constraint");
    using TVC = TotalVolumeConstraint<double>;
    py::class_<TVC, C_, std::shared_ptr<TVC>>(m, "TotalVolumeConstraint")
        .def(py::init<double>(), py::arg("volumeFraction"))
        .def_readwrite("volumeFraction", &TVC::m_volumeFraction);

    enum class NumberType { DOUBLE, FLOAT };
    py::enum_<NumberType>(m, "NumberType").value("DOUBLE", NumberType::DOUBLE).value("FLOAT", NumberType::FLOAT);
    // Factory masquerading as a class (VoxelFEM.cc:422-430): only the instantiations the reference registers, <double, 1, 1[, 1]> (:301-308)
    m.def("TensorProductSimulator", [](const std::vector<size_t> &degrees, const std::vector<std::vector<double>> &bbox, const std::vector<size_t> &ne, NumberType nt) -> py::object {
            const bool q1 = std::all_of(degrees.begin(), degrees.end(), [](size_t d) { return d == 1; });
            if (nt != NumberType::DOUBLE || !q1 || (degrees.size() != 2 && degrees.size() != 3) || bbox.size() != 2 || ne.size() != degrees.size())
                throw std::runtime_error("No template instantiation matching degreesPerDimension/number type!");
            if (degrees.size() == 2) { using T = TensorProductSimulator<double, 1, 1>; typename T::BBoxN b{to_vnd<2>(bbox[0]), to_vnd<2>(bbox[1])}; return py::cast(std::make_shared<T>(b, to_idx<2>(ne))); }
            using T = TensorProductSimulator<double, 1, 1, 1>; typename T::BBoxN b{to_vnd<3>(bbox[0]), to_vnd<3>(bbox[1])}; return py::cast(std::make_shared<T>(b, to_idx<3>(ne)));
        }, py::arg("degreesPerDimension"), py::arg("domainBBox"), py::arg("elementsPerDimension"), py::arg("numberType") = NumberType::DOUBLE);
}
