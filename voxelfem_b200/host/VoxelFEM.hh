// VoxelFEM.hh -- C++ host classes of the B200-native VoxelFEM hot path.
//
// Header-only veneer over the C ABI (include/voxelfem_b200.h) with the reference's class and method names, so that C++
// code written against the reference's headers for the MG-PCG / topopt path recompiles against this one:
//   TensorProductSimulator<double, 1, 1[, 1]>     TensorProductSimulator.hh:170-2182
//   MultigridSolver<double, 1, 1[, 1]>            MultigridSolver.hh:23-1176
//   Filter / SmoothingFilter / ProjectionFilter   TopologyOptimizationFilter.hh:20-400
//   TotalVolumeConstraint                         TopologyOptimizationConstraint.hh:24-40
//   MultigridComplianceObjective                  TopologyOptimizationObjective.hh:60-105
//   TopologyOptimizationProblem                   TopologyOptimizationProblem.hh:17-155
//   OCOptimizer                                   OptimalityCriterion.hh:38-149
//   LayerByLayerEvaluator                         LayerByLayer.hh:25-309
//   MMA                                           MethodOfMovingAsymptotes.hh:28-470
// Eigen is not available here, so the Eigen types of the reference's signatures are replaced by two minimal owning
// arrays with the same storage order: VField (numNodes x N, column-major == the C ABI's component-major VField,
// TensorProductSimulator.hh:180) and VXd (std::vector<double>).
// Every method is one C-ABI call; all arithmetic runs in CUDA kernels of libvoxelfem_b200.so.  Errors arrive as a
// status + vf_last_error() and are rethrown as the exception types the reference throws (std::runtime_error, and
// std::logic_error for the PCG's NaN guard, MultigridSolver.hh:1083).
#pragma once
#include <array>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <exception>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/voxelfem_b200.h"

namespace voxelfem_b200 {

inline void check(int rc) {
    if (rc == 0) return;
    const std::string msg = vf_last_error();
    if (msg.rfind("NaN encountered", 0) == 0) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

using VXd = std::vector<double>;

// (rows x N) column-major nodal vector field
class VField {
public:
    VField() = default;
    VField(size_t rows, size_t cols) : m_rows(rows), m_cols(cols), m_d(rows * cols, 0.0) {}
    void setZero(size_t rows, size_t cols) { m_rows = rows; m_cols = cols; m_d.assign(rows * cols, 0.0); }
    void setZero() { std::fill(m_d.begin(), m_d.end(), 0.0); }
    size_t rows() const { return m_rows; }
    size_t cols() const { return m_cols; }
    size_t size() const { return m_d.size(); }
    double &operator()(size_t i, size_t c) { return m_d[c * m_rows + i]; }
    double operator()(size_t i, size_t c) const { return m_d[c * m_rows + i]; }
    double *data() { return m_d.data(); }
    const double *data() const { return m_d.data(); }
    double squaredNorm() const { double s = 0; for (double v : m_d) s += v * v; return s; }
    double norm() const { return std::sqrt(squaredNorm()); }
    double dot(const VField &o) const { double s = 0; for (size_t i = 0; i < m_d.size(); ++i) s += m_d[i] * o.m_d[i]; return s; }
private:
    size_t m_rows = 0, m_cols = 0;
    std::vector<double> m_d;
};

enum class InterpolationLaw { SIMP = VF_LAW_SIMP, RAMP = VF_LAW_RAMP };

namespace detail {
// Minimal reader for the .bc files the voxel simulator accepts (MeshFEM BoundaryConditions.cc:219-380 restricted to
// "dirichlet[xyz]*" / "force" box regions, TensorProductSimulator.hh:600-652): {"regions": [{"type", "value", "box%"|"box"}]}
struct BCRegion { int kind = 0, cmask = 7; double value[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; bool relative = false; };
class MiniJSON {
public:
    explicit MiniJSON(const std::string &s) : m_s(s) {}
    std::vector<BCRegion> regions() {
        std::vector<BCRegion> out;
        expect('{');
        while (true) {
            const std::string key = str(); expect(':');
            if (key == "regions") {
                expect('[');
                if (peek() == ']') { ++m_p; }
                else while (true) { out.push_back(region()); if (peek() == ',') { ++m_p; continue; } expect(']'); break; }
            } else skipValue();
            if (peek() == ',') { ++m_p; continue; }
            expect('}'); break;
        }
        return out;
    }
private:
    const std::string &m_s; size_t m_p = 0;
    char peek() { while (m_p < m_s.size() && std::isspace((unsigned char)m_s[m_p])) ++m_p; if (m_p >= m_s.size()) throw std::runtime_error("unexpected end of .bc file"); return m_s[m_p]; }
    void expect(char c) { if (peek() != c) throw std::runtime_error(std::string("malformed .bc file: expected '") + c + "'"); ++m_p; }
    std::string str() { expect('"'); std::string r; while (m_p < m_s.size() && m_s[m_p] != '"') r += m_s[m_p++]; ++m_p; return r; }
    double num() { peek(); size_t used = 0; const double v = std::stod(m_s.substr(m_p), &used); m_p += used; return v; }
    void vec(double (&v)[3]) { expect('['); int i = 0; if (peek() == ']') { ++m_p; return; } while (true) { const double x = num(); if (i < 3) v[i++] = x; if (peek() == ',') { ++m_p; continue; } expect(']'); break; } }
    void skipValue() {
        const char c = peek();
        if (c == '"') { str(); return; }
        if (c == '{' || c == '[') { int depth = 0; do { const char d = m_s[m_p++]; if (d == '{' || d == '[') ++depth; else if (d == '}' || d == ']') --depth; else if (d == '"') { while (m_s[m_p] != '"') ++m_p; ++m_p; } } while (depth > 0); return; }
        while (m_p < m_s.size() && m_s[m_p] != ',' && m_s[m_p] != '}' && m_s[m_p] != ']') ++m_p;
    }
    void box(BCRegion &r) {
        expect('{');
        while (true) {
            const std::string k = str(); expect(':');
            if (k == "minCorner") vec(r.lo); else if (k == "maxCorner") vec(r.hi); else skipValue();
            if (peek() == ',') { ++m_p; continue; } expect('}'); break;
        }
    }
    BCRegion region() {
        BCRegion r; bool haveBox = false;
        expect('{');
        while (true) {
            const std::string k = str(); expect(':');
            if (k == "type") {
                const std::string t = str();
                if (t.rfind("dirichlet", 0) == 0) {
                    r.kind = 0; const std::string comp = t.substr(9);
                    if (!comp.empty()) { r.cmask = 0; for (char ch : comp) { if (ch < 'x' || ch > 'z') throw std::runtime_error("Invalid type '" + t + "'"); r.cmask |= 1 << (ch - 'x'); } }
                } else if (t == "force") r.kind = 1;
                else throw std::runtime_error("Illegal constraint type, only \"dirichlet\" and \"force\" accepted");
            } else if (k == "value") vec(r.value);
            else if (k == "box%") { r.relative = true; box(r); haveBox = true; }
            else if (k == "box") { r.relative = false; box(r); haveBox = true; }
            else skipValue();
            if (peek() == ',') { ++m_p; continue; } expect('}'); break;
        }
        if (!haveBox) throw std::runtime_error("only box / box% regions are supported");
        return r;
    }
};
} // namespace detail

namespace detail {
// applyDisplacementsAndLoadsFromFile (TensorProductSimulator.hh:654-658; JSON regions of MeshFEM BoundaryConditions.cc:219-380) on a
// simulator handle: `box%` regions are fractions of the GLOBAL domain (also for a slab window)
inline void apply_bc_file(vf_sim *h, int N, const double *domMin, const double *domMax, const std::string &bcPath) {
    std::ifstream f(bcPath); if (!f) throw std::runtime_error("Couldn't open " + bcPath);
    std::stringstream ss; ss << f.rdbuf(); const std::string s = ss.str();
    const auto regs = MiniJSON(s).regions();
    std::vector<int32_t> kind, cmask; std::vector<double> val, lo, hi;
    for (const auto &r : regs) {
        kind.push_back(r.kind); cmask.push_back(r.cmask);
        for (int d = 0; d < 3; ++d) {
            val.push_back(r.value[d]);
            const bool act = d < N;
            const double a = act ? domMin[d] : 0.0, b = act ? domMax[d] : 0.0;
            lo.push_back(r.relative && act ? a + r.lo[d] * (b - a) : r.lo[d]);
            hi.push_back(r.relative && act ? a + r.hi[d] * (b - a) : r.hi[d]);
        }
    }
    check(vf_sim_apply_bc_regions(h, (int)regs.size(), kind.data(), cmask.data(), val.data(), lo.data(), hi.data()));
}
} // namespace detail

template<typename Real_, size_t... Degrees> class MultigridSolver;

// ---------------------------------------------------------------------------------------------------------------------
template<typename Real_, size_t... Degrees>
class TensorProductSimulator {
public:
    static constexpr size_t N = sizeof...(Degrees);
    static_assert(std::is_same<Real_, double>::value, "only the double instantiations exist (python_bindings/VoxelFEM.cc:301-308)");
    static_assert((N == 2 || N == 3) && ((Degrees == 1) && ...), "only Q1 elements in 2D and 3D are instantiated");
    using Scalar = Real_;
    using VNd = std::array<double, N>;
    using EigenNDIndex = std::array<size_t, N>;
    using VField = voxelfem_b200::VField;
    using VXd = voxelfem_b200::VXd;
    struct BBoxN { VNd minCorner, maxCorner; };

    TensorProductSimulator(const BBoxN &domain, const EigenNDIndex &elementsPerDimension) : m_domain(domain), m_ne(elementsPerDimension) {
        int64_t ne[3] = {1, 1, 1}; double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
        for (size_t d = 0; d < N; ++d) { ne[d] = (int64_t)m_ne[d]; lo[d] = domain.minCorner[d]; hi[d] = domain.maxCorner[d]; }
        check(vf_sim_create((int)N, ne, lo, hi, &m_h));
    }
    ~TensorProductSimulator() { if (m_h) vf_sim_destroy(m_h); }
    TensorProductSimulator(const TensorProductSimulator &) = delete;
    TensorProductSimulator &operator=(const TensorProductSimulator &) = delete;

    size_t numNodes() const { return (size_t)vf_sim_num_nodes(m_h); }
    size_t numElements() const { return (size_t)vf_sim_num_elements(m_h); }
    const EigenNDIndex &NbElementsPerDimension() const { return m_ne; }
    EigenNDIndex NbNodesPerDimension() const { EigenNDIndex r = m_ne; for (auto &v : r) ++v; return r; }
    const BBoxN &domain() const { return m_domain; }

    // material: isotropic ETensor(E, nu) (MeshFEM ElasticityTensor.hh:100-115) or the flattened (N(N+1)/2)^2 tensor
    void setIsotropicETensor(double E, double nu) { check(vf_sim_set_isotropic(m_h, E, nu)); m_et = ETensorState{true, E, nu, {}}; }
    void setETensor(const std::vector<double> &flattened) { check(vf_sim_set_elasticity_tensor(m_h, flattened.data())); m_et = ETensorState{false, 0, 0, flattened}; }
    void readMaterial(const std::string &materialPath) {   // isotropic_material files (MeshFEM Materials.cc:291-311)
        std::ifstream f(materialPath); if (!f) throw std::runtime_error("Couldn't open material " + materialPath);
        std::stringstream ss; ss << f.rdbuf(); const std::string s = ss.str();
        auto field = [&](const char *k) { const size_t p = s.find(std::string("\"") + k + "\""); if (p == std::string::npos) throw std::runtime_error(std::string("material file lacks ") + k); return std::stod(s.substr(s.find(':', p) + 1)); };
        if (s.find("isotropic_material") == std::string::npos) throw std::runtime_error("only isotropic_material files are supported");
        setIsotropicETensor(field("young"), field("poisson"));
    }
    std::vector<double> fullDensityElementStiffnessMatrix() const { const size_t k = N << N; std::vector<double> r(k * k); check(vf_sim_get_K0(m_h, r.data())); return r; }

    InterpolationLaw interpolationLaw() const { return m_law; }
    double E_0() const { return m_E0; }  double E_min() const { return m_Emin; }
    double SIMPExponent() const { return m_gamma; }  double RAMPFactor() const { return m_q; }
    void setInterpolationLaw(InterpolationLaw law) { m_law = law; pushInterp(); }
    void setE_0(double v) { m_E0 = v; pushInterp(); }        void setE_min(double v) { m_Emin = v; pushInterp(); }
    void setSIMPExponent(double v) { m_gamma = v; pushInterp(); }  void setRAMPFactor(double v) { m_q = v; pushInterp(); }
    void setGravity(const VNd &g) { double v[3] = {0, 0, 0}; for (size_t d = 0; d < N; ++d) v[d] = g[d]; m_gravity = g; check(vf_sim_set_gravity(m_h, v)); }
    const VNd &getGravity() const { return m_gravity; }

    void setDensities(const VXd &rho) { if (rho.size() != numElements()) throw std::runtime_error("Density vector size mismatch"); check(vf_sim_set_densities(m_h, rho.data())); }
    void setUniformDensities(double density) { check(vf_sim_set_uniform_density(m_h, density)); }
    VXd getDensities() const { VXd r(numElements()); check(vf_sim_get_densities(m_h, r.data())); return r; }
    VXd getYoungModulusScaleFactor() const { VXd r(numElements()); check(vf_sim_get_young_moduli(m_h, r.data())); return r; }
    void setFabricationMaskHeightByLayer(size_t l) { check(vf_sim_set_mask_layer(m_h, (int64_t)l)); }

    void applyDisplacementsAndLoadsFromFile(const std::string &bcPath) {
        double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (size_t d = 0; d < N; ++d) { lo[d] = m_domain.minCorner[d]; hi[d] = m_domain.maxCorner[d]; }
        detail::apply_bc_file(m_h, (int)N, lo, hi, bcPath);
    }
    void addDirichletCondition(const VNd &u, const VNd &minCorner, const VNd &maxCorner, const std::string &componentMask = "xyz") {
        double uu[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; int cm = 0;
        for (size_t d = 0; d < N; ++d) { uu[d] = u[d]; lo[d] = minCorner[d]; hi[d] = maxCorner[d]; }
        for (char c : componentMask) cm |= 1 << (std::tolower(c) - 'x');
        check(vf_sim_add_dirichlet_box(m_h, uu, lo, hi, cm));
    }
    std::vector<uint8_t> getDirichletMask() const { std::vector<uint8_t> m(numNodes()); check(vf_sim_get_dirichlet_mask(m_h, m.data())); return m; }
    VField buildLoadVector() const { VField f(numNodes(), N); check(vf_sim_build_load_vector(m_h, f.data())); return f; }

    VField applyK(const VField &u) const { VField r(numNodes(), N); check(vf_sim_apply_K(m_h, u.data(), r.data(), 1, 0)); return r; }
    template<bool ZeroInit = true, bool Negate = false>
    void applyK(const VField &u, VField &result) const { if (ZeroInit) result.setZero(numNodes(), N); check(vf_sim_apply_K(m_h, u.data(), result.data(), ZeroInit, Negate)); }
    VField solve(const VField &f) const { VField u(numNodes(), N); check(vf_sim_solve(m_h, f.data(), u.data())); return u; }
    VXd complianceGradientFlattened(const VField &u) const { VXd g(numElements()); check(vf_sim_compliance_gradient(m_h, u.data(), g.data(), 0)); return g; }
    VXd elementEnergyDensity(const VField &u) const { VXd e(numElements()); check(vf_sim_element_energy_density(m_h, u.data(), e.data())); return e; }

    // ---- element / node bookkeeping (TensorProductSimulator.hh:1532-1651; NDVector.hh:249-275) ----
    VNd getStretchings() const { VNd r; for (size_t d = 0; d < N; ++d) r[d] = (m_domain.maxCorner[d] - m_domain.minCorner[d]) / double(m_ne[d]); return r; }
    double elementVolume(size_t /* ei */ = 0) const { double v = 1; for (double h : getStretchings()) v *= h; return v; }
    EigenNDIndex unflattenNode(size_t ni) const { EigenNDIndex r; for (size_t d = N; d-- > 0;) { r[d] = ni % (m_ne[d] + 1); ni /= (m_ne[d] + 1); } return r; }
    EigenNDIndex unflattenElement(size_t ei) const { EigenNDIndex r; for (size_t d = N; d-- > 0;) { r[d] = ei % m_ne[d]; ei /= m_ne[d]; } return r; }
    VNd nodePosition(size_t ni) const { const auto c = unflattenNode(ni); const auto h = getStretchings(); VNd p; for (size_t d = 0; d < N; ++d) p[d] = m_domain.minCorner[d] + double(c[d]) * h[d]; return p; }
    size_t elementIndexForGridCell(const EigenNDIndex &cellIdxs) const { size_t r = 0; for (size_t d = 0; d < N; ++d) r = r * m_ne[d] + cellIdxs[d]; return r; }
    size_t elemNodeGlobalIndex(size_t ei, size_t n) const {
        const auto e = unflattenElement(ei); size_t r = 0;
        for (size_t d = 0; d < N; ++d) r = r * (m_ne[d] + 1) + e[d] + ((n >> (N - 1 - d)) & 1);
        return r;
    }
    std::vector<size_t> elementNodes(size_t ei) const { std::vector<size_t> r(size_t(1) << N); for (size_t n = 0; n < r.size(); ++n) r[n] = elemNodeGlobalIndex(ei, n); return r; }
    double elementDensity(size_t ei) const { return getDensities().at(ei); }
    double elementYoungModulusScaleFactor(size_t ei) const { return getYoungModulusScaleFactor().at(ei); }
    std::vector<double> elementStiffnessMatrix(size_t ei) const { auto K = fullDensityElementStiffnessMatrix(); const double E = elementYoungModulusScaleFactor(ei); for (double &v : K) v *= E; return K; }
    void setDensity(size_t ei, double value) { auto rho = getDensities(); rho.at(ei) = value; setDensities(rho); }
    void setDensitiesFromCoarseGrid(size_t upscalingFactor, const VXd &rho) {   // (:309-320 of the reference's density setters)
        EigenNDIndex cs; size_t nc = 1; for (size_t d = 0; d < N; ++d) { cs[d] = m_ne[d] / upscalingFactor; nc *= cs[d]; }
        if (rho.size() != nc) throw std::runtime_error("Density vector size mismatch");
        VXd fine(numElements());
        for (size_t e = 0; e < fine.size(); ++e) { const auto c = unflattenElement(e); size_t ci = 0; for (size_t d = 0; d < N; ++d) ci = ci * cs[d] + c[d] / upscalingFactor; fine[e] = rho[ci]; }
        setDensities(fine);
    }
    double getFabricationMaskHeight() const { int64_t a = 0, b = 0; double h = 0; check(vf_sim_get_mask_info(m_h, &a, &b, &h)); return h; }
    void applySymmetryConditions(const std::array<bool, N> &symmetry_axes, const std::array<bool, N> &minMaxFace = {}) {
        int axes = 0, faces = 0; for (size_t d = 0; d < N; ++d) { axes |= int(symmetry_axes[d]) << d; faces |= int(minMaxFace[d]) << d; }
        check(vf_sim_apply_symmetry_conditions(m_h, axes, faces));
    }
    // material as set (needed to configure derived simulators)
    struct ETensorState { bool isotropic = true; double E = 1, nu = 0; std::vector<double> D; };
    const ETensorState &getETensor() const { return m_et; }
    void setETensor(const ETensorState &et) { if (et.isotropic) setIsotropicETensor(et.E, et.nu); else setETensor(et.D); }

    // ---- layer-by-layer helpers (:1852-1923) and downsampling (:1926-1992) ----
    static constexpr size_t BUILD_DIRECTION = 1;
    std::shared_ptr<TensorProductSimulator> getIntermediateFabricationShape(double hfrac, bool validateBoundaryConditions = true, InterpolationLaw law = InterpolationLaw::SIMP) const {
        if (hfrac < 0 || hfrac > 1) throw std::runtime_error("hfrac is out of bounds");
        const size_t full = m_ne[BUILD_DIRECTION]; const size_t nh = (size_t)std::llround(hfrac * double(full));
        if (std::abs(hfrac * double(full) - double(nh)) > 1e-10) throw std::runtime_error("hfrac chops off a noninteger number of element layers");
        auto ne = m_ne; ne[BUILD_DIRECTION] = nh;
        auto sub = m_domain; sub.maxCorner[BUILD_DIRECTION] = sub.minCorner[BUILD_DIRECTION] + hfrac * (sub.maxCorner[BUILD_DIRECTION] - sub.minCorner[BUILD_DIRECTION]);
        auto r = std::make_shared<TensorProductSimulator>(sub, ne);
        r->setInterpolationLaw(law);
        r->setETensor(m_et);
        transferDensitiesToIntermediateFabricationShape(*r);
        r->setE_min(m_Emin); r->setE_0(m_E0); r->setSIMPExponent(m_gamma);           // the RAMP factor keeps its default (:1881-1883)
        const std::runtime_error unexpectedBC("Original simulator has unexpected boundary conditions for layer-by-layer simulation");
        double g2 = 0; for (double g : m_gravity) g2 += g * g;
        if (g2 == 0) { if (validateBoundaryConditions) throw unexpectedBC; VNd g{}; g[BUILD_DIRECTION] = -1; r->setGravity(g); }
        else r->setGravity(m_gravity);
        if (validateBoundaryConditions) {                                                // (:1893-1907)
            if (vf_sim_num_force_nodes(m_h) != 0 || vf_sim_num_nonzero_dirichlet_values(m_h) != 0) throw unexpectedBC;
            const auto mask = getDirichletMask(); const uint8_t fullMask = uint8_t((1u << N) - 1u);
            for (size_t ni = 0; ni < mask.size(); ++ni) { const bool base = unflattenNode(ni)[BUILD_DIRECTION] == 0; if (base ? mask[ni] != fullMask : mask[ni] != 0) throw unexpectedBC; }
        }
        double ext = 0; for (size_t d = 0; d < N; ++d) ext = std::max(ext, sub.maxCorner[d] - sub.minCorner[d]);
        VNd lo = sub.minCorner, hi = sub.maxCorner, zero{}; const double eps = 1e-9 * ext;
        for (size_t d = 0; d < N; ++d) { lo[d] -= eps; hi[d] += eps; }
        hi[BUILD_DIRECTION] = sub.minCorner[BUILD_DIRECTION] + eps;
        r->addDirichletCondition(zero, lo, hi);                                          // build platform fully clamped (:1913-1920)
        return r;
    }
    void transferDensitiesToIntermediateFabricationShape(TensorProductSimulator &inter) const {
        const auto rho = getDensities(); VXd out(inter.numElements());
        for (size_t e = 0; e < out.size(); ++e) out[e] = rho[elementIndexForGridCell(inter.unflattenElement(e))];
        inter.setDensities(out);
    }
    std::shared_ptr<TensorProductSimulator> downsample(size_t downsamplingLevels) const {
        const size_t f = size_t(1) << downsamplingLevels; auto ne = m_ne;
        for (auto &n : ne) { if (n % f) throw std::runtime_error("Grid size must be divisible by 2^downsamplingLevels"); n /= f; }
        auto r = std::make_shared<TensorProductSimulator>(m_domain, ne);
        r->setETensor(m_et); r->setE_min(m_Emin); r->setE_0(m_E0); r->setSIMPExponent(m_gamma);
        return r;
    }
    size_t downsamplingFactor(const TensorProductSimulator &coarse) const {
        const size_t f = m_ne[0] / coarse.m_ne[0];
        for (size_t d = 0; d < N; ++d) if (coarse.m_ne[d] * f != m_ne[d]) throw std::runtime_error("Invalid downsampled simulator");
        return f;
    }
    void downsampleDensityFieldTo(const VXd &densities, TensorProductSimulator &coarse) const {
        const size_t f = downsamplingFactor(coarse);
        if (densities.size() != numElements()) throw std::runtime_error("Invalid input densities size (" + std::to_string(densities.size()) + " vs " + std::to_string(numElements()) + ")");
        VXd c(coarse.numElements(), 0.0); const double w = 1.0 / std::pow(double(f), double(N));
        for (size_t e = 0; e < densities.size(); ++e) { auto q = unflattenElement(e); for (auto &v : q) v /= f; c[coarse.elementIndexForGridCell(q)] += densities[e]; }
        for (double &v : c) v *= w;
        coarse.setDensities(c);
    }
    VXd upsampleDensityGradientFrom(const TensorProductSimulator &coarse, const VXd &g_coarse) const {
        const size_t f = downsamplingFactor(coarse);
        if (g_coarse.size() != coarse.numElements()) throw std::runtime_error("Invalid coarse gradient size");
        VXd g(numElements()); const double w = 1.0 / std::pow(double(f), double(N));
        for (size_t e = 0; e < g.size(); ++e) { auto q = unflattenElement(e); for (auto &v : q) v /= f; g[e] = g_coarse[coarse.elementIndexForGridCell(q)] * w; }
        return g;
    }

    vf_sim *handle() const { return m_h; }
private:
    ETensorState m_et{true, 1.0, 0.0, {}};                         // ETensor(1, 0), TensorProductSimulator.hh:2114
    void pushInterp() { check(vf_sim_set_interpolation(m_h, (int)m_law, m_E0, m_Emin, m_gamma, m_q)); }
    BBoxN m_domain; EigenNDIndex m_ne; vf_sim *m_h = nullptr; VNd m_gravity{};
    InterpolationLaw m_law = InterpolationLaw::SIMP; double m_E0 = 1, m_Emin = 1e-4, m_gamma = 3, m_q = 3;   // TensorProductSimulator.hh:2160-2166
};

// ---------------------------------------------------------------------------------------------------------------------
template<typename Real_, size_t... Degrees>
class MultigridSolver {
public:
    using TPS = TensorProductSimulator<Real_, Degrees...>;
    static constexpr size_t N = TPS::N;
    using VField = voxelfem_b200::VField;
    using PCGCallback = std::function<void(size_t, const VField &, const VField &)>;   // (it, x, r), MultigridSolver.hh:1043-1045
    using MGCallback = std::function<void(size_t, const VField &)>;

    MultigridSolver(std::shared_ptr<TPS> fineSimulator, size_t numCoarseningLevels) : m_fine(std::move(fineSimulator)) {
        check(vf_mg_create(m_fine->handle(), (int)numCoarseningLevels, &m_h));
    }
    ~MultigridSolver() { if (m_h) vf_mg_destroy(m_h); }
    MultigridSolver(const MultigridSolver &) = delete;
    MultigridSolver &operator=(const MultigridSolver &) = delete;

    TPS &getSimulator(size_t l = 0) { if (l != 0) throw std::runtime_error("coarse-level simulators live on the device; use numNodes(l)"); return *m_fine; }
    std::shared_ptr<TPS> getSimulatorPtr() const { return m_fine; }
    size_t numLevels() const { return (size_t)vf_mg_num_levels(m_h); }
    size_t numNodes(size_t l) const { return (size_t)vf_mg_level_num_nodes(m_h, (int)l); }

    void updateStiffnessMatrices() { check(vf_mg_update_stiffness_matrices(m_h)); }
    void setSymmetricGaussSeidel(bool symmetric) { check(vf_mg_set_symmetric_gauss_seidel(m_h, symmetric)); }
    void setFabricationMaskHeightByLayer(size_t l) { check(vf_mg_set_mask_layer(m_h, (int64_t)l)); }
    void decrementFabricationMaskHeightByLayer(size_t inc) { check(vf_mg_decrement_mask(m_h, (int)inc)); }

    VField applyK(size_t l, const VField &u) { VField r(numNodes(l), N); check(vf_mg_apply_K(m_h, (int)l, u.data(), r.data())); return r; }
    void computeResidual(size_t l, const VField &u, const VField &b, VField &r) { r.setZero(numNodes(l), N); check(vf_mg_compute_residual(m_h, (int)l, u.data(), b.data(), r.data())); }

    // solve (:546-573); the callback sees the iterate after every cycle
    VField solve(const VField &u, const VField &f, size_t numSteps, size_t numSmoothingSteps, bool stiffnessUpdated = false, bool zeroDirichlet = false,
                 MGCallback it_callback = nullptr, bool fmg = false) {
        VField x(u.rows(), u.cols());
        if (!it_callback) { check(vf_mg_solve(m_h, u.data(), f.data(), (int)numSteps, (int)numSmoothingSteps, stiffnessUpdated, zeroDirichlet, fmg, x.data())); return x; }
        VField cur = u;
        for (size_t i = 0; i < numSteps; ++i) {
            check(vf_mg_solve(m_h, cur.data(), f.data(), 1, (int)numSmoothingSteps, stiffnessUpdated || i > 0, zeroDirichlet, fmg && i == 0, x.data()));
            it_callback(i, x); cur = x;
        }
        return x;
    }

    // preconditionedConjugateGradient (:1047-1152): x is the initial guess on entry and the solution on exit
    void preconditionedConjugateGradient(VField &x, const VField &b, size_t maxIter, double tol, PCGCallback cb = nullptr,
                                         size_t mgIterations = 1, size_t mgSmoothingIterations = 1, bool fmg = false) {
        struct Tramp { MultigridSolver *self; PCGCallback *cb; size_t n; } t{this, &cb, x.rows()};
        auto call = [](int it, double, void *user) {
            Tramp &tr = *static_cast<Tramp *>(user);
            VField xi(tr.n, N), ri(tr.n, N);   // the fields are fetched only because a callback asked for them
            check(vf_mg_get_pcg_iterate(tr.self->m_h, xi.data())); check(vf_mg_get_pcg_residual(tr.self->m_h, ri.data()));
            (*tr.cb)((size_t)it, xi, ri);
        };
        int iters = 0;
        check(vf_mg_pcg(m_h, x.data(), b.data(), (int)maxIter, tol, (int)mgIterations, (int)mgSmoothingIterations, fmg, 0, &iters, nullptr,
                        cb ? static_cast<vf_pcg_callback>(call) : nullptr, &t));
        m_lastIters = (size_t)iters;
    }
    size_t lastPCGIterations() const { return m_lastIters; }
    VField pcgResidual() { VField r(numNodes(0), N); check(vf_mg_get_pcg_residual(m_h, r.data())); return r; }

    // zeroOutDirichletComponents (:511-524) on a host field of level l
    void zeroOutDirichletComponents(size_t l, VField &u) const {
        std::vector<uint8_t> m(numNodes(l)); check(vf_mg_level_dirichlet_mask(m_h, (int)l, m.data()));
        for (size_t n = 0; n < m.size(); ++n) for (size_t c = 0; c < N; ++c) if ((m[n] >> c) & 1) u(n, c) = 0.0;
    }
    std::vector<uint8_t> levelDirichletMask(size_t l) const { std::vector<uint8_t> m(numNodes(l)); check(vf_mg_level_dirichlet_mask(m_h, (int)l, m.data())); return m; }
    VField debug_get_x(size_t l) { VField r(numNodes(l), N); check(vf_mg_debug_get(m_h, 0, (int)l, r.data())); return r; }
    VField debug_get_b(size_t l) { VField r(numNodes(l), N); check(vf_mg_debug_get(m_h, 1, (int)l, r.data())); return r; }
    std::vector<int32_t> debugMulticolorVisit() { std::vector<int32_t> r(numNodes(0)); check(vf_mg_debug_multicolor_visit(m_h, r.data())); return r; }

    vf_mg *handle() const { return m_h; }
private:
    std::shared_ptr<TPS> m_fine; vf_mg *m_h = nullptr; size_t m_lastIters = 0;
};

// ---------------------------------------------------------------------------------------------------------------------
// Filters (TopologyOptimizationFilter.hh:17-712).  A filter maps a flat row-major grid array to another; inside a
// TopologyOptimizationProblem the chain runs on the device (vf_top_*, described there by spec()); apply / backprop here are the
// stand-alone host-array entry points (vf_filter_*), which is what FilterChain and the python bindings use.
using GridDims = std::vector<size_t>;
template<typename Real_> struct Filter {
    using VXd = voxelfem_b200::VXd;
    virtual ~Filter() = default;
    virtual std::array<double, 4> spec() const = 0;
    virtual VXd apply(const VXd &in) = 0;                                       // Filter::apply (:25)
    virtual VXd backprop(const VXd &d_dout, const VXd &vars) const = 0;        // Filter::backprop (:31)
    const GridDims &inputDimensions() const { return m_inputDims; }
    const GridDims &outputDimensions() const { return m_outputDims; }
    void setInputDimensions(const GridDims &dims) { m_setInputDimensions(dims); m_gridDimsAreSet = true; }
    void setOutputDimensions(const GridDims &dims) { m_setOutputDimensions(dims); m_gridDimsAreSet = true; }
    void checkGridDimensionsAreSet() const { if (!m_gridDimsAreSet) throw std::runtime_error("Filter grid dimensions not set. Initialize a TopologyOptimizationProblem object with this filter before using it."); }
    static size_t numEntries(const GridDims &d) { size_t n = 1; for (size_t v : d) n *= v; return n; }
protected:
    virtual void m_setInputDimensions(const GridDims &dims) { m_inputDims = m_outputDims = dims; }
    virtual void m_setOutputDimensions(const GridDims &dims) { m_inputDims = m_outputDims = dims; }
    std::vector<int64_t> dims64(const GridDims &d) const { return std::vector<int64_t>(d.begin(), d.end()); }
    void checkIn(const VXd &x, const GridDims &d) const { checkGridDimensionsAreSet(); if (x.size() != numEntries(d)) throw std::runtime_error("Input dimension mismatch"); }
    GridDims m_inputDims, m_outputDims; bool m_gridDimsAreSet = false;
};
template<typename Real_> struct SmoothingFilter : Filter<Real_> {   // TopologyOptimizationFilter.hh:283-400
    using VXd = voxelfem_b200::VXd;
    enum class Type { Const = VF_SMOOTH_CONST, Linear = VF_SMOOTH_LINEAR };
    SmoothingFilter(size_t r = 1, Type t = Type::Const) : radius(r), type(t) {}
    size_t radius; Type type;
    std::array<double, 4> spec() const override { return {double(VF_FILTER_SMOOTH), double(radius), double(int(type)), 0.0}; }
    VXd apply(const VXd &in) override { this->checkIn(in, this->m_inputDims); VXd out(in.size()); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_smooth((int)d.size(), d.data(), (int)radius, int(type), in.data(), out.data())); return out; }
    VXd backprop(const VXd &g, const VXd &) const override { this->checkIn(g, this->m_outputDims); VXd out(g.size()); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_smooth((int)d.size(), d.data(), (int)radius, int(type), g.data(), out.data())); return out; }   // symmetric operator (:297-310)
};
template<typename Real_> struct ProjectionFilter : Filter<Real_> {  // TopologyOptimizationFilter.hh:189-245
    using VXd = voxelfem_b200::VXd;
    explicit ProjectionFilter(Real_ beta = 1) { setBeta(beta); }
    Real_ getBeta() const { return m_beta; }
    void setBeta(Real_ beta) { if (beta <= 0) throw std::runtime_error("Beta parameter has to be positive (received beta = " + std::to_string(beta) + ")"); m_beta = beta; }
    Real_ invert(Real_ filteredValue) const {
        if (filteredValue > 1.0 || filteredValue < 0.0) throw std::runtime_error("ProjectionFilter::invert domain error: target density for inversion is outside [0, 1].");
        return std::atanh((2 * filteredValue - 1) * std::tanh(0.5 * m_beta)) / m_beta + 0.5;
    }
    std::array<double, 4> spec() const override { return {double(VF_FILTER_PROJECT), 0.0, 0.0, double(m_beta)}; }
    VXd apply(const VXd &in) override { this->checkIn(in, this->m_inputDims); VXd out(in.size()); check(vf_filter_project((int64_t)in.size(), m_beta, in.data(), out.data())); return out; }
    VXd backprop(const VXd &g, const VXd &vars) const override { VXd out(g.size()); check(vf_filter_project_backprop((int64_t)g.size(), m_beta, g.data(), vars.data(), out.data())); return out; }
private:
    Real_ m_beta = 1;
};
template<typename Real_> struct UpsampleFilter : Filter<Real_> {    // TopologyOptimizationFilter.hh:418-523
    using VXd = voxelfem_b200::VXd;
    explicit UpsampleFilter(size_t factor = 2) : m_factor(factor) {}
    std::array<double, 4> spec() const override { return {double(VF_FILTER_UPSAMPLE), double(m_factor), 0.0, 0.0}; }
    VXd apply(const VXd &in) override { this->checkIn(in, this->m_inputDims); VXd out(this->numEntries(this->m_outputDims)); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_upsample((int)d.size(), d.data(), (int)m_factor, in.data(), out.data())); return out; }
    VXd backprop(const VXd &g, const VXd &) const override { this->checkIn(g, this->m_outputDims); VXd out(this->numEntries(this->m_inputDims)); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_upsample_backprop((int)d.size(), d.data(), (int)m_factor, g.data(), out.data())); return out; }
protected:
    void m_setInputDimensions(const GridDims &dims) override {
        for (size_t v : dims) if (v < 2) throw std::runtime_error("Interpolation can only be applied to a 2^d grid or larger");
        this->m_inputDims = dims; this->m_outputDims = dims; for (auto &v : this->m_outputDims) v = (v - 1) * m_factor + 1;
    }
    void m_setOutputDimensions(const GridDims &dims) override {
        for (size_t v : dims) if (v < 2) throw std::runtime_error("Interpolation can only be applied to a 2^d grid or larger");
        this->m_outputDims = dims; this->m_inputDims = dims;
        for (size_t d = 0; d < dims.size(); ++d) { this->m_inputDims[d] = (dims[d] - 1) / m_factor + 1; if ((this->m_inputDims[d] - 1) * m_factor + 1 != dims[d]) throw std::runtime_error("Output size is not divisible by factor"); }
    }
private:
    size_t m_factor;
};
template<typename Real_> struct VertexToCellFilter : Filter<Real_> {   // TopologyOptimizationFilter.hh:528-598
    using VXd = voxelfem_b200::VXd;
    std::array<double, 4> spec() const override { return {double(VF_FILTER_VERTEX_TO_CELL), 0.0, 0.0, 0.0}; }
    VXd apply(const VXd &in) override { this->checkIn(in, this->m_inputDims); VXd out(this->numEntries(this->m_outputDims)); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_vertex_to_cell((int)d.size(), d.data(), in.data(), out.data())); return out; }
    VXd backprop(const VXd &g, const VXd &) const override { this->checkIn(g, this->m_outputDims); VXd out(this->numEntries(this->m_inputDims)); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_vertex_to_cell_backprop((int)d.size(), d.data(), g.data(), out.data())); return out; }
protected:
    void m_setInputDimensions(const GridDims &dims) override { for (size_t v : dims) if (v < 2) throw std::runtime_error("Input grid must be 2^d or larger."); this->m_inputDims = dims; this->m_outputDims = dims; for (auto &v : this->m_outputDims) v -= 1; }
    void m_setOutputDimensions(const GridDims &dims) override { this->m_outputDims = dims; this->m_inputDims = dims; for (auto &v : this->m_inputDims) v += 1; }
};
template<typename Real_> struct LangelaarFilter : Filter<Real_> {      // TopologyOptimizationFilter.hh:601-712
    using VXd = voxelfem_b200::VXd;
    std::array<double, 4> spec() const override { return {double(VF_FILTER_LANGELAAR), 0.0, 0.0, 0.0}; }
    VXd apply(const VXd &in) override {
        this->checkIn(in, this->m_inputDims); const auto d = this->dims64(this->m_inputDims);
        if (m_cachedFiltered.size() != in.size()) { m_cachedFiltered.assign(in.size(), 0.0); m_cachedSmax.assign(in.size(), 0.0); }
        check(vf_filter_langelaar((int)d.size(), d.data(), in.data(), m_cachedFiltered.data(), m_cachedSmax.data()));   // output array is in/out (header note)
        return m_cachedFiltered;
    }
    VXd backprop(const VXd &g, const VXd &vars) const override {
        if (m_cachedFiltered.size() != g.size()) throw std::runtime_error("LangelaarFilter::backprop before apply");
        VXd out(g.size()); const auto d = this->dims64(this->m_inputDims);
        check(vf_filter_langelaar_backprop((int)d.size(), d.data(), g.data(), vars.data(), m_cachedFiltered.data(), m_cachedSmax.data(), out.data())); return out;
    }
private:
    VXd m_cachedFiltered, m_cachedSmax;
};
template<typename Real_> struct PythonFilter : Filter<Real_> {         // TopologyOptimizationFilter.hh:247-275
    using VXd = voxelfem_b200::VXd;
    using ApplyCallback = std::function<void(const VXd &in, VXd &out)>;
    using BackpropCallback = std::function<void(const VXd &in, const VXd &vars, VXd &out)>;
    std::array<double, 4> spec() const override { return {double(VF_FILTER_PYTHON), 0.0, 0.0, 0.0}; }
    VXd apply(const VXd &in) override { if (!apply_cb) throw std::runtime_error("Apply callback must be configured"); VXd out(this->numEntries(this->m_outputDims), 0.0); apply_cb(in, out); return out; }
    VXd backprop(const VXd &g, const VXd &vars) const override { if (!backprop_cb) throw std::runtime_error("Backprop callback must be configured"); VXd out(this->numEntries(this->m_inputDims), 0.0); backprop_cb(g, vars, out); return out; }
    ApplyCallback apply_cb; BackpropCallback backprop_cb;
};
// FilterChain (TopologyOptimizationFilter.hh:90-187) on host arrays
template<typename Real_> struct FilterChain {
    using VXd = voxelfem_b200::VXd;
    using Filters = std::vector<std::shared_ptr<Filter<Real_>>>;
    FilterChain(const Filters &f, const GridDims &outGridDims) : m_filters(f) {
        GridDims dims = outGridDims;
        for (size_t i = m_filters.size(); i-- > 0;) { m_filters[i]->setOutputDimensions(dims); dims = m_filters[i]->inputDimensions(); }
        m_inDims = dims; m_outDims = outGridDims;
        m_vars.assign(m_filters.size() + 1, VXd());
        m_vars[0].assign(Filter<Real_>::numEntries(dims), 0.0);
        for (size_t i = 0; i < m_filters.size(); ++i) m_vars[i + 1].assign(Filter<Real_>::numEntries(m_filters[i]->outputDimensions()), 0.0);
    }
    size_t numVars() const { return m_vars.front().size(); }
    size_t numPhysicalVars() const { return m_vars.back().size(); }
    const GridDims &gridDims() const { return m_inDims; }
    const GridDims &physicalGridDims() const { return m_outDims; }
    void setDesignVars(const VXd &x) { if (x.size() != numVars()) throw std::runtime_error("Variable size mismatch"); m_vars[0] = x; for (size_t i = 0; i < m_filters.size(); ++i) m_vars[i + 1] = m_filters[i]->apply(m_vars[i]); }
    VXd backprop(VXd g) const { if (g.size() != numPhysicalVars()) throw std::runtime_error("Size mismatch"); for (size_t i = m_filters.size(); i-- > 0;) g = m_filters[i]->backprop(g, m_vars[i]); return g; }
    const VXd &designVars() const { return m_vars.front(); }
    const VXd &physicalVars() const { return m_vars.back(); }
    const Filters &filters() const { return m_filters; }
private:
    Filters m_filters; std::vector<VXd> m_vars; GridDims m_inDims, m_outDims;
};
template<typename Real_> struct Constraint { virtual ~Constraint() = default; };
template<typename Real_> struct TotalVolumeConstraint : Constraint<Real_> {   // TopologyOptimizationConstraint.hh:24-40
    explicit TotalVolumeConstraint(Real_ volumeFraction) : m_volumeFraction(volumeFraction) {}
    Real_ m_volumeFraction;
};

template<typename Sim> struct Objective { virtual ~Objective() = default; };
template<typename Sim> class TopologyOptimizationProblem;
template<typename Sim>
struct MultigridComplianceObjective : Objective<Sim> {            // TopologyOptimizationObjective.hh:60-105
    template<typename MGS> explicit MultigridComplianceObjective(std::shared_ptr<MGS> mg_) : m_handle(mg_->handle()), m_keep(mg_) {}
    size_t cgIter = 100; double tol = 1e-5; size_t mgIterations = 1, mgSmoothingIterations = 2; bool fullMultigrid = true, zeroInit = false;
    std::function<void(size_t, const VField &)> residual_cb;       // (:104) invoked per PCG iteration of updateCache with (it, r)
    vf_mg *mgHandle() const { return m_handle; }
    std::shared_ptr<void> mgHolder() const { return m_keep; }
    // compliance / u / f / gradient of the attached problem (ComplianceObjective, :27-57)
    double compliance() const { return problem().TopologyOptimizationProblem<Sim>::evaluateObjective(); }
    VField u() const { return problem().displacement(); }
    VField f() const { return problem().getSimulator().buildLoadVector(); }
    VXd gradient() const { return problem().getSimulator().complianceGradientFlattened(u()); }   // dJ/d(rho_phys) (:45-47)
    void attach(const TopologyOptimizationProblem<Sim> *p) { m_problem = p; }
private:
    const TopologyOptimizationProblem<Sim> &problem() const { if (!m_problem) throw std::runtime_error("objective is not attached to a TopologyOptimizationProblem yet"); return *m_problem; }
    vf_mg *m_handle; std::shared_ptr<void> m_keep; const TopologyOptimizationProblem<Sim> *m_problem = nullptr;
};

template<typename Sim>
class TopologyOptimizationProblem {                               // TopologyOptimizationProblem.hh:17-155
public:
    using VXd = voxelfem_b200::VXd;
    using ObjectivePtr = std::shared_ptr<MultigridComplianceObjective<Sim>>;
    using ConstraintsList = std::vector<std::shared_ptr<Constraint<double>>>;
    using FiltersList = std::vector<std::shared_ptr<Filter<double>>>;
    TopologyOptimizationProblem(Sim &simulator, ObjectivePtr objective, const ConstraintsList &constraints, const FiltersList &filters)
        : m_sim(simulator), m_objective(std::move(objective)), m_constraints(constraints), m_filters(filters) {
        if (constraints.size() != 1) throw std::runtime_error("exactly one TotalVolumeConstraint is supported (OptimalityCriterion.hh:43-45)");
        auto tvc = std::dynamic_pointer_cast<TotalVolumeConstraint<double>>(constraints[0]);
        if (!tvc) throw std::runtime_error("exactly one TotalVolumeConstraint is supported (OptimalityCriterion.hh:43-45)");
        GridDims dims(simulator.NbElementsPerDimension().begin(), simulator.NbElementsPerDimension().end());
        for (size_t i = filters.size(); i-- > 0;) { filters[i]->setOutputDimensions(dims); dims = filters[i]->inputDimensions(); }   // FilterChain::setOutputDimensions (:117-132)
        std::vector<double> spec;
        for (const auto &f : filters) { const auto sp = f->spec(); spec.insert(spec.end(), sp.begin(), sp.end()); }
        if (spec.empty()) spec.push_back(0.0);
        check(vf_top_create(m_objective->mgHandle(), (int)filters.size(), spec.data(), tvc->m_volumeFraction, &m_h));
        for (size_t i = 0; i < filters.size(); ++i) {              // PythonFilter: host callbacks inside the device chain
            auto *pf = dynamic_cast<PythonFilter<double> *>(filters[i].get());
            if (!pf) continue;
            auto ap = [](const double *in, int64_t nin, double *out, int64_t nout, void *user) -> int {
                try { auto &f = *static_cast<PythonFilter<double> *>(user); if (!f.apply_cb) throw std::runtime_error("Apply callback must be configured");
                      VXd o((size_t)nout, 0.0); f.apply_cb(VXd(in, in + nin), o); std::copy(o.begin(), o.end(), out); return 0; } catch (...) { s_pending() = std::current_exception(); return 1; } };
            auto bp = [](const double *g, int64_t nout, const double *vars, int64_t nin, double *out, void *user) -> int {
                try { auto &f = *static_cast<PythonFilter<double> *>(user); if (!f.backprop_cb) throw std::runtime_error("Backprop callback must be configured");
                      VXd o((size_t)nin, 0.0); f.backprop_cb(VXd(g, g + nout), VXd(vars, vars + nin), o); std::copy(o.begin(), o.end(), out); return 0; } catch (...) { s_pending() = std::current_exception(); return 1; } };
            check(vf_top_set_python_filter(m_h, (int)i, ap, bp, pf));
        }
        m_objective->attach(this);
    }
    virtual ~TopologyOptimizationProblem() { if (m_h) vf_top_destroy(m_h); }
    // true for subclasses that override setVars / evaluateObjective / evaluateObjectiveGradientAndReturn on the host (the python trampoline)
    virtual bool hasHostOverrides() const { return false; }
    TopologyOptimizationProblem(const TopologyOptimizationProblem &) = delete;
    TopologyOptimizationProblem &operator=(const TopologyOptimizationProblem &) = delete;

    size_t numVars() const { return (size_t)vf_top_num_vars(m_h); }
    size_t numPhysicalVars() const { return (size_t)vf_top_num_physical_vars(m_h); }
    virtual bool setVars(const VXd &x, bool /* forceUpdate */ = false) { syncSolver(); if (x.size() != numVars()) throw std::runtime_error("Size mismatch"); call(vf_top_set_vars(m_h, x.data())); return true; }
    VXd getVars() const { VXd r(numVars()); check(vf_top_get_vars(m_h, 0, r.data())); return r; }
    VXd getDensities() const { VXd r(numPhysicalVars()); check(vf_top_get_vars(m_h, 1, r.data())); return r; }
    virtual double evaluateObjective() const { double v = 0; check(vf_top_compliance(m_h, &v)); return v; }
    virtual VXd evaluateObjectiveGradientAndReturn() const { VXd g(numVars()); call(vf_top_objective_gradient(m_h, g.data())); return g; }
    VXd evaluateConstraints() const { double v = 0; check(vf_top_constraint(m_h, &v)); return VXd(1, v); }
    VXd evaluateConstraintsJacobianAndReturn() const { VXd g(numVars()); call(vf_top_constraint_jacobian(m_h, g.data())); return g; }
    VField displacement() const { VField u(m_sim.numNodes(), Sim::N); check(vf_top_get_u(m_h, u.data())); return u; }
    int lastPCGIterations() const { return vf_top_last_pcg_iterations(m_h); }
    ObjectivePtr getObjective() const { return m_objective; }
    const FiltersList &getFilters() const { return m_filters; }
    const ConstraintsList &getConstraints() const { return m_constraints; }
    GridDims gridDims(bool physical = false) const { std::vector<int64_t> d(Sim::N); check(vf_top_get_grid_dims(m_h, physical, d.data())); return GridDims(d.begin(), d.end()); }
    // back-propagation of an arbitrary physical-space gradient through the problem's filters at the current design variables
    VXd backpropThroughFilters(const VXd &g) const { FilterChain<double> fc(m_filters, gridDims(true)); fc.setDesignVars(getVars()); return fc.backprop(g); }
    Sim &getSimulator() const { return m_sim; }
    void syncSolver() const {
        const auto &o = *m_objective;
        auto rcb = [](int it, double, void *user) {                  // residual_cb(it, r): the residual is fetched only because a callback is set
            const auto &self = *static_cast<const TopologyOptimizationProblem *>(user);
            if (!self.m_objective->residual_cb) return;
            VField r(self.m_sim.numNodes(), Sim::N); check(vf_mg_get_pcg_residual(self.m_objective->mgHandle(), r.data()));
            try { self.m_objective->residual_cb((size_t)it, r); } catch (...) { s_pending() = std::current_exception(); }
        };
        check(vf_top_set_residual_callback(m_h, o.residual_cb ? static_cast<vf_pcg_callback>(rcb) : nullptr, const_cast<TopologyOptimizationProblem *>(this)));
        check(vf_top_set_solver(m_h, (int)o.cgIter, o.tol, (int)o.mgIterations, (int)o.mgSmoothingIterations, o.fullMultigrid, o.zeroInit));
    }
    vf_top *handle() const { return m_h; }
    // a C-ABI call during which PythonFilter callbacks may run: their exception (if any) is rethrown instead of the generic status
    static void call(int rc) { if (s_pending()) { auto e = s_pending(); s_pending() = nullptr; std::rethrow_exception(e); } check(rc); }
private:
    static std::exception_ptr &s_pending() { static thread_local std::exception_ptr p; return p; }
    Sim &m_sim; ObjectivePtr m_objective; ConstraintsList m_constraints; FiltersList m_filters; vf_top *m_h = nullptr;
};

template<typename Problem>
class OCOptimizer {                                               // OptimalityCriterion.hh:38-149
public:
    explicit OCOptimizer(Problem &p) : m_p(p) {}
    // inplace: gradient, multiplier search and setVars all stay on the device (vf_top_oc_step).  !inplace (:57-60): the gradient comes
    // from the problem's virtual evaluateObjectiveGradientAndReturn(), the search runs on the device, and the stepped variables go
    // through the problem's virtual setVars (:133) -- the path a subclassed problem (trampoline, VoxelFEM.cc:58-66) needs.
    void step(double m = 0.2, double p = 0.5, double ctol = 1e-6, bool inplace = true) {
        m_p.syncSolver(); int evals = 0;
        if (inplace && !m_p.hasHostOverrides()) Problem::call(vf_top_oc_step(m_p.handle(), m, p, ctol, &evals));
        else {
            const auto dJ = m_p.evaluateObjectiveGradientAndReturn();
            std::vector<double> stepped(m_p.numVars());
            Problem::call(vf_top_oc_search(m_p.handle(), dJ.data(), m, p, ctol, stepped.data(), &evals));
            m_p.setVars(stepped);
        }
        m_lastEvals = evals;
    }
    int lastConstraintEvaluations() const { return m_lastEvals; }
private:
    Problem &m_p; int m_lastEvals = 0;
};

template<typename TPS>
class LayerByLayerEvaluator {                                     // LayerByLayer.hh:25-309
public:
    using VXd = voxelfem_b200::VXd;
    using VField = voxelfem_b200::VField;
    using LBLCallback = std::function<void(size_t, double, const VXd &, const VField &)>;   // cb(l, compliance, grad_compliance, u) (:222)
    explicit LayerByLayerEvaluator(std::shared_ptr<TPS> lblSim) : m_sim(std::move(lblSim)) {}
    ~LayerByLayerEvaluator() { if (m_h) vf_lbl_destroy(m_h); }
    void selectInitMethod(const std::string &method) { m_method = method; if (m_h) check(vf_lbl_select_init_method(m_h, method.c_str())); }
    // run (:223-296); it_callback(it, x, r) is handed to every layer's PCG (:265)
    template<typename MG>
    void run(MG &solver, bool zeroInit, size_t layerIncrement, size_t maxIter, double tol, typename MG::PCGCallback it_callback = nullptr,
             size_t mgIterations = 1, size_t mgSmoothingIterations = 1, bool fullMultigrid = false, bool verbose = false, LBLCallback lblCallback = nullptr) {
        if (!m_h || m_mg != solver.handle()) {
            if (m_h) vf_lbl_destroy(m_h);
            m_h = nullptr; check(vf_lbl_create(solver.handle(), &m_h)); m_mg = solver.handle();
            check(vf_lbl_select_init_method(m_h, m_method.c_str()));
        }
        struct Ctx { LayerByLayerEvaluator *self; LBLCallback *cb; bool verbose; typename MG::PCGCallback *pcg; vf_mg *mg; size_t nn; std::exception_ptr err; }
            ctx{this, &lblCallback, verbose, &it_callback, solver.handle(), m_sim->numNodes(), nullptr};
        auto layerCb = [](int64_t layer, double c, int, void *user) {
            Ctx &x = *static_cast<Ctx *>(user);
            if (x.err) return;
            try {
                if (x.verbose) std::cout << "Layer " << layer << ": " << c << std::endl;   // (:275-276)
                if (*x.cb) {   // the two fields cross the bus only because a callback asked for them
                    VXd g(x.self->m_sim->numElements()); VField u(x.nn, TPS::N);
                    check(vf_lbl_get_layer_gradient(x.self->m_h, g.data())); check(vf_lbl_get_layer_u(x.self->m_h, u.data()));
                    (*x.cb)((size_t)layer, c, g, u);
                }
            } catch (...) { x.err = std::current_exception(); }
        };
        auto pcgCb = [](int it, double, void *user) {
            Ctx &x = *static_cast<Ctx *>(user);
            if (x.err) return;
            try {
                VField xi(x.nn, TPS::N), ri(x.nn, TPS::N);
                check(vf_mg_get_pcg_iterate(x.mg, xi.data())); check(vf_mg_get_pcg_residual(x.mg, ri.data()));
                (*x.pcg)((size_t)it, xi, ri);
            } catch (...) { x.err = std::current_exception(); }
        };
        check(vf_lbl_run(m_h, zeroInit, (int64_t)layerIncrement, (int)maxIter, tol, (int)mgIterations, (int)mgSmoothingIterations, fullMultigrid,
                         (lblCallback || verbose) ? static_cast<vf_lbl_callback>(layerCb) : nullptr, &ctx,
                         it_callback ? static_cast<vf_pcg_callback>(pcgCb) : nullptr, &ctx));
        if (ctx.err) std::rethrow_exception(ctx.err);
    }
    double objective() const { double v = 0; check(vf_lbl_objective(m_h, &v)); return v; }
    VXd gradient() const { VXd g(m_sim->numElements()); check(vf_lbl_gradient(m_h, g.data())); return g; }
private:
    std::shared_ptr<TPS> m_sim; vf_lbl *m_h = nullptr; vf_mg *m_mg = nullptr; std::string m_method = "N=3";
};

class MMA {                                                       // MethodOfMovingAsymptotes.hh:28-470
public:
    using VXd = voxelfem_b200::VXd;
    using F = std::function<VXd(const VXd &)>;      // m + 1 values
    using DF = std::function<VXd(const VXd &)>;     // (m + 1) x n, row-major
    MMA(int numVars, int numConstr, const VXd &xmin, const VXd &xmax, F f, DF df_dx) : m_n(numVars), m_m(numConstr), m_f(std::move(f)), m_df(std::move(df_dx)) {
        check(vf_mma_create(numVars, numConstr, xmin.data(), xmax.data(), &m_h));
    }
    ~MMA() { if (m_h) vf_mma_destroy(m_h); }
    MMA(const MMA &) = delete;
    MMA &operator=(const MMA &) = delete;
    void enableGCMMA(bool enable) { check(vf_mma_enable_gcmma(m_h, enable)); }
    void setInitialVar(const VXd &x0) { check(vf_mma_set_initial_var(m_h, x0.data())); }
    void step() {
        auto f = [](const double *x, double *out, void *user) -> int {
            MMA &s = *static_cast<MMA *>(user);
            try { const VXd v = s.m_f(VXd(x, x + s.m_n)); std::copy(v.begin(), v.end(), out); return 0; } catch (...) { return 1; }
        };
        auto df = [](const double *x, double *out, void *user) -> int {
            MMA &s = *static_cast<MMA *>(user);
            try { const VXd v = s.m_df(VXd(x, x + s.m_n)); std::copy(v.begin(), v.end(), out); return 0; } catch (...) { return 1; }
        };
        check(vf_mma_step(m_h, f, df, this, 0));
    }
    VXd getOptimalVar() const { VXd x(m_n); check(vf_mma_get_optimal_var(m_h, x.data())); return x; }
private:
    int m_n, m_m; F m_f; DF m_df; vf_mma *m_h = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
// Degree-2 elements: TensorProductSimulator<double, 2, 2[, 2]> on the reference's generic element path (TPSStencils.hh:163-185),
// vf_q2_* of the C ABI.  Node grid (2 ne + 1)^N; no multigrid (the reference has no degree-2 instantiation of it either).
// ---------------------------------------------------------------------------------------------------------------------
template<size_t N_>
class TensorProductSimulatorQ2 {
public:
    static constexpr size_t N = N_;
    static_assert(N == 2 || N == 3, "2D and 3D");
    using VNd = std::array<double, N>; using EigenNDIndex = std::array<size_t, N>;
    using VField = voxelfem_b200::VField; using VXd = voxelfem_b200::VXd;
    struct BBoxN { VNd minCorner, maxCorner; };
    TensorProductSimulatorQ2(const BBoxN &domain, const EigenNDIndex &elementsPerDimension) : m_ne(elementsPerDimension) {
        int64_t ne[3] = {1, 1, 1}; double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
        for (size_t d = 0; d < N; ++d) { ne[d] = (int64_t)m_ne[d]; lo[d] = domain.minCorner[d]; hi[d] = domain.maxCorner[d]; }
        check(vf_q2_create((int)N, ne, lo, hi, &m_h));
    }
    ~TensorProductSimulatorQ2() { if (m_h) vf_q2_destroy(m_h); }
    TensorProductSimulatorQ2(const TensorProductSimulatorQ2 &) = delete;
    TensorProductSimulatorQ2 &operator=(const TensorProductSimulatorQ2 &) = delete;
    size_t numNodes() const { return (size_t)vf_q2_num_nodes(m_h); }
    size_t numElements() const { return (size_t)vf_q2_num_elements(m_h); }
    EigenNDIndex NbNodesPerDimension() const { EigenNDIndex r = m_ne; for (auto &v : r) v = 2 * v + 1; return r; }
    void setIsotropicETensor(double E, double nu) { check(vf_q2_set_isotropic(m_h, E, nu)); }
    void setInterpolation(InterpolationLaw law, double E_0, double E_min, double gamma, double q) { check(vf_q2_set_interpolation(m_h, (int)law, E_0, E_min, gamma, q)); }
    void setDensities(const VXd &rho) { if (rho.size() != numElements()) throw std::runtime_error("Incorrect rho size"); check(vf_q2_set_densities(m_h, rho.data())); }
    std::vector<double> fullDensityElementStiffnessMatrix() const { size_t k = N; for (size_t d = 0; d < N; ++d) k *= 3; std::vector<double> K(k * k); check(vf_q2_get_K0(m_h, K.data())); return K; }
    template<bool ZeroInit = true, bool Negate = false>
    void applyK(const VField &u, VField &result) const { if (ZeroInit) result.setZero(numNodes(), N); check(vf_q2_apply_K(m_h, u.data(), result.data(), ZeroInit, Negate)); }
    VField applyK(const VField &u) const { VField r; applyK<true, false>(u, r); return r; }
    VXd elementEnergies(const VField &u) const { VXd e(numElements()); check(vf_q2_element_energies(m_h, u.data(), e.data())); return e; }
    // Jacobi-preconditioned CG with the flagged components (component-major, numNodes * N) clamped to zero; returns the iteration count
    int solve(VField &x, const VField &b, const std::vector<uint8_t> &fixed, int maxIter, double tol, double *relResidual = nullptr) const {
        int it = 0; check(vf_q2_pcg(m_h, x.data(), b.data(), fixed.data(), maxIter, tol, &it, relResidual)); return it;
    }
    vf_q2 *handle() const { return m_h; }
private:
    EigenNDIndex m_ne; vf_q2 *m_h = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
// Compliance topology optimization on a 3D grid cut into slabs along axis 0 (BASELINE.json configs[3]): TopologyOptimizationProblem +
// MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer over vf_group_top_* (no reference equivalent: the reference is
// single-address-space).  Local group: all slabs live in this process on the current device (one per entry of `slabs`); NCCL rank:
// one slab per process, the 128-byte id of vf_nccl_unique_id() broadcast by the caller.  Design variables and gradients cross this
// interface as arrays over the WHOLE grid.
// ---------------------------------------------------------------------------------------------------------------------
class SlabTopologyOptimizationProblem {
public:
    using VXd = voxelfem_b200::VXd;
    struct Filter { int kind; int radius; int type; double beta; };      // kind 0: SmoothingFilter(radius, type), 1: ProjectionFilter(beta)
    struct Setup {
        std::array<double, 3> domainMin{{0, 0, 0}}, domainMax{{1, 1, 1}}; std::array<int64_t, 3> elements{{2, 2, 2}};
        std::string bcPath; double young = 1.0, poisson = 0.3, E_0 = 1.0, E_min = 1e-4, gamma = 3.0;
        int numCoarseningLevels = 1, firstReplicatedLevel = 1;
        std::vector<Filter> filters; double volumeFraction = 0.3;
    };
    // local group over the element-layer ranges `slabs` (ordered, contiguous, multiples of 2^firstReplicatedLevel)
    SlabTopologyOptimizationProblem(const Setup &s, const std::vector<std::pair<int64_t, int64_t>> &slabs) : m_ne(s.elements) {
        for (const auto &r : slabs) addPart(s, r.first, r.second, m_sims.empty() ? nullptr : m_sims[0]);
        check(vf_group_create_local((int)m_mgs.size(), m_mgs.data(), &m_grp));
        createProblem(s);
    }
    // one rank of an NCCL group
    SlabTopologyOptimizationProblem(const Setup &s, int64_t slabBegin, int64_t slabEnd, int rank, int world, const void *uniqueId128) : m_ne(s.elements) {
        addPart(s, slabBegin, slabEnd, nullptr);
        check(vf_group_create_nccl(m_mgs[0], rank, world, uniqueId128, &m_grp));
        createProblem(s);
    }
    ~SlabTopologyOptimizationProblem() {
        if (m_top) vf_group_top_destroy(m_top);
        if (m_grp) vf_group_destroy(m_grp);
        for (size_t i = m_mgs.size(); i-- > 0;) vf_mg_destroy(m_mgs[i]);
        for (size_t i = m_sims.size(); i-- > 0;) vf_sim_destroy(m_sims[i]);      // the first part owns the stream the others share
    }
    SlabTopologyOptimizationProblem(const SlabTopologyOptimizationProblem &) = delete;
    SlabTopologyOptimizationProblem &operator=(const SlabTopologyOptimizationProblem &) = delete;
    size_t numVars() const { return (size_t)(m_ne[0] * m_ne[1] * m_ne[2]); }
    void setSolver(int cgIter, double tol, int mgIterations, int mgSmoothingIterations, bool fullMultigrid, bool zeroInit) {
        check(vf_group_top_set_solver(m_top, cgIter, tol, mgIterations, mgSmoothingIterations, fullMultigrid, zeroInit));
    }
    void setVars(const VXd &x) { if (x.size() != numVars()) throw std::runtime_error("Incorrect number of variables"); check(vf_group_top_set_vars(m_top, x.data())); }
    VXd getVars() const { VXd x(numVars()); check(vf_group_top_get_vars(m_top, 0, x.data())); return x; }
    VXd getDensities() const { VXd x(numVars()); check(vf_group_top_get_vars(m_top, 1, x.data())); return x; }
    double evaluateObjective() const { double v = 0; check(vf_group_top_compliance(m_top, &v)); return v; }
    double evaluateConstraint() const { double v = 0; check(vf_group_top_constraint(m_top, &v)); return v; }
    VXd evaluateObjectiveGradient() const { VXd g(numVars()); check(vf_group_top_objective_gradient(m_top, g.data())); return g; }
    VXd evaluateConstraintsJacobian() const { VXd g(numVars()); check(vf_group_top_constraint_jacobian(m_top, g.data())); return g; }
    int ocStep(double m = 0.2, double p = 0.5, double ctol = 1e-6) { int n = 0; check(vf_group_top_oc_step(m_top, m, p, ctol, &n)); return n; }   // OCOptimizer::step; returns the constraint evaluations
    int lastPCGIterations() const { return vf_group_top_last_pcg_iterations(m_top); }
    int64_t filterHaloLayers() const { return vf_group_top_halo_layers(m_top); }
private:
    void addPart(const Setup &s, int64_t b, int64_t e, vf_sim *shareStreamWith) {
        vf_sim *sim = nullptr;
        check(vf_sim_create_slab(3, s.elements.data(), s.domainMin.data(), s.domainMax.data(), b, e, shareStreamWith, &sim));
        m_sims.push_back(sim);
        check(vf_sim_set_isotropic(sim, s.young, s.poisson));
        check(vf_sim_set_interpolation(sim, (int)InterpolationLaw::SIMP, s.E_0, s.E_min, s.gamma, 3.0));
        if (!s.bcPath.empty()) detail::apply_bc_file(sim, 3, s.domainMin.data(), s.domainMax.data(), s.bcPath);
        check(vf_sim_set_uniform_density(sim, 1.0));
        vf_mg *mg = nullptr;
        check(vf_mg_create_slab(sim, s.numCoarseningLevels, s.firstReplicatedLevel, &mg));
        m_mgs.push_back(mg);
    }
    void createProblem(const Setup &s) {
        std::vector<double> spec;
        for (const Filter &f : s.filters) { spec.push_back(f.kind); spec.push_back(f.radius); spec.push_back(f.type); spec.push_back(f.beta); }
        if (spec.empty()) spec.push_back(0.0);
        check(vf_group_top_create(m_grp, (int)s.filters.size(), spec.data(), s.volumeFraction, &m_top));
    }
    std::array<int64_t, 3> m_ne; std::vector<vf_sim *> m_sims; std::vector<vf_mg *> m_mgs; vf_group *m_grp = nullptr; vf_gtop *m_top = nullptr;
};

} // namespace voxelfem_b200
