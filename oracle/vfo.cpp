// oracle/vfo.cpp -- CPU ORACLE: TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
//
// A plain C++/OpenMP restatement (no Eigen/TBB/CHOLMOD) of the reference VoxelFEM
// algorithm for the MG-PCG / topology-optimization hot path, used (a) as the parity
// checker for the CUDA path in tests/ and __graft_entry__.smoke(), and (b) as the timed
// `cpu_baseline` / `--impl reference` arm of bench.py.  Nothing under voxelfem_b200/
// may include, link or call this file.
//
// PARITY STATUS: "parity unpinned" by the reference's own tests -- VoxelFEM ships no
// tests, golden vectors or stored outputs for this path and
// cannot be compiled here (Eigen 3.3.7, TBB, CHOLMOD, Boost, nlohmann-json absent).
// The oracle is instead pinned by analytic invariants and by an independent
// numpy/scipy assembly + sparse direct solve in tests/test_oracle_*.py.
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference).
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace vfo {

using idx = int64_t;
static thread_local std::string g_err;

constexpr int SIMD_WIDTH = 4;          // VOXELFEM_SIMD_WIDTH, TPSStencils.hh:18-20
constexpr int BUILD_DIRECTION = 1;     // TensorProductSimulator.hh:1851
constexpr int LAYER_MASK_NONE = INT_MAX; // TensorProductSimulator.hh:2179

// ---------------------------------------------------------------------------
// Dense banded Cholesky: stands in for CHOLMOD (SparseMatrices.hh:1984-2131) at
// the coarsest level. Any backward-stable SPD solve agrees to ~cond*eps.
// ---------------------------------------------------------------------------
struct BandChol {
    idx n = 0, bw = 0;            // bw = half bandwidth (number of sub-diagonals)
    std::vector<double> L;        // L(i, j) for i - bw <= j <= i stored at L[i * (bw + 1) + (j - i + bw)]
    double &at(idx i, idx j) { return L[i * (bw + 1) + (j - i + bw)]; }
    double  at(idx i, idx j) const { return L[i * (bw + 1) + (j - i + bw)]; }
    void init(idx n_, idx bw_) { n = n_; bw = bw_; L.assign(size_t(n) * (bw + 1), 0.0); }
    // In-place factorization A = L L^T (lower band of A stored in L on entry).
    void factor() {
        for (idx j = 0; j < n; ++j) {
            double d = at(j, j);
            idx k0 = std::max<idx>(0, j - bw);
            for (idx k = k0; k < j; ++k) d -= at(j, k) * at(j, k);
            if (!(d > 0)) throw std::runtime_error("Cholesky failure: matrix not positive definite");
            d = std::sqrt(d);
            at(j, j) = d;
            idx iend = std::min(n, j + bw + 1);
            #pragma omp parallel for schedule(static) if (iend - j > 256)
            for (idx i = j + 1; i < iend; ++i) {
                double s = at(i, j);
                idx kk0 = std::max<idx>(std::max<idx>(0, i - bw), k0);
                for (idx k = kk0; k < j; ++k) s -= at(i, k) * at(j, k);
                at(i, j) = s / d;
            }
        }
    }
    void solve(std::vector<double> &x) const {
        for (idx i = 0; i < n; ++i) {
            double s = x[i];
            for (idx k = std::max<idx>(0, i - bw); k < i; ++k) s -= at(i, k) * x[k];
            x[i] = s / at(i, i);
        }
        for (idx i = n - 1; i >= 0; --i) {
            double s = x[i];
            idx kend = std::min(n, i + bw + 1);
            for (idx k = i + 1; k < kend; ++k) s -= at(k, i) * x[k];
            x[i] = s / at(i, i);
        }
    }
};

// ---------------------------------------------------------------------------
// Simulator (TensorProductSimulator<double, 1, 1[, 1]>)
// ---------------------------------------------------------------------------
struct Sim {
    int N = 3;
    int npe = 8, ke = 24;                // nodes per element, Ke size
    idx ne[3] = {1, 1, 1}, nn[3] = {1, 1, 1};
    double dmin[3] = {0, 0, 0}, dmax[3] = {1, 1, 1};
    double stretch[3] = {1, 1, 1}, spacing[3] = {1, 1, 1};
    idx numNodes = 0, numElems = 0;
    idx ninc[3] = {0, 0, 0}, einc[3] = {0, 0, 0};
    idx refNodes[8];
    double D[6][6];                      // flattened elasticity tensor (ElasticityTensor.hh:118-131)
    std::vector<double> K0;              // ke x ke (symmetric)
    std::vector<double> rho, E;
    int law = 0;                         // 0 SIMP, 1 RAMP
    double E0 = 1, Emin = 1e-4, gamma = 3, q = 3; // TensorProductSimulator.hh:2137-2141
    double gravity[3] = {0, 0, 0};

    std::vector<idx> dirNodes; std::vector<uint8_t> dirMask; std::vector<double> dirVals;
    std::vector<uint8_t> nodeDirMask;    // per node: bits of constrained components
    std::vector<idx> forceNodes; std::vector<double> forceVals;

    double maskHeight = std::numeric_limits<double>::infinity();
    int firstMasked = LAYER_MASK_NONE, firstDetached = LAYER_MASK_NONE;

    // coarse-level operator storage (MultigridSolver.hh:874-886)
    std::vector<double> stencil;         // [node][3^N][N*N]  ("blockK", full storage)
    bool hasStencil = false;
    std::vector<double> KeCache;         // [elem][ke*ke]
    bool hasKeCache = false;
    // direct solver state (TensorProductSimulator.hh:1198-1230)
    BandChol chol; bool factorOK = false; std::vector<idx> freeVars; std::vector<uint8_t> isFixedCache;

    int nstencil() const { return N == 3 ? 27 : 9; }

    Sim(int N_, const idx *ne_, const double *dmin_, const double *dmax_) {
        // TensorProductSimulator.hh:209-279
        N = N_; npe = 1 << N; ke = N * npe;
        numNodes = 1; numElems = 1;
        for (int d = 0; d < N; ++d) {
            ne[d] = ne_[d]; nn[d] = ne_[d] + 1;
            dmin[d] = dmin_[d]; dmax[d] = dmax_[d];
            numNodes *= nn[d]; numElems *= ne[d];
            spacing[d] = (dmax[d] - dmin[d]) / (double(nn[d]) - 1.0);
            stretch[d] = (dmax[d] - dmin[d]) / double(ne[d]);
        }
        // row-major increments, NDVector.hh:256-264
        idx ni = 1, ei = 1;
        for (int d = N - 1; d >= 0; --d) { ninc[d] = ni; einc[d] = ei; ni *= nn[d]; ei *= ne[d]; }
        for (int n = 0; n < npe; ++n) {
            idx off = 0;
            for (int d = 0; d < N; ++d) off += idx((n >> (N - 1 - d)) & 1) * ninc[d];
            refNodes[n] = off;
        }
        rho.assign(numElems, 0.0); // NDVector default-constructs to zero
        setIsotropic(1.0, 0.0);    // TensorProductSimulator.hh:2114
        updateYoungModuli();
        nodeDirMask.assign(numNodes, 0);
    }

    // ---- indexing (TensorProductSimulator.hh:1532-1651) ----
    idx flatNode(const idx *n) const { idx r = n[0]; for (int d = 1; d < N; ++d) r = r * nn[d] + n[d]; return r; }
    idx flatElem(const idx *e) const { idx r = e[0]; for (int d = 1; d < N; ++d) r = r * ne[d] + e[d]; return r; }
    void ndNode(idx n, idx *out) const { for (int d = N - 1; d >= 0; --d) { out[d] = n % nn[d]; n /= nn[d]; } }
    void ndElem(idx e, idx *out) const { for (int d = N - 1; d >= 0; --d) { out[d] = e % ne[d]; e /= ne[d]; } }
    idx firstNodeOfElem(const idx *e) const { idx r = 0; for (int d = 0; d < N; ++d) r += e[d] * ninc[d]; return r; }
    void nodePosition(const idx *n, double *p) const { for (int d = 0; d < N; ++d) p[d] = dmin[d] + double(n[d]) * spacing[d]; } // :349-351

    // ---- material / K0 ----
    // ElasticityTensor.hh:100-131 (2D = plane stress)
    void setIsotropic(double Ey, double nu) {
        double lambda = (nu * Ey) / ((1.0 + nu) * (1.0 - 2.0 * nu));
        double mu = Ey / (2.0 + 2.0 * nu);
        if (N == 2) lambda = (nu * Ey) / (1.0 - nu * nu);
        std::memset(D, 0, sizeof(D));
        if (N == 3) {
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) D[i][j] = lambda;
            for (int i = 0; i < 3; ++i) D[i][i] = lambda + 2 * mu;
            D[3][3] = D[4][4] = D[5][5] = mu;
        } else {
            D[0][0] = D[1][1] = lambda + 2 * mu; D[0][1] = D[1][0] = lambda; D[2][2] = mu;
        }
        updateK0();
        factorOK = false;
    }
    void setD(const double *Din) { // row-major flatLen x flatLen
        int fl = (N == 3) ? 6 : 3;
        std::memset(D, 0, sizeof(D));
        for (int i = 0; i < fl; ++i) for (int j = 0; j < fl; ++j) D[i][j] = Din[i * fl + j];
        updateK0(); factorOK = false;
    }
    // flattened symmetric index: (0,0)(1,1)(2,2)(1,2)(0,2)(0,1) in 3D; (0,0)(1,1)(0,1) in 2D (SymmetricMatrix.hh flattenIndices)
    int symIdx(int i, int j) const {
        if (i == j) return i;
        if (N == 2) return 2;
        return 6 - i - j; // (1,2)->3, (0,2)->4, (0,1)->5
    }
    // Element_T::Stiffness (TensorProductSimulator.hh:67-80), Strains::getStrains
    // (TensorProductPolynomialInterpolant.hh:204-231), 2-point Gauss rule on [0,1]
    // (TensorProductQuadrature.hh:134-143), m_updateK0 (:2078-2086).
    void updateK0() {
        K0.assign(size_t(ke) * ke, 0.0);
        const int fl = (N == 3) ? 6 : 3;
        const double gp[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
        double vol = 1; for (int d = 0; d < N; ++d) vol *= stretch[d];
        const int nq = 1 << N;
        std::vector<double> strain(size_t(ke) * fl); // strain of each vector basis function at the quadrature point
        for (int qi = 0; qi < nq; ++qi) {
            double xi[3] = {0, 0, 0};
            for (int d = 0; d < N; ++d) xi[d] = gp[(qi >> (N - 1 - d)) & 1];
            double w = 1.0 / nq; // each 1D weight is 0.5
            for (int j = 0; j < npe; ++j) {
                double g[3] = {0, 0, 0};
                for (int c = 0; c < N; ++c) {
                    double v = 1;
                    for (int d = 0; d < N; ++d) {
                        int bit = (j >> (N - 1 - d)) & 1;
                        if (d == c) v *= (bit ? 1.0 : -1.0) / stretch[d];
                        else        v *= bit ? xi[d] : (1.0 - xi[d]);
                    }
                    g[c] = v;
                }
                for (int c = 0; c < N; ++c) {
                    double *s = &strain[size_t(j * N + c) * fl];
                    for (int t = 0; t < fl; ++t) s[t] = 0;
                    for (int i = 0; i < N; ++i) s[symIdx(c, i)] = 0.5 * g[i];
                    s[symIdx(c, c)] = g[c];
                }
            }
            for (int a = 0; a < ke; ++a) {
                // sigma = D * shearDoubled(strain_b); a : sigma with off-diagonals doubled
                for (int b = a; b < ke; ++b) {
                    const double *sa = &strain[size_t(a) * fl], *sb = &strain[size_t(b) * fl];
                    double acc = 0;
                    for (int i = 0; i < fl; ++i) {
                        double sig = 0;
                        for (int j = 0; j < fl; ++j) sig += D[i][j] * (j >= N ? 2.0 : 1.0) * sb[j];
                        acc += (i >= N ? 2.0 : 1.0) * sa[i] * sig;
                    }
                    K0[size_t(a) * ke + b] += w * acc;
                }
            }
        }
        for (int a = 0; a < ke; ++a) for (int b = a; b < ke; ++b) {
            K0[size_t(a) * ke + b] *= vol;
            K0[size_t(b) * ke + a] = K0[size_t(a) * ke + b];
        }
    }
    double k0(int a, int b) const { return K0[size_t(a) * ke + b]; }

    // ---- densities / Young's moduli (TensorProductSimulator.hh:2055-2060, 2088-2102) ----
    double unmaskedE(idx e) const {
        if (law == 0) return Emin + std::pow(rho[e], gamma) * (E0 - Emin);
        return Emin + rho[e] * (E0 - Emin) / (1 + q * (1 - rho[e]));
    }
    bool elemMasked(const idx *e) const { return int(e[BUILD_DIRECTION]) >= firstMasked; }   // :383-385
    bool nodeDetached(const idx *n) const { return int(n[BUILD_DIRECTION]) >= firstDetached; } // :364-366
    idx elemLayer(idx e) const { idx nd[3]; ndElem(e, nd); return nd[BUILD_DIRECTION]; }
    void updateYoungModuli() {
        E.resize(numElems);
        const bool masked = maskHeight < dmax[BUILD_DIRECTION];
        #pragma omp parallel for schedule(static)
        for (idx e = 0; e < numElems; ++e) {
            idx nd[3]; ndElem(e, nd);
            E[e] = (masked && elemMasked(nd)) ? 0.0 : unmaskedE(e);
        }
        factorOK = false;
    }
    idx nondetachedNodes(int d) const { // :371-375
        if (d != BUILD_DIRECTION) return nn[d];
        return std::min<idx>(firstDetached, nn[d]);
    }
    idx nonmaskedElems(int d) const { // :377-381
        if (d != BUILD_DIRECTION) return ne[d];
        return std::min<idx>(firstMasked, ne[d]);
    }
    // :290-309
    void setMaskHeight(double h, bool updateE = true) {
        if (h < 0 || h > dmax[BUILD_DIRECTION]) throw std::runtime_error("Fabrication height out of range");
        maskHeight = h;
        firstMasked = int(std::ceil(maskHeight / stretch[BUILD_DIRECTION] - 1e-10));
        firstDetached = firstMasked * 1 + 1;
        if (updateE) updateYoungModuli();
    }
    void setMaskHeightByLayer(idx l) { setMaskHeight(spacing[BUILD_DIRECTION] * double(l)); } // :327-329
    // :311-324
    void decrementMaskByLayer(int inc) {
        if (idx(firstMasked) > ne[BUILD_DIRECTION]) throw std::runtime_error("Mask must already be applied");
        if (firstMasked < inc) throw std::runtime_error("Mask decrement of bounds");
        maskHeight -= inc * spacing[BUILD_DIRECTION];
        firstMasked -= inc;
        firstDetached = firstMasked + 1;
        #pragma omp parallel for schedule(static)
        for (idx e = 0; e < numElems; ++e) {
            idx l = elemLayer(e);
            if (l >= firstMasked && l < firstMasked + inc) E[e] = 0.0;
        }
    }

    // ---- boundary conditions (BCBuilder, TensorProductSimulator.hh:464-566) ----
    struct BCBuilder {
        Sim &s; std::vector<double> forces, dvals; std::vector<uint8_t> dmask;
        BCBuilder(Sim &s_) : s(s_), forces(size_t(s_.numNodes) * s_.N, 0.0), dvals(size_t(s_.numNodes) * s_.N, 0.0), dmask(s_.numNodes, 0) {
            for (size_t f = 0; f < s.forceNodes.size(); ++f)
                for (int c = 0; c < s.N; ++c) forces[s.forceNodes[f] * s.N + c] = s.forceVals[f * s.N + c];
            for (size_t k = 0; k < s.dirNodes.size(); ++k) {
                for (int c = 0; c < s.N; ++c) dvals[s.dirNodes[k] * s.N + c] = s.dirVals[k * s.N + c];
                dmask[s.dirNodes[k]] = s.dirMask[k];
            }
        }
        void setDirichlet(idx ni, const double *val, uint8_t cm) { // :492-507
            for (int c = 0; c < s.N; ++c) {
                if (!(cm >> c & 1)) continue;
                if (!(dmask[ni] >> c & 1)) { dmask[ni] |= uint8_t(1 << c); dvals[ni * s.N + c] = val[c]; }
                else if (std::abs(dvals[ni * s.N + c] - val[c]) > 1e-10) throw std::runtime_error("Conflicting dirichlet displacements.");
            }
        }
        void setDirichletComponent(idx ni, int d, double v) { dmask[ni] |= uint8_t(1 << d); dvals[ni * s.N + d] = v; }
        void setForce(idx ni, const double *f) { for (int c = 0; c < s.N; ++c) forces[ni * s.N + c] = f[c]; }
        void apply() { // :516-562
            s.dirNodes.clear(); s.dirMask.clear(); s.dirVals.clear();
            const uint8_t full = uint8_t((1 << s.N) - 1);
            s.nodeDirMask.assign(s.numNodes, 0);
            for (idx ni = 0; ni < s.numNodes; ++ni) {
                uint8_t m = dmask[ni] & full; // ComponentMask::count(dim) ignores z in 2D
                if (m) {
                    s.dirNodes.push_back(ni); s.dirMask.push_back(m);
                    for (int c = 0; c < s.N; ++c) s.dirVals.push_back(dvals[ni * s.N + c]);
                    s.nodeDirMask[ni] = m;
                }
            }
            s.forceNodes.clear(); s.forceVals.clear();
            for (idx ni = 0; ni < s.numNodes; ++ni) {
                double sq = 0; for (int c = 0; c < s.N; ++c) sq += forces[ni * s.N + c] * forces[ni * s.N + c];
                if (sq != 0.0) { s.forceNodes.push_back(ni); for (int c = 0; c < s.N; ++c) s.forceVals.push_back(forces[ni * s.N + c]); }
            }
        }
    };
    bool hasDirichlet(idx n) const { return nodeDirMask[n] != 0; }
    bool hasFullDirichlet(idx n) const { return nodeDirMask[n] == uint8_t((1 << N) - 1); }

    // applyDisplacementsAndLoads (:600-652). Regions are axis-aligned boxes in absolute
    // coordinates (the "box%" -> absolute conversion of BoundaryConditions.cc:310-316 is
    // done by the caller); kind 0 = dirichlet (with component mask), 1 = force.
    void applyBCs(int nreg, const int *kind, const int *cmask, const double *values, const double *bmin, const double *bmax) {
        if (dirNodes.size() + forceNodes.size() > 0) throw std::runtime_error("Boundary condition updates unsupported");
        BCBuilder b(*this);
        for (int r = 0; r < nreg; ++r) {
            const double *lo = bmin + 3 * r, *hi = bmax + 3 * r, *val = values + 3 * r;
            std::vector<idx> inRegion;
            for (idx ni = 0; ni < numNodes; ++ni) {
                idx nd[3]; double p[3]; ndNode(ni, nd); nodePosition(nd, p);
                bool in = true;
                for (int d = 0; d < N; ++d) in = in && (p[d] >= lo[d]) && (p[d] <= hi[d]); // Geometry.hh:276-279
                if (in) inRegion.push_back(ni);
            }
            if (kind[r] == 1) {
                if (inRegion.empty()) throw std::runtime_error("Force constraint region unmatched");
                double f[3]; for (int c = 0; c < N; ++c) f[c] = val[c] / double(inRegion.size()); // :629-630
                for (idx ni : inRegion) b.setForce(ni, f);
            } else {
                if (inRegion.empty()) throw std::runtime_error("Dirichlet region unmatched");
                for (idx ni : inRegion) b.setDirichlet(ni, val, uint8_t(cmask[r]));
                factorOK = false;
            }
        }
        b.apply();
    }
    void addDirichletCondition(const double *u, const double *lo, const double *hi, int cmask) { // :660-671
        BCBuilder b(*this);
        for (idx ni = 0; ni < numNodes; ++ni) {
            idx nd[3]; double p[3]; ndNode(ni, nd); nodePosition(nd, p);
            bool in = true;
            for (int d = 0; d < N; ++d) in = in && (p[d] >= lo[d]) && (p[d] <= hi[d]);
            if (in) b.setDirichlet(ni, u, uint8_t(cmask));
        }
        b.apply(); factorOK = false;
    }
    void zeroOutDirichlet(double *u) const { // :575-583
        for (size_t k = 0; k < dirNodes.size(); ++k)
            for (int d = 0; d < N; ++d) if (dirMask[k] >> d & 1) u[d * numNodes + dirNodes[k]] = 0;
    }
    void enforceDirichlet(double *u) const { // :585-593
        for (size_t k = 0; k < dirNodes.size(); ++k)
            for (int d = 0; d < N; ++d) if (dirMask[k] >> d & 1) u[d * numNodes + dirNodes[k]] = dirVals[k * N + d];
    }

    // buildLoadVector (:1269-1288); integratedShapeFunctions (:1257-1263) = 1/2^N for Q1
    void buildLoadVector(double *f) const {
        std::fill(f, f + size_t(numNodes) * N, 0.0);
        for (size_t k = 0; k < forceNodes.size(); ++k)
            for (int c = 0; c < N; ++c) f[c * numNodes + forceNodes[k]] = forceVals[k * N + c];
        double gsq = 0; for (int c = 0; c < N; ++c) gsq += gravity[c] * gravity[c];
        if (gsq != 0) {
            double vol = 1; for (int d = 0; d < N; ++d) vol *= stretch[d];
            const double intPhi = 1.0 / double(npe);
            for (idx e = 0; e < numElems; ++e) {
                idx nd[3]; ndElem(e, nd);
                if (elemMasked(nd)) continue;
                idx off = firstNodeOfElem(nd);
                for (int l = 0; l < npe; ++l)
                    for (int c = 0; c < N; ++c) f[c * numNodes + refNodes[l] + off] += gravity[c] * (intPhi * rho[e] * vol);
            }
        }
    }

    // ---- masked vector helpers (TensorProductSimulator.hh:413-459) ----
    // A "block" is the contiguous run of non-detached nodes for one outer (axis 0) index.
    idx maskedBlockSize(idx margin = 0) const {
        idx height = std::min<idx>(firstDetached, nn[BUILD_DIRECTION]);
        height = std::min<idx>(height + margin, nn[BUILD_DIRECTION]);
        idx inner = 1; for (int d = 2; d < N; ++d) inner *= nn[d];
        return height * inner;
    }
    template<class F> void maskedVisit(const F &f, idx margin = 0) const {
        idx bs = maskedBlockSize(margin);
        #pragma omp parallel for schedule(static)
        for (idx i = 0; i < nn[0]; ++i) f(i * ninc[0], bs);
    }
    double maskedDot(const double *u, const double *v) const {
        idx bs = maskedBlockSize(0); double tot = 0;
        #pragma omp parallel for schedule(static) reduction(+:tot)
        for (idx i = 0; i < nn[0]; ++i) {
            double s = 0;
            for (int c = 0; c < N; ++c) { const double *a = u + c * numNodes + i * ninc[0], *b = v + c * numNodes + i * ninc[0]; for (idx k = 0; k < bs; ++k) s += a[k] * b[k]; }
            tot += s;
        }
        return tot;
    }
    void maskedCopy(const double *in, double *out, idx margin = 0) const {
        maskedVisit([&](idx start, idx n) { for (int c = 0; c < N; ++c) std::memcpy(out + c * numNodes + start, in + c * numNodes + start, size_t(n) * sizeof(double)); }, margin);
    }
    void maskedZero(double *out, idx margin = 0) const {
        maskedVisit([&](idx start, idx n) { for (int c = 0; c < N; ++c) std::memset(out + c * numNodes + start, 0, size_t(n) * sizeof(double)); }, margin);
    }

    // ---- level-0 matrix-free stiffness apply ----
    // SpecializedTPSStencils<Real,1,1[,1]>::applyK<ZeroInit, Negate> (TPSStencils.hh:231-396, 431-728):
    //   f_n (=, +=, -=) sum_{incident e} E_e * K0[:, N*ln(e,n)+c]^T u_e,
    // visiting only non-detached nodes; detached entries are zero-filled when ZeroInit.
    // Row-vectorised over the fastest axis like the reference's SIMD formulation.
    void applyK0(const double *u, double *out, bool zeroInit, bool negate) const {
        const idx nyv = (N >= 2) ? nondetachedNodes(1) : 1;
        const idx nzr = nn[N - 1];     // row length (fastest axis)
        const idx nrows = numNodes / nzr;
        const int No = N - 1;          // number of outer dims
        #pragma omp parallel
        {
            std::vector<double> tmp(size_t(N) * nzr), acc(size_t(N) * nzr);
            #pragma omp for schedule(static)
            for (idx row = 0; row < nrows; ++row) {
                idx o[2] = {0, 0}; // outer nd index (x[, y])
                if (N == 3) { o[0] = row / nn[1]; o[1] = row % nn[1]; } else { o[0] = row; }
                const idx rowStart = row * nzr;
                const bool detached = (N == 3) ? (o[1] >= nyv) : false;
                if (N == 2) {
                    // In 2D the build direction is the fastest axis: only k < nyv visited.
                }
                if (detached) { if (zeroInit) for (int c = 0; c < N; ++c) std::fill(out + c * numNodes + rowStart, out + c * numNodes + rowStart + nzr, 0.0); continue; }
                const idx kvis = (N == 2) ? nyv : nzr; // visited nodes along the row
                for (int c = 0; c < N; ++c) {
                    double *a = &acc[size_t(c) * nzr];
                    if (zeroInit) std::fill(a, a + nzr, 0.0);
                    else std::memcpy(a, out + c * numNodes + rowStart, size_t(nzr) * sizeof(double));
                }
                // loop over incident element offsets in the outer dims and along the row
                const int nOuterOff = 1 << No;
                for (int oo = 0; oo < nOuterOff; ++oo) {
                    idx eo[2]; bool valid = true; int lnOuter = 0;
                    for (int d = 0; d < No; ++d) {
                        int minus = (oo >> (No - 1 - d)) & 1;   // 1 => element at offset -1 along d
                        eo[d] = o[d] - minus;
                        if (eo[d] < 0 || eo[d] >= ne[d]) valid = false;
                        lnOuter = (lnOuter << 1) | minus;       // local node bit = 1 when element is at -1
                    }
                    if (!valid) continue;
                    idx erow = 0, nrow0 = 0;
                    for (int d = 0; d < No; ++d) { erow += eo[d] * einc[d]; nrow0 += eo[d] * ninc[d]; }
                    for (int oz = 0; oz < 2; ++oz) { // oz = 1 => element at offset -1 along the row
                        const int ln = (lnOuter << 1) | oz;
                        const idx k0v = oz;                                   // first node with a valid element
                        const idx k1v = std::min<idx>(kvis, ne[N - 1] + oz);  // one past last
                        if (k1v <= k0v) continue;
                        for (int c = 0; c < N; ++c) std::fill(&tmp[size_t(c) * nzr + k0v], &tmp[size_t(c) * nzr + k1v], 0.0);
                        for (int m = 0; m < npe; ++m) {
                            // node m of the element whose first node (along row) is k - oz
                            idx noff = nrow0 + refNodes[m] - oz;
                            for (int dcomp = 0; dcomp < N; ++dcomp) {
                                const double *us = u + dcomp * numNodes + noff;
                                for (int c = 0; c < N; ++c) {
                                    const double kv = k0(N * m + dcomp, N * ln + c);
                                    double *t = &tmp[size_t(c) * nzr];
                                    for (idx k = k0v; k < k1v; ++k) t[k] += kv * us[k];
                                }
                            }
                        }
                        const double *Er = E.data() + erow - oz;
                        for (int c = 0; c < N; ++c) {
                            double *a = &acc[size_t(c) * nzr]; const double *t = &tmp[size_t(c) * nzr];
                            if (negate) for (idx k = k0v; k < k1v; ++k) a[k] -= Er[k] * t[k];
                            else        for (idx k = k0v; k < k1v; ++k) a[k] += Er[k] * t[k];
                        }
                    }
                }
                for (int c = 0; c < N; ++c) {
                    double *o_ = out + c * numNodes + rowStart; const double *a = &acc[size_t(c) * nzr];
                    std::memcpy(o_, a, size_t(kvis) * sizeof(double));
                    if (zeroInit && kvis < nzr) std::fill(o_ + kvis, o_ + nzr, 0.0);
                }
            }
        }
    }

    // Element-scatter apply with cached per-element Ke (TensorProductSimulator.hh:1418-1434);
    // serial-colour order replaced by a gather-free serial loop per colour.
    void applyKeCache(const double *u, double *out, bool zeroInit, bool negate) const {
        if (zeroInit) std::fill(out, out + size_t(numNodes) * N, 0.0);
        for (int color = 0; color < npe; ++color) {
            idx cnt[3] = {1, 1, 1}, off[3] = {0, 0, 0}; bool any = true;
            for (int d = 0; d < N; ++d) { off[d] = (color >> (N - 1 - d)) & 1; if (off[d] >= ne[d]) any = false; else cnt[d] = (ne[d] - 1 - off[d]) / 2 + 1; }
            if (!any) continue;
            const idx tot = cnt[0] * cnt[1] * cnt[2];
            #pragma omp parallel for schedule(static)
            for (idx t = 0; t < tot; ++t) {
                idx e[3]; idx r = t;
                for (int d = N - 1; d >= 0; --d) { e[d] = 2 * (r % cnt[d]) + off[d]; r /= cnt[d]; }
                const idx ei = flatElem(e), noff = firstNodeOfElem(e);
                const double *Ke = &KeCache[size_t(ei) * ke * ke];
                double ul[24], fl[24];
                for (int m = 0; m < npe; ++m) for (int c = 0; c < N; ++c) ul[N * m + c] = u[c * numNodes + refNodes[m] + noff];
                for (int a = 0; a < ke; ++a) { double s = 0; for (int b = 0; b < ke; ++b) s += Ke[size_t(a) * ke + b] * ul[b]; fl[a] = s; }
                for (int m = 0; m < npe; ++m) for (int c = 0; c < N; ++c) {
                    double &o = out[c * numNodes + refNodes[m] + noff];
                    if (negate) o -= fl[N * m + c]; else o += fl[N * m + c];
                }
            }
        }
    }

    // applyBlockK (TensorProductSimulator.hh:1500-1504 -> CSCMatrix::applyTransposeParallel,
    // SparseMatrices.hh:1613-1677): out_n (=,+=,-=) sum_delta S[n][delta] u_{n+delta}
    void applyStencil(const double *u, double *out, bool zeroInit, bool negate) const {
        const int ns = nstencil(), NN = N * N;
        #pragma omp parallel for schedule(static)
        for (idx n = 0; n < numNodes; ++n) {
            idx nd[3]; ndNode(n, nd);
            double acc[3] = {0, 0, 0};
            for (int s = 0; s < ns; ++s) {
                idx m = 0; bool ok = true; int r = s;
                idx dlt[3];
                for (int d = N - 1; d >= 0; --d) { dlt[d] = (r % 3) - 1; r /= 3; }
                for (int d = 0; d < N; ++d) { idx q = nd[d] + dlt[d]; if (q < 0 || q >= nn[d]) ok = false; m = m * nn[d] + q; }
                if (!ok) continue;
                const double *B = &stencil[(size_t(n) * ns + s) * NN];
                for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) acc[a] += B[a * N + b] * u[b * numNodes + m];
            }
            for (int a = 0; a < N; ++a) {
                double &o = out[a * numNodes + n];
                if (zeroInit) o = negate ? -acc[a] : acc[a];
                else if (negate) o -= acc[a]; else o += acc[a];
            }
        }
    }

    // computeComplianceGradient (:972-1005); g = d(1/2 f.u)/d rho
    void complianceGradient(const double *u, double *g, bool accumulate) const {
        double gsq = 0; for (int c = 0; c < N; ++c) gsq += gravity[c] * gravity[c];
        const bool selfWeight = gsq != 0;
        double vol = 1; for (int d = 0; d < N; ++d) vol *= stretch[d];
        const double intPhi = 1.0 / double(npe);
        #pragma omp parallel for schedule(static)
        for (idx e = 0; e < numElems; ++e) {
            idx nd[3]; ndElem(e, nd);
            if (elemMasked(nd)) { if (!accumulate) g[e] = 0; continue; } // accumulate variant visits unmasked layers only (:1016)
            idx off = firstNodeOfElem(nd);
            double ue[24];
            for (int l = 0; l < npe; ++l) for (int c = 0; c < N; ++c) ue[l * N + c] = u[c * numNodes + refNodes[l] + off];
            double uKu = 0;
            for (int a = 0; a < ke; ++a) { double s = 0; for (int b = 0; b < ke; ++b) s += k0(a, b) * ue[b]; uKu += ue[a] * s; }
            double val;
            if (law == 0) val = -0.5 * gamma * std::pow(rho[e], gamma - 1.0) * (E0 - Emin) * uKu;
            else          val = -0.5 * (1 + q) * (E0 - Emin) / std::pow(1 + q * (1 - rho[e]), 2) * uKu;
            if (selfWeight)
                for (int l = 0; l < npe; ++l) { double gd = 0; for (int c = 0; c < N; ++c) gd += gravity[c] * ue[l * N + c]; val += intPhi * gd * vol; }
            if (accumulate) g[e] += val; else g[e] = val;
        }
    }
    // elementEnergyDensity (:1057-1073)
    void elementEnergyDensity(const double *u, double *out) const {
        #pragma omp parallel for schedule(static)
        for (idx e = 0; e < numElems; ++e) {
            idx nd[3]; ndElem(e, nd); idx off = firstNodeOfElem(nd);
            double ue[24];
            for (int l = 0; l < npe; ++l) for (int c = 0; c < N; ++c) ue[l * N + c] = u[c * numNodes + refNodes[l] + off];
            double uKu = 0;
            for (int a = 0; a < ke; ++a) { double s = 0; for (int b = 0; b < ke; ++b) s += k0(a, b) * ue[b]; uKu += ue[a] * s; }
            out[e] = 0.5 * E[e] * uKu;
        }
    }

    // Per-element stiffness matrix (:1080-1083)
    void elementStiffness(idx e, double *Ke) const {
        if (hasKeCache) { std::memcpy(Ke, &KeCache[size_t(e) * ke * ke], sizeof(double) * ke * ke); return; }
        for (int i = 0; i < ke * ke; ++i) Ke[i] = E[e] * K0[i];
    }

    // findFixedVars (:1181-1195) + assembled K with fixed rows/cols removed + Cholesky (:1198-1230).
    // DOF order is N*node + component; the band is what remains of the 27-point coupling.
    void factorize() {
        const idx ndof = numNodes * N;
        isFixedCache.assign(ndof, 0);
        for (size_t k = 0; k < dirNodes.size(); ++k)
            for (int c = 0; c < N; ++c) if (dirMask[k] >> c & 1) {
                if (dirVals[k * N + c] != 0) throw std::runtime_error("Nonzero Dirichlet constraints currently unsupported");
                isFixedCache[dirNodes[k] * N + c] = 1;
            }
        for (idx n = 0; n < numNodes; ++n) { idx nd[3]; ndNode(n, nd); if (nodeDetached(nd)) for (int c = 0; c < N; ++c) isFixedCache[n * N + c] = 1; }
        std::vector<idx> red(ndof, -1); freeVars.clear();
        for (idx i = 0; i < ndof; ++i) if (!isFixedCache[i]) { red[i] = idx(freeVars.size()); freeVars.push_back(i); }
        // bandwidth in full DOF numbering (reduced numbering can only shrink distances)
        idx maxNodeOff = 0; for (int d = 0; d < N; ++d) maxNodeOff += ninc[d];
        idx bw = maxNodeOff * N + (N - 1);
        idx nfree = idx(freeVars.size());
        bw = std::min(bw, std::max<idx>(nfree - 1, 0));
        chol.init(nfree, bw);
        std::vector<double> Ke(size_t(ke) * ke);
        for (idx e = 0; e < numElems; ++e) {
            idx nd[3]; ndElem(e, nd); idx off = firstNodeOfElem(nd);
            elementStiffness(e, Ke.data());
            for (int i = 0; i < npe; ++i) for (int ci = 0; ci < N; ++ci) {
                idx gi = red[(refNodes[i] + off) * N + ci]; if (gi < 0) continue;
                for (int j = 0; j < npe; ++j) for (int cj = 0; cj < N; ++cj) {
                    idx gj = red[(refNodes[j] + off) * N + cj]; if (gj < 0 || gj > gi) continue;
                    chol.at(gi, gj) += Ke[size_t(N * i + ci) * ke + (N * j + cj)];
                }
            }
        }
        chol.factor();
        factorOK = true;
    }
    // TPS::solve (:1198-1230): fixed DOFs get zero, rhs entries of fixed DOFs dropped.
    void solve(const double *f, double *x) {
        if (!factorOK) factorize();
        std::vector<double> rhs(freeVars.size());
        for (size_t i = 0; i < freeVars.size(); ++i) { idx dof = freeVars[i]; rhs[i] = f[(dof % N) * numNodes + dof / N]; }
        chol.solve(rhs);
        std::fill(x, x + size_t(numNodes) * N, 0.0);
        for (size_t i = 0; i < freeVars.size(); ++i) { idx dof = freeVars[i]; x[(dof % N) * numNodes + dof / N] = rhs[i]; }
    }
};

// ---------------------------------------------------------------------------
// Multigrid solver (MultigridSolver.hh)
// ---------------------------------------------------------------------------
struct MG {
    int N;
    std::vector<std::shared_ptr<Sim>> sims;
    std::vector<std::vector<double>> x, b, r;
    std::vector<double> Ad, dvec;
    std::vector<double> cK0[8];      // coarsenedFineK0s (MultigridSolver.hh:116-120)
    double phi[8][8][8];             // phi[fi][fine_n][coarse_n] (:664-687)
    bool symmetricGS = true;
    size_t cachedStiffnessLayer = LAYER_MASK_NONE; bool bandedUpdatesEnabled = true; // :825-844
    std::vector<double> lastResiduals; int lastIters = 0;

    MG(std::shared_ptr<Sim> fine, int numCoarseningLevels) {
        // MultigridSolver.hh:35-121
        N = fine->N;
        idx ne[3] = {fine->ne[0], fine->ne[1], fine->ne[2]};
        for (int l = 0; l <= numCoarseningLevels; ++l) {
            std::shared_ptr<Sim> tps = fine;
            if (l > 0) {
                for (int d = 0; d < N; ++d) {
                    if (ne[d] % 2 == 1) throw std::runtime_error("Grid size currently must be divisible by 2^numCoarseningLevels (nonuniform coarsening not yet implemented)");
                    ne[d] /= 2;
                }
                tps = std::make_shared<Sim>(N, ne, fine->dmin, fine->dmax);
                std::memcpy(tps->D, fine->D, sizeof(fine->D)); tps->updateK0();
                const Sim &finer = *sims.back(); Sim &coarser = *tps;
                Sim::BCBuilder cbc(coarser);
                const double zero[3] = {0, 0, 0};
                for (size_t k = 0; k < finer.dirNodes.size(); ++k) {
                    idx fnd[3]; finer.ndNode(finer.dirNodes[k], fnd);
                    double p[3]; finer.nodePosition(fnd, p);
                    // getElementAndReferenceCoordinates (TensorProductSimulator.hh:1731-1767)
                    idx ec[3]; double ref[3];
                    for (int d = 0; d < N; ++d) {
                        double fi = (p[d] - coarser.dmin[d]) / coarser.stretch[d];
                        if (std::abs(fi - double(coarser.ne[d])) < 1e-10) ec[d] = coarser.ne[d] - 1;
                        else ec[d] = idx(fi);
                        if (ec[d] >= coarser.ne[d]) throw std::runtime_error("Point out of bounds");
                        double first = coarser.dmin[d] + double(ec[d]) * coarser.spacing[d];
                        ref[d] = (p[d] - first) / coarser.stretch[d];
                    }
                    int onB[3];
                    for (int d = 0; d < N; ++d) onB[d] = std::abs(ref[d]) < 1e-9 ? 0 : (std::abs(ref[d] - 1.0) < 1e-9 ? 1 : -1);
                    for (int ln = 0; ln < coarser.npe; ++ln) {
                        bool match = true; idx cn[3];
                        for (int d = 0; d < N; ++d) {
                            int bit = (ln >> (N - 1 - d)) & 1;
                            if (onB[d] != -1 && onB[d] != bit) match = false;
                            cn[d] = ec[d] + bit;
                        }
                        if (!match) continue;
                        cbc.setDirichlet(coarser.flatNode(cn), zero, finer.dirMask[k]);
                    }
                }
                cbc.apply();
            }
            x.emplace_back(size_t(tps->numNodes) * N, 0.0);
            b.emplace_back(size_t(tps->numNodes) * N, 0.0);
            r.emplace_back(size_t(tps->numNodes) * N, 0.0);
            sims.push_back(tps);
        }
        buildPhis();
        const Sim &s0 = *sims[0];
        for (int fi = 0; fi < (1 << N); ++fi) {
            cK0[fi].assign(size_t(s0.ke) * s0.ke, 0.0);
            accumulateCoarsened(fi, s0.K0.data(), cK0[fi].data());
        }
    }

    int numLevels() const { return int(sims.size()); }

    // getCompressedElementInterpolationOperator (:664-687)
    void buildPhis() {
        const int npe = 1 << N;
        for (int fi = 0; fi < npe; ++fi) for (int fn = 0; fn < npe; ++fn) for (int cn = 0; cn < npe; ++cn) {
            double v = 1;
            for (int d = 0; d < N; ++d) {
                double pos = 0.5 * ((fn >> (N - 1 - d)) & 1) + 0.5 * ((fi >> (N - 1 - d)) & 1);
                v *= ((cn >> (N - 1 - d)) & 1) ? pos : (1.0 - pos);
            }
            phi[fi][fn][cn] = v;
        }
    }
    // accumulateCoarsenedStiffnessMatrix (:711-722): Ke_c += I^T Ke_f I
    void accumulateCoarsened(int fi, const double *Kf, double *Kc) const {
        const int npe = 1 << N, ke = N * npe;
        double T[24 * 24]; // T = Kf * I  (ke x ke)
        for (int a = 0; a < ke; ++a) for (int j = 0; j < npe; ++j) for (int dcomp = 0; dcomp < N; ++dcomp) {
            double s = 0;
            for (int i = 0; i < npe; ++i) s += Kf[size_t(a) * ke + (N * i + dcomp)] * phi[fi][i][j];
            T[a * ke + (N * j + dcomp)] = s;
        }
        for (int j = 0; j < npe; ++j) for (int c = 0; c < N; ++c) for (int bcol = 0; bcol < ke; ++bcol) {
            double s = 0;
            for (int i = 0; i < npe; ++i) s += phi[fi][i][j] * T[(N * i + c) * ke + bcol];
            Kc[size_t(N * j + c) * ke + bcol] += s;
        }
    }
    // visitFineElementsInside (:694-702)
    idx fineElemInside(const Sim &finer, const idx *ec, int fi) const {
        idx ef[3];
        for (int d = 0; d < N; ++d) ef[d] = 2 * ec[d] + ((fi >> (N - 1 - d)) & 1);
        return finer.flatElem(ef);
    }
    // m_firstLevelCoarsenedStiffnessMatrix (:724-732)
    void firstLevelKe(const idx *ec, double *Ke) const {
        const Sim &finest = *sims[0]; const int kk = finest.ke * finest.ke;
        for (int fi = 0; fi < (1 << N); ++fi) {
            double Ef = finest.E[fineElemInside(finest, ec, fi)];
            if (fi == 0) for (int i = 0; i < kk; ++i) Ke[i]  = Ef * cK0[fi][i];
            else         for (int i = 0; i < kk; ++i) Ke[i] += Ef * cK0[fi][i];
        }
    }
    static int stencilSlot(int N, const int *dlt) { int s = 0; for (int d = 0; d < N; ++d) s = s * 3 + (dlt[d] + 1); return s; }
    // accumulate a per-element Ke into the level's block stencil (:782-814); sign=-1 for banded subtraction (:990-1011)
    void accumToStencil(Sim &sim, const idx *ec, const double *Ke, double sign) const {
        const int npe = sim.npe, ke = sim.ke, ns = sim.nstencil(), NN = N * N;
        idx off = sim.firstNodeOfElem(ec);
        for (int i = 0; i < npe; ++i) for (int j = 0; j < npe; ++j) {
            int dlt[3];
            for (int d = 0; d < N; ++d) dlt[d] = ((j >> (N - 1 - d)) & 1) - ((i >> (N - 1 - d)) & 1);
            double *B = &sim.stencil[(size_t(sim.refNodes[i] + off) * ns + stencilSlot(N, dlt)) * NN];
            for (int a = 0; a < N; ++a) for (int c = 0; c < N; ++c) B[a * N + c] += sign * Ke[size_t(N * i + a) * ke + (N * j + c)];
        }
    }
    // m_getCoarsenedStiffnessMatrix (:744-819)
    void getCoarsened(int l, const idx *ec, double *result) {
        Sim &coarser = *sims[l]; const int kk = coarser.ke * coarser.ke;
        if (coarser.elemMasked(ec)) std::fill(result, result + kk, 0.0);
        else if (l == 1) firstLevelKe(ec, result);
        else {
            std::fill(result, result + kk, 0.0);
            std::vector<double> child(kk);
            for (int fi = 0; fi < (1 << N); ++fi) {
                idx ef[3]; for (int d = 0; d < N; ++d) ef[d] = 2 * ec[d] + ((fi >> (N - 1 - d)) & 1);
                getCoarsened(l - 1, ef, child.data());
                accumulateCoarsened(fi, child.data(), result);
            }
        }
        const int nl = numLevels(); const bool toBlock = l < nl - 1;
        if (l == 1 && toBlock) return;
        if (toBlock) accumToStencil(coarser, ec, result, 1.0);
        else std::memcpy(&coarser.KeCache[size_t(coarser.flatElem(ec)) * kk], result, sizeof(double) * kk);
    }
    template<class F> void visitElementsMulticolored(const Sim &sim, const F &f) const { // TensorProductSimulator.hh:1444-1457
        const int npe = sim.npe;
        for (int color = 0; color < npe; ++color) {
            idx cnt[3] = {1, 1, 1}, off[3] = {0, 0, 0}; bool any = true;
            for (int d = 0; d < N; ++d) { off[d] = (color >> (N - 1 - d)) & 1; if (off[d] >= sim.ne[d]) any = false; else cnt[d] = (sim.ne[d] - 1 - off[d]) / 2 + 1; }
            if (!any) continue;
            const idx tot = cnt[0] * cnt[1] * cnt[2];
            #pragma omp parallel for schedule(dynamic, 1)
            for (idx t = 0; t < tot; ++t) {
                idx e[3]; idx rr = t;
                for (int d = N - 1; d >= 0; --d) { e[d] = 2 * (rr % cnt[d]) + off[d]; rr /= cnt[d]; }
                f(e);
            }
        }
    }
    // updateStiffnessMatrices (:846-905)
    void updateStiffnessMatrices() {
        size_t currentLayer = size_t(sims[0]->firstMasked);
        const bool partial = bandedUpdatesEnabled && cachedStiffnessLayer != size_t(LAYER_MASK_NONE) && currentLayer < cachedStiffnessLayer;
        if (partial) { partialUpdate(); return; }
        const bool full = !bandedUpdatesEnabled || cachedStiffnessLayer == size_t(LAYER_MASK_NONE) || currentLayer == size_t(LAYER_MASK_NONE);
        if (!full) return; // "WARNING: entirely skipping stiffness matrix update" (:863-866)
        cachedStiffnessLayer = currentLayer;
        const int nl = numLevels();
        for (int l = 1; l < nl; ++l) {
            const bool useBlock = l < nl - 1;
            if (l == 1 && useBlock) continue;
            Sim &sim = *sims[l];
            if (useBlock) { sim.stencil.assign(size_t(sim.numNodes) * sim.nstencil() * N * N, 0.0); sim.hasStencil = true; sim.KeCache.clear(); sim.hasKeCache = false; }
            else { sim.KeCache.assign(size_t(sim.numElems) * sim.ke * sim.ke, 0.0); sim.hasKeCache = true; }
        }
        if (nl == 1) return;
        Sim &coarsest = *sims[nl - 1];
        visitElementsMulticolored(coarsest, [&](const idx *e) { double Ke[24 * 24]; getCoarsened(nl - 1, e, Ke); });
        coarsest.factorOK = false;
    }
    // m_partialStiffnessMatrixUpdate (:907-938) + m_computeAndSubtractCoarsenedStiffnessMatrixBand (:947-1017)
    void bandKe(int l, const idx *ec, const std::vector<std::pair<size_t, size_t>> &bands, double *result) {
        Sim &coarser = *sims[l]; const int kk = coarser.ke * coarser.ke;
        if (size_t(ec[BUILD_DIRECTION]) < bands[l].first || size_t(ec[BUILD_DIRECTION]) >= bands[l].second) { std::fill(result, result + kk, 0.0); return; }
        if (l == 1) {
            const Sim &finest = *sims[0];
            for (int fi = 0; fi < (1 << N); ++fi) {
                idx ef[3]; for (int d = 0; d < N; ++d) ef[d] = 2 * ec[d] + ((fi >> (N - 1 - d)) & 1);
                double sf = 0.0;
                if (size_t(ef[BUILD_DIRECTION]) >= bands[0].first && size_t(ef[BUILD_DIRECTION]) < bands[0].second) sf = finest.unmaskedE(finest.flatElem(ef));
                if (fi == 0) for (int i = 0; i < kk; ++i) result[i]  = sf * cK0[fi][i];
                else         for (int i = 0; i < kk; ++i) result[i] += sf * cK0[fi][i];
            }
        } else {
            std::fill(result, result + kk, 0.0);
            std::vector<double> child(kk);
            for (int fi = 0; fi < (1 << N); ++fi) {
                idx ef[3]; for (int d = 0; d < N; ++d) ef[d] = 2 * ec[d] + ((fi >> (N - 1 - d)) & 1);
                bandKe(l - 1, ef, bands, child.data());
                accumulateCoarsened(fi, child.data(), result);
            }
        }
        const int nl = numLevels(); const bool toBlock = l < nl - 1;
        if (l == 1 && toBlock) return;
        if (toBlock) accumToStencil(coarser, ec, result, -1.0);
        else { double *K = &coarser.KeCache[size_t(coarser.flatElem(ec)) * kk]; for (int i = 0; i < kk; ++i) K[i] -= result[i]; }
    }
    void partialUpdate() {
        size_t currentLayer = size_t(sims[0]->firstMasked);
        const int L = numLevels() - 1;
        std::vector<std::pair<size_t, size_t>> bands(L + 1);
        bands[0] = {currentLayer, cachedStiffnessLayer};
        size_t cb = currentLayer, ce = cachedStiffnessLayer;
        for (int i = 1; i <= L; ++i) { cb /= 2; ce = (ce + 1) / 2; bands[i] = {cb, ce}; }
        if (L >= 1) {
            Sim &coarsest = *sims[L];
            visitElementsMulticolored(coarsest, [&](const idx *e) { double Ke[24 * 24]; bandKe(L, e, bands, Ke); });
            coarsest.factorOK = false;
        }
        cachedStiffnessLayer = currentLayer;
    }
    // setFabricationMaskHeightByLayer / decrement (:1022-1036)
    void setMaskByLayer(idx l) {
        sims[0]->setMaskHeightByLayer(l); double h = sims[0]->maskHeight;
        for (size_t i = 1; i < sims.size(); ++i) sims[i]->setMaskHeight(h, false);
        cachedStiffnessLayer = LAYER_MASK_NONE;
    }
    void decrementMaskByLayer(int inc) {
        sims[0]->decrementMaskByLayer(inc); double h = sims[0]->maskHeight;
        for (size_t i = 1; i < sims.size(); ++i) sims[i]->setMaskHeight(h, false);
    }

    // ---- transfer operators ----
    // interpolation / accum_interpolation (:178-212): multilinear, weights {1, 1/2}^N
    void interpolate(int lf, const double *vc, double *vf, bool accumulate) const {
        const Sim &finer = *sims[lf], &coarser = *sims[lf + 1];
        idx lim[3] = {1, 1, 1};
        for (int d = 0; d < N; ++d) lim[d] = accumulate ? finer.nondetachedNodes(d) : finer.nn[d]; // :192 vs :211
        #pragma omp parallel for schedule(static)
        for (idx i0 = 0; i0 < lim[0]; ++i0) {
            idx F[3] = {i0, 0, 0};
            for (F[1] = 0; F[1] < lim[1]; ++F[1]) for (F[2] = 0; F[2] < (N == 3 ? lim[2] : 1); ++F[2]) {
                idx nf = finer.flatNode(F);
                double res[3] = {0, 0, 0};
                // visitInterpolationOperatorRow (:162-176)
                const int cnt = 1 << N;
                for (int cn = 0; cn < cnt; ++cn) {
                    double w = 1; idx C[3]; bool skip = false;
                    for (int d = 0; d < N; ++d) {
                        int bit = (cn >> (N - 1 - d)) & 1;
                        idx e_c = std::min<idx>(F[d] / 2, coarser.ne[d]);
                        idx lf_ = F[d] - 2 * e_c; // 0 or 1
                        double wd = (lf_ == 0) ? (bit ? 0.0 : 1.0) : 0.5;
                        if (wd == 0) { skip = true; break; }
                        w *= wd; C[d] = e_c + bit;
                    }
                    if (skip) continue;
                    idx nc = coarser.flatNode(C);
                    for (int c = 0; c < N; ++c) res[c] += w * vc[c * coarser.numNodes + nc];
                }
                for (int c = 0; c < N; ++c) { if (accumulate) vf[c * finer.numNodes + nf] += res[c]; else vf[c * finer.numNodes + nf] = res[c]; }
            }
        }
    }
    // restriction (:216-262) with TPSStencils::fineNodesInSupport weights (TPSStencils.hh:82-128)
    void restrict_(int lf, const double *vf, double *vc) const {
        const Sim &finer = *sims[lf], &coarser = *sims[lf + 1];
        idx cs[3] = {1, 1, 1}, fs[3] = {1, 1, 1};
        for (int d = 0; d < N; ++d) { cs[d] = coarser.nondetachedNodes(d); fs[d] = finer.nondetachedNodes(d); }
        #pragma omp parallel for schedule(static)
        for (idx i0 = 0; i0 < cs[0]; ++i0) {
            idx C[3] = {i0, 0, 0};
            for (C[1] = 0; C[1] < cs[1]; ++C[1]) for (C[2] = 0; C[2] < (N == 3 ? cs[2] : 1); ++C[2]) {
                double val[3] = {0, 0, 0};
                idx lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
                for (int d = 0; d < N; ++d) { lo[d] = std::max<idx>(2 * C[d] - 1, 0); hi[d] = std::min<idx>(2 * C[d] + 2, fs[d]); }
                idx F[3] = {0, 0, 0};
                for (F[0] = lo[0]; F[0] < hi[0]; ++F[0]) for (F[1] = lo[1]; F[1] < hi[1]; ++F[1]) for (F[2] = (N == 3 ? lo[2] : 0); F[2] < (N == 3 ? hi[2] : 1); ++F[2]) {
                    double w = 1; for (int d = 0; d < N; ++d) w *= 1.0 - 0.5 * double(std::llabs(F[d] - 2 * C[d]));
                    idx nf = finer.flatNode(F);
                    for (int c = 0; c < N; ++c) val[c] += w * vf[c * finer.numNodes + nf];
                }
                idx nc = coarser.flatNode(C);
                for (int c = 0; c < N; ++c) vc[c * coarser.numNodes + nc] = val[c];
            }
        }
        if (cs[BUILD_DIRECTION] != coarser.nn[BUILD_DIRECTION]) { // zero first detached layer (:251-262)
            for (idx n = 0; n < coarser.numNodes; ++n) { idx nd[3]; coarser.ndNode(n, nd); if (nd[BUILD_DIRECTION] == cs[BUILD_DIRECTION]) for (int c = 0; c < N; ++c) vc[c * coarser.numNodes + n] = 0; }
        }
    }

    // ---- smoother ----
    // visitIncidentElements (TPSStencils.hh:145-161, 411-427): incident element i of 2^N, offset -1 along d iff bit d of i is 0 (bit 0 <-> axis 0).
    // Calls f(e_nd, localIndex).
    template<class F> void visitIncident(const Sim &sim, const idx *g, const F &f) const {
        for (int i = 0; i < (1 << N); ++i) {
            idx e[3]; int ln = 0; bool ok = true;
            for (int d = 0; d < N; ++d) {
                bool minus = !((i >> d) & 1);
                e[d] = g[d] - (minus ? 1 : 0);
                if (e[d] < 0 || e[d] >= sim.ne[d]) ok = false;
                ln |= (minus ? 1 : 0) << (N - 1 - d);
            }
            if (ok) f(e, ln);
        }
    }
    static void inv3(const double *M, double *inv, int N) { // Eigen fixed-size inverse (cofactors), MultigridSolver.hh:370
        if (N == 2) {
            double det = M[0] * M[3] - M[1] * M[2];
            inv[0] = M[3] / det; inv[1] = -M[1] / det; inv[2] = -M[2] / det; inv[3] = M[0] / det; return;
        }
        double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
        double det = M[0] * c00 + M[1] * c01 + M[2] * c02, id = 1.0 / det;
        inv[0] = c00 * id; inv[1] = (M[2] * M[7] - M[1] * M[8]) * id; inv[2] = (M[1] * M[5] - M[2] * M[4]) * id;
        inv[3] = c01 * id; inv[4] = (M[0] * M[8] - M[2] * M[6]) * id; inv[5] = (M[2] * M[3] - M[0] * M[5]) * id;
        inv[6] = c02 * id; inv[7] = (M[1] * M[6] - M[0] * M[7]) * id; inv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    }
    // m_smoothNode (:347-378) with the three stencil builders (:277-334)
    void smoothNode(int l, const idx *g, double *u, const double *bb, bool forward) {
        Sim &sim = *sims[l];
        const idx n = sim.flatNode(g), nn_ = sim.numNodes;
        if (sim.hasFullDirichlet(n)) return;
        double rhs[3] = {0, 0, 0}, M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < N; ++c) rhs[c] = bb[c * nn_ + n];
        const int ke = sim.ke, npe = sim.npe;
        if (l == 0 && !sim.hasKeCache) {               // NodeSmoothStencilFinest (:277-292)
            visitIncident(sim, g, [&](const idx *e, int ln) {
                double Ee = sim.E[sim.flatElem(e)]; idx off = sim.firstNodeOfElem(e);
                double ul[24];
                for (int m = 0; m < npe; ++m) for (int c = 0; c < N; ++c) ul[N * m + c] = u[c * nn_ + sim.refNodes[m] + off];
                for (int c = 0; c < N; ++c) { double s = 0; for (int a = 0; a < ke; ++a) s += sim.k0(a, N * ln + c) * ul[a]; rhs[c] -= Ee * s; }
                for (int a = 0; a < N; ++a) for (int c = 0; c < N; ++c) M[a * N + c] += Ee * sim.k0(N * ln + a, N * ln + c);
            });
        } else if (l == 1 && numLevels() > 2) {        // NodeSmoothStencilSecondFinest (:294-321)
            const Sim &finest = *sims[0];
            visitIncident(sim, g, [&](const idx *e, int ln) {
                idx off = sim.firstNodeOfElem(e);
                double ul[24], cols[24 * 3];
                for (int m = 0; m < npe; ++m) for (int c = 0; c < N; ++c) ul[N * m + c] = u[c * nn_ + sim.refNodes[m] + off];
                for (int fi = 0; fi < (1 << N); ++fi) {
                    double Ef = finest.E[fineElemInside(finest, e, fi)];
                    const double *K = cK0[fi].data();
                    if (fi == 0) for (int a = 0; a < ke; ++a) for (int c = 0; c < N; ++c) cols[a * N + c]  = Ef * K[size_t(a) * ke + N * ln + c];
                    else         for (int a = 0; a < ke; ++a) for (int c = 0; c < N; ++c) cols[a * N + c] += Ef * K[size_t(a) * ke + N * ln + c];
                }
                for (int c = 0; c < N; ++c) { double s = 0; for (int a = 0; a < ke; ++a) s += cols[a * N + c] * ul[a]; rhs[c] -= s; }
                for (int a = 0; a < N; ++a) for (int c = 0; c < N; ++c) M[a * N + c] += cols[(N * ln + a) * N + c];
            });
        } else if (sim.hasStencil) {                   // NodeSmoothStencilBlockK (:323-334)
            const int ns = sim.nstencil(), NN = N * N;
            for (int s = 0; s < ns; ++s) {
                int rr = s; idx m = 0; bool ok = true; idx dl[3];
                for (int d = N - 1; d >= 0; --d) { dl[d] = (rr % 3) - 1; rr /= 3; }
                for (int d = 0; d < N; ++d) { idx qd = g[d] + dl[d]; if (qd < 0 || qd >= sim.nn[d]) ok = false; m = m * sim.nn[d] + qd; }
                if (!ok) continue;
                const double *B = &sim.stencil[(size_t(n) * ns + s) * NN];
                for (int a = 0; a < N; ++a) for (int c = 0; c < N; ++c) rhs[a] -= B[a * N + c] * u[c * nn_ + m];
                if (m == n) for (int i = 0; i < NN; ++i) M[i] = B[i];
            }
        } else {
            // Coarsest-level (or single-level hierarchies) never get smoothed in the reference; cached-Ke smoothing is unused.
            throw std::logic_error("smoothNode: no operator representation at this level");
        }
        if (sim.hasDirichlet(n)) { // partial Dirichlet: point GS on the free components (:358-365)
            double ud[3] = {0, 0, 0}; uint8_t dc = sim.nodeDirMask[n];
            auto upd = [&](int i) { double s = rhs[i]; for (int j = 0; j < N; ++j) s -= M[i * N + j] * ud[j]; ud[i] = s * (double(!((dc >> i) & 1)) / M[i * N + i]); };
            if (forward) for (int i = 0; i < N; ++i) upd(i); else for (int i = N - 1; i >= 0; --i) upd(i);
            for (int c = 0; c < N; ++c) u[c * nn_ + n] += ud[c];
        } else {
            double inv[9]; inv3(M, inv, N);
            for (int a = 0; a < N; ++a) { double s = 0; for (int c = 0; c < N; ++c) s += inv[a * N + c] * rhs[c]; u[a * nn_ + n] += s; }
        }
    }
    // visitNodesMulticolored (:408-442) + smoothingMulticoloredGS (:452-458)
    template<class F> void visitNodesMulticolored(int l, const F &f, bool forward, bool parallel = true, bool skipDetached = true) const {
        const Sim &sim = *sims[l]; const int nc = 1 << N;
        for (int i = 0; i < nc; ++i) {
            int lni = forward ? i : (nc - i - 1);
            idx cnt[3] = {1, 1, 1}, off[3] = {0, 0, 0}; bool any = true;
            for (int d = 0; d < N; ++d) {
                off[d] = (lni >> (N - 1 - d)) & 1;
                idx lim = skipDetached ? sim.nondetachedNodes(d) : sim.nn[d];
                if (lim - 1 - off[d] < 0) { any = false; break; }
                cnt[d] = (lim - 1 - off[d]) / 2 + 1;
            }
            if (!any) continue;
            const idx tot = cnt[0] * cnt[1] * cnt[2];
            if (parallel) {
                #pragma omp parallel for schedule(static)
                for (idx t = 0; t < tot; ++t) { idx g[3] = {0, 0, 0}; idx rr = t; for (int d = N - 1; d >= 0; --d) { g[d] = off[d] + 2 * (rr % cnt[d]); rr /= cnt[d]; } f(g); }
            } else {
                for (idx t = 0; t < tot; ++t) { idx g[3] = {0, 0, 0}; idx rr = t; for (int d = N - 1; d >= 0; --d) { g[d] = off[d] + 2 * (rr % cnt[d]); rr /= cnt[d]; } f(g); }
            }
        }
    }
    void smoothMulticolored(int l, double *u, const double *bb, bool forward) {
        visitNodesMulticolored(l, [&](const idx *g) { smoothNode(l, g, u, bb, forward); }, forward);
    }

    // ---- per-level operator apply (:471-504) ----
    void applyK(int l, const double *u, double *out, bool zeroInit = true, bool negate = false) {
        Sim &sim = *sims[l];
        if (l == 0) { if (sim.hasKeCache) sim.applyKeCache(u, out, zeroInit, negate); else sim.applyK0(u, out, zeroInit, negate); return; }
        if (l == 1 && !(sim.hasKeCache || sim.hasStencil)) {
            // on-the-fly first-level Galerkin operator (:475-499); note: l == 1 < coarsest stores nothing
            if (numLevels() > 2) {
                if (zeroInit) sim.maskedZero(out);
                for (int color = 0; color < sim.npe; ++color) {
                    idx cnt[3] = {1, 1, 1}, off[3] = {0, 0, 0}; bool any = true;
                    for (int d = 0; d < N; ++d) { off[d] = (color >> (N - 1 - d)) & 1; idx lim = sim.nonmaskedElems(d); if (off[d] >= lim) any = false; else cnt[d] = (lim - 1 - off[d]) / 2 + 1; }
                    if (!any) continue;
                    const idx tot = cnt[0] * cnt[1] * cnt[2];
                    #pragma omp parallel for schedule(static)
                    for (idx t = 0; t < tot; ++t) {
                        idx e[3] = {0, 0, 0}; idx rr = t; for (int d = N - 1; d >= 0; --d) { e[d] = 2 * (rr % cnt[d]) + off[d]; rr /= cnt[d]; }
                        double Ke[24 * 24], ul[24]; firstLevelKe(e, Ke);
                        idx noff = sim.firstNodeOfElem(e);
                        for (int m = 0; m < sim.npe; ++m) for (int c = 0; c < N; ++c) ul[N * m + c] = u[c * sim.numNodes + sim.refNodes[m] + noff];
                        for (int a = 0; a < sim.ke; ++a) {
                            double s = 0; for (int bq = 0; bq < sim.ke; ++bq) s += Ke[a * sim.ke + bq] * ul[bq];
                            double &o = out[(a % N) * sim.numNodes + sim.refNodes[a / N] + noff];
                            if (negate) o -= s; else o += s;
                        }
                    }
                }
                return;
            }
        }
        if (!sim.hasKeCache && !sim.hasStencil) updateStiffnessMatrices();
        if (sim.hasStencil) sim.applyStencil(u, out, zeroInit, negate);
        else sim.applyKeCache(u, out, zeroInit, negate);
    }
    // computeResidual (:527-541)
    void computeResidual(int l, const double *u, const double *bb, double *res) {
        sims[l]->maskedCopy(bb, res, SIMD_WIDTH);
        applyK(l, u, res, false, true);
        sims[l]->zeroOutDirichlet(res);
    }
    // vcycle (:617-658)
    void vcycle(int l, int nsmooth, bool residualSystem) {
        const int coarsest = numLevels() - 1;
        if (l == coarsest) { sims[l]->solve(b[l].data(), x[l].data()); return; }
        if (residualSystem) sims[l]->zeroOutDirichlet(x[l].data()); else sims[l]->enforceDirichlet(x[l].data());
        for (int i = 0; i < nsmooth; ++i) smoothMulticolored(l, x[l].data(), b[l].data(), true);
        computeResidual(l, x[l].data(), b[l].data(), r[l].data());
        restrict_(l, r[l].data(), b[l + 1].data());
        sims[l + 1]->maskedZero(x[l + 1].data(), SIMD_WIDTH);
        vcycle(l + 1, nsmooth, true);
        interpolate(l, x[l + 1].data(), x[l].data(), true);
        for (int i = 0; i < nsmooth; ++i) smoothMulticolored(l, x[l].data(), b[l].data(), !symmetricGS);
    }
    // fullMultigrid (:587-609)
    void fullMultigrid(int l, int nsmooth, bool residualSystem) {
        const int coarsest = numLevels() - 1;
        if (l == coarsest) { sims[l]->solve(b[l].data(), x[l].data()); return; }
        restrict_(l, b[l].data(), b[l + 1].data());
        fullMultigrid(l + 1, nsmooth, residualSystem);
        interpolate(l, x[l + 1].data(), x[l].data(), false);
        vcycle(l, nsmooth, residualSystem);
    }
    // solve (:546-573)
    const std::vector<double> &solve(const double *u, const double *f, int numSteps, int nsmooth, bool stiffnessUpdated, bool zeroDirichlet, bool fmg,
                                     const std::function<void(int)> &cb = nullptr) {
        if (!stiffnessUpdated) updateStiffnessMatrices();
        if (u != x[0].data()) std::memcpy(x[0].data(), u, x[0].size() * sizeof(double));
        if (numSteps == 0) return x[0];
        if (f != b[0].data()) std::memcpy(b[0].data(), f, b[0].size() * sizeof(double));
        int start = 0;
        if (fmg) { fullMultigrid(0, nsmooth, zeroDirichlet); if (cb) cb(0); start = 1; }
        for (int i = start; i < numSteps; ++i) { vcycle(0, nsmooth, zeroDirichlet); if (cb) cb(i); }
        return x[0];
    }
    // preconditionedConjugateGradient (:1047-1152)
    // stiffnessPrebuilt: test/bench infrastructure only -- skip the per-call hierarchy rebuild of the reference (:1104-1107) when the
    // caller knows the operators are current (bench.py times the rebuild and the iterations separately).
    bool stiffnessPrebuilt = false;
    void pcg(double *xx, const double *bb, int maxIter, double tol, int mgIterations, int mgSmoothing, bool fmg, bool dirichletOK,
             const std::function<void(int, const double *, const double *)> &cb = nullptr) {
        Sim &fine = *sims[0];
        const size_t len = size_t(fine.numNodes) * N;
        lastResiduals.clear(); lastIters = 0;
        if (sims.size() == 1) { // :1057-1065
            double *rr = b[0].data();
            computeResidual(0, xx, bb, rr);
            double rs = 0, bs = 0; for (size_t i = 0; i < len; ++i) { rs += rr[i] * rr[i]; bs += bb[i] * bb[i]; }
            if (rs < tol * tol * bs) return;
            fine.solve(bb, xx);
            computeResidual(0, xx, bb, rr);
            lastIters = 1; if (cb) cb(1, xx, rr);
            return;
        }
        double *rr = b[0].data(); double *s = x[0].data();
        if (Ad.size() != len) Ad.assign(len, 0.0);
        if (dvec.size() != len) dvec.assign(len, 0.0);
        double bNormSq, rSq, rMr = 0;
        bool stiffUpdated = stiffnessPrebuilt;
        if (!dirichletOK) fine.enforceDirichlet(xx);
        bNormSq = fine.maskedDot(bb, bb);
        computeResidual(0, xx, bb, rr);
        rSq = fine.maskedDot(rr, rr);
        if (std::isnan(rSq)) throw std::logic_error("NaN encountered");
        int i = 0; double *d = nullptr;
        while ((i++ < maxIter) && (rSq > tol * tol * bNormSq)) {
            if (mgIterations > 0) {
                if (!stiffUpdated) { updateStiffnessMatrices(); stiffUpdated = true; }
                if (d == s) { d = dvec.data(); fine.maskedCopy(s, d, SIMD_WIDTH); }
                if (mgSmoothing > 0) { // applyPreconditionerInv (:577-580): zero initial guess, r is the RHS (already in b[0])
                    std::fill(x[0].begin(), x[0].end(), 0.0);
                    solve(x[0].data(), rr, mgIterations, mgSmoothing, true, true, fmg);
                } else std::memcpy(s, rr, len * sizeof(double)); // returns r itself
            } else std::memcpy(s, rr, len * sizeof(double));
            fine.zeroOutDirichlet(s);
            double rMrOld = rMr;
            rMr = fine.maskedDot(rr, s);
            if (d) { // scaleAndAddInPlace (ParallelVectorOps.hh:76-84): d = beta * d + s
                double beta = rMr / rMrOld;
                #pragma omp parallel for schedule(static)
                for (idx k = 0; k < idx(len); ++k) d[k] = beta * d[k] + s[k];
            } else d = s;
            applyK(0, d, Ad.data(), true, false);
            fine.zeroOutDirichlet(Ad.data());
            double alpha = rMr / fine.maskedDot(d, Ad.data());
            const double *Adp = Ad.data();
            fine.maskedVisit([&](idx start, idx cnt) { for (int c = 0; c < N; ++c) { double *xc = xx + c * fine.numNodes + start; const double *dc = d + c * fine.numNodes + start; for (idx k = 0; k < cnt; ++k) xc[k] += alpha * dc[k]; } });
            fine.maskedVisit([&](idx start, idx cnt) { for (int c = 0; c < N; ++c) { double *rc = rr + c * fine.numNodes + start; const double *ac = Adp + c * fine.numNodes + start; for (idx k = 0; k < cnt; ++k) rc[k] -= alpha * ac[k]; } });
            rSq = fine.maskedDot(rr, rr);
            if (std::isnan(rSq)) throw std::logic_error("NaN encountered at iteration" + std::to_string(i));
            lastIters = i; lastResiduals.push_back(std::sqrt(rSq));
            if (cb) cb(i, xx, rr);
        }
    }
};

// ---------------------------------------------------------------------------
// Filters / constraint / OC (TopologyOptimizationFilter.hh, ...Constraint.hh, OptimalityCriterion.hh)
// ---------------------------------------------------------------------------
// SmoothingFilter::ApplyImpl (TopologyOptimizationFilter.hh:328-395)
static void smoothingFilter(int N, const idx *sizes, int radius, int type, const double *in, double *out) {
    const double rp1 = 1.0 + radius;
    auto reflect = [](idx i, idx s) { while (i < 0 || i >= s) { if (i >= s) i = 2 * s - i - 1; if (i < 0) i = -i - 1; } return i; };
    idx tot = 1; for (int d = 0; d < N; ++d) tot *= sizes[d];
    const int w = 2 * radius + 1; int nOff = 1; for (int d = 0; d < N; ++d) nOff *= w;
    std::vector<double> wts(nOff);
    for (int o = 0; o < nOff; ++o) { int rr = o; double sq = 0; for (int d = N - 1; d >= 0; --d) { int q = rr % w - radius; rr /= w; sq += double(q) * q; } wts[o] = type == 1 ? (rp1 - std::sqrt(sq)) : 1.0; }
    #pragma omp parallel for schedule(static)
    for (idx e = 0; e < tot; ++e) {
        idx c[3]; idx rr = e; for (int d = N - 1; d >= 0; --d) { c[d] = rr % sizes[d]; rr /= sizes[d]; }
        double acc = 0, tw = 0;
        for (int o = 0; o < nOff; ++o) {
            double wt = wts[o]; if (wt <= 0) continue;
            int r2 = o; idx k = 0; idx nb[3];
            for (int d = N - 1; d >= 0; --d) { nb[d] = reflect(c[d] + (r2 % w - radius), sizes[d]); r2 /= w; }
            for (int d = 0; d < N; ++d) k = k * sizes[d] + nb[d];
            acc += wt * in[k]; tw += wt;
        }
        out[e] = acc / tw;
    }
}
// ProjectionFilter (TopologyOptimizationFilter.hh:199-232)
static void projectionApply(idx n, double beta, const double *in, double *out) {
    double th = std::tanh(0.5 * beta);
    #pragma omp parallel for schedule(static)
    for (idx i = 0; i < n; ++i) out[i] = (th + std::tanh(beta * (in[i] - 0.5))) / (2 * th);
}
static void projectionBackprop(idx n, double beta, const double *in, const double *vars, double *out) {
    double scale = 1.0 / (2 * std::tanh(0.5 * beta) / beta);
    #pragma omp parallel for schedule(static)
    for (idx i = 0; i < n; ++i) { double t = std::tanh(beta * (vars[i] - 0.5)); out[i] = in[i] * (1.0 - t * t) * scale; }
}

// ---------------------------------------------------------------------------------------------------------------------------
// UpsampleFilter / VertexToCellFilter / LangelaarFilter (TopologyOptimizationFilter.hh:418-712), restated loop for loop.
// Flat row-major arrays, last axis fastest (NDVector.hh:256-264).
// ---------------------------------------------------------------------------------------------------------------------------
namespace filt {
static idx flat(int N, const idx *sz, const idx *c) { idx r = 0; for (int d = 0; d < N; ++d) r = r * sz[d] + c[d]; return r; }
static void unflat(int N, const idx *sz, idx i, idx *c) { for (int d = N - 1; d >= 0; --d) { c[d] = i % sz[d]; i /= sz[d]; } }
static idx total(int N, const idx *sz) { idx n = 1; for (int d = 0; d < N; ++d) n *= sz[d]; return n; }

// UpsampleFilter::ApplyImpl (:463-489): every coarse entry that is the min corner of a cell samples that cell's multilinear
// interpolant at the (factor + 1)^N fine nodes of the cell (faces shared with the next cell are written twice with the same value)
static void upsample(int N, const idx *cs, int factor, const double *in, double *out) {
    idx fs[3]; for (int d = 0; d < N; ++d) fs[d] = (cs[d] - 1) * factor + 1;
    const idx nc = total(N, cs);
    for (idx i = 0; i < nc; ++i) {
        idx mc[3]; unflat(N, cs, i, mc);
        bool valid = true; for (int d = 0; d < N; ++d) if (mc[d] + 1 >= cs[d]) valid = false;
        if (!valid) continue;
        double coeff[8];
        for (int b = 0; b < (1 << N); ++b) { idx q[3]; for (int d = 0; d < N; ++d) q[d] = mc[d] + ((b >> (N - 1 - d)) & 1); coeff[b] = in[flat(N, cs, q)]; }
        idx li[3] = {0, 0, 0};
        const idx per = factor + 1; idx cnt = 1; for (int d = 0; d < N; ++d) cnt *= per;
        for (idx k = 0; k < cnt; ++k) {
            idx r = k; for (int d = N - 1; d >= 0; --d) { li[d] = r % per; r /= per; }
            double v = 0.0;
            for (int b = 0; b < (1 << N); ++b) {
                double w = 1.0;
                for (int d = 0; d < N; ++d) { const double t = double(li[d]) / factor; w *= ((b >> (N - 1 - d)) & 1) ? t : 1.0 - t; }
                v += coeff[b] * w;
            }
            idx q[3]; for (int d = 0; d < N; ++d) q[d] = mc[d] * factor + li[d];
            out[flat(N, fs, q)] = v;
        }
    }
}
// UpsampleFilter::BackpropImpl (:497-523)
static void upsampleBackprop(int N, const idx *cs, int factor, const double *dout, double *din) {
    idx fs[3]; for (int d = 0; d < N; ++d) fs[d] = (cs[d] - 1) * factor + 1;
    const idx nc = total(N, cs);
    for (idx i = 0; i < nc; ++i) {
        idx n_c[3]; unflat(N, cs, i, n_c);
        idx n_f[3], lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
        for (int d = 0; d < N; ++d) {
            n_f[d] = factor * n_c[d]; lo[d] = n_f[d]; hi[d] = n_f[d] + 1;
            if (n_c[d] > 0) lo[d] -= (factor - 1);
            if (n_c[d] < cs[d] - 1) hi[d] += (factor - 1);
        }
        double acc = 0.0;
        idx q[3];
        for (q[0] = lo[0]; q[0] < hi[0]; ++q[0]) for (q[1] = lo[1]; q[1] < hi[1]; ++q[1]) for (q[2] = (N == 3 ? lo[2] : 0); q[2] < (N == 3 ? hi[2] : 1); ++q[2]) {
            double phi = 1.0;
            for (int d = 0; d < N; ++d) phi *= 1.0 - double(std::max(q[d], n_f[d]) - std::min(q[d], n_f[d])) / factor;
            acc += phi * dout[flat(N, fs, q)];
        }
        din[i] = acc;
    }
}
// VertexToCellFilter (:543-584)
static void v2c(int N, const idx *vs, const double *in, double *out) {
    idx es[3]; for (int d = 0; d < N; ++d) es[d] = vs[d] - 1;
    const double weight = std::pow(2.0, -N); const idx ne = total(N, es);
    for (idx e = 0; e < ne; ++e) {
        idx mc[3]; unflat(N, es, e, mc);
        double a = 0.0;
        for (int b = 0; b < (1 << N); ++b) { idx q[3]; for (int d = 0; d < N; ++d) q[d] = mc[d] + ((b >> (N - 1 - d)) & 1); a += in[flat(N, vs, q)]; }
        out[e] = a * weight;
    }
}
static void v2cBackprop(int N, const idx *vs, const double *dout, double *din) {
    idx es[3]; for (int d = 0; d < N; ++d) es[d] = vs[d] - 1;
    const double weight = std::pow(2.0, -N); const idx nv = total(N, vs);
    for (idx v = 0; v < nv; ++v) {
        idx c[3]; unflat(N, vs, v, c);
        double a = 0.0;
        for (int b = 0; b < (1 << N); ++b) {
            bool ok = true; idx q[3];
            for (int d = 0; d < N; ++d) { const idx e = c[d] - 1 + ((b >> (N - 1 - d)) & 1); if (e < 0 || e >= es[d]) ok = false; q[d] = e; }
            if (ok) a += dout[flat(N, es, q)];
        }
        din[v] = a * weight;
    }
}
// NDVector::visitLayer (NDVector.hh:187-209) and visitSupportingRegion (:211-229) -- note the loop over the FIRST N - 1 axes
template<class F> static void visitLayer(int N, const idx *sz, idx layer, F &&cb) {
    if (N == 2) { for (idx i = 0; i < sz[0]; ++i) { idx c[2] = {i, layer}; cb(flat(2, sz, c)); } }
    else { for (idx i = 0; i < sz[0]; ++i) for (idx j = 0; j < sz[2]; ++j) { idx c[3] = {i, layer, j}; cb(flat(3, sz, c)); } }
}
template<class F> static void visitSupport(int N, const idx *sz, const idx *voxel, F &&cb) {
    idx center[3] = {0, 0, 0}; for (int d = 0; d < N; ++d) center[d] = voxel[d];
    center[1] -= 1;
    cb(flat(N, sz, center));
    for (int d = 0; d < N - 1; ++d) {
        idx cur[3] = {center[0], center[1], center[2]};
        cur[d] -= 1;
        bool in = true; for (int a = 0; a < N; ++a) if (cur[a] < 0 || cur[a] >= sz[a]) in = false;
        if (in) cb(flat(N, sz, cur));
        cur[d] += 2;
        in = true; for (int a = 0; a < N; ++a) if (cur[a] < 0 || cur[a] >= sz[a]) in = false;
        if (in) cb(flat(N, sz, cur));
    }
}
static const double LG_P = 40, LG_Q = 40 - 1.58, LG_EPS = 1e-4;   // (:703-711)
static double lgSmin(double x1, double x2) { return 0.5 * (x1 + x2 - std::pow((x1 - x2) * (x1 - x2) + LG_EPS, 0.5) + std::pow(LG_EPS, 0.5)); }
static double lgDsminDx1(double x1, double x2) { return 0.5 * (1 - (x1 - x2) * std::pow((x1 - x2) * (x1 - x2) + LG_EPS, -0.5)); }
static double lgDsminDx2(double x1, double x2) { return 0.5 * (1 + (x1 - x2) * std::pow((x1 - x2) * (x1 - x2) + LG_EPS, -0.5)); }
// LangelaarFilter::apply (:609-623): out is in/out (see visitSupport), smaxCache = m_cachedSmax
static void langelaar(int N, const idx *sz, const double *in, double *out, double *smaxCache) {
    visitLayer(N, sz, 0, [&](idx i) { out[i] = in[i]; });
    for (idx layer = 1; layer < sz[1]; ++layer)
        visitLayer(N, sz, layer, [&](idx i) {
            idx c[3]; unflat(N, sz, i, c);
            double sum = 0; visitSupport(N, sz, c, [&](idx k) { sum += std::pow(out[k], LG_P); });
            smaxCache[i] = std::pow(sum, 1 / LG_Q);
            out[i] = lgSmin(in[i], smaxCache[i]);
        });
}
// LangelaarFilter::backprop (:625-634) with computeLagrangeMultipliers (:643-661)
static void langelaarBackprop(int N, const idx *sz, const double *in, const double *vars, const double *filtered, const double *smaxCache, double *out) {
    const idx n = total(N, sz);
    std::vector<double> lambdas(n, 0.0);
    auto smaxDerivative = [&](const idx *indices, idx der) {
        double sum = 0; visitSupport(N, sz, indices, [&](idx i) { sum += std::pow(filtered[i], LG_P); });
        return LG_P * std::pow(filtered[der], LG_P - 1) / LG_Q * std::pow(sum, 1 / LG_Q - 1);
    };
    for (idx layer = sz[1] - 1; layer >= 0; --layer) {
        visitLayer(N, sz, layer, [&](idx i) { lambdas[i] = in[i]; });
        if (layer < sz[1] - 1)
            visitLayer(N, sz, layer + 1, [&](idx i) {
                idx c[3]; unflat(N, sz, i, c);
                visitSupport(N, sz, c, [&](idx k) { lambdas[k] += lambdas[i] * (lgDsminDx2(vars[i], smaxCache[i]) * smaxDerivative(c, k)); });
            });
    }
    visitLayer(N, sz, 0, [&](idx i) { out[i] = lambdas[i]; });
    for (idx layer = 1; layer < sz[1]; ++layer) visitLayer(N, sz, layer, [&](idx i) { out[i] = lambdas[i] * lgDsminDx1(vars[i], smaxCache[i]); });
}
} // namespace filt

// A filter chain restricted to the in-scope filters: kind 0 = Smoothing(radius, type), 1 = Projection(beta)
struct FilterSpec { int kind; int radius; int type; double beta; };
struct Problem { // TopologyOptimizationProblem + MultigridComplianceObjective + TotalVolumeConstraint
    std::shared_ptr<MG> mg; Sim *sim;
    std::vector<FilterSpec> filters; double volFrac;
    std::vector<std::vector<double>> vars; // vars[0] = design ... vars.back() = physical (FilterChain m_vars, :111-132)
    std::vector<double> u, f;
    int cgIter = 100; double tol = 1e-5; int mgIterations = 1, mgSmoothing = 2; bool fullMG = true, zeroInit = false; // TopologyOptimizationObjective.hh:99-103
    double lamMin = 1, lamMax = 2; // OptimalityCriterion.hh:46-49
    int lastPcgIters = 0;
    Problem(std::shared_ptr<MG> mg_, const std::vector<FilterSpec> &fl, double V) : mg(mg_), sim(mg_->sims[0].get()), filters(fl), volFrac(V) {
        vars.assign(filters.size() + 1, std::vector<double>(sim->numElems, 0.0));
        f.resize(size_t(sim->numNodes) * sim->N); sim->buildLoadVector(f.data());
        u.assign(f.size(), 0.0);
        updateCache(sim->rho.data()); // MultigridComplianceObjective ctor (TopologyOptimizationObjective.hh:82-86)
    }
    void applyFilter(const FilterSpec &fs, const double *in, double *out) const {
        if (fs.kind == 0) smoothingFilter(sim->N, sim->ne, fs.radius, fs.type, in, out);
        else projectionApply(sim->numElems, fs.beta, in, out);
    }
    void updateCache(const double *xPhys) { // TopologyOptimizationObjective.hh:88-96
        if (xPhys != sim->rho.data()) std::copy(xPhys, xPhys + sim->numElems, sim->rho.begin());
        sim->updateYoungModuli();
        if (zeroInit) std::fill(u.begin(), u.end(), 0.0);
        mg->pcg(u.data(), f.data(), cgIter, tol, mgIterations, mgSmoothing, fullMG, false);
        lastPcgIters = mg->lastIters;
    }
    void setVars(const double *xd) { // TopologyOptimizationProblem.hh:41-50, FilterChain::setDesignVars (:142-152)
        std::copy(xd, xd + sim->numElems, vars[0].begin());
        for (size_t i = 0; i < filters.size(); ++i) applyFilter(filters[i], vars[i].data(), vars[i + 1].data());
        updateCache(vars.back().data());
    }
    double compliance() const { double s = 0; for (size_t i = 0; i < f.size(); ++i) s += f[i] * u[i]; return 0.5 * s; } // :41-43
    void backprop(std::vector<double> &g) const { // FilterChain::backprop (:162-170)
        std::vector<double> scratch(g.size());
        for (size_t i = filters.size(); i-- > 0;) {
            if (filters[i].kind == 0) smoothingFilter(sim->N, sim->ne, filters[i].radius, filters[i].type, g.data(), scratch.data());
            else projectionBackprop(sim->numElems, filters[i].beta, g.data(), vars[i].data(), scratch.data());
            g.swap(scratch);
        }
    }
    void objectiveGradient(std::vector<double> &g) const { g.resize(sim->numElems); sim->complianceGradient(u.data(), g.data(), false); backprop(g); }
    double constraintValue(const std::vector<double> &xPhys) const { double s = 0; for (double v : xPhys) s += v; return 1.0 - (s / double(xPhys.size())) / volFrac; } // TopologyOptimizationConstraint.hh:30-32
    void constraintJacobian(std::vector<double> &dc) const { dc.assign(sim->numElems, -1.0 / (volFrac * double(sim->numElems))); backprop(dc); } // :34-36
    double evalOCConstraint(std::vector<double> &xv, std::vector<double> &scratch) const { // TopologyOptimizationProblem.hh:58-66
        for (const auto &fs : filters) { applyFilter(fs, xv.data(), scratch.data()); xv.swap(scratch); }
        return constraintValue(xv);
    }
    // OCOptimizer::step (OptimalityCriterion.hh:51-134)
    int ocStep(double m, double p, double ctol) {
        std::vector<double> dJ, dc; objectiveGradient(dJ); constraintJacobian(dc);
        const std::vector<double> x0 = vars[0];
        const idx n = idx(x0.size());
        std::vector<double> stepped(n), xv(n), scratch(n);
        int nevals = 0;
        auto ceval = [&](double lambda) {
            #pragma omp parallel for schedule(static)
            for (idx i = 0; i < n; ++i) {
                double res = x0[i] * std::pow(dJ[i] / (dc[i] * lambda), p);
                res = std::min(std::max(std::min(std::max(res, x0[i] - m), x0[i] + m), 0.0), 1.0);
                if (!std::isfinite(res)) res = x0[i];
                stepped[i] = res;
            }
            xv = stepped; ++nevals;
            return evalOCConstraint(xv, scratch);
        };
        const double dilation = 32;
        double mid = 0.5 * (lamMin + lamMax);
        lamMax = dilation * lamMax + (1 - dilation) * mid;
        lamMin = std::max(dilation * lamMin + (1 - dilation) * mid, 0.01);
        const int guard = 100; int nit = 0;
        for (; nit < guard; ++nit) { if (ceval(lamMin) < 0) break; lamMax = lamMin; lamMin /= 2; }
        if (nit == guard) throw std::runtime_error("Bracketing constraint(lambda_min) < 0 failed (100 times).");
        if (nit == 0) for (; nit < guard; ++nit) { if (ceval(lamMax) > 0) break; lamMin = lamMax; lamMax *= 2; }
        if (nit == guard) throw std::runtime_error("Bracketing constraint(lambda_max) > 0 failed (100 times).");
        double violation;
        do {
            mid = 0.5 * (lamMin + lamMax);
            violation = ceval(mid);
            if (std::abs(violation) <= ctol) break;
            ++nit;
            if (violation < 0) lamMin = mid;
            if (violation > 0) lamMax = mid;
        } while (true);
        setVars(stepped.data());
        return nevals;
    }
};


// ---------------------------------------------------------------------------
// LayerByLayerEvaluator (LayerByLayer.hh:25-309) with the Zero / FD / Subspace(N=k) initial-guess
// generators (:56-209) and the band-limited recurrences for A = U^T K U, b = U^T f (:149-202,
// TensorProductSimulator.hh:1292-1406).
// ---------------------------------------------------------------------------
// Eigen::JacobiSVD(A).solve(b) for a small symmetric matrix (LayerByLayer.hh:120-121): minimum-norm least-squares
// solution, singular values below eps * k * sigma_max dropped.  One-sided Jacobi on the columns of A.
static std::vector<double> svdSolveSmall(const std::vector<double> &Ain, const std::vector<double> &b, int k) {
    std::vector<double> U = Ain, V(size_t(k) * k, 0.0);     // U (k x k, row-major) converges to U * Sigma
    for (int i = 0; i < k; ++i) V[i * k + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < k; ++p) for (int q = p + 1; q < k; ++q) {
            double app = 0, aqq = 0, apq = 0;
            for (int r = 0; r < k; ++r) { app += U[r * k + p] * U[r * k + p]; aqq += U[r * k + q] * U[r * k + q]; apq += U[r * k + p] * U[r * k + q]; }
            if (std::abs(apq) <= 1e-300 || std::abs(apq) <= 1e-16 * std::sqrt(app * aqq)) continue;
            rotated = true;
            const double zeta = (aqq - app) / (2 * apq);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1 + zeta * zeta));
            const double c = 1 / std::sqrt(1 + t * t), sn = c * t;
            for (int r = 0; r < k; ++r) {
                const double up = U[r * k + p], uq = U[r * k + q]; U[r * k + p] = c * up - sn * uq; U[r * k + q] = sn * up + c * uq;
                const double vp = V[r * k + p], vq = V[r * k + q]; V[r * k + p] = c * vp - sn * vq; V[r * k + q] = sn * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    std::vector<double> sig(k), x(k, 0.0);
    double smax = 0;
    for (int j = 0; j < k; ++j) { double s = 0; for (int r = 0; r < k; ++r) s += U[r * k + j] * U[r * k + j]; sig[j] = std::sqrt(s); smax = std::max(smax, sig[j]); }
    const double thresh = std::numeric_limits<double>::epsilon() * k * smax;
    for (int j = 0; j < k; ++j) {
        if (sig[j] <= thresh) continue;
        double proj = 0; for (int r = 0; r < k; ++r) proj += (U[r * k + j] / sig[j]) * b[r];
        for (int r = 0; r < k; ++r) x[r] += V[r * k + j] * proj / sig[j];
    }
    return x;
}

struct LBL {
    std::shared_ptr<MG> mg; Sim *sim;
    int method = 2; size_t maxHist = 3;                  // selectInitMethod("N=3") (:33-36)
    std::vector<std::vector<double>> hist;               // front = most recent (m_storage, :86)
    std::vector<double> A, b;                            // subspace system (:205-206); A is s x s row-major
    std::vector<double> uFull, totalGrad;
    double totalCompliance = 0; idx layersAccumulated = 0;
    std::vector<int> layerIters; std::vector<double> layerCompliance;
    explicit LBL(std::shared_ptr<MG> m) : mg(std::move(m)), sim(mg->sims[0].get()) {}

    void selectInitMethod(const std::string &m) { // :214-220
        if (m == "zero") { method = 0; maxHist = 0; }
        else if (m == "constant") { method = 1; maxHist = 1; }
        else if (m == "fd") { method = 1; maxHist = 2; }
        else if (m.substr(0, 2) == "N=") { method = 2; maxHist = size_t(std::stoi(m.substr(2))); }
        else throw std::runtime_error("Unrecognized method " + m);
        hist.clear(); A.clear(); b.clear();
    }
    void checkGravity() const { // TensorProductSimulator.hh:1293-1294
        double g2 = 0; for (int c = 0; c < sim->N; ++c) g2 += sim->gravity[c] * sim->gravity[c];
        if (g2 == 0 || std::abs(g2 - sim->gravity[BUILD_DIRECTION] * sim->gravity[BUILD_DIRECTION]) > 1e-10) throw std::runtime_error("Unexpected gravity vector");
    }
    template<class F> void visitLayerElements(idx lbegin, idx lend, const F &f) const {
        for (idx e = 0; e < sim->numElems; ++e) { idx nd[3] = {0, 0, 0}; sim->ndElem(e, nd); if (nd[BUILD_DIRECTION] >= lbegin && nd[BUILD_DIRECTION] < lend) f(e, nd); }
    }
    // addLayerRemovalDeltaLoadVector (TensorProductSimulator.hh:1292-1305)
    void addLayerRemovalDeltaLoad(idx lbegin, idx lend, double *f) const {
        checkGravity();
        double vol = 1; for (int d = 0; d < sim->N; ++d) vol *= sim->stretch[d];
        const double intPhi = 1.0 / double(sim->npe);
        visitLayerElements(lbegin, lend, [&](idx e, const idx *nd) {
            const idx off = sim->firstNodeOfElem(nd);
            const double w = sim->gravity[BUILD_DIRECTION] * (sim->rho[e] * vol);
            for (int l = 0; l < sim->npe; ++l) f[BUILD_DIRECTION * sim->numNodes + sim->refNodes[l] + off] -= w * intPhi;
        });
    }
    // dotLayerRemovalDeltaLoadVector (:1309-1341)
    std::vector<double> dotLayerRemovalDeltaLoad(idx lbegin, idx lend) const {
        checkGravity();
        double vol = 1; for (int d = 0; d < sim->N; ++d) vol *= sim->stretch[d];
        const double intPhi = 1.0 / double(sim->npe);
        std::vector<double> out(hist.size(), 0.0);
        visitLayerElements(lbegin, lend, [&](idx e, const idx *nd) {
            const idx off = sim->firstNodeOfElem(nd);
            const double w = sim->gravity[BUILD_DIRECTION] * (sim->rho[e] * vol);
            for (size_t i = 0; i < hist.size(); ++i) {
                double dot = 0; for (int l = 0; l < sim->npe; ++l) dot += hist[i][BUILD_DIRECTION * sim->numNodes + sim->refNodes[l] + off] * intPhi;
                out[i] -= w * dot;
            }
        });
        return out;
    }
    // layerRemovalDeltaUKU (:1346-1406): lower triangle of  result(i, j) += u_i . (delta K) u_j  for voiding [lbegin, lend)
    void layerRemovalDeltaUKU(idx lbegin, idx lend, std::vector<double> &result, int s) const {
        const int N = sim->N, ke = sim->ke, npe = sim->npe;
        std::vector<double> dKu(size_t(sim->numNodes) * N);
        for (int i = 0; i < s; ++i) {
            std::fill(dKu.begin(), dKu.end(), 0.0);
            const std::vector<double> &u = hist[i];
            visitLayerElements(lbegin, lend, [&](idx e, const idx *nd) {
                const idx off = sim->firstNodeOfElem(nd);
                double ue[24], Ku[24];
                for (int l = 0; l < npe; ++l) for (int c = 0; c < N; ++c) ue[l * N + c] = u[c * sim->numNodes + sim->refNodes[l] + off];
                for (int a = 0; a < ke; ++a) { double t = 0; for (int bb = 0; bb < ke; ++bb) t += sim->k0(a, bb) * ue[bb]; Ku[a] = t * sim->E[e]; }
                for (int l = 0; l < npe; ++l) for (int c = 0; c < N; ++c) dKu[c * sim->numNodes + sim->refNodes[l] + off] -= Ku[l * N + c];
            });
            for (int j = 0; j <= i; ++j) { double t = 0; for (size_t k = 0; k < dKu.size(); ++k) t += dKu[k] * hist[j][k]; result[i * s + j] += t; }
        }
    }
    void addToHistory(std::vector<double> &u) { // m_addToHistory (:72-84)
        if (maxHist == 0) return;
        if (hist.size() == maxHist) hist.pop_back();
        hist.insert(hist.begin(), std::vector<double>());
        std::swap(u, hist.front());
    }
    void constructGuess(std::vector<double> &g) const {
        const size_t len = size_t(sim->numNodes) * sim->N, s = hist.size();
        g.assign(len, 0.0);
        if (method == 0 || s == 0) return;                               // :56-61, :96, :115-118
        if (method == 1) {                                                // InitGenFD (:95-100)
            if (s == 1) g = hist[0];
            else if (s == 2) for (size_t k = 0; k < len; ++k) g[k] = 2 * hist[0][k] - hist[1][k];
            else throw std::runtime_error("Unimplemented");
            return;
        }
        const std::vector<double> c = svdSolveSmall(A, b, int(s));       // :120-121
        sim->maskedVisit([&](idx start, idx n) {                          // :124-128 (detached rows stay zero, :129-146)
            for (int comp = 0; comp < sim->N; ++comp) for (idx k = 0; k < n; ++k) {
                const size_t at = size_t(comp) * sim->numNodes + start + k;
                double t = c[0] * hist[0][at]; for (size_t i = 1; i < s; ++i) t += c[i] * hist[i][at]; g[at] = t;
            }
        });
    }
    // InitGenSubspace::finalizeLayer (:149-202) / InitializationGenerator::finalizeLayer (:44-46)
    void finalizeLayer(idx lbegin, idx lend, std::vector<double> &u, const double *r, double compliance) {
        addToHistory(u);
        if (method != 2) return;
        const int s = int(hist.size());
        const std::vector<double> bOld = b, AOld = A; const int so = int(bOld.size());
        b.assign(s, 0.0); b[0] = compliance;
        for (int i = 1; i < s; ++i) b[i] = bOld[i - 1];
        const std::vector<double> db = dotLayerRemovalDeltaLoad(lbegin, lend);
        for (int i = 0; i < s; ++i) b[i] += db[i];
        A.assign(size_t(s) * s, 0.0);
        A[0] = compliance;
        for (int i = 1; i < s; ++i) A[i * s] = bOld[i - 1];
        for (int i = 0; i < s; ++i) A[i * s] -= sim->maskedDot(r, hist[i].data());   // U^T r (:182-197)
        for (int i = 1; i < s; ++i) for (int j = 1; j < s; ++j) A[i * s + j] = AOld[(i - 1) * so + (j - 1)];
        layerRemovalDeltaUKU(lbegin, lend, A, s);
        for (int i = 0; i < s; ++i) for (int j = i + 1; j < s; ++j) A[i * s + j] = A[j * s + i];
    }
    // run (:223-296)
    void run(bool zeroInit, idx layerIncrement, int maxIter, double tol, int mgIterations, int mgSmoothing, bool fmg) {
        const idx numLayers = sim->ne[BUILD_DIRECTION];
        const size_t len = size_t(sim->numNodes) * sim->N;
        std::vector<double> f(len), u;
        if (!zeroInit) u = uFull;
        hist.clear(); A.clear(); b.clear();
        layersAccumulated = 0; totalCompliance = 0; totalGrad.assign(sim->numElems, 0.0); layerIters.clear(); layerCompliance.clear();
        mg->setMaskByLayer(numLayers);
        sim->buildLoadVector(f.data());
        for (idx l = numLayers; l > 0; l -= std::min(layerIncrement, l)) {
            if (l < numLayers) { mg->decrementMaskByLayer(int(layerIncrement)); addLayerRemovalDeltaLoad(l, l + layerIncrement, f.data()); }
            if (l < numLayers || u.size() != len) constructGuess(u);
            mg->pcg(u.data(), f.data(), maxIter, tol, mgIterations, mgSmoothing, fmg, true);
            layerIters.push_back(mg->lastIters);
            const double compliance = sim->maskedDot(f.data(), u.data());
            layerCompliance.push_back(compliance);
            totalCompliance += compliance;
            sim->complianceGradient(u.data(), totalGrad.data(), true);
            ++layersAccumulated;
            if (l == numLayers) uFull = u;
            if (l >= layerIncrement) finalizeLayer(l - layerIncrement, l, u, mg->b[0].data(), compliance);
        }
    }
};


// ---------------------------------------------------------------------------
// MMA / GCMMA (MethodOfMovingAsymptotes.hh:28-469): Svanberg's method of moving asymptotes with the primal-dual
// interior-point subproblem solver.  Row i of the (m+1) x n arrays is function i (0 = objective).
// ---------------------------------------------------------------------------
typedef void (*mma_f_cb)(const double *x, double *f, void *user);          // f: m + 1 values
typedef void (*mma_df_cb)(const double *x, double *df, void *user);        // df: (m + 1) x n, row-major
// Solve the small dense system M sol = rhs (MethodOfMovingAsymptotes.hh:351 uses colPivHouseholderQr; any backward-stable
// solve agrees to rounding): Gaussian elimination with partial pivoting.
static std::vector<double> solveDense(std::vector<double> M, std::vector<double> rhs, int k) {
    for (int c = 0; c < k; ++c) {
        int piv = c; for (int r = c + 1; r < k; ++r) if (std::abs(M[r * k + c]) > std::abs(M[piv * k + c])) piv = r;
        if (piv != c) { for (int j = 0; j < k; ++j) std::swap(M[c * k + j], M[piv * k + j]); std::swap(rhs[c], rhs[piv]); }
        for (int r = c + 1; r < k; ++r) {
            const double fct = M[r * k + c] / M[c * k + c];
            for (int j = c; j < k; ++j) M[r * k + j] -= fct * M[c * k + j];
            rhs[r] -= fct * rhs[c];
        }
    }
    std::vector<double> x(k);
    for (int r = k - 1; r >= 0; --r) { double t = rhs[r]; for (int j = r + 1; j < k; ++j) t -= M[r * k + j] * x[j]; x[r] = t / M[r * k + r]; }
    return x;
}
struct MMA {
    using V = std::vector<double>;
    int m, n; V xmin, xmax, xdiff, a, d, c; const double a0 = 1;
    mma_f_cb f; mma_df_cb df; void *user;
    bool enableInner = false; int outerIter = 0, innerIter = 0;
    std::vector<V> xhist;                         // front = x^k (FixedSizeDeque<AXd>{3})
    V l, u, alpha, beta, xInner, rho, p, q, r, fcur, dfp, dfm, diffFSubf;
    struct IV { V umx, xml, umx2, xml2; } cur, old;
    const double raa0 = 1e-5, albefa = 0.1, move = 0.5, asyinit = 0.5; // :196
    // subproblem state (:452-465)
    struct Vars { double z, zeta; V x, xi, eta, y, lam, mu, s; } data, delta;
    V Dx, G, dpsi, plam, qlam, gvec;
    long subsolveNewtonIters = 0;

    MMA(int n_, int m_, const double *xmin_, const double *xmax_, mma_f_cb f_, mma_df_cb df_, void *user_)
        : m(m_), n(n_), xmin(xmin_, xmin_ + n_), xmax(xmax_, xmax_ + n_), xdiff(n_), a(m_, 0.0), d(m_, 1.0), c(m_, 1000.0), f(f_), df(df_), user(user_) {
        for (int j = 0; j < n; ++j) xdiff[j] = xmax[j] - xmin[j];
        l.assign(n, 0); u.assign(n, 0); alpha.assign(n, 0); beta.assign(n, 0);
        p.assign(size_t(m + 1) * n, 0); q.assign(size_t(m + 1) * n, 0); dfp.assign(size_t(m + 1) * n, 0); dfm.assign(size_t(m + 1) * n, 0);
        cur.umx.assign(n, 0); cur.xml.assign(n, 0); cur.umx2.assign(n, 0); cur.xml2.assign(n, 0);
    }
    void setInitialVar(const double *x) { addToHistory(V(x, x + n)); }   // :54-56
    void addToHistory(V x) { xhist.insert(xhist.begin(), std::move(x)); if (xhist.size() > 3) xhist.pop_back(); }
    const V &xcur() const { return xhist[0]; }
    // g_i(x) = sum_j p_ij / (u_j - x_j) + q_ij / (x_j - l_j)   (:160-180); rows [first, m]
    V subG(bool allRows, bool useOld = false) const {
        const IV &v = useOld ? old : cur; const int first = allRows ? 0 : 1;
        V res(m + 1 - first);
        for (int i = first; i <= m; ++i) { double t = 0; for (int j = 0; j < n; ++j) t += p[size_t(i) * n + j] / v.umx[j] + q[size_t(i) * n + j] / v.xml[j]; res[i - first] = t; }
        return res;
    }
    void refreshCur(const V &x) { for (int j = 0; j < n; ++j) { cur.umx[j] = u[j] - x[j]; cur.umx2[j] = cur.umx[j] * cur.umx[j]; cur.xml[j] = x[j] - l[j]; cur.xml2[j] = cur.xml[j] * cur.xml[j]; } }
    void buildPQ(const V &rhoRow, const IV &v) { // :108-114 / :123-129
        for (int i = 0; i <= m; ++i) for (int j = 0; j < n; ++j) {
            const double rod = rhoRow[i] / xdiff[j], dp = dfp[size_t(i) * n + j], dm = dfm[size_t(i) * n + j];
            p[size_t(i) * n + j] = v.umx2[j] * (1.001 * dp + 0.001 * dm + rod);
            q[size_t(i) * n + j] = v.xml2[j] * (0.001 * dp + 1.001 * dm + rod);
        }
    }
    V nextRho() const { // :136-156
        V res(m + 1);
        if (innerIter == 0) {
            for (int i = 0; i <= m; ++i) { double t = 0; for (int j = 0; j < n; ++j) t += 0.1 / n * (dfp[size_t(i) * n + j] + dfm[size_t(i) * n + j]) * xdiff[j]; res[i] = std::max(t, 1e-6); }
            return res;
        }
        double dd = 0;
        for (int j = 0; j < n; ++j) { const double dx = xInner[j] - xcur()[j]; dd += (u[j] - l[j]) * dx * dx / (cur.umx[j] * cur.xml[j] * xdiff[j]); }
        for (int i = 0; i <= m; ++i) { const double del = diffFSubf[i] / dd; res[i] = del < 0 ? rho[i] : std::min(1.1 * (rho[i] + del), 10 * rho[i]); }
        return res;
    }
    bool isFeasible() { // :190-193
        V fx(m + 1); f(xInner.data(), fx.data(), user);
        const V g = subG(true); diffFSubf.assign(m + 1, 0.0);
        double mx = -std::numeric_limits<double>::infinity();
        for (int i = 0; i <= m; ++i) { diffFSubf[i] = fx[i] - (g[i] + r[i]); mx = std::max(mx, diffFSubf[i]); }
        return mx < 0;
    }
    void step() { // :63-133
        if (xhist.empty()) throw std::runtime_error("Must specify an initial value");
        ++outerIter;
        const V &x = xcur();
        if (outerIter <= 2) for (int j = 0; j < n; ++j) { l[j] = x[j] - asyinit * xdiff[j]; u[j] = x[j] + asyinit * xdiff[j]; }
        else for (int j = 0; j < n; ++j) {
            const double diff = (xhist[0][j] - xhist[1][j]) * (xhist[1][j] - xhist[2][j]);
            const double gam = diff > 0 ? 1.2 : 0.7;
            l[j] = std::max(std::min(xhist[0][j] - gam * (xhist[1][j] - l[j]), x[j] - 0.01 * xdiff[j]), x[j] - 10 * xdiff[j]);
            u[j] = std::min(std::max(xhist[0][j] + gam * (u[j] - xhist[1][j]), x[j] + 0.01 * xdiff[j]), x[j] + 10 * xdiff[j]);
        }
        for (int j = 0; j < n; ++j) {
            alpha[j] = std::max(std::max(xmin[j], l[j] + albefa * (x[j] - l[j])), x[j] - move * xdiff[j]);
            beta[j]  = std::min(std::min(xmax[j], u[j] - albefa * (u[j] - x[j])), x[j] + move * xdiff[j]);
        }
        fcur.assign(m + 1, 0.0); f(x.data(), fcur.data(), user);
        df(x.data(), dfp.data(), user);
        for (size_t k = 0; k < dfp.size(); ++k) { dfm[k] = std::max(-dfp[k], 0.0); dfp[k] = std::max(dfp[k], 0.0); }
        refreshCur(x);
        if (enableInner) {
            old = cur;
            do {
                rho = nextRho();
                buildPQ(rho, old);
                const V g = subG(true, true); r.assign(m + 1, 0.0); for (int i = 0; i <= m; ++i) r[i] = fcur[i] - g[i];
                xInner = subsolve();
                ++innerIter;
            } while (!isFeasible());
            addToHistory(xInner); innerIter = 0;
        } else {
            buildPQ(V(m + 1, raa0), cur);
            const V g = subG(true); r.assign(m + 1, 0.0); for (int i = 0; i <= m; ++i) r[i] = fcur[i] - g[i];
            addToHistory(subsolve());
        }
    }
    // ---- Subproblem (:240-466) ----
    void refreshDual() { // the recomputation block of init_vars (:303-312) and squared_residual (:401-411)
        refreshCur(data.x);
        for (int j = 0; j < n; ++j) {
            double pl = p[j], ql = q[j];
            for (int i = 0; i < m; ++i) { pl += p[size_t(i + 1) * n + j] * data.lam[i]; ql += q[size_t(i + 1) * n + j] * data.lam[i]; }
            plam[j] = pl; qlam[j] = ql; dpsi[j] = pl / cur.umx2[j] - ql / cur.xml2[j];
        }
        gvec = subG(false);
    }
    void initVars() { // :281-313
        data.x.assign(n, 0); data.xi.assign(n, 0); data.eta.assign(n, 0);
        data.y.assign(m, 1.0); data.z = 1; data.zeta = 1; data.lam.assign(m, 1.0); data.s.assign(m, 1.0); data.mu.assign(m, 0.0);
        for (int i = 0; i < m; ++i) data.mu[i] = std::max(c[i] / 2, 1.0);
        Dx.assign(n, 0); G.assign(size_t(m) * n, 0); delta.x.assign(n, 0); delta.xi.assign(n, 0); delta.eta.assign(n, 0);
        dpsi.assign(n, 0); plam.assign(n, 0); qlam.assign(n, 0);
        for (int j = 0; j < n; ++j) {
            data.x[j] = 0.5 * (alpha[j] + beta[j]);
            data.xi[j] = std::max(1.0 / (data.x[j] - alpha[j]), 1.0);
            data.eta[j] = std::max(1.0 / (beta[j] - data.x[j]), 1.0);
        }
        refreshDual();
    }
    void newtonDirection(double eps) { // :315-363
        for (int j = 0; j < n; ++j) {
            const double xa = data.x[j] - alpha[j], bx = beta[j] - data.x[j];
            delta.x[j] = dpsi[j] - eps / xa + eps / bx;
            Dx[j] = 2 * plam[j] / (cur.umx[j] * cur.umx2[j]) + 2 * qlam[j] / (cur.xml[j] * cur.xml2[j]) + data.xi[j] / xa + data.eta[j] / bx;
            for (int i = 0; i < m; ++i) G[size_t(i) * n + j] = p[size_t(i + 1) * n + j] / cur.umx2[j] - q[size_t(i + 1) * n + j] / cur.xml2[j];
        }
        V Dy(m), dy(m);
        for (int i = 0; i < m; ++i) { Dy[i] = d[i] + data.mu[i] / data.y[i]; dy[i] = c[i] + d[i] * data.y[i] - data.lam[i] - eps / data.y[i]; }
        const int k = m + 1; V M(size_t(k) * k, 0.0), rhs(k, 0.0);
        for (int ci = 0; ci < m; ++ci) for (int cj = ci; cj < m; ++cj) { double t = 0; for (int j = 0; j < n; ++j) t += G[size_t(ci) * n + j] * G[size_t(cj) * n + j] / Dx[j]; M[ci * k + cj] = t; }
        for (int i = 0; i < m; ++i) { M[i * k + i] += data.s[i] / data.lam[i] + 1 / Dy[i]; M[i * k + m] = a[i]; }
        M[m * k + m] = -data.zeta / data.z;
        for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j) M[j * k + i] = M[i * k + j];
        for (int i = 0; i < m; ++i) {
            rhs[i] = gvec[i] - a[i] * data.z - data.y[i] + r[i + 1] + eps / data.lam[i] + dy[i] / Dy[i];
            double t = 0; for (int j = 0; j < n; ++j) t += G[size_t(i) * n + j] * (delta.x[j] / Dx[j]);
            rhs[i] -= t;
        }
        double la = 0; for (int i = 0; i < m; ++i) la += data.lam[i] * a[i];
        rhs[m] = a0 - la - eps / data.z;
        const V sol = solveDense(M, rhs, k);
        delta.lam.assign(sol.begin(), sol.begin() + m); delta.z = sol[m];
        for (int j = 0; j < n; ++j) {
            const double xa = data.x[j] - alpha[j], bx = beta[j] - data.x[j];
            double gl = 0; for (int i = 0; i < m; ++i) gl += G[size_t(i) * n + j] * delta.lam[i];
            delta.x[j] = -(delta.x[j] + gl) / Dx[j];
            delta.xi[j] = -data.xi[j] + eps / xa - data.xi[j] * delta.x[j] / xa;
            delta.eta[j] = -data.eta[j] + eps / bx + data.eta[j] * delta.x[j] / bx;
        }
        delta.y.assign(m, 0); delta.mu.assign(m, 0); delta.s.assign(m, 0);
        for (int i = 0; i < m; ++i) {
            delta.y[i] = delta.lam[i] / Dy[i] - dy[i] / Dy[i];
            delta.mu[i] = (eps - data.mu[i] * delta.y[i]) / data.y[i] - data.mu[i];
            delta.s[i] = (eps - data.s[i] * delta.lam[i]) / data.lam[i] - data.s[i];
        }
        delta.zeta = (eps - data.zeta * delta.z) / data.z - data.zeta;
    }
    double stepSatisfyKKT() const { // :384-397
        double mn = std::numeric_limits<double>::infinity();
        for (int j = 0; j < n; ++j) {
            mn = std::min(mn, delta.x[j] / (data.x[j] - alpha[j])); mn = std::min(mn, delta.x[j] / (data.x[j] - beta[j]));
            mn = std::min(mn, delta.xi[j] / data.xi[j]); mn = std::min(mn, delta.eta[j] / data.eta[j]);
        }
        for (int i = 0; i < m; ++i) { mn = std::min(mn, delta.y[i] / data.y[i]); mn = std::min(mn, delta.s[i] / data.s[i]); mn = std::min(mn, delta.mu[i] / data.mu[i]); mn = std::min(mn, delta.lam[i] / data.lam[i]); }
        mn = std::min(mn, delta.z / data.z); mn = std::min(mn, delta.zeta / data.zeta);
        return 1 / std::max(-1.01 * mn, 1.0);
    }
    void newtonStep(double t) { // :426-438
        for (int j = 0; j < n; ++j) { data.x[j] += t * delta.x[j]; data.xi[j] += t * delta.xi[j]; data.eta[j] += t * delta.eta[j]; }
        for (int i = 0; i < m; ++i) { data.y[i] += t * delta.y[i]; data.lam[i] += t * delta.lam[i]; data.mu[i] += t * delta.mu[i]; data.s[i] += t * delta.s[i]; }
        data.z += t * delta.z; data.zeta += t * delta.zeta;
    }
    // eq_a .. eq_i (:440-450): squared norm and max-norm of the perturbed KKT residual
    void kktResidual(double eps, const V &g, double &sq, double &mx) const {
        sq = 0; mx = 0;
        auto acc = [&](double v) { sq += v * v; mx = std::max(mx, std::abs(v)); };
        for (int j = 0; j < n; ++j) { acc(dpsi[j] - data.xi[j] + data.eta[j]); acc(data.xi[j] * (data.x[j] - alpha[j]) - eps); acc(data.eta[j] * (beta[j] - data.x[j]) - eps); }
        double la = 0;
        for (int i = 0; i < m; ++i) {
            acc(c[i] + d[i] * data.y[i] - data.lam[i] - data.mu[i]);
            acc(g[i] - a[i] * data.z - data.y[i] + data.s[i] + r[i + 1]);
            acc(data.mu[i] * data.y[i] - eps); acc(data.lam[i] * data.s[i] - eps);
            la += data.lam[i] * a[i];
        }
        acc(a0 - data.zeta - la); acc(data.zeta * data.z - eps);
    }
    double squaredResidual(double eps, bool init = false) { if (!init) refreshDual(); double sq, mx; kktResidual(eps, gvec, sq, mx); return sq; } // :399-424
    double kktInfNorm(double eps) const { double sq, mx; kktResidual(eps, subG(false), sq, mx); return mx; }                               // :261-279
    double backtrack(double eps, double resOld) { // :365-376
        double t = stepSatisfyKKT();
        newtonStep(t);
        double res = squaredResidual(eps);
        while (res > resOld) { t /= 2; newtonStep(-t); res = squaredResidual(eps); }
        return res;
    }
    V subsolve() { // :242-258
        initVars();
        double eps = 1;
        while (eps > 1e-7) {
            double resOld = 0;
            for (int i = 0; i < 10; ++i) {
                newtonDirection(eps); ++subsolveNewtonIters;
                if (i == 0) resOld = squaredResidual(eps, true);
                resOld = backtrack(eps, resOld);
                if (kktInfNorm(eps) <= 0.9 * eps) break;
            }
            eps *= 0.1;
        }
        return data.x;
    }
};

} // namespace vfo

// ---------------------------------------------------------------------------
// C ABI for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------
using namespace vfo;
struct SimHandle { std::shared_ptr<Sim> s; };
struct MGHandle { std::shared_ptr<MG> m; };
struct ProbHandle { std::unique_ptr<Problem> p; };

#define VFO_TRY try {
#define VFO_CATCH } catch (const std::exception &e) { g_err = e.what(); return -1; } return 0;

extern "C" {
const char *vfo_last_error() { return g_err.c_str(); }
int vfo_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void vfo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void *vfo_sim_create(int N, const int64_t *ne, const double *dmin, const double *dmax) { auto *h = new SimHandle; h->s = std::make_shared<Sim>(N, ne, dmin, dmax); return h; }
void vfo_sim_destroy(void *h) { delete static_cast<SimHandle *>(h); }
static Sim &S(void *h) { return *static_cast<SimHandle *>(h)->s; }
int64_t vfo_sim_num_nodes(void *h) { return S(h).numNodes; }
int64_t vfo_sim_num_elements(void *h) { return S(h).numElems; }
void vfo_sim_set_isotropic(void *h, double E, double nu) { S(h).setIsotropic(E, nu); }
void vfo_sim_set_D(void *h, const double *D) { S(h).setD(D); }
void vfo_sim_get_K0(void *h, double *out) { std::copy(S(h).K0.begin(), S(h).K0.end(), out); }
void vfo_sim_set_interp(void *h, int law, double E0, double Emin, double gamma, double q) { Sim &s = S(h); s.law = law; s.E0 = E0; s.Emin = Emin; s.gamma = gamma; s.q = q; s.updateYoungModuli(); }
void vfo_sim_set_gravity(void *h, const double *g) { for (int c = 0; c < S(h).N; ++c) S(h).gravity[c] = g[c]; }
int vfo_sim_set_densities(void *h, const double *rho) { VFO_TRY Sim &s = S(h); for (idx e = 0; e < s.numElems; ++e) if (rho[e] > 1.0 || rho[e] < 0) {} std::copy(rho, rho + s.numElems, s.rho.begin()); s.updateYoungModuli(); VFO_CATCH }
int vfo_sim_set_uniform_density(void *h, double v) { VFO_TRY if (v > 1.0 || v < 0) throw std::runtime_error("Density value has to be in between 0 and 1"); Sim &s = S(h); std::fill(s.rho.begin(), s.rho.end(), v); s.updateYoungModuli(); VFO_CATCH }
void vfo_sim_get_E(void *h, double *out) { std::copy(S(h).E.begin(), S(h).E.end(), out); }
void vfo_sim_get_densities(void *h, double *out) { std::copy(S(h).rho.begin(), S(h).rho.end(), out); }
int vfo_sim_apply_bcs(void *h, int nreg, const int *kind, const int *cmask, const double *values, const double *bmin, const double *bmax) { VFO_TRY S(h).applyBCs(nreg, kind, cmask, values, bmin, bmax); VFO_CATCH }
int vfo_sim_add_dirichlet(void *h, const double *u, const double *lo, const double *hi, int cmask) { VFO_TRY S(h).addDirichletCondition(u, lo, hi, cmask); VFO_CATCH }
void vfo_sim_get_dirichlet_mask(void *h, uint8_t *out) { std::copy(S(h).nodeDirMask.begin(), S(h).nodeDirMask.end(), out); }
int64_t vfo_sim_num_force_nodes(void *h) { return int64_t(S(h).forceNodes.size()); }
void vfo_sim_build_load(void *h, double *f) { S(h).buildLoadVector(f); }
void vfo_sim_apply_K(void *h, const double *u, double *out, int zeroInit, int negate) { S(h).applyK0(u, out, zeroInit != 0, negate != 0); }
int vfo_sim_set_mask_layer(void *h, int64_t l) { VFO_TRY S(h).setMaskHeightByLayer(l); VFO_CATCH }
void vfo_sim_mask_info(void *h, int64_t *firstMasked, int64_t *firstDetached) { *firstMasked = S(h).firstMasked; *firstDetached = S(h).firstDetached; }
void vfo_sim_compliance_gradient(void *h, const double *u, double *g, int accumulate) { S(h).complianceGradient(u, g, accumulate != 0); }
void vfo_sim_energy_density(void *h, const double *u, double *out) { S(h).elementEnergyDensity(u, out); }
int vfo_sim_solve(void *h, const double *f, double *u) { VFO_TRY S(h).solve(f, u); VFO_CATCH }
void vfo_sim_zero_dirichlet(void *h, double *u) { S(h).zeroOutDirichlet(u); }
double vfo_sim_masked_dot(void *h, const double *a, const double *b) { return S(h).maskedDot(a, b); }

void *vfo_mg_create(void *simh, int levels) { try { auto *h = new MGHandle; h->m = std::make_shared<MG>(static_cast<SimHandle *>(simh)->s, levels); return h; } catch (const std::exception &e) { g_err = e.what(); return nullptr; } }
void vfo_mg_destroy(void *h) { delete static_cast<MGHandle *>(h); }
static MG &M(void *h) { return *static_cast<MGHandle *>(h)->m; }
int vfo_mg_num_levels(void *h) { return M(h).numLevels(); }
void *vfo_mg_get_sim(void *h, int l) { auto *sh = new SimHandle; sh->s = M(h).sims.at(l); return sh; } // caller frees with vfo_sim_destroy
void vfo_mg_get_coarsened_fine_K0(void *h, int fi, double *out) { std::copy(M(h).cK0[fi].begin(), M(h).cK0[fi].end(), out); }
int vfo_mg_update_stiffness(void *h) { VFO_TRY M(h).updateStiffnessMatrices(); VFO_CATCH }
int vfo_mg_apply_K(void *h, int l, const double *u, double *out) { VFO_TRY M(h).applyK(l, u, out, true, false); VFO_CATCH }
int vfo_mg_residual(void *h, int l, const double *u, const double *b, double *r) { VFO_TRY M(h).computeResidual(l, u, b, r); VFO_CATCH }
int vfo_mg_smooth(void *h, int l, double *u, const double *b, int forward) { VFO_TRY M(h).smoothMulticolored(l, u, b, forward != 0); VFO_CATCH }
void vfo_mg_restrict(void *h, int lf, const double *fine, double *coarse) { M(h).restrict_(lf, fine, coarse); }
void vfo_mg_interpolate(void *h, int lf, const double *coarse, double *fine, int accumulate) { M(h).interpolate(lf, coarse, fine, accumulate != 0); }
// assembled block stencil of level l >= 1 ([node][3^N][N*N]); levels that do not store one get it built from per-element matrices here (test helper).
int vfo_mg_get_stencil(void *h, int l, double *out) {
    VFO_TRY
    MG &mg = M(h); Sim &sim = *mg.sims.at(l);
    const size_t len = size_t(sim.numNodes) * sim.nstencil() * sim.N * sim.N;
    if (sim.hasStencil) { std::copy(sim.stencil.begin(), sim.stencil.end(), out); return 0; }
    std::vector<double> saved; saved.swap(sim.stencil); sim.stencil.assign(len, 0.0);
    std::vector<double> Ke(size_t(sim.ke) * sim.ke);
    for (idx e = 0; e < sim.numElems; ++e) {
        idx nd[3]; sim.ndElem(e, nd);
        if (l == 0) sim.elementStiffness(e, Ke.data());
        else if (sim.hasKeCache) sim.elementStiffness(e, Ke.data());
        else if (l == 1) { if (sim.elemMasked(nd)) std::fill(Ke.begin(), Ke.end(), 0.0); else mg.firstLevelKe(nd, Ke.data()); }
        else throw std::runtime_error("no operator at this level; call update_stiffness first");
        mg.accumToStencil(sim, nd, Ke.data(), 1.0);
    }
    std::copy(sim.stencil.begin(), sim.stencil.end(), out); sim.stencil.swap(saved);
    VFO_CATCH
}
int vfo_mg_coarse_solve(void *h, const double *f, double *x) { VFO_TRY MG &mg = M(h); mg.sims.back()->solve(f, x); VFO_CATCH }
int vfo_mg_solve(void *h, const double *u, const double *f, int numSteps, int nsmooth, int stiffnessUpdated, int zeroDirichlet, int fmg, double *out) {
    VFO_TRY const auto &res = M(h).solve(u, f, numSteps, nsmooth, stiffnessUpdated != 0, zeroDirichlet != 0, fmg != 0); std::copy(res.begin(), res.end(), out); VFO_CATCH }
void vfo_mg_set_stiffness_prebuilt(void *h, int on) { M(h).stiffnessPrebuilt = on != 0; }
int vfo_mg_pcg(void *h, double *x, const double *b, int maxIter, double tol, int mgIt, int mgSmooth, int fmg, int dirichletOK, int *outIters, double *outResiduals) {
    VFO_TRY MG &mg = M(h); mg.pcg(x, b, maxIter, tol, mgIt, mgSmooth, fmg != 0, dirichletOK != 0);
    if (outIters) *outIters = mg.lastIters;
    if (outResiduals) std::copy(mg.lastResiduals.begin(), mg.lastResiduals.end(), outResiduals);
    VFO_CATCH }
void vfo_mg_get_pcg_residual(void *h, double *out) { std::copy(M(h).b[0].begin(), M(h).b[0].end(), out); }
void vfo_mg_set_symmetric_gs(void *h, int s) { M(h).symmetricGS = s != 0; }
int vfo_mg_set_mask_layer(void *h, int64_t l) { VFO_TRY M(h).setMaskByLayer(l); VFO_CATCH }
int vfo_mg_decrement_mask(void *h, int inc) { VFO_TRY M(h).decrementMaskByLayer(inc); VFO_CATCH }
void vfo_mg_debug_get(void *h, int which, int l, double *out) { MG &mg = M(h); const auto &v = which == 0 ? mg.x.at(l) : (which == 1 ? mg.b.at(l) : mg.r.at(l)); std::copy(v.begin(), v.end(), out); }
void vfo_mg_debug_multicolor_visit(void *h, int32_t *out) { MG &mg = M(h); const Sim &s = *mg.sims[0]; int32_t i = 0; mg.visitNodesMulticolored(0, [&](const idx *g) { out[s.flatNode(g)] = i++; }, true, false); }

void vfo_smoothing_filter(int N, const int64_t *sizes, int radius, int type, const double *in, double *out) { smoothingFilter(N, sizes, radius, type, in, out); }
void vfo_projection_apply(int64_t n, double beta, const double *in, double *out) { projectionApply(n, beta, in, out); }
void vfo_projection_backprop(int64_t n, double beta, const double *in, const double *vars, double *out) { projectionBackprop(n, beta, in, vars, out); }
void vfo_filter_upsample(int N, const int64_t *cs, int factor, const double *in, double *out) { filt::upsample(N, cs, factor, in, out); }
void vfo_filter_upsample_backprop(int N, const int64_t *cs, int factor, const double *dout, double *din) { filt::upsampleBackprop(N, cs, factor, dout, din); }
void vfo_filter_v2c(int N, const int64_t *vs, const double *in, double *out) { filt::v2c(N, vs, in, out); }
void vfo_filter_v2c_backprop(int N, const int64_t *vs, const double *dout, double *din) { filt::v2cBackprop(N, vs, dout, din); }
void vfo_filter_langelaar(int N, const int64_t *sz, const double *in, double *out, double *smax) { filt::langelaar(N, sz, in, out, smax); }
void vfo_filter_langelaar_backprop(int N, const int64_t *sz, const double *in, const double *vars, const double *filtered, const double *smax, double *out) { filt::langelaarBackprop(N, sz, in, vars, filtered, smax, out); }

// filters: flat array of (kind, radius, type, beta) quadruples as doubles
void *vfo_problem_create(void *mgh, int nfilters, const double *fspec, double volFrac) {
    try {
        std::vector<FilterSpec> fl;
        for (int i = 0; i < nfilters; ++i) fl.push_back({int(fspec[4 * i]), int(fspec[4 * i + 1]), int(fspec[4 * i + 2]), fspec[4 * i + 3]});
        auto *h = new ProbHandle; h->p = std::make_unique<Problem>(static_cast<MGHandle *>(mgh)->m, fl, volFrac); return h;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void vfo_problem_destroy(void *h) { delete static_cast<ProbHandle *>(h); }
static Problem &P(void *h) { return *static_cast<ProbHandle *>(h)->p; }
void vfo_problem_set_solver(void *h, int cgIter, double tol, int mgIt, int mgSmooth, int fmg, int zeroInit) { Problem &p = P(h); p.cgIter = cgIter; p.tol = tol; p.mgIterations = mgIt; p.mgSmoothing = mgSmooth; p.fullMG = fmg != 0; p.zeroInit = zeroInit != 0; }
int vfo_problem_set_vars(void *h, const double *x) { VFO_TRY P(h).setVars(x); VFO_CATCH }
void vfo_problem_get_vars(void *h, int which, double *out) { Problem &p = P(h); const auto &v = which == 0 ? p.vars.front() : p.vars.back(); std::copy(v.begin(), v.end(), out); }
double vfo_problem_compliance(void *h) { return P(h).compliance(); }
double vfo_problem_constraint(void *h) { return P(h).constraintValue(P(h).vars.back()); }
void vfo_problem_objective_gradient(void *h, double *g) { std::vector<double> v; P(h).objectiveGradient(v); std::copy(v.begin(), v.end(), g); }
void vfo_problem_constraint_jacobian(void *h, double *g) { std::vector<double> v; P(h).constraintJacobian(v); std::copy(v.begin(), v.end(), g); }
void vfo_problem_get_u(void *h, double *u) { std::copy(P(h).u.begin(), P(h).u.end(), u); }
int vfo_problem_last_pcg_iters(void *h) { return P(h).lastPcgIters; }
int vfo_problem_oc_step(void *h, double m, double p, double ctol, int *nevals) { VFO_TRY int n = P(h).ocStep(m, p, ctol); if (nevals) *nevals = n; VFO_CATCH }
void vfo_problem_get_lambda(void *h, double *lo, double *hi) { *lo = P(h).lamMin; *hi = P(h).lamMax; }
}

// ---- layer-by-layer evaluator ----
struct LBLHandle { std::unique_ptr<LBL> l; };
static LBL &LB(void *h) { return *static_cast<LBLHandle *>(h)->l; }
extern "C" {
void *vfo_lbl_create(void *mgh) { auto *h = new LBLHandle; h->l = std::make_unique<LBL>(static_cast<MGHandle *>(mgh)->m); return h; }
void vfo_lbl_destroy(void *h) { delete static_cast<LBLHandle *>(h); }
int vfo_lbl_select_init_method(void *h, const char *m) { VFO_TRY LB(h).selectInitMethod(m); VFO_CATCH }
int vfo_lbl_run(void *h, int zeroInit, int64_t inc, int maxIter, double tol, int mgIt, int mgSmooth, int fmg) { VFO_TRY LB(h).run(zeroInit != 0, inc, maxIter, tol, mgIt, mgSmooth, fmg != 0); VFO_CATCH }
double vfo_lbl_objective(void *h) { return 0.5 * LB(h).totalCompliance / double(LB(h).layersAccumulated); }
void vfo_lbl_gradient(void *h, double *g) { LBL &l = LB(h); for (size_t i = 0; i < l.totalGrad.size(); ++i) g[i] = l.totalGrad[i] / double(l.layersAccumulated); }
int vfo_lbl_num_layers_run(void *h) { return int(LB(h).layerIters.size()); }
void vfo_lbl_layer_info(void *h, int32_t *iters, double *compliance) { LBL &l = LB(h); for (size_t i = 0; i < l.layerIters.size(); ++i) { iters[i] = l.layerIters[i]; compliance[i] = l.layerCompliance[i]; } }
// ---- MMA ----
void *vfo_mma_create(int n, int m, const double *xmin, const double *xmax, mma_f_cb f, mma_df_cb df, void *user) { return new MMA(n, m, xmin, xmax, f, df, user); }
void vfo_mma_destroy(void *h) { delete static_cast<MMA *>(h); }
void vfo_mma_enable_gcmma(void *h, int e) { static_cast<MMA *>(h)->enableInner = e != 0; }
void vfo_mma_set_initial_var(void *h, const double *x) { static_cast<MMA *>(h)->setInitialVar(x); }
int vfo_mma_step(void *h) { VFO_TRY static_cast<MMA *>(h)->step(); VFO_CATCH }
void vfo_mma_get_optimal_var(void *h, double *x) { MMA &M_ = *static_cast<MMA *>(h); std::copy(M_.xhist[0].begin(), M_.xhist[0].end(), x); }
int64_t vfo_mma_newton_iterations(void *h) { return static_cast<MMA *>(h)->subsolveNewtonIters; }
}
