// vf_api.cu -- host side of libvoxelfem_b200: the C ABI of include/voxelfem_b200.h.
//
// Owns all device state (densities, moduli, per-level fields and stencils, the dense coarse
// inverse) and drives the kernels of vf_l0.cu / vf_stencil.cu / vf_vec.cu / vf_top.cu on one CUDA
// stream per simulator.  Control flow follows MultigridSolver.hh (vcycle :617-658, fullMultigrid
// :587-609, solve :546-573, preconditionedConjugateGradient :1047-1152) and
// TensorProductSimulator.hh; each function cites the lines it restates.
#include "vf_internal.cuh"
#include "../../include/voxelfem_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace vf {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_last_error(const std::string &msg) { g_err = msg; }

// ---------------------------------------------------------------------------
// Profiler: CUDA-event pairs around each launch, resolved lazily
// ---------------------------------------------------------------------------
struct Profiler {
    bool enabled = false;
    struct Pending { int cat; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    long long launches[PC_COUNT] = {0};
    double ms[PC_COUNT] = {0}, units[PC_COUNT] = {0};
    cudaEvent_t curStart = nullptr;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; VF_CUDA(cudaEventCreate(&e)); return e;
    }
    void resolve() {
        for (auto &p : pending) {
            cudaEventSynchronize(p.b);
            float t = 0; cudaEventElapsedTime(&t, p.a, p.b);
            ms[p.cat] += t; pool.push_back(p.a); pool.push_back(p.b);
        }
        pending.clear();
    }
    void reset() { resolve(); for (int i = 0; i < PC_COUNT; ++i) { launches[i] = 0; ms[i] = 0; units[i] = 0; } }
    ~Profiler() { for (auto &p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); } for (auto e : pool) cudaEventDestroy(e); }
};
void prof_begin(const LaunchCtx &ctx, int cat, double units) {
    Profiler *p = ctx.prof;
    if (!p || !p->enabled) return;
    p->curStart = p->get();
    cudaEventRecord(p->curStart, ctx.stream);
    p->launches[cat]++; p->units[cat] += units;
}
void prof_end(const LaunchCtx &ctx, int cat) {
    Profiler *p = ctx.prof;
    if (!p || !p->enabled || !p->curStart) return;
    cudaEvent_t b = p->get();
    cudaEventRecord(b, ctx.stream);
    p->pending.push_back({cat, p->curStart, b});
    p->curStart = nullptr;
    if (p->pending.size() > 8192) p->resolve();
}
static const char *kProfNames[PC_COUNT] = {"apply_l0", "residual_l0", "gs_l0", "apply_stencil", "residual_stencil", "gs_stencil",
                                           "restrict", "prolong", "coarse_solve", "vector_ops", "coarsen", "topopt", "other",
                                           "apply_stencil_small", "residual_stencil_small", "gs_stencil_small"};

// ---------------------------------------------------------------------------
// Small RAII device buffer
// ---------------------------------------------------------------------------
template<class T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete; DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    // The solver streams are non-blocking, i.e. NOT ordered against the legacy default stream the memset runs on: the memset is
    // completed on the host side before anything else can be enqueued that touches the buffer.
    static void zero_now(void *q, size_t bytes) { VF_CUDA(cudaMemset(q, 0, bytes)); VF_CUDA(cudaStreamSynchronize(0)); }
    void alloc(size_t count, bool zero = true) {
        if (count == n && p) { if (zero) zero_now(p, n * sizeof(T)); return; }
        release();
        if (count == 0) return;
        VF_CUDA(cudaMalloc(&p, count * sizeof(T))); n = count;
        if (zero) zero_now(p, count * sizeof(T));
    }
    void upload(const T *h, size_t count, cudaStream_t s) { VF_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
    void download(T *h, size_t count, cudaStream_t s) const { VF_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s)); VF_CUDA(cudaStreamSynchronize(s)); }
};

static void ensure_device() {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) throw std::runtime_error("voxelfem_b200: no usable CUDA device (this library has no CPU fallback)");
}

static GridDesc make_grid(int N, const int64_t *ne) {
    GridDesc g; std::memset(&g, 0, sizeof(g));
    g.N = N;
    if (N == 3) { for (int d = 0; d < 3; ++d) { g.ne[d] = (int)ne[d]; g.nn[d] = (int)ne[d] + 1; } g.bd = 1; }
    else { g.ne[0] = 1; g.nn[0] = 1; g.ne[1] = (int)ne[0]; g.nn[1] = (int)ne[0] + 1; g.ne[2] = (int)ne[1]; g.nn[2] = (int)ne[1] + 1; g.bd = 2; }
    g.ns[2] = 1; g.ns[1] = g.nn[2]; g.ns[0] = (long long)g.nn[1] * g.nn[2];
    g.es[2] = 1; g.es[1] = g.ne[2]; g.es[0] = (long long)g.ne[1] * g.ne[2];
    g.numNodes = (long long)g.nn[0] * g.nn[1] * g.nn[2];
    g.numElems = (long long)g.ne[0] * g.ne[1] * g.ne[2];
    g.nActive = g.nn[g.bd]; g.neActive = g.ne[g.bd];
    long long base = 0;
    for (int c = 0; c < 8; ++c) {
        long long cnt = 1;
        for (int a = 0; a < 3; ++a) {
            const int off = (c >> (2 - a)) & 1;
            g.ccnt[c][a] = (g.nn[a] - 1 - off >= 0) ? (g.nn[a] - 1 - off) / 2 + 1 : 0;
            cnt *= g.ccnt[c][a];
        }
        g.cbase[c] = base; base += (cnt + kStencilTile - 1) / kStencilTile * kStencilTile;
    }
    g.numPos = base;
    g.xoff = 0; g.ownLo = 0; g.ownHi = g.nn[0]; g.cmpLo = 0; g.cmpHi = g.nn[0]; g.oeLo = 0; g.oeHi = g.ne[0];
    return g;
}
static void set_mask_limits(GridDesc &g, int firstMasked, int firstDetached) {
    g.nActive = std::min<long long>(firstDetached, g.nn[g.bd]);
    g.neActive = std::min<long long>(firstMasked, g.ne[g.bd]);
}

// ---------------------------------------------------------------------------
// Dense SPD solver on the GPU (stands in for CHOLMOD, TensorProductSimulator.hh:1198-1230): Cholesky factor L and the explicit
// inverse of the TRIANGULAR FACTOR -- not the full inverse.  A solve is two bandwidth-bound triangular mat-vecs
// x = L^-T (L^-1 b) on a matrix that stays L2-resident, so the coarse solve on the V-cycle's critical path is two small kernels
// with no dependent-block latency chain.  The factorization runs on this library's own kernels (vf_dense.cu: a tiled FP64 matrix
// product and a single-block Cholesky + inversion of 64 x 64 diagonal blocks); no cuSOLVER / cuBLAS.
// ---------------------------------------------------------------------------
struct DenseSolver {
    DevBuf<double> A, W, Lb, Lip, Tbuf, work, rhs, y; DevBuf<int> info, red, freeDofs;
    int nfree = 0; bool ok = false;
    // fixed: per-DOF flags in (node*N + c) order
    void factor(const LaunchCtx &ctx, const GridDesc &g, const double *S, const std::vector<uint8_t> &fixed) {
        const long long ndof = g.numNodes * g.N;
        // Free DOFs are numbered plane by plane along the longest grid axis: nodes of non-adjacent planes do not couple, so the
        // matrix of the free DOFs is block tridiagonal with one block per node plane (blockOff).
        int ax = 3 - g.N;
        for (int a = 3 - g.N; a < 3; ++a) if (g.nn[a] > g.nn[ax]) ax = a;
        std::vector<int> redH(ndof, -1), freeH, blockOff(1, 0);
        for (int p = 0; p < g.nn[ax]; ++p) {
            for (long long n = 0; n < g.numNodes; ++n) {
                const int c = (int)((n / g.ns[ax]) % g.nn[ax]);
                if (c != p) continue;
                for (int k = 0; k < g.N; ++k) { const long long i = n * g.N + k; if (!fixed[i]) { redH[i] = (int)freeH.size(); freeH.push_back((int)i); } }
            }
            blockOff.push_back((int)freeH.size());
        }
        nfree = (int)freeH.size();
        red.alloc(ndof, false); red.upload(redH.data(), ndof, ctx.stream);
        freeDofs.alloc(std::max(nfree, 1), false); if (nfree) freeDofs.upload(freeH.data(), nfree, ctx.stream);
        VF_CUDA(cudaStreamSynchronize(ctx.stream)); // redH / freeH are stack-owned
        if (nfree == 0) { ok = true; return; }
        if (A.n != (size_t)nfree * nfree) A.alloc((size_t)nfree * nfree, false);   // A.p stays put: captured CUDA graphs hold it
        if (W.n != (size_t)nfree * nfree) W.alloc((size_t)nfree * nfree, false);
        VF_CUDA(cudaMemsetAsync(A.p, 0, sizeof(double) * A.n, ctx.stream));       // receives L^-1: zeros above the diagonal
        VF_CUDA(cudaMemsetAsync(W.p, 0, sizeof(double) * W.n, ctx.stream));
        rhs.alloc(nfree, true); y.alloc(nfree, true);
        info.alloc(1, true);
        int maxBlock = 0, nBlocks = 0;
        for (size_t b = 0; b + 1 < blockOff.size(); ++b) { const int m = blockOff[b + 1] - blockOff[b]; maxBlock = std::max(maxBlock, m); nBlocks += m > 0; }
        static const bool noBlockTri = [] { const char *e = std::getenv("VF_COARSE_DENSE"); return e && e[0] == '1'; }();
        launch_stencil_to_dense(ctx, g, S, red.p, nfree, W.p);
        if (!noBlockTri && nBlocks >= 3 && maxBlock >= 96) factor_block_tridiagonal(ctx, blockOff, maxBlock);
        else factor_dense(ctx);
        check_info(ctx);
        // The factorization is written in column-major terms: L^-1(r, c), r >= c, sits at A[c * n + r] -- in our row-major reading
        // that is row c, column r: the upper triangle, i.e. L^-T.  Mirror it so that the lower triangle holds L^-1 row by row.
        launch_symmetrize_upper_to_lower(ctx);
        ok = true;
    }
    void check_info(const LaunchCtx &ctx) {
        int h = 0; info.download(&h, 1, ctx.stream);
        if (h != 0) throw std::runtime_error("Cholesky factorization failed: coarse stiffness matrix is not positive definite (pivot " + std::to_string(h - 1) + ")");
    }
    // Dense path (small or unstructured coarse grids): blocked Cholesky + inversion of the whole matrix.
    void factor_dense(const LaunchCtx &ctx) {
        const int n = nfree;
        if (Lb.n < (size_t)n * n) Lb.alloc((size_t)n * n, false);
        if (work.n < (size_t)kDiagBlockHost * n) work.alloc((size_t)kDiagBlockHost * n, false);
        potrf_inv_blocked(ctx, W.p, n, n, Lb.p, n, A.p, n, work.p, info.p, 0);
    }
    // Block-tridiagonal path: block Cholesky
    //   L_ii L_ii^T = A_ii - L_ip L_ip^T,  L_ip = A_ip L_pp^-T   (p = i - 1)
    // followed by the block forward substitution L X = I, X_ii = L_ii^-1, X_i,: = -L_ii^-1 L_ip X_p,: , which leaves the dense
    // lower-triangular L^-1 that solve() applies as two bandwidth-bound triangular mat-vecs.  n bw^2 + n^2 bw flops in
    // matrix products instead of the 2/3 n^3 of a dense factorization + inversion.
    // The factor chain (L_ip, Schur update, factorization + inversion of the diagonal block) is sequential over the blocks; the block
    // rows of L^-1 left of the diagonal only feed the NEXT block row, so they run on a side stream behind the chain (forked / joined
    // with events; L_ip and the product buffer are double-buffered between the two streams).
    cudaStream_t side = nullptr; cudaEvent_t evChain[2] = {nullptr, nullptr}, evRow[2] = {nullptr, nullptr}, evJoin = nullptr;
    ~DenseSolver() {
        for (cudaEvent_t e : {evChain[0], evChain[1], evRow[0], evRow[1], evJoin}) if (e) cudaEventDestroy(e);
        if (side) cudaStreamDestroy(side);
    }
    void factor_block_tridiagonal(const LaunchCtx &ctx, const std::vector<int> &off, int maxBlock) {
        const int n = nfree, nb = (int)off.size() - 1;   // W: the matrix (lower triangle), A: receives L^-1
        auto M = [&](int i, int j) { return W.p + (size_t)off[j] * n + off[i]; };   // block (i, j), leading dimension n
        auto X = [&](int i, int j) { return A.p + (size_t)off[j] * n + off[i]; };
        if (Lip.n < (size_t)2 * maxBlock * maxBlock) Lip.alloc((size_t)2 * maxBlock * maxBlock, false);
        if (Lb.n < (size_t)maxBlock * maxBlock) Lb.alloc((size_t)maxBlock * maxBlock, false);
        if (Tbuf.n < (size_t)maxBlock * n) Tbuf.alloc((size_t)maxBlock * n, false);
        if (work.n < (size_t)kDiagBlockHost * maxBlock) work.alloc((size_t)kDiagBlockHost * maxBlock, false);
        if (!side) {
            VF_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
            for (cudaEvent_t *e : {&evChain[0], &evChain[1], &evRow[0], &evRow[1], &evJoin}) VF_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        LaunchCtx sctx = ctx; sctx.stream = side; sctx.prof = nullptr;
        // the side stream starts behind everything the chain's stream has done so far (the memsets of A, the assembly of W)
        VF_CUDA(cudaEventRecord(evJoin, ctx.stream)); VF_CUDA(cudaStreamWaitEvent(side, evJoin, 0));
        int prev = -1, step = 0; bool rowPending[2] = {false, false};
        for (int i = 0; i < nb; ++i) {
            const int m = off[i + 1] - off[i];
            if (m == 0) continue;
            const bool coupled = prev == i - 1 && prev >= 0;
            const int mp = coupled ? off[prev + 1] - off[prev] : 0;
            const int slot = step & 1;
            double *lip = Lip.p + (size_t)slot * maxBlock * maxBlock;
            if (coupled) {   // L_ip = A_ip L_pp^-T = A_ip X_pp^T;  A_ii -= L_ip L_ip^T
                if (rowPending[slot]) { VF_CUDA(cudaStreamWaitEvent(ctx.stream, evRow[slot], 0)); rowPending[slot] = false; }   // the row update two steps back still reads this L_ip buffer
                launch_dgemm(ctx, true, m, mp, mp, 1.0, M(i, prev), n, X(prev, prev), n, 0.0, lip, m, false);
                launch_dgemm(ctx, true, m, m, mp, -1.0, lip, m, lip, m, 1.0, M(i, i), n, true);
            }
            potrf_inv_blocked(ctx, M(i, i), n, m, Lb.p, maxBlock, X(i, i), n, work.p, info.p, off[i]);
            if (coupled && off[i] > 0) {   // the block row left of the diagonal: X_i,: = -X_ii (L_ip X_p,:)
                const int w = off[i];
                VF_CUDA(cudaEventRecord(evChain[slot], ctx.stream)); VF_CUDA(cudaStreamWaitEvent(side, evChain[slot], 0));
                launch_dgemm(sctx, false, m, w, mp, 1.0, lip, m, X(prev, 0), n, 0.0, Tbuf.p, m, false);
                launch_dgemm(sctx, false, m, w, m, -1.0, X(i, i), n, Tbuf.p, m, 0.0, X(i, 0), n, false);
                VF_CUDA(cudaEventRecord(evRow[slot], side)); rowPending[slot] = true;
            }
            prev = i; ++step;
        }
        VF_CUDA(cudaEventRecord(evJoin, side)); VF_CUDA(cudaStreamWaitEvent(ctx.stream, evJoin, 0));
    }
    void launch_symmetrize_upper_to_lower(const LaunchCtx &ctx);
    // x = A^-1 f on free DOFs, zero on fixed DOFs (TensorProductSimulator.hh:1227-1229, 1243-1252)
    void solve(const LaunchCtx &ctx, const GridDesc &g, const double *f, double *x) {
        VF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * g.numNodes * g.N, ctx.stream));
        if (nfree == 0) return;
        launch_gather_free(ctx, f, freeDofs.p, nfree, g.numNodes, g.N, rhs.p);
        launch_dense_trmv(ctx, A.p, nfree, rhs.p, y.p, true);    // y = L^-1 rhs   (rows of the lower triangle)
        launch_dense_trmv(ctx, A.p, nfree, y.p, rhs.p, false);   // z = L^-T y     (rows of the upper triangle)
        launch_scatter_free(ctx, rhs.p, freeDofs.p, nfree, g.numNodes, g.N, x);
    }
};

} // namespace vf

// kernels local to this file ----------------------------------------------------------------
__global__ void k_sym_upper_to_lower(double *A, int n) { // row-major: copy A[i][j] (j > i) into A[j][i]
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < n && j < n && j > i) A[(size_t)j * n + i] = A[(size_t)i * n + j];
}
namespace vf {
void DenseSolver::launch_symmetrize_upper_to_lower(const LaunchCtx &ctx) {
    // cuSOLVER (column-major, FILL_MODE_LOWER) holds element (r, c), r >= c, at A[c * n + r]; in our row-major reading
    // that is row c, column r >= c: the upper triangle.  Mirror it into the lower triangle.
    ProfScope ps(ctx, PC_COARSEN, (double)nfree * nfree);
    dim3 block(32, 8), grid((nfree + 31) / 32, (nfree + 7) / 8);
    k_sym_upper_to_lower<<<grid, block, 0, ctx.stream>>>(A.p, nfree);
    VF_KERNEL_CHECK();
}
} // namespace vf

using namespace vf;

// ---------------------------------------------------------------------------
// vf_sim
// ---------------------------------------------------------------------------
struct vf_sim {
    int N = 3;
    int64_t ne[3] = {1, 1, 1}, nn[3] = {1, 1, 1};
    double dmin[3] = {0, 0, 0}, dmax[3] = {1, 1, 1}, stretch[3] = {1, 1, 1}, spacing[3] = {1, 1, 1};
    GridDesc g;
    double D[6][6];
    std::vector<double> K0; K0Param K0p; DevBuf<double> K0dev;
    int law = VF_LAW_SIMP; double E0 = 1, Emin = 1e-4, gamma = 3, q = 3;
    double gravity[3] = {0, 0, 0};
    std::vector<int64_t> dirNodes; std::vector<uint8_t> dirMask; std::vector<double> dirVals; std::vector<uint8_t> nodeMask;
    std::vector<int64_t> forceNodes; std::vector<double> forceVals;
    DevBuf<double> rho, E, dirValsDev, scratch, scalars, tmpU, tmpV;
    DevBuf<uint8_t> dmaskDev, dirMaskDev; DevBuf<long long> dirNodesDev;
    double maskHeight = std::numeric_limits<double>::infinity();
    int firstMasked = INT_MAX, firstDetached = INT_MAX;
    cudaStream_t stream = nullptr; LaunchCtx ctx; Profiler prof;
    uint64_t version = 1; // bumped whenever E, the mask, K0 or the Dirichlet set changes
    uint64_t structVersion = 1; // bumped when K0, the Dirichlet set or the mask change (not on density updates): keys captured CUDA graphs
    DenseSolver direct; uint64_t directVersion = 0; DevBuf<double> directStencil;
    // Slab window (multi-GPU): this simulator stores node planes [xoff, xoff + nn[0] - 1] of a grid with gne0 element layers
    // along axis 0, of which element layers [slabBegin, slabEnd) are owned.  dmin/dmax/spacing/stretch are the global ones.
    bool window = false, ownsStream = true; int64_t gne0 = 0, xoff = 0, slabBegin = 0, slabEnd = 0;

    long long numNodes() const { return g.numNodes; }
    long long numElems() const { return g.numElems; }
    double elemVolume() const { double v = 1; for (int d = 0; d < N; ++d) v *= stretch[d]; return v; }
    bool maskActive() const { return maskHeight < dmax[1]; }
    void touch() { ++version; }

    int symIdx(int i, int j) const { if (i == j) return i; if (N == 2) return 2; return 6 - i - j; }
    // Element_T::Stiffness (:67-80) with Strains::getStrains (TensorProductPolynomialInterpolant.hh:204-231) and the
    // 2-point Gauss rule on [0,1] (TensorProductQuadrature.hh:134-143); m_updateK0 (:2078-2086)
    void updateK0() {
        const int npe = 1 << N, ke = N * npe, fl = (N == 3) ? 6 : 3, nq = 1 << N;
        K0.assign((size_t)ke * ke, 0.0);
        const double gp[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
        std::vector<double> strain((size_t)ke * fl);
        for (int qi = 0; qi < nq; ++qi) {
            double xi[3] = {0, 0, 0};
            for (int d = 0; d < N; ++d) xi[d] = gp[(qi >> (N - 1 - d)) & 1];
            const double w = 1.0 / nq;
            for (int j = 0; j < npe; ++j) {
                double gr[3] = {0, 0, 0};
                for (int c = 0; c < N; ++c) {
                    double v = 1;
                    for (int d = 0; d < N; ++d) {
                        const int bit = (j >> (N - 1 - d)) & 1;
                        v *= (d == c) ? ((bit ? 1.0 : -1.0) / stretch[d]) : (bit ? xi[d] : 1.0 - xi[d]);
                    }
                    gr[c] = v;
                }
                for (int c = 0; c < N; ++c) {
                    double *s = &strain[(size_t)(j * N + c) * fl];
                    for (int t = 0; t < fl; ++t) s[t] = 0;
                    for (int i = 0; i < N; ++i) s[symIdx(c, i)] = 0.5 * gr[i];
                    s[symIdx(c, c)] = gr[c];
                }
            }
            for (int a = 0; a < ke; ++a) for (int b = a; b < ke; ++b) {
                const double *sa = &strain[(size_t)a * fl], *sb = &strain[(size_t)b * fl];
                double acc = 0;
                for (int i = 0; i < fl; ++i) {
                    double sig = 0;
                    for (int j = 0; j < fl; ++j) sig += D[i][j] * (j >= N ? 2.0 : 1.0) * sb[j]; // doubleContract with shear doubling
                    acc += (i >= N ? 2.0 : 1.0) * sa[i] * sig;
                }
                K0[(size_t)a * ke + b] += w * acc;
            }
        }
        const double vol = elemVolume();
        for (int a = 0; a < ke; ++a) for (int b = a; b < ke; ++b) { K0[(size_t)a * ke + b] *= vol; K0[(size_t)b * ke + a] = K0[(size_t)a * ke + b]; }
        std::memset(&K0p, 0, sizeof(K0p));
        std::copy(K0.begin(), K0.end(), K0p.v);
        finalize_k0_param(K0p, N);
        K0dev.alloc(K0.size(), false); K0dev.upload(K0.data(), K0.size(), stream);
        VF_CUDA(cudaStreamSynchronize(stream));
        touch(); ++structVersion;
    }
    void setIsotropic(double Ey, double nu) {
        double lambda = (nu * Ey) / ((1.0 + nu) * (1.0 - 2.0 * nu));
        const double mu = Ey / (2.0 + 2.0 * nu);
        if (N == 2) lambda = (nu * Ey) / (1.0 - nu * nu);
        std::memset(D, 0, sizeof(D));
        if (N == 3) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) D[i][j] = lambda; for (int i = 0; i < 3; ++i) D[i][i] = lambda + 2 * mu; D[3][3] = D[4][4] = D[5][5] = mu; }
        else { D[0][0] = D[1][1] = lambda + 2 * mu; D[0][1] = D[1][0] = lambda; D[2][2] = mu; }
        updateK0();
    }
    void refreshGridMask() { set_mask_limits(g, firstMasked, firstDetached); }
    void updateModuli() { // m_updateYoungModuli (:2088-2102)
        launch_update_moduli(ctx, g, rho.p, E.p, law, E0, Emin, gamma, q, maskActive());
        touch();
    }
    void uploadBCs() {
        dmaskDev.alloc(g.numNodes, false); dmaskDev.upload(nodeMask.data(), g.numNodes, stream);
        const size_t nd = dirNodes.size();
        dirNodesDev.alloc(std::max<size_t>(nd, 1), false); dirMaskDev.alloc(std::max<size_t>(nd, 1), false); dirValsDev.alloc(std::max<size_t>(nd * N, 1), false);
        if (nd) {
            std::vector<long long> tmp(dirNodes.begin(), dirNodes.end());
            dirNodesDev.upload(tmp.data(), nd, stream); dirMaskDev.upload(dirMask.data(), nd, stream); dirValsDev.upload(dirVals.data(), nd * N, stream);
            VF_CUDA(cudaStreamSynchronize(stream));
        }
        VF_CUDA(cudaStreamSynchronize(stream));
        touch(); ++structVersion;
    }
    // node index ranges per axis covered by an inclusive box (Geometry.hh:276-279 applied to nodePosition, :349-351)
    // idx: local indices inside this simulator's window; *globalCount (optional): number of nodes of the whole grid in the box
    bool boxRanges(const double *lo, const double *hi, std::vector<int64_t> (&idx)[3], double *globalCount = nullptr) const {
        double cnt = 1;
        for (int d = 0; d < N; ++d) {
            idx[d].clear();
            const int64_t gn = (d == 0 && window) ? gne0 + 1 : nn[d], off = (d == 0 && window) ? xoff : 0;
            int64_t c = 0;
            for (int64_t i = 0; i < gn; ++i) {
                const double p = dmin[d] + double(i) * spacing[d];
                if (p >= lo[d] && p <= hi[d]) { ++c; if (i - off >= 0 && i - off < nn[d]) idx[d].push_back(i - off); }
            }
            if (c == 0) return false;
            cnt *= double(c);
        }
        for (int d = N; d < 3; ++d) idx[d].assign(1, 0);
        if (globalCount) *globalCount = cnt;
        return true;
    }
    int64_t flatNode(int64_t i, int64_t j, int64_t k) const { return N == 3 ? (i * nn[1] + j) * nn[2] + k : i * nn[1] + j; }
};

// BCBuilder (TensorProductSimulator.hh:464-566): dense staging of the sparse BC lists
struct BCBuilder {
    vf_sim &s; std::vector<double> forces, dvals; std::vector<uint8_t> dmask;
    explicit BCBuilder(vf_sim &s_) : s(s_), forces((size_t)s_.g.numNodes * s_.N, 0.0), dvals((size_t)s_.g.numNodes * s_.N, 0.0), dmask(s_.g.numNodes, 0) {
        for (size_t f = 0; f < s.forceNodes.size(); ++f) for (int c = 0; c < s.N; ++c) forces[s.forceNodes[f] * s.N + c] = s.forceVals[f * s.N + c];
        for (size_t k = 0; k < s.dirNodes.size(); ++k) { for (int c = 0; c < s.N; ++c) dvals[s.dirNodes[k] * s.N + c] = s.dirVals[k * s.N + c]; dmask[s.dirNodes[k]] = s.dirMask[k]; }
    }
    void setDirichlet(int64_t ni, const double *val, unsigned cm) {
        for (int c = 0; c < s.N; ++c) {
            if (!((cm >> c) & 1u)) continue;
            if (!((dmask[ni] >> c) & 1u)) { dmask[ni] |= uint8_t(1u << c); dvals[ni * s.N + c] = val[c]; }
            else if (std::abs(dvals[ni * s.N + c] - val[c]) > 1e-10) throw std::runtime_error("Conflicting dirichlet displacements.");
        }
    }
    void setDirichletComponent(int64_t ni, int d, double v) { dmask[ni] |= uint8_t(1u << d); dvals[ni * s.N + d] = v; }
    void setForce(int64_t ni, const double *f) { for (int c = 0; c < s.N; ++c) forces[ni * s.N + c] = f[c]; }
    void apply() {
        s.dirNodes.clear(); s.dirMask.clear(); s.dirVals.clear();
        const uint8_t full = uint8_t((1u << s.N) - 1u);
        s.nodeMask.assign(s.g.numNodes, 0);
        for (int64_t ni = 0; ni < s.g.numNodes; ++ni) {
            const uint8_t m = dmask[ni] & full;
            if (m) { s.dirNodes.push_back(ni); s.dirMask.push_back(m); for (int c = 0; c < s.N; ++c) s.dirVals.push_back(dvals[ni * s.N + c]); s.nodeMask[ni] = m; }
        }
        s.forceNodes.clear(); s.forceVals.clear();
        for (int64_t ni = 0; ni < s.g.numNodes; ++ni) {
            double sq = 0; for (int c = 0; c < s.N; ++c) sq += forces[ni * s.N + c] * forces[ni * s.N + c];
            if (sq != 0.0) { s.forceNodes.push_back(ni); for (int c = 0; c < s.N; ++c) s.forceVals.push_back(forces[ni * s.N + c]); }
        }
        s.uploadBCs();
    }
};

// ---------------------------------------------------------------------------
// vf_mg
// ---------------------------------------------------------------------------
struct vf_group;
struct MGLevel {
    GridDesc g; int64_t ne[3]; double stretchBD = 1;
    // slab windows: this part owns global element layers [sb, se) of the level (node planes sb..se); replicated levels hold the whole grid
    int64_t sb = 0, se = 0, gne0 = 0; bool windowed = false;
    std::vector<uint8_t> nodeMask; DevBuf<uint8_t> dmask;
    DevBuf<double> x, b, r, S;
    DevBuf<unsigned long long> posTab;   // latency-bound levels >= 1: position -> node coordinates for the tile kernels
    int firstMasked = INT_MAX, firstDetached = INT_MAX;
};
struct vf_mg {
    vf_sim *sim = nullptr; int N = 3;
    std::vector<std::unique_ptr<MGLevel>> lv;
    std::vector<double> cK0; DevBuf<double> cK0dev; // [fi][KE][KE]
    bool symmetricGS = true;
    DenseSolver coarse;
    uint64_t stiffnessVersion = 0; // sim->version the coarse operators were built for
    // Pending banded update: since the hierarchy was last built, only the moduli of fine element layers [bandLo, bandHi) along
    // the build direction changed (mask decrements), and sim->version == bandVersion (MultigridSolver.hh:907-1017).
    bool bandActive = false; int bandLo = 0, bandHi = 0; uint64_t bandVersion = 0;
    DevBuf<double> coarsenScratch;   // intermediates of the separable Galerkin product
    DevBuf<unsigned> gridBar;        // arrival counter + generation of the persistent small-level sweep (k_stencil_sweep)
    DevBuf<double> Ad, d, scalars, scratch, tmpA, tmpB, tmpC;
    double *hostScalars = nullptr; // pinned
    std::vector<double> lastResiduals; int lastIters = 0; const double *pcgX = nullptr; // device iterate of the running / last PCG
    LaunchCtx ctx;
    vf_group *grp = nullptr; int firstRep = INT_MAX; // slab group this solver is a part of; first replicated (non-windowed) level
    std::vector<vf_mg *> self;
    DevBuf<double> stage[2];                          // staging buffers of the slab exchanges
    // The preconditioner application (one FMG / V-cycle: a fixed sequence of ~350 launches, most of them tiny coarse-level
    // kernels) is captured once into a CUDA graph and replayed every PCG iteration.
    struct PrecondGraph { cudaGraphExec_t exec = nullptr; uint64_t version = 0; int nActive = -1, mgIt = 0, nsmooth = 0; bool fmg = false, sym = true; long long launches = 0; } pg, sg;   // sg: V-cycle over the replicated levels of an NCCL rank
    bool inSubCapture = false;
    bool useGraphs = true;
    // Reference-literal mode (vf_mg_set_rebuild_every_solve): every PCG call rebuilds the coarse hierarchy (MultigridSolver.hh:1104-1107) instead
    // of only when the moduli changed.  rebuiltForThisSolve: vf_mg_pcg_io already did it, overlapped with its host -> device copies.
    bool rebuildEverySolve = false, rebuiltForThisSolve = false;
    cudaStream_t ioStream = nullptr; cudaEvent_t ioEvA = nullptr, ioEvB = nullptr;   // copy stream of the host-buffer entry points
    ~vf_mg() {
        if (hostScalars) cudaFreeHost(hostScalars); if (pg.exec) cudaGraphExecDestroy(pg.exec); if (sg.exec) cudaGraphExecDestroy(sg.exec);
        if (ioEvA) cudaEventDestroy(ioEvA); if (ioEvB) cudaEventDestroy(ioEvB); if (ioStream) cudaStreamDestroy(ioStream);
    }
    int numLevels() const { return (int)lv.size(); }
    const uint8_t *dmask(int l) const { return l == 0 ? sim->dmaskDev.p : lv[l]->dmask.p; }
    const GridDesc &grid(int l) const { return l == 0 ? sim->g : lv[l]->g; }
};

// ---------------------------------------------------------------------------
// Slab groups: the grid is cut into slabs along axis 0, one vf_mg "part" per slab.  A group is either
//   - local: all parts live in this process on one device and one stream (exchanges are device copies) -- used to test the
//     partitioned algorithm on a single GPU and to drive several slabs from one process, or
//   - NCCL:  one part per process / GPU; halo planes travel with ncclSend/ncclRecv, scalars and the replicated coarse
//     data with ncclAllReduce, all enqueued on the solver's stream.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded, else the system one).
// ---------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (nccl.h)
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    static NcclApi &get() {
        static NcclApi api;
        if (!api.lib) {
            api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!api.lib) throw std::runtime_error(std::string("cannot load NCCL: ") + dlerror());
#define VF_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name)); if (!api.field) throw std::runtime_error("NCCL symbol missing: " name);
            VF_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") VF_NCCL_SYM(CommInitRank, "ncclCommInitRank") VF_NCCL_SYM(CommDestroy, "ncclCommDestroy")
            VF_NCCL_SYM(Send, "ncclSend") VF_NCCL_SYM(Recv, "ncclRecv") VF_NCCL_SYM(AllReduce, "ncclAllReduce")
            VF_NCCL_SYM(GroupStart, "ncclGroupStart") VF_NCCL_SYM(GroupEnd, "ncclGroupEnd") VF_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef VF_NCCL_SYM
        }
        return api;
    }
    void check(int rc, const char *what) { if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "?")); }
};
constexpr int kNcclDouble = 8, kNcclSum = 0; // ncclFloat64, ncclSum (nccl.h)

// ---------------------------------------------------------------------------------------------------------------------------------
// Device-initiated ghost-plane exchange between the ranks of an NCCL group (one process per GPU on one NVSwitch box).
//
// Every rank owns, per neighbour, a ring of kP2PSlots MAILBOXES in its own HBM plus two words: `ready` (written by the neighbour:
// sequence number of the last message it deposited) and `consumed` (written by the neighbour: sequence number of the last message
// of MINE it has unpacked).  Mailboxes and words are exported with CUDA IPC, so the neighbour's kernels store into them directly
// over NVLink.  An exchange is ONE kernel (k_p2p_exchange) on the solver's stream and NO host-side communication call; its blocks play
// up to four roles:
//   send:       waits (rarely) until the slot it is about to overwrite has been consumed, packs the boundary plane straight into the
//               neighbour's mailbox, fences at system scope and publishes `ready = seq`;
//   receive:    spins on the local `ready` word, copies the mailbox into the ghost plane and publishes `consumed = seq` to the sender.
// Sequence numbers live in device memory, so the kernels replay unchanged from a captured CUDA graph.  A spin that lasts longer
// than kP2PTimeoutNs sets an error word instead of hanging the device.  NCCL remains the transport of the all-reduces (PCG
// scalars, replicated coarse fields) and the fallback of the exchange (VF_P2P=0).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int kP2PSlots = 4;
constexpr long long kP2PTimeoutNs = 5LL * 1000 * 1000 * 1000;
struct P2PWords {                          // one cache line each: written by different agents
    unsigned long long ready[16];          // [0]: written by the NEIGHBOUR: last sequence number deposited in my mailbox ring
    unsigned long long consumed[16];       // [0]: written by the NEIGHBOUR: last sequence number of mine it unpacked
    unsigned long long sendSeq[16];        // [0]: local: messages sent so far on this link
    unsigned long long recvSeq[16];        // [0]: local: messages received so far
    unsigned long long sendDone[16];       // [0]: local: blocks of the running send kernel that finished their part
    unsigned long long recvDone[16];
    unsigned long long error[16];          // [0]: non-zero after a timed-out wait
};
struct P2PLink {
    bool active = false;
    double *myBox = nullptr; P2PWords *myWords = nullptr;        // local allocations (neighbour writes box, ready, consumed)
    double *peerBox = nullptr; P2PWords *peerWords = nullptr;    // the neighbour's, mapped through CUDA IPC
    size_t slotDoubles = 0;
};
struct vf_group {
    std::vector<vf_mg *> parts;   // local parts, ordered by slab
    int rank = 0, world = 1;      // NCCL mode: this process' slab index and the number of slabs
    void *comm = nullptr;         // ncclComm_t
    P2PLink link[2];              // NCCL mode: 0 = lower neighbour (rank - 1), 1 = upper neighbour (rank + 1)
    bool p2p = false;
    ~vf_group() {
        for (P2PLink &l : link) { if (l.peerBox) cudaIpcCloseMemHandle(l.peerBox); if (l.peerWords) cudaIpcCloseMemHandle(l.peerWords); if (l.myBox) cudaFree(l.myBox); if (l.myWords) cudaFree(l.myWords); }
        if (comm) NcclApi::get().CommDestroy(comm);
    }
};

namespace {

enum Scalar { SC_RMR_A = 0, SC_RMR_B, SC_DAD, SC_RSQ, SC_BSQ, SC_TMP, SC_COUNT = 8 };

void mg_sync_level_masks(vf_mg &mg) {
    // coarse simulators share the physical mask height (MultigridSolver.hh:1022-1036)
    for (int l = 1; l < mg.numLevels(); ++l) {
        MGLevel &L = *mg.lv[l];
        if (std::isinf(mg.sim->maskHeight)) { L.firstMasked = INT_MAX; L.firstDetached = INT_MAX; }
        else { L.firstMasked = (int)std::ceil(mg.sim->maskHeight / L.stretchBD - 1e-10); L.firstDetached = L.firstMasked + 1; } // TensorProductSimulator.hh:297-301
        set_mask_limits(L.g, L.firstMasked, L.firstDetached);
    }
}

std::vector<vf_mg *> &parts_of(vf_mg &mg) {
    if (mg.grp) return mg.grp->parts;
    if (mg.self.empty()) mg.self.assign(1, &mg);
    return mg.self;
}
MGLevel &level(vf_mg &mg, int l) { return *mg.lv[l]; }
bool level_windowed(vf_mg &mg, int l) { return mg.grp && l < mg.firstRep; }
double *stage_buf(vf_mg &mg, int which, size_t n) { if (mg.stage[which].n < n) mg.stage[which].alloc(n, false); return mg.stage[which].p; }

__global__ void k_sum_into_all(int np, double *p0, double *p1, double *p2, double *p3, double *p4, double *p5, double *p6, double *p7, long long n) {
    double *p[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < np; ++k) s += p[k][i];
        for (int k = 0; k < np; ++k) p[k][i] = s;
    }
}
// Sum a buffer (same length on every part) over all parts of the group; every part ends up with the total.
template<class Sel> void grp_allreduce(vf_mg &lead, Sel sel, size_t n) {
    if (!lead.grp || n == 0) return;
    vf_group &G = *lead.grp;
    if (G.comm) {
        NcclApi &A = NcclApi::get(); double *p = sel(*G.parts[0]);
        count_launch();
        A.check(A.AllReduce(p, p, n, kNcclDouble, kNcclSum, G.comm, lead.ctx.stream), "ncclAllReduce");
        return;
    }
    const int np = (int)G.parts.size();
    if (np < 2) return;
    if (np > 8) throw std::runtime_error("a local slab group holds at most 8 parts");
    double *p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < np; ++k) p[k] = sel(*G.parts[k]);
    count_launch();
    const long long blocks = std::min<long long>((long long)(n + 255) / 256, 148LL * 8);
    k_sum_into_all<<<(unsigned)blocks, 256, 0, lead.ctx.stream>>>(np, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], (long long)n);
    VF_KERNEL_CHECK();
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long global_timer_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// wait until *word >= target (system scope); returns false on timeout
__device__ __forceinline__ bool p2p_wait(const unsigned long long *word, unsigned long long target) {
    if (ld_acquire_sys(word) >= target) return true;
    const long long t0 = global_timer_ns();
    while (ld_acquire_sys(word) < target) { if (global_timer_ns() - t0 > kP2PTimeoutNs) return false; __nanosleep(200); }
    return true;
}
// One ghost-plane exchange of a rank = ONE kernel: its blocks are split into up to four roles (send to / receive from the lower and
// the upper neighbour).  The send blocks never wait for this rank's receive blocks (only, rarely, for a free mailbox slot), the
// receive blocks spin on the neighbour's `ready` word; all blocks of the launch are resident (<= 256 blocks), so no role starves.
struct P2PRole {
    double *field;            // send: first component plane to pack; receive: first ghost component plane to fill
    double *box;              // send: the neighbour's mailbox ring; receive: my mailbox ring
    P2PWords *mine, *peer;
    long long compStride, n; size_t slotDoubles; int ncomp, nblocks, recv;
};
struct P2PExchange { P2PRole role[4]; int nroles; };
// pack ncomp component planes (n doubles each, component stride compStride) into the neighbour's mailbox slot and publish it
__device__ __forceinline__ void p2p_send_body(const P2PRole &R, int blk) {
    __shared__ unsigned long long s_seq; __shared__ int s_ok;
    if (threadIdx.x == 0) {
        const unsigned long long seq = R.mine->sendSeq[0] + 1;
        s_seq = seq;
        // the slot is free once the neighbour has unpacked the message that used it last (kP2PSlots messages ago)
        s_ok = (seq <= (unsigned long long)kP2PSlots) || p2p_wait(&R.mine->consumed[0], seq - kP2PSlots);
        if (!s_ok) R.mine->error[0] = 1;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    double *dst = R.box + (size_t)(seq % kP2PSlots) * R.slotDoubles;
    if (s_ok) {
        const long long total = R.n * R.ncomp;
        for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < total; i += (long long)R.nblocks * blockDim.x) {
            const long long c = i / R.n, k = i - c * R.n;
            dst[i] = R.field[c * R.compStride + k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(&R.mine->sendDone[0], 1ULL) + 1;
        if (done == (unsigned long long)R.nblocks) {   // last block: everything of this message is visible system-wide
            R.mine->sendDone[0] = 0;
            __threadfence_system();
            st_release_sys(&R.peer->ready[0], seq);
            R.mine->sendSeq[0] = seq;
        }
    }
}
__device__ __forceinline__ void p2p_recv_body(const P2PRole &R, int blk) {
    __shared__ unsigned long long r_seq; __shared__ int r_ok;
    if (threadIdx.x == 0) {
        const unsigned long long seq = R.mine->recvSeq[0] + 1;
        r_seq = seq;
        r_ok = p2p_wait(&R.mine->ready[0], seq);
        if (!r_ok) R.mine->error[0] = 2;
    }
    __syncthreads();
    const unsigned long long seq = r_seq;
    const double *box = R.box + (size_t)(seq % kP2PSlots) * R.slotDoubles;
    if (r_ok) {
        const long long total = R.n * R.ncomp;
        for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < total; i += (long long)R.nblocks * blockDim.x) {
            const long long c = i / R.n, k = i - c * R.n;
            R.field[c * R.compStride + k] = __ldcg(box + i);     // written by the neighbour: bypass L1
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(&R.mine->recvDone[0], 1ULL) + 1;
        if (done == (unsigned long long)R.nblocks) {
            R.mine->recvDone[0] = 0;
            __threadfence_system();
            st_release_sys(&R.peer->consumed[0], seq);
            R.mine->recvSeq[0] = seq;
        }
    }
}
__global__ void __launch_bounds__(256) k_p2p_exchange(const __grid_constant__ P2PExchange X) {
    int blk = blockIdx.x;
    #pragma unroll 1
    for (int r = 0; r < X.nroles; ++r) {
        if (blk < X.role[r].nblocks) { if (X.role[r].recv) p2p_recv_body(X.role[r], blk); else p2p_send_body(X.role[r], blk); return; }
        blk -= X.role[r].nblocks;
    }
}
static int p2p_blocks(long long n, int ncomp) { return (int)std::min<long long>((n * ncomp + 255) / 256, 64); }
// sendL / sendR: first component plane to send to the lower / upper neighbour (nullptr: none); recvL / recvR: ghost planes to fill
static void p2p_exchange(vf_group &G, double *sendL, double *sendR, double *recvL, double *recvR, long long compStride, long long n, int ncomp, cudaStream_t stream) {
    P2PExchange X; X.nroles = 0;
    int total = 0;
    auto add = [&](int side, double *field, bool recv) {
        if (!field) return;
        P2PLink &L = G.link[side];
        if ((size_t)(n * ncomp) > L.slotDoubles) throw std::runtime_error("ghost plane larger than the peer mailbox");
        P2PRole &R = X.role[X.nroles++];
        R.field = field; R.box = recv ? L.myBox : L.peerBox; R.mine = L.myWords; R.peer = L.peerWords;
        R.compStride = compStride; R.n = n; R.slotDoubles = L.slotDoubles; R.ncomp = ncomp; R.nblocks = p2p_blocks(n, ncomp); R.recv = recv ? 1 : 0;
        total += R.nblocks;
    };
    // sends first in block order: they are scheduled no later than the receives of the same launch
    add(0, sendL, false); add(1, sendR, false); add(0, recvL, true); add(1, recvR, true);
    if (!X.nroles) return;
    k_p2p_exchange<<<total, 256, 0, stream>>>(X);
    VF_KERNEL_CHECK();
}
// a timed-out wait leaves an error word behind: turned into an exception at the solver's synchronisation points
static void p2p_check(vf_group &G, cudaStream_t stream) {
    if (!G.p2p) return;
    for (int sd = 0; sd < 2; ++sd) {
        if (!G.link[sd].active) continue;
        unsigned long long e = 0;
        VF_CUDA(cudaMemcpyAsync(&e, &G.link[sd].myWords->error[0], sizeof(e), cudaMemcpyDeviceToHost, stream));
        VF_CUDA(cudaStreamSynchronize(stream));
        if (e) throw std::runtime_error("peer-to-peer ghost-plane exchange timed out waiting for rank " + std::to_string(G.rank + (sd ? 1 : -1)) + (e == 1 ? " (mailbox never consumed)" : " (message never arrived)"));
    }
}
// Sets up the mailboxes of an NCCL group: allocate, export with CUDA IPC, swap handles with the neighbours over NCCL, map.
static void p2p_setup(vf_group &G, vf_mg &m) {
    static const bool disabled = [] { const char *e = std::getenv("VF_P2P"); return e && e[0] == '0'; }();
    if (disabled || G.world < 2) return;
    NcclApi &A = NcclApi::get();
    const GridDesc &g0 = m.grid(0);
    const size_t slot = ((size_t)g0.ns[0] * m.N + 31) / 32 * 32;
    cudaStream_t st = m.ctx.stream;
    struct Handles { cudaIpcMemHandle_t box, words; };
    Handles mine[2], theirs[2]; std::memset(mine, 0, sizeof(mine)); std::memset(theirs, 0, sizeof(theirs));
    const bool has[2] = {G.rank > 0, G.rank + 1 < G.world};
    for (int sd = 0; sd < 2; ++sd) {
        if (!has[sd]) continue;
        P2PLink &L = G.link[sd];
        L.slotDoubles = slot;
        VF_CUDA(cudaMalloc(&L.myBox, sizeof(double) * slot * kP2PSlots));
        VF_CUDA(cudaMalloc(&L.myWords, sizeof(P2PWords)));
        VF_CUDA(cudaMemset(L.myWords, 0, sizeof(P2PWords)));
        VF_CUDA(cudaIpcGetMemHandle(&mine[sd].box, L.myBox));
        VF_CUDA(cudaIpcGetMemHandle(&mine[sd].words, L.myWords));
    }
    VF_CUDA(cudaDeviceSynchronize());
    // handles travel as bytes through device buffers (NCCL moves device memory)
    char *dsend = nullptr, *drecv = nullptr;
    VF_CUDA(cudaMalloc(&dsend, 2 * sizeof(Handles))); VF_CUDA(cudaMalloc(&drecv, 2 * sizeof(Handles)));
    VF_CUDA(cudaMemcpy(dsend, mine, 2 * sizeof(Handles), cudaMemcpyHostToDevice));
    A.check(A.GroupStart(), "ncclGroupStart");
    for (int sd = 0; sd < 2; ++sd) {
        if (!has[sd]) continue;
        const int peer = G.rank + (sd ? 1 : -1);
        A.check(A.Send(dsend + sd * sizeof(Handles), sizeof(Handles), /* ncclInt8 */ 0, peer, G.comm, st), "ncclSend");
        A.check(A.Recv(drecv + sd * sizeof(Handles), sizeof(Handles), 0, peer, G.comm, st), "ncclRecv");
    }
    A.check(A.GroupEnd(), "ncclGroupEnd");
    VF_CUDA(cudaStreamSynchronize(st));
    VF_CUDA(cudaMemcpy(theirs, drecv, 2 * sizeof(Handles), cudaMemcpyDeviceToHost));
    cudaFree(dsend); cudaFree(drecv);
    for (int sd = 0; sd < 2; ++sd) {
        if (!has[sd]) continue;
        P2PLink &L = G.link[sd];
        // what the neighbour exported for ITS link towards me
        if (cudaIpcOpenMemHandle((void **)&L.peerBox, theirs[sd].box, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle((void **)&L.peerWords, theirs[sd].words, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return;                                    // no peer mapping (e.g. no NVLink / IPC namespace): NCCL send/recv stays the transport
        }
        L.active = true;
    }
    G.p2p = true;
}
// Ghost-plane exchange of a nodal field of level l: every part receives its ghost planes (global planes sb - 1 and se + 1)
// from the neighbour that owns them.  parity >= 0: only ghost planes whose global index has that parity (after a colour pass).
template<class Sel> void grp_exchange(vf_mg &lead, int l, Sel sel, int parity = -1) {
    if (!level_windowed(lead, l)) return;
    vf_group &G = *lead.grp;
    const int N = lead.N;
    auto planePtr = [&](vf_mg &m, int c, int64_t globalPlane) {
        const GridDesc &g = m.grid(l);
        return sel(m) + (size_t)c * g.numNodes + (size_t)(globalPlane - g.xoff) * g.ns[0];
    };
    if (G.comm) {
        vf_mg &m = *G.parts[0]; MGLevel &L = level(m, l); const GridDesc &g = m.grid(l);
        const bool hasLeft = g.xoff < L.sb, hasRight = g.xoff + g.nn[0] - 1 > L.se;
        const bool doLeft = hasLeft && (parity < 0 || ((L.sb - 1) & 1) == parity), doRight = hasRight && (parity < 0 || ((L.se + 1) & 1) == parity);
        if (!doLeft && !doRight) return;
        if (G.p2p) {   // device-initiated: the neighbour's kernels deposit the planes in this rank's mailboxes (see P2PLink)
            count_launch();
            p2p_exchange(G, doLeft ? planePtr(m, 0, L.sb + 1) : nullptr, doRight ? planePtr(m, 0, L.se - 1) : nullptr,
                         doLeft ? planePtr(m, 0, L.sb - 1) : nullptr, doRight ? planePtr(m, 0, L.se + 1) : nullptr, g.numNodes, g.ns[0], N, lead.ctx.stream);
            return;
        }
        NcclApi &A = NcclApi::get();
        count_launch();
        A.check(A.GroupStart(), "ncclGroupStart");
        for (int c = 0; c < N; ++c) {
            if (doLeft)  { A.check(A.Send(planePtr(m, c, L.sb + 1), g.ns[0], kNcclDouble, G.rank - 1, G.comm, lead.ctx.stream), "ncclSend");
                           A.check(A.Recv(planePtr(m, c, L.sb - 1), g.ns[0], kNcclDouble, G.rank - 1, G.comm, lead.ctx.stream), "ncclRecv"); }
            if (doRight) { A.check(A.Send(planePtr(m, c, L.se - 1), g.ns[0], kNcclDouble, G.rank + 1, G.comm, lead.ctx.stream), "ncclSend");
                           A.check(A.Recv(planePtr(m, c, L.se + 1), g.ns[0], kNcclDouble, G.rank + 1, G.comm, lead.ctx.stream), "ncclRecv"); }
        }
        A.check(A.GroupEnd(), "ncclGroupEnd");
        return;
    }
    for (size_t i = 0; i + 1 < G.parts.size(); ++i) {
        vf_mg &a = *G.parts[i], &b = *G.parts[i + 1];
        MGLevel &La = level(a, l); const size_t bytes = (size_t)a.grid(l).ns[0] * sizeof(double);
        const int64_t shared = La.se;              // == level(b, l).sb
        if (parity >= 0 && ((shared + 1) & 1) != parity) continue;   // shared - 1 and shared + 1 have the same parity
        for (int c = 0; c < N; ++c) {
            count_launch();
            VF_CUDA(cudaMemcpyAsync(planePtr(a, c, shared + 1), planePtr(b, c, shared + 1), bytes, cudaMemcpyDeviceToDevice, lead.ctx.stream));
            VF_CUDA(cudaMemcpyAsync(planePtr(b, c, shared - 1), planePtr(a, c, shared - 1), bytes, cudaMemcpyDeviceToDevice, lead.ctx.stream));
        }
    }
}
// Completion of the sub-assembled stencil of a windowed level: the rows of a plane shared by two slabs are the sum of
// both parts' sub-assemblies (each part assembled only the element layers it owns).
void grp_complete_stencil(vf_mg &lead, int l) {
    vf_group &G = *lead.grp;
    const int NE = (lead.N == 3 ? 27 : 9) * lead.N * lead.N;
    if (G.comm) {
        vf_mg &m = *G.parts[0]; MGLevel &L = level(m, l); const GridDesc &g = L.g;
        const bool hasLeft = g.xoff < L.sb, hasRight = g.xoff + g.nn[0] - 1 > L.se;
        const size_t n = (size_t)stencil_plane_rows(g, 0) * NE;
        NcclApi &A = NcclApi::get();
        double *sendL = stage_buf(m, 0, 4 * n), *sendR = sendL + n, *recvL = sendR + n, *recvR = recvL + n;
        if (hasLeft) launch_stencil_plane_pack(m.ctx, g, L.S.p, (int)(L.sb - g.xoff), sendL);
        if (hasRight) launch_stencil_plane_pack(m.ctx, g, L.S.p, (int)(L.se - g.xoff), sendR);
        A.check(A.GroupStart(), "ncclGroupStart");
        if (hasLeft)  { A.check(A.Send(sendL, n, kNcclDouble, G.rank - 1, G.comm, m.ctx.stream), "ncclSend"); A.check(A.Recv(recvL, n, kNcclDouble, G.rank - 1, G.comm, m.ctx.stream), "ncclRecv"); }
        if (hasRight) { A.check(A.Send(sendR, n, kNcclDouble, G.rank + 1, G.comm, m.ctx.stream), "ncclSend"); A.check(A.Recv(recvR, n, kNcclDouble, G.rank + 1, G.comm, m.ctx.stream), "ncclRecv"); }
        A.check(A.GroupEnd(), "ncclGroupEnd");
        if (hasLeft) launch_stencil_plane_add(m.ctx, g, L.S.p, (int)(L.sb - g.xoff), recvL);
        if (hasRight) launch_stencil_plane_add(m.ctx, g, L.S.p, (int)(L.se - g.xoff), recvR);
        return;
    }
    for (size_t i = 0; i + 1 < G.parts.size(); ++i) {
        vf_mg &a = *G.parts[i], &b = *G.parts[i + 1];
        MGLevel &La = level(a, l), &Lb = level(b, l);
        const size_t n = (size_t)stencil_plane_rows(La.g, 0) * NE;
        double *bufA = stage_buf(a, 0, n), *bufB = stage_buf(b, 0, n);
        const int pa = (int)(La.se - La.g.xoff), pb = (int)(Lb.sb - Lb.g.xoff);
        launch_stencil_plane_pack(a.ctx, La.g, La.S.p, pa, bufA);
        launch_stencil_plane_pack(b.ctx, Lb.g, Lb.S.p, pb, bufB);
        launch_stencil_plane_add(a.ctx, La.g, La.S.p, pa, bufB);
        launch_stencil_plane_add(b.ctx, Lb.g, Lb.S.p, pb, bufA);
    }
}
// zero the node planes of a replicated-level field that this part does not own (before summing the parts' contributions)
void zero_unowned_planes(vf_mg &m, int l, double *f) {
    const GridDesc &g = m.grid(l);
    for (int c = 0; c < m.N; ++c) {
        double *p = f + (size_t)c * g.numNodes;
        if (g.ownLo > 0) VF_CUDA(cudaMemsetAsync(p, 0, (size_t)g.ownLo * g.ns[0] * sizeof(double), m.ctx.stream));
        if (g.ownHi < g.nn[0]) VF_CUDA(cudaMemsetAsync(p + (size_t)g.ownHi * g.ns[0], 0, (size_t)(g.nn[0] - g.ownHi) * g.ns[0] * sizeof(double), m.ctx.stream));
    }
}

// updateStiffnessMatrices (MultigridSolver.hh:846-905): rebuild all coarse operators for the current moduli/mask.
// The reference's banded partial update (:907-1017) yields the same operators; here the (cheap) full rebuild is
// always used and is skipped only when nothing changed since the last build.
// Slab groups: every part sub-assembles the windowed levels from the element layers it owns (Galerkin coarsening is
// additive over elements), the rows on planes shared by two slabs are then completed by one exchange-add per level, and
// the first replicated level is the all-reduced sum of the parts' sub-assemblies; deeper levels are coarsened redundantly.
void mg_update_stiffness(vf_mg &lead, bool force = false) {
    std::vector<vf_mg *> &P = parts_of(lead);
    bool stale = force;
    for (vf_mg *m : P) { mg_sync_level_masks(*m); stale = stale || m->stiffnessVersion != m->sim->version; }
    if (!stale) return;
    TraceScope ts("updateStiffnessMatrices");                  // MultigridSolver.hh:854
    const int nl = lead.numLevels();
    const int T = lead.grp ? lead.firstRep : 0;   // levels 1..T are sub-assembled per part, then completed
    // A pending banded update recomputes only the coarse rows whose support meets the changed fine element layers: the range
    // of affected node layers halves (and widens by one node on either side) from level to level.
    const bool banded = !force && !lead.grp && lead.bandActive && lead.bandVersion == lead.sim->version;
    std::vector<int> bLo(nl, 0), bHi(nl, 0x7fffffff);
    if (banded) {
        int lo = lead.bandLo, hi = lead.bandHi;             // fine node layers lo .. hi touch the changed elements [lo, hi)
        for (int l = 1; l < nl; ++l) { lo = std::max(lo / 2 - 1, 0); hi = (hi + 1) / 2 + 1; bLo[l] = lo; bHi[l] = hi; }
    }
    lead.bandActive = false;
    static const bool noSeparable = [] { const char *e = std::getenv("VF_COARSEN_ONESHOT"); return e && e[0] == '1'; }();
    auto coarsen = [&](vf_mg &mg, int l) {
        MGLevel &L = *mg.lv[l];
        const size_t len = (size_t)L.g.numPos * (mg.N == 3 ? 27 : 9) * mg.N * mg.N;
        if (L.S.n != len) L.S.alloc(len, true);
        if (l == 1) launch_coarsen_from_moduli(mg.ctx, L.g, mg.sim->g, mg.sim->E.p, mg.cK0dev.p, L.S.p, bLo[l], bHi[l], mg.cK0.data());
        else if (mg.N == 3 && !banded && !noSeparable) {
            const GridDesc &gf = mg.lv[l - 1]->g;
            const size_t need = coarsen_separable_scratch(L.g, gf);
            if (mg.coarsenScratch.n < need) mg.coarsenScratch.alloc(need, false);
            launch_coarsen_stencil_separable(mg.ctx, L.g, gf, mg.lv[l - 1]->S.p, L.S.p, mg.coarsenScratch.p);
        }
        else        launch_coarsen_stencil(mg.ctx, L.g, mg.lv[l - 1]->g, mg.lv[l - 1]->S.p, L.S.p, bLo[l], bHi[l]);
    };
    for (int l = 1; l < nl && l <= T; ++l) for (vf_mg *m : P) coarsen(*m, l);
    if (lead.grp) {
        for (int l = 1; l < nl && l < T; ++l) grp_complete_stencil(lead, l);
        if (T < nl) grp_allreduce(lead, [&](vf_mg &m) { return m.lv[T]->S.p; }, lead.lv[T]->S.n);
    }
    for (int l = std::max(T + 1, 1); l < nl; ++l) for (vf_mg *m : P) coarsen(*m, l);
    if (nl > 1) {
        for (vf_mg *m : P) {
            vf_mg &mg = *m; MGLevel &C = *mg.lv[nl - 1];
            // findFixedVars (TensorProductSimulator.hh:1181-1195): Dirichlet components and detached nodes
            std::vector<uint8_t> fixed((size_t)C.g.numNodes * mg.N, 0);
            for (long long n = 0; n < C.g.numNodes; ++n) {
                const int cbd = (C.g.bd == 2) ? (int)(n % C.g.nn[2]) : (int)((n / C.g.nn[2]) % C.g.nn[1]);
                const bool det = cbd >= C.g.nActive;
                for (int c = 0; c < mg.N; ++c) if (det || ((C.nodeMask[n] >> c) & 1)) fixed[n * mg.N + c] = 1;
            }
            mg.coarse.factor(mg.ctx, C.g, C.S.p, fixed);
        }
    }
    for (vf_mg *m : P) m->stiffnessVersion = m->sim->version;
}

double *lx(vf_mg &mg, int l) { return mg.lv[l]->x.p; }
double *lb(vf_mg &mg, int l) { return mg.lv[l]->b.p; }
double *lr(vf_mg &mg, int l) { return mg.lv[l]->r.p; }
enum FieldId { F_X = 0, F_B, F_R, F_D, F_AD, F_USER };
// Field selector shared by all parts of a group ("the same field on every part")
struct Field {
    int id, l; double *user0 = nullptr; // user pointer: only valid for single-part calls
    double *operator()(vf_mg &m) const {
        switch (id) { case F_X: return lx(m, l); case F_B: return lb(m, l); case F_R: return lr(m, l); case F_D: return m.d.p; case F_AD: return m.Ad.p; default: return user0; }
    }
};
Field fu(const double *p) { return Field{F_USER, 0, const_cast<double *>(p)}; }
Field fx(int l) { return Field{F_X, l}; } Field fb(int l) { return Field{F_B, l}; } Field fr(int l) { return Field{F_R, l}; }

// out (=, +=, -=) K u  or  out = b - K u  on one part
void part_apply_K(vf_mg &mg, int l, const double *u, const double *b, double *out, int mode, bool zeroDirichlet, double *dotOut = nullptr) {
    if (l == 0) launch_apply_l0(mg.ctx, mg.sim->g, mg.sim->K0p, u, mg.sim->E.p, b, zeroDirichlet ? mg.dmask(0) : nullptr, out, mode, dotOut, mg.scratch.p);
    else launch_apply_stencil(mg.ctx, mg.lv[l]->g, mg.lv[l]->S.p, u, b, zeroDirichlet ? mg.dmask(l) : nullptr, out, mode, mg.lv[l]->posTab.p);
}
void mg_apply_K(vf_mg &lead, int l, Field u, Field b, Field out, int mode, bool zeroDirichlet, int dotSlot = -1) {
    if (l > 0) mg_update_stiffness(lead);
    for (vf_mg *m : parts_of(lead)) part_apply_K(*m, l, u(*m), mode == APPLY_RESIDUAL ? b(*m) : nullptr, out(*m), mode, zeroDirichlet, dotSlot >= 0 ? m->scalars.p + dotSlot : nullptr);
    if (dotSlot >= 0) grp_allreduce(lead, [&](vf_mg &m) { return m.scalars.p + dotSlot; }, 1);
}
// computeResidual (:527-541): r = b - K u on non-detached nodes, Dirichlet components zeroed.  The ghost planes of r are
// refreshed because the restriction that follows reads them.
void mg_residual(vf_mg &lead, int l, Field u, Field b, Field r) {
    mg_apply_K(lead, l, u, b, r, APPLY_RESIDUAL, true);
    grp_exchange(lead, l, r);
}

// smoothingMulticoloredGS (:452-458): 2^N colour passes, colours reversed for backward sweeps (:417).  In a slab window the
// colour of a node is the parity class of its GLOBAL index; after the passes that updated the parity of the ghost planes those
// planes are received from the neighbours.
// res != nullptr (stored-stencil levels of an undivided solver, gs_residual_fusable): the sweep leaves b - K u of its final iterate
// in *res, Dirichlet components not yet zeroed.
void mg_smooth(vf_mg &lead, int l, Field u, Field b, bool forward, const Field *res = nullptr) {
    const int nc = 1 << lead.N;
    if (l > 0) mg_update_stiffness(lead);
    if (l == 0 && lead.N == 3) {
        // row-unit kernel (vf_gs0.cu): one launch per (x, y) parity class, both z-colours of a row inside the launch -- the same
        // visiting order.  All parts of a group share the material and the z extent, so they take the same path.
        bool rows = true;
        for (vf_mg *mp : parts_of(lead)) rows = rows && gs_rows_supported(mp->grid(0), mp->sim->K0p);
        if (rows) {
            for (int i = 0; i < 4; ++i) {
                const int gcls = forward ? i : 3 - i;          // global (x parity, y parity) class
                for (vf_mg *mp : parts_of(lead)) {
                    vf_mg &mg = *mp; const GridDesc &g = mg.grid(0);
                    launch_gs_rows_l0(mg.ctx, g, mg.sim->K0p, u(mg), b(mg), mg.sim->E.p, mg.dmask(0), gcls ^ ((g.xoff & 1) << 1), forward);
                }
                if (i & 1) grp_exchange(lead, l, u, (gcls >> 1) & 1);   // all planes of this x parity are final
            }
            return;
        }
    }
    if (l > 0 && (!lead.grp || l >= lead.firstRep) && stencil_sweep_fused(lead.grid(l))) {
        // small level without ghost planes: the 2^N colour passes in one persistent launch (grid barrier between the colours)
        for (vf_mg *mp : parts_of(lead)) {
            vf_mg &mg = *mp; const GridDesc &g = mg.grid(l);
            launch_gs_stencil_sweep(mg.ctx, g, mg.lv[l]->S.p, u(mg), b(mg), mg.dmask(l), forward, g.xoff & 1, mg.gridBar.p);
        }
        return;
    }
    for (int i = 0; i < nc; ++i) {
        const int color = forward ? i : (nc - 1 - i);
        for (vf_mg *mp : parts_of(lead)) {
            vf_mg &mg = *mp; const GridDesc &g = mg.grid(l);
            const int lc = (lead.N == 3) ? (color ^ ((g.xoff & 1) << 2)) : color;   // local parity class of this global colour
            if (l == 0 && mg.N == 3) launch_gs3_color_l0(mg.ctx, g, mg.sim->K0p, u(mg), b(mg), mg.sim->E.p, mg.dmask(0), lc, forward);
            else if (l == 0)         launch_gs_l0(mg.ctx, g, mg.sim->K0p, u(mg), b(mg), mg.sim->E.p, mg.dmask(0), lc, forward);
            else                     launch_gs_stencil(mg.ctx, g, mg.lv[l]->S.p, u(mg), b(mg), mg.dmask(l), lc, forward, /* chained */ i > 0 && !lead.grp, res ? (*res)(mg) : nullptr, mg.lv[l]->posTab.p);
        }
        // The four passes of one x parity only read planes of the other parity besides their own plane, so a ghost plane (one parity)
        // has to be current only when the passes of the OTHER parity start: one exchange per parity group instead of one per pass.
        if (lead.N == 3 && (i & 3) == 3) grp_exchange(lead, l, u, (color >> 2) & 1);
    }
}
// coarsest-level solve on the replicated coarsest grid
void mg_coarse_solve(vf_mg &lead, Field f, Field x) {
    mg_update_stiffness(lead);
    for (vf_mg *m : parts_of(lead)) m->coarse.solve(m->ctx, m->grid(m->numLevels() - 1), f(*m), x(*m));
}
void mg_enforce_dirichlet(vf_mg &lead, int l, Field u, bool zero) { // (:521-524)
    for (vf_mg *mp : parts_of(lead)) {
        vf_mg &mg = *mp;
        if (zero || l > 0) launch_zero_dirichlet(mg.ctx, mg.grid(l), mg.dmask(l), u(mg));
        else launch_enforce_dirichlet(mg.ctx, mg.sim->g.numNodes, mg.N, (int)mg.sim->dirNodes.size(), mg.sim->dirNodesDev.p, mg.sim->dirMaskDev.p, mg.sim->dirValsDev.p, u(mg));
    }
}
// restriction (:216-262) onto level l + 1.  Onto the first replicated level every part restricts the coarse planes it owns
// and the parts' contributions are summed; between windowed levels the coarse ghost planes are received.
void mg_restrict(vf_mg &lead, int l, Field fine, Field coarse) {
    for (vf_mg *m : parts_of(lead)) launch_restrict(m->ctx, m->grid(l), m->grid(l + 1), fine(*m), coarse(*m));
    if (!lead.grp) return;
    if (l + 1 == lead.firstRep) {
        for (vf_mg *m : parts_of(lead)) zero_unowned_planes(*m, l + 1, coarse(*m));
        grp_allreduce(lead, coarse, (size_t)lead.grid(l + 1).numNodes * lead.N);
    } else grp_exchange(lead, l + 1, coarse);
}
void mg_prolong(vf_mg &lead, int l, Field coarse, Field fine, bool accumulate) { // interpolation / accum_interpolation (:178-212), level l + 1 -> l
    for (vf_mg *m : parts_of(lead)) launch_prolong(m->ctx, m->grid(l), m->grid(l + 1), coarse(*m), fine(*m), accumulate);
}

// vcycle (:617-658)
void mg_vcycle(vf_mg &lead, int l, int nsmooth, bool residualSystem) {
    const int coarsest = lead.numLevels() - 1;
    if (l == coarsest) { mg_coarse_solve(lead, fb(l), fx(l)); return; }
    // NCCL rank: the V-cycle over the replicated levels (first replicated level downwards) needs no communication and is entered
    // firstRep + 1 times per FMG cycle with ~100 tiny launches each: replay it as a captured CUDA graph.  (The windowed levels stay
    // eager: capturing the NCCL ghost-plane exchanges deadlocked, see mg_pcg.)
    if (lead.grp && lead.grp->comm && l == lead.firstRep && l > 0 && residualSystem && !lead.inSubCapture && lead.useGraphs &&
        !(lead.ctx.prof && lead.ctx.prof->enabled) && !trace_enabled() && parts_of(lead).size() == 1) {
        static const bool enabled = [] { const char *e = std::getenv("VF_SUBGRAPH"); return !(e && e[0] == '0'); }();
        if (enabled) {
            mg_update_stiffness(lead);
            vf_mg::PrecondGraph &sg = lead.sg;
            const bool valid = sg.exec && sg.version == lead.sim->structVersion && sg.nActive == lead.sim->g.nActive && sg.nsmooth == nsmooth && sg.sym == lead.symmetricGS;
            if (!valid) {
                if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
                const long long before = g_launches.load();
                cudaGraph_t graph = nullptr;
                lead.inSubCapture = true;
                VF_CUDA(cudaStreamBeginCapture(lead.ctx.stream, cudaStreamCaptureModeThreadLocal));
                try { mg_vcycle(lead, l, nsmooth, true); } catch (...) { lead.inSubCapture = false; cudaStreamEndCapture(lead.ctx.stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
                lead.inSubCapture = false;
                VF_CUDA(cudaStreamEndCapture(lead.ctx.stream, &graph));
                VF_CUDA(cudaGraphInstantiate(&sg.exec, graph, 0));
                cudaGraphDestroy(graph);
                sg.launches = g_launches.load() - before; g_launches.fetch_sub(sg.launches);
                sg.version = lead.sim->structVersion; sg.nActive = lead.sim->g.nActive; sg.nsmooth = nsmooth; sg.sym = lead.symmetricGS;
            }
            VF_CUDA(cudaGraphLaunch(sg.exec, lead.ctx.stream));
            g_launches.fetch_add(sg.launches);
            return;
        }
    }
    TraceScope ts("V Cycle " + std::to_string(l));             // MultigridSolver.hh:618
    mg_enforce_dirichlet(lead, l, fx(l), residualSystem);
    // On a stored-stencil level the last pre-smoothing sweep emits the residual itself (k_stencil_tile<RES>): the stencil, 97 % of the
    // level's bytes, is streamed twice per visit instead of three times.
    bool fusedRes = l > 0 && nsmooth > 0;
    for (vf_mg *m : parts_of(lead)) fusedRes = fusedRes && gs_residual_fusable(m->grid(l));
    const Field rl = fr(l);
    for (int i = 0; i < nsmooth; ++i) mg_smooth(lead, l, fx(l), fb(l), true, (fusedRes && i == nsmooth - 1) ? &rl : nullptr);
    if (fusedRes) {
        for (vf_mg *mp : parts_of(lead)) {
            vf_mg &m = *mp; const GridDesc &g = m.grid(l); MGLevel &L = *m.lv[l];
            // slab window: the shared planes next to a ghost plane did not see the neighbouring part's updates -- their residual is
            // formed directly from the final iterate (the ghost planes are current after the sweep's last exchange)
            if (g.cmpLo > 0)        launch_residual_stencil_plane(m.ctx, g, L.S.p, lx(m, l), lb(m, l), nullptr, lr(m, l), g.cmpLo, L.posTab.p);
            if (g.cmpHi < g.nn[0])  launch_residual_stencil_plane(m.ctx, g, L.S.p, lx(m, l), lb(m, l), nullptr, lr(m, l), g.cmpHi - 1, L.posTab.p);
            launch_zero_dirichlet(m.ctx, g, m.dmask(l), lr(m, l));   // computeResidual's mask (:538)
        }
        grp_exchange(lead, l, fr(l));   // as mg_residual: the restriction reads the ghost planes of r
    }
    else mg_residual(lead, l, fx(l), fb(l), fr(l));
    mg_restrict(lead, l, fr(l), fb(l + 1));
    for (vf_mg *m : parts_of(lead)) launch_masked_zero(m->ctx, m->grid(l + 1), lx(*m, l + 1), 4 /* VOXELFEM_SIMD_WIDTH margin (:644) */);
    mg_vcycle(lead, l + 1, nsmooth, true);
    mg_prolong(lead, l, fx(l + 1), fx(l), true);
    for (int i = 0; i < nsmooth; ++i) mg_smooth(lead, l, fx(l), fb(l), !lead.symmetricGS);
}
// fullMultigrid (:587-609)
void mg_fmg(vf_mg &lead, int l, int nsmooth, bool residualSystem) {
    const int coarsest = lead.numLevels() - 1;
    if (l == coarsest) { mg_coarse_solve(lead, fb(l), fx(l)); return; }
    mg_restrict(lead, l, fb(l), fb(l + 1));
    mg_fmg(lead, l + 1, nsmooth, residualSystem);
    mg_prolong(lead, l, fx(l + 1), fx(l), false);
    mg_vcycle(lead, l, nsmooth, residualSystem);
}
// solve (:546-573) operating on lv[0].x (initial guess already there) and lv[0].b
void mg_solve_inplace(vf_mg &lead, int numSteps, int nsmooth, bool zeroDirichlet, bool fmg) {
    if (numSteps == 0) return;
    TraceScope ts("MG Solver");                                // MultigridSolver.hh:549
    int start = 0;
    if (fmg) { mg_fmg(lead, 0, nsmooth, zeroDirichlet); start = 1; }
    for (int i = start; i < numSteps; ++i) mg_vcycle(lead, 0, nsmooth, zeroDirichlet);
}

double read_scalar(vf_mg &mg, int slot) {
    VF_CUDA(cudaMemcpyAsync(mg.hostScalars + slot, mg.scalars.p + slot, sizeof(double), cudaMemcpyDeviceToHost, mg.ctx.stream));
    VF_CUDA(cudaStreamSynchronize(mg.ctx.stream));
    return mg.hostScalars[slot];
}
// masked dot product over the owned nodes of all parts -> scalar slot (every part gets the total)
void mg_dot(vf_mg &lead, Field a, Field b, int slot) {
    for (vf_mg *m : parts_of(lead)) launch_masked_dot(m->ctx, m->sim->g, a(*m), b(*m), m->scalars.p + slot, m->scratch.p);
    grp_allreduce(lead, [&](vf_mg &m) { return m.scalars.p + slot; }, 1);
}

void sim_direct_solve(vf_sim &s, const double *fDev, double *xDev); // below

// preconditionedConjugateGradient (:1047-1152), device-resident x and b (one pointer per part of the group)
void mg_pcg(vf_mg &lead, double *const *xs, const double *const *bs, int maxIter, double tol, int mgIterations, int mgSmoothing, bool fmg, bool dirichletOK,
            vf_pcg_callback cb, void *user) {
    std::vector<vf_mg *> &P = parts_of(lead);
    const int N = lead.N;
    lead.lastResiduals.clear(); lead.lastIters = 0;
    const bool prebuilt = lead.rebuiltForThisSolve; lead.rebuiltForThisSolve = false;
    for (vf_mg *m : P) mg_sync_level_masks(*m);
    const Field r = fb(0), s = fx(0), d{F_D, 0}, Ad{F_AD, 0};
    auto X = [&](vf_mg &m) { for (size_t i = 0; i < P.size(); ++i) if (P[i] == &m) return xs[i]; return (double *)nullptr; };
    auto B = [&](vf_mg &m) { for (size_t i = 0; i < P.size(); ++i) if (P[i] == &m) return const_cast<double *>(bs[i]); return (double *)nullptr; };
    if (lead.numLevels() == 1) { // (:1057-1065)
        vf_sim &sim = *lead.sim; const GridDesc &g = sim.g; double *sc = lead.scalars.p;
        part_apply_K(lead, 0, xs[0], bs[0], lb(lead, 0), APPLY_RESIDUAL, true);
        launch_dot_plain(lead.ctx, g.numNodes * N, lb(lead, 0), lb(lead, 0), sc + SC_RSQ, lead.scratch.p);
        launch_dot_plain(lead.ctx, g.numNodes * N, bs[0], bs[0], sc + SC_BSQ, lead.scratch.p);
        const double rs = read_scalar(lead, SC_RSQ), bsq1 = read_scalar(lead, SC_BSQ);
        if (rs < tol * tol * bsq1) return;
        sim_direct_solve(sim, bs[0], xs[0]);
        part_apply_K(lead, 0, xs[0], bs[0], lb(lead, 0), APPLY_RESIDUAL, true);
        launch_dot_plain(lead.ctx, g.numNodes * N, lb(lead, 0), lb(lead, 0), sc + SC_RSQ, lead.scratch.p);
        const double rn = std::sqrt(read_scalar(lead, SC_RSQ));
        lead.lastIters = 1; lead.lastResiduals.push_back(rn);
        if (cb) cb(1, rn, user);
        return;
    }
    TraceScope tsCG("CG Iterations");                          // MultigridSolver.hh:1067
    trace_push("Preamble");                                    // (:1077)
    for (vf_mg *m : P) {
        if (!dirichletOK) launch_enforce_dirichlet(m->ctx, m->sim->g.numNodes, N, (int)m->sim->dirNodes.size(), m->sim->dirNodesDev.p, m->sim->dirMaskDev.p, m->sim->dirValsDev.p, X(*m));
        launch_masked_dot(m->ctx, m->sim->g, B(*m), B(*m), m->scalars.p + SC_BSQ, m->scratch.p);
        part_apply_K(*m, 0, X(*m), B(*m), r(*m), APPLY_RESIDUAL, true);       // computeResidual(0, x, b, r) (:1080)
    }
    grp_allreduce(lead, [&](vf_mg &m) { return m.scalars.p + SC_BSQ; }, 1);
    grp_exchange(lead, 0, r);
    mg_dot(lead, r, r, SC_RSQ);
    const double bsq = read_scalar(lead, SC_BSQ);
    double rsq = read_scalar(lead, SC_RSQ);
    trace_pop("Preamble");
    if (std::isnan(rsq)) throw std::logic_error("NaN encountered");
    int i = 0; bool first = true; int cur = SC_RMR_A, old = SC_RMR_B;
    while ((i++ < maxIter) && (rsq > tol * tol * bsq)) {
        if (mgIterations > 0 && mgSmoothing > 0) {
            mg_update_stiffness(lead, lead.rebuildEverySolve && !prebuilt && i == 1); // lazily, first iteration (:1104-1107)
            // applyPreconditionerInv: zero initial guess (:577-580); the FMG cycle overwrites s by interpolation (:600)
            auto precond = [&]() {
                if (!fmg) for (vf_mg *m : P) VF_CUDA(cudaMemsetAsync(s(*m), 0, sizeof(double) * m->sim->g.numNodes * N, m->ctx.stream));
                mg_solve_inplace(lead, mgIterations, mgSmoothing, true, fmg);
            };
            // Slab groups: capturing NCCL send/recv pairs into the graph deadlocked on 2 GPUs (round 1).  With the device-initiated
            // ghost-plane exchange (P2PLink) the windowed levels hold no host-side communication call any more, only the NCCL
            // all-reduce onto the first replicated level, which NCCL supports inside a capture: an NCCL rank replays the whole
            // preconditioner as one graph (VF_GROUP_GRAPH=0 keeps it eager); local groups and the send/recv transport stay eager.
            static const bool groupGraph = [] { const char *e = std::getenv("VF_GROUP_GRAPH"); return !(e && e[0] == '0'); }();
            const bool groupOk = !lead.grp || (groupGraph && lead.grp->comm && lead.grp->p2p && parts_of(lead).size() == 1);
            const bool graphable = lead.useGraphs && groupOk && !(lead.ctx.prof && lead.ctx.prof->enabled) && !trace_enabled();   // section timers drain the device: not capturable
            vf_mg::PrecondGraph &pg = lead.pg;
            const bool valid = pg.exec && pg.version == lead.sim->structVersion && pg.nActive == lead.sim->g.nActive && pg.mgIt == mgIterations &&
                               pg.nsmooth == mgSmoothing && pg.fmg == fmg && pg.sym == lead.symmetricGS;
            if (!graphable) precond();
            else {
                if (!valid) {
                    if (pg.exec) { cudaGraphExecDestroy(pg.exec); pg.exec = nullptr; }
                    const long long before = g_launches.load();
                    cudaGraph_t graph = nullptr;
                    VF_CUDA(cudaStreamBeginCapture(lead.ctx.stream, cudaStreamCaptureModeThreadLocal));
                    const bool wasSub = lead.inSubCapture; lead.inSubCapture = true;   // no nested capture of the replicated-level V-cycle
                    try { precond(); } catch (...) { lead.inSubCapture = wasSub; cudaStreamEndCapture(lead.ctx.stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
                    lead.inSubCapture = wasSub;
                    VF_CUDA(cudaStreamEndCapture(lead.ctx.stream, &graph));
                    VF_CUDA(cudaGraphInstantiate(&pg.exec, graph, 0));
                    cudaGraphDestroy(graph);
                    pg.launches = g_launches.load() - before; g_launches.fetch_sub(pg.launches);
                    pg.version = lead.sim->structVersion; pg.nActive = lead.sim->g.nActive; pg.mgIt = mgIterations; pg.nsmooth = mgSmoothing; pg.fmg = fmg; pg.sym = lead.symmetricGS;
                }
                VF_CUDA(cudaGraphLaunch(pg.exec, lead.ctx.stream));
                g_launches.fetch_add(pg.launches);
            }
        } else {
            for (vf_mg *m : P) VF_CUDA(cudaMemcpyAsync(s(*m), r(*m), sizeof(double) * m->sim->g.numNodes * N, cudaMemcpyDeviceToDevice, m->ctx.stream)); // s = r (:578, :1118)
        }
        for (vf_mg *m : P) launch_zero_dirichlet(m->ctx, m->sim->g, m->dmask(0), s(*m));            // (:1122)
        std::swap(cur, old);
        trace_push("CG direction");                                                                 // (:1121)
        mg_dot(lead, r, s, cur);                                                                    // r_Minv_r (:1124)
        for (vf_mg *m : P) launch_cg_direction(m->ctx, m->sim->g, s(*m), m->d.p, m->scalars.p + cur, m->scalars.p + old, first); // d = s + beta d (:1125-1126)
 first = false;
        trace_pop("CG direction");
        trace_push("CG update");                                                                    // (:1133)
        mg_apply_K(lead, 0, d, d, Ad, APPLY_SET, true);                                             // Ad = K d, zero Dirichlet (:1129-1130)
        mg_dot(lead, d, Ad, SC_DAD);                                                                // d . Ad (:1134); a reduction fused into the apply kernel measured slower (1.79 vs 1.10 + 0.11 ms at 256^3)
        grp_exchange(lead, 0, Ad);                                                                  // keeps r consistent on the ghost planes (r feeds the next restriction)
        for (vf_mg *m : P) launch_cg_update(m->ctx, m->sim->g, X(*m), m->d.p, r(*m), m->Ad.p, m->scalars.p + cur, m->scalars.p + SC_DAD, m->scalars.p + SC_RSQ, m->scratch.p); // (:1134-1143)
        grp_allreduce(lead, [&](vf_mg &m) { return m.scalars.p + SC_RSQ; }, 1);
        rsq = read_scalar(lead, SC_RSQ);
        if (std::isnan(rsq)) throw std::logic_error("NaN encountered at iteration" + std::to_string(i));
        lead.lastIters = i; lead.lastResiduals.push_back(std::sqrt(rsq));
        trace_pop("CG update");
        if (cb) { TraceScope tcb("Callback"); cb(i, std::sqrt(rsq), user); }                        // (:1148)
    }
}
void mg_pcg(vf_mg &mg, double *x, const double *b, int maxIter, double tol, int mgIterations, int mgSmoothing, bool fmg, bool dirichletOK,
            vf_pcg_callback cb, void *user) {
    if (mg.grp && mg.grp->parts.size() != 1) throw std::runtime_error("use vf_group_pcg_dev for a multi-part slab group");
    double *xs[1] = {x}; const double *bs[1] = {b};
    mg.pcgX = x;
    mg_pcg(mg, xs, bs, maxIter, tol, mgIterations, mgSmoothing, fmg, dirichletOK, cb, user);
}

// TPS::solve at level 0 (TensorProductSimulator.hh:1198-1230) through the dense GPU solver
void sim_direct_solve(vf_sim &s, const double *fDev, double *xDev) {
    if (s.directVersion != s.version || !s.direct.ok) {
        std::vector<uint8_t> fixed((size_t)s.g.numNodes * s.N, 0);
        long long nfree = 0;
        for (long long n = 0; n < s.g.numNodes; ++n) {
            const int cbd = (s.g.bd == 2) ? (int)(n % s.g.nn[2]) : (int)((n / s.g.nn[2]) % s.g.nn[1]);
            const bool det = cbd >= s.g.nActive;
            for (int c = 0; c < s.N; ++c) { if (det || ((s.nodeMask[n] >> c) & 1)) fixed[n * s.N + c] = 1; else ++nfree; }
        }
        for (size_t k = 0; k < s.dirNodes.size(); ++k) for (int c = 0; c < s.N; ++c)
            if (((s.dirMask[k] >> c) & 1) && s.dirVals[k * s.N + c] != 0) throw std::runtime_error("Nonzero Dirichlet constraints currently unsupported");
        if (nfree > VF_MAX_DIRECT_DOFS) throw std::runtime_error("direct solve requested for " + std::to_string(nfree) + " free variables; use the multigrid solver (limit " + std::to_string(VF_MAX_DIRECT_DOFS) + ")");
        const size_t len = (size_t)s.g.numPos * (s.N == 3 ? 27 : 9) * s.N * s.N;
        s.directStencil.alloc(len, true);
        launch_stencil_from_moduli_l0(s.ctx, s.g, s.E.p, s.K0dev.p, s.directStencil.p);
        s.direct.factor(s.ctx, s.g, s.directStencil.p, fixed);
        s.directVersion = s.version;
    }
    s.direct.solve(s.ctx, s.g, fDev, xDev);
}

void sim_build_load_dev(vf_sim &s, double *f) { // buildLoadVector (:1269-1288)
    VF_CUDA(cudaMemsetAsync(f, 0, sizeof(double) * s.g.numNodes * s.N, s.stream));
    const size_t nf = s.forceNodes.size();
    if (nf) { // scatter the point loads: reuse the Dirichlet "set value" kernel with a full component mask
        DevBuf<long long> nodes; DevBuf<uint8_t> masks; DevBuf<double> vals;
        std::vector<long long> hn(s.forceNodes.begin(), s.forceNodes.end()); std::vector<uint8_t> hm(nf, uint8_t((1u << s.N) - 1u));
        nodes.alloc(nf, false); masks.alloc(nf, false); vals.alloc(nf * s.N, false);
        nodes.upload(hn.data(), nf, s.stream); masks.upload(hm.data(), nf, s.stream); vals.upload(s.forceVals.data(), nf * s.N, s.stream);
        launch_enforce_dirichlet(s.ctx, s.g.numNodes, s.N, (int)nf, nodes.p, masks.p, vals.p, f);
        VF_CUDA(cudaStreamSynchronize(s.stream));
    }
    if (s.gravity[0] != 0 || s.gravity[1] != 0 || s.gravity[2] != 0)
        launch_self_weight_load(s.ctx, s.g, s.rho.p, s.gravity, s.elemVolume(), f, 0, s.g.neActive, 1.0);
}

} // namespace

#define VF_TRY try {
#define VF_CATCH } catch (const std::logic_error &e) { vf::g_err = std::string("logic_error: ") + e.what(); return 2; } \
                   catch (const std::exception &e) { vf::g_err = e.what(); return 1; } return 0;

// scratch VField (device) of the simulator for host-pointer entry points
static double *sim_tmp(vf_sim *s, int which) {
    DevBuf<double> &b = which == 0 ? s->tmpU : s->tmpV;
    const size_t len = (size_t)s->g.numNodes * s->N;
    if (b.n != len) b.alloc(len, true);
    return b.p;
}
static double *mg_tmp(vf_mg *mg, int which, size_t len) {
    DevBuf<double> &b = which == 0 ? mg->tmpA : (which == 1 ? mg->tmpB : mg->tmpC);
    if (b.n < len) b.alloc(len, true);
    return b.p;
}
static void h2d(double *dev, const double *host, size_t n, cudaStream_t s) { VF_CUDA(cudaMemcpyAsync(dev, host, n * sizeof(double), cudaMemcpyHostToDevice, s)); }
static void d2h(double *host, const double *dev, size_t n, cudaStream_t s) { VF_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, s)); VF_CUDA(cudaStreamSynchronize(s)); }

extern "C" {

const char *vf_last_error(void) { return vf::g_err.c_str(); }
int vf_version(void) { return 100; }
int vf_device_count(int *count) { VF_TRY VF_CUDA(cudaGetDeviceCount(count)); VF_CATCH }
int vf_set_device(int device) { VF_TRY VF_CUDA(cudaSetDevice(device)); VF_CATCH }
int64_t vf_kernel_launch_count(void) { return vf::g_launches.load(); }
void vf_reset_kernel_launch_count(void) { vf::g_launches.store(0); }
int vf_measure_fp64_peak(double *tera_dfma_per_s) { VF_TRY ensure_device(); *tera_dfma_per_s = vf::measure_dfma_peak(nullptr); VF_CATCH }

// ---- simulator ----------------------------------------------------------------------------
static vf_sim *sim_create_common(int dim, const int64_t *gne, const double *dmin, const double *dmax, bool window, int64_t slabBegin, int64_t slabEnd, vf_sim *share) {
    if (dim != 2 && dim != 3) throw std::runtime_error("dim must be 2 or 3");
    ensure_device();
    auto s = std::make_unique<vf_sim>();
    s->N = dim;
    int64_t ne[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
        if (gne[d] < 1) throw std::runtime_error("grid must have at least one element per dimension");
        ne[d] = gne[d]; s->dmin[d] = dmin[d]; s->dmax[d] = dmax[d];
        s->spacing[d] = (dmax[d] - dmin[d]) / (double(gne[d] + 1) - 1.0);
        s->stretch[d] = (dmax[d] - dmin[d]) / double(gne[d]);
    }
    int64_t wlo = 0;
    if (window) {
        if (dim != 3) throw std::runtime_error("slab windows are implemented for 3D grids");
        if (slabBegin < 0 || slabEnd > gne[0] || slabEnd - slabBegin < 2) throw std::runtime_error("slab must hold at least two element layers inside the grid");
        wlo = std::max<int64_t>(slabBegin - 1, 0);
        const int64_t whi = std::min<int64_t>(slabEnd + 1, gne[0]);
        ne[0] = whi - wlo;
        s->window = true; s->gne0 = gne[0]; s->xoff = wlo; s->slabBegin = slabBegin; s->slabEnd = slabEnd;
    }
    for (int d = 0; d < dim; ++d) { s->ne[d] = ne[d]; s->nn[d] = ne[d] + 1; }
    s->g = make_grid(dim, ne);
    if (window) {
        GridDesc &g = s->g;
        g.xoff = (int)wlo;
        g.ownLo = (int)(slabBegin - wlo); g.ownHi = (int)((slabEnd == gne[0] ? slabEnd + 1 : slabEnd) - wlo);
        g.cmpLo = (int)(slabBegin - wlo); g.cmpHi = (int)(slabEnd + 1 - wlo);
        g.oeLo = (int)(slabBegin - wlo);  g.oeHi = (int)(slabEnd - wlo);
    }
    if (share) { s->stream = share->stream; s->ownsStream = false; }
    else VF_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->ctx.stream = s->stream; s->ctx.prof = &s->prof;
    s->rho.alloc(s->g.numElems, true); s->E.alloc(s->g.numElems, true);
    s->scratch.alloc(reduce_scratch_doubles(), true); s->scalars.alloc(SC_COUNT, true);
    s->nodeMask.assign(s->g.numNodes, 0);
    s->uploadBCs();
    s->setIsotropic(1.0, 0.0);  // TensorProductSimulator.hh:2114
    s->updateModuli();
    return s.release();
}
int vf_sim_create(int dim, const int64_t *ne, const double *dmin, const double *dmax, vf_sim **out) {
    VF_TRY *out = sim_create_common(dim, ne, dmin, dmax, false, 0, 0, nullptr); VF_CATCH
}
int vf_sim_create_slab(int dim, const int64_t *ne_global, const double *dmin, const double *dmax, int64_t slab_begin, int64_t slab_end,
                       vf_sim *share_stream_with, vf_sim **out) {
    VF_TRY *out = sim_create_common(dim, ne_global, dmin, dmax, true, slab_begin, slab_end, share_stream_with); VF_CATCH
}
int vf_sim_window(const vf_sim *s, int64_t *plane_lo, int64_t *plane_hi, int64_t *own_lo, int64_t *own_hi) {
    VF_TRY *plane_lo = s->xoff; *plane_hi = s->xoff + s->nn[0] - 1; *own_lo = s->xoff + s->g.ownLo; *own_hi = s->xoff + s->g.ownHi - 1; VF_CATCH
}
int vf_sim_destroy(vf_sim *s) { VF_TRY if (s) { if (s->stream) { cudaStreamSynchronize(s->stream); } cudaStream_t st = s->ownsStream ? s->stream : nullptr; delete s; if (st) cudaStreamDestroy(st); } VF_CATCH }
int64_t vf_sim_num_nodes(const vf_sim *s) { return s->g.numNodes; }
int64_t vf_sim_num_elements(const vf_sim *s) { return s->g.numElems; }
int vf_sim_set_elasticity_tensor(vf_sim *s, const double *D) {
    VF_TRY const int fl = s->N == 3 ? 6 : 3; std::memset(s->D, 0, sizeof(s->D));
    for (int i = 0; i < fl; ++i) for (int j = 0; j < fl; ++j) s->D[i][j] = D[i * fl + j];
    s->updateK0(); VF_CATCH
}
int vf_sim_set_isotropic(vf_sim *s, double young, double poisson) { VF_TRY s->setIsotropic(young, poisson); VF_CATCH }
int vf_sim_get_K0(const vf_sim *s, double *out) { VF_TRY std::copy(s->K0.begin(), s->K0.end(), out); VF_CATCH }
int vf_sim_set_interpolation(vf_sim *s, int law, double E_0, double E_min, double gamma, double q) {
    VF_TRY s->law = law; s->E0 = E_0; s->Emin = E_min; s->gamma = gamma; s->q = q; s->updateModuli(); VF_CATCH
}
int vf_sim_set_gravity(vf_sim *s, const double *g) { VF_TRY for (int c = 0; c < s->N; ++c) s->gravity[c] = g[c]; VF_CATCH }
int vf_sim_set_densities(vf_sim *s, const double *rho) {
    VF_TRY h2d(s->rho.p, rho, s->g.numElems, s->stream); s->updateModuli(); VF_CUDA(cudaStreamSynchronize(s->stream)); VF_CATCH
}
int vf_sim_set_uniform_density(vf_sim *s, double rho) {
    VF_TRY if (rho > 1.0 || rho < 0) throw std::runtime_error("Density value (" + std::to_string(rho) + ") has to be in between 0 and 1");
    launch_fill(s->ctx, s->g.numElems, rho, s->rho.p); s->updateModuli(); VF_CATCH
}
int vf_sim_get_densities(const vf_sim *s, double *rho) { VF_TRY d2h(rho, s->rho.p, s->g.numElems, s->stream); VF_CATCH }
int vf_sim_get_young_moduli(const vf_sim *s, double *E) { VF_TRY d2h(E, s->E.p, s->g.numElems, s->stream); VF_CATCH }

int vf_sim_apply_bc_regions(vf_sim *s, int nreg, const int32_t *kind, const int32_t *cmask, const double *values, const double *bmin, const double *bmax) {
    VF_TRY
    if (s->dirNodes.size() + s->forceNodes.size() > 0) throw std::runtime_error("Boundary condition updates unsupported");
    BCBuilder b(*s);
    for (int r = 0; r < nreg; ++r) {
        std::vector<int64_t> idx[3];
        double cnt = 0;
        const bool any = s->boxRanges(bmin + 3 * r, bmax + 3 * r, idx, &cnt);
        const double *val = values + 3 * r;
        if (kind[r] == 1) {
            if (!any) throw std::runtime_error("Force constraint region unmatched");
            double f[3]; for (int c = 0; c < s->N; ++c) f[c] = val[c] / cnt;
            for (int64_t i : idx[0]) for (int64_t j : idx[1]) for (int64_t k : idx[2]) b.setForce(s->flatNode(i, j, k), f);
        } else if (kind[r] == 0) {
            if (!any) throw std::runtime_error("Dirichlet region unmatched");
            for (int64_t i : idx[0]) for (int64_t j : idx[1]) for (int64_t k : idx[2]) b.setDirichlet(s->flatNode(i, j, k), val, (unsigned)cmask[r]);
        } else throw std::runtime_error("Illegal constraint type, only \"dirichlet\" and \"force\" accepted");
    }
    b.apply();
    VF_CATCH
}
int vf_sim_add_dirichlet_box(vf_sim *s, const double *u, const double *bmin, const double *bmax, int cmask) {
    VF_TRY BCBuilder b(*s); std::vector<int64_t> idx[3];
    if (s->boxRanges(bmin, bmax, idx)) for (int64_t i : idx[0]) for (int64_t j : idx[1]) for (int64_t k : idx[2]) b.setDirichlet(s->flatNode(i, j, k), u, (unsigned)cmask);
    b.apply(); VF_CATCH
}
int vf_sim_apply_symmetry_conditions(vf_sim *s, int axes_mask, int max_face_mask) {
    VF_TRY BCBuilder b(*s);
    for (int d = 0; d < s->N; ++d) {
        if (!((axes_mask >> d) & 1)) continue;
        const double target = ((max_face_mask >> d) & 1) ? s->dmax[d] : s->dmin[d];
        for (int64_t n = 0; n < s->g.numNodes; ++n) {
            int64_t c[3]; int64_t r = n; for (int a = s->N - 1; a >= 0; --a) { c[a] = r % s->nn[a]; r /= s->nn[a]; }
            if (std::abs(s->dmin[d] + double(c[d] + (d == 0 ? s->xoff : 0)) * s->spacing[d] - target) < 1e-10) b.setDirichletComponent(n, d, 0.0);
        }
    }
    b.apply(); VF_CATCH
}
// number of Dirichlet components with a non-zero prescribed value (m_dirichletNodeDisplacements, TensorProductSimulator.hh:575-593)
int64_t vf_sim_num_nonzero_dirichlet_values(const vf_sim *s) {
    int64_t n = 0;
    for (size_t k = 0; k < s->dirNodes.size(); ++k) for (int c = 0; c < s->N; ++c) if (((s->dirMask[k] >> c) & 1) && s->dirVals[k * s->N + c] != 0) ++n;
    return n;
}
int vf_sim_get_dirichlet_mask(const vf_sim *s, uint8_t *m) { VF_TRY std::copy(s->nodeMask.begin(), s->nodeMask.end(), m); VF_CATCH }
int64_t vf_sim_num_force_nodes(const vf_sim *s) { return (int64_t)s->forceNodes.size(); }
int64_t vf_sim_num_dirichlet_nodes(const vf_sim *s) { return (int64_t)s->dirNodes.size(); }
// nodes[n], masks[n] (bit c: component c constrained), values[n * N] in the order the conditions are stored (ascending node index)
int vf_sim_get_dirichlet_conditions(const vf_sim *s, int64_t *nodes, uint8_t *masks, double *values) {
    VF_TRY
    std::copy(s->dirNodes.begin(), s->dirNodes.end(), nodes); std::copy(s->dirMask.begin(), s->dirMask.end(), masks);
    std::copy(s->dirVals.begin(), s->dirVals.end(), values);
    VF_CATCH
}
// nodes[n], forces[n * N]: the per-node forces of the "force" / "traction" conditions (TensorProductSimulator.hh:617-632)
int vf_sim_get_force_nodes(const vf_sim *s, int64_t *nodes, double *forces) {
    VF_TRY
    std::copy(s->forceNodes.begin(), s->forceNodes.end(), nodes); std::copy(s->forceVals.begin(), s->forceVals.end(), forces);
    VF_CATCH
}
int vf_sim_build_load_vector_dev(vf_sim *s, double *f_dev) { VF_TRY sim_build_load_dev(*s, f_dev); VF_CATCH }
int vf_sim_build_load_vector(vf_sim *s, double *f) {
    VF_TRY double *t = sim_tmp(s, 0); sim_build_load_dev(*s, t); d2h(f, t, (size_t)s->g.numNodes * s->N, s->stream); VF_CATCH
}
int vf_sim_apply_K(vf_sim *s, const double *u, double *out, int zero_init, int negate) {
    VF_TRY
    const size_t len = (size_t)s->g.numNodes * s->N;
    double *du = sim_tmp(s, 0), *dout = sim_tmp(s, 1);
    h2d(du, u, len, s->stream);
    if (!zero_init) h2d(dout, out, len, s->stream);
    launch_apply_l0(s->ctx, s->g, s->K0p, du, s->E.p, nullptr, nullptr, dout, zero_init ? APPLY_SET : (negate ? APPLY_SUB : APPLY_ADD), nullptr, nullptr);
    if (zero_init && negate) launch_scale(s->ctx, (long long)len, -1.0, dout);
    d2h(out, dout, len, s->stream);
    VF_CATCH
}
int vf_sim_set_mask_layer(vf_sim *s, int64_t layer) {
    VF_TRY // setFabricationMaskHeightByLayer -> setFabricationMaskHeight (:290-309, 327-329)
    const double h = s->spacing[1] * double(layer);
    if (h < 0 || h > s->dmax[1]) throw std::runtime_error("Fabrication height (" + std::to_string(h) + ") has to be in between 0 and " + std::to_string(s->dmax[1]));
    s->maskHeight = h;
    s->firstMasked = (int)std::ceil(h / s->stretch[1] - 1e-10);
    s->firstDetached = s->firstMasked + 1;
    s->refreshGridMask(); s->updateModuli(); ++s->structVersion;
    VF_CATCH
}
int vf_sim_get_mask_info(const vf_sim *s, int64_t *fm, int64_t *fd, double *h) { VF_TRY if (fm) *fm = s->firstMasked; if (fd) *fd = s->firstDetached; if (h) *h = s->maskHeight; VF_CATCH }
int vf_sim_compliance_gradient(vf_sim *s, const double *u, double *g, int accumulate) {
    VF_TRY
    const size_t len = (size_t)s->g.numNodes * s->N;
    double *du = sim_tmp(s, 0); h2d(du, u, len, s->stream);
    DevBuf<double> dg; dg.alloc(s->g.numElems, !accumulate);
    if (accumulate) h2d(dg.p, g, s->g.numElems, s->stream);
    launch_compliance_gradient(s->ctx, s->g, s->K0p, du, s->rho.p, dg.p, s->law, s->E0, s->Emin, s->gamma, s->q, s->gravity, s->elemVolume(), accumulate != 0);
    d2h(g, dg.p, s->g.numElems, s->stream);
    VF_CATCH
}
int vf_sim_element_energy_density(vf_sim *s, const double *u, double *out) {
    VF_TRY
    double *du = sim_tmp(s, 0); h2d(du, u, (size_t)s->g.numNodes * s->N, s->stream);
    DevBuf<double> dg; dg.alloc(s->g.numElems, false);
    launch_energy_density(s->ctx, s->g, s->K0p, du, s->E.p, dg.p);
    d2h(out, dg.p, s->g.numElems, s->stream);
    VF_CATCH
}
int vf_sim_solve(vf_sim *s, const double *f, double *u) {
    VF_TRY
    const size_t len = (size_t)s->g.numNodes * s->N;
    double *df = sim_tmp(s, 0), *du = sim_tmp(s, 1);
    h2d(df, f, len, s->stream);
    sim_direct_solve(*s, df, du);
    d2h(u, du, len, s->stream);
    VF_CATCH
}

// ---- multigrid -----------------------------------------------------------------------------
// Dirichlet coarsening (MultigridSolver.hh:58-103): a fine Dirichlet node constrains, with the same component mask and
// zero value, every coarse node on the vertex/edge/face/cell of the coarse element it lies on.  Index arithmetic is done
// on global plane indices so that it also serves slab windows (coarse nodes outside the window are skipped).
static void coarsen_dirichlet_mask(int N, const MGLevel &F, const std::vector<uint8_t> &fineMask, MGLevel &L) {
    L.nodeMask.assign(L.g.numNodes, 0);
    int64_t fnn[3] = {1, 1, 1}, cnn[3] = {1, 1, 1};
    for (int d = 0; d < N; ++d) { fnn[d] = F.ne[d] + 1; cnn[d] = L.ne[d] + 1; }
    const int64_t fo = F.g.xoff, co = L.g.xoff;
    for (int64_t fn = 0; fn < F.g.numNodes; ++fn) {
        const uint8_t m = fineMask[fn]; if (!m) continue;
        int64_t c[3] = {0, 0, 0}; { int64_t r = fn; for (int d = N - 1; d >= 0; --d) { c[d] = r % fnn[d]; r /= fnn[d]; } }
        if (N == 3) c[0] += fo;
        for (int k = 0; k < (1 << N); ++k) {
            int64_t cn = 0; bool ok = true;
            for (int d = 0; d < N; ++d) {
                const int bit = (k >> d) & 1;
                if ((c[d] & 1) == 0 && bit) { ok = false; break; }
                int64_t q = (c[d] >> 1) + bit;
                if (N == 3 && d == 0) q -= co;
                if (q < 0 || q >= cnn[d]) { ok = false; break; }
                cn = cn * cnn[d] + q;
            }
            if (ok) L.nodeMask[cn] |= m;
        }
    }
}

static vf_mg *mg_create_common(vf_sim *fine, int levels, int firstRep) {
    if (levels < 0) throw std::runtime_error("numCoarseningLevels must be >= 0");
    auto mg = std::make_unique<vf_mg>();
    mg->sim = fine; mg->N = fine->N; mg->ctx = fine->ctx;
    const int N = fine->N;
    const bool slab = fine->window;
    if (slab) {
        if (levels < 1) throw std::runtime_error("a slab-partitioned solver needs at least one coarsening level");
        firstRep = std::min(std::max(firstRep, 1), levels);     // the coarsest level is always replicated
        mg->firstRep = firstRep;
    }
    int64_t gne[3] = {slab ? fine->gne0 : fine->ne[0], fine->ne[1], fine->ne[2]};
    for (int l = 0; l <= levels; ++l) {
        auto L = std::make_unique<MGLevel>();
        if (l > 0) {
            for (int d = 0; d < N; ++d) {
                if (gne[d] % 2 == 1) throw std::runtime_error("Grid size currently must be divisible by 2^numCoarseningLevels (nonuniform coarsening not yet implemented)");
                gne[d] /= 2;
            }
        }
        int64_t ne[3] = {gne[0], gne[1], gne[2]};
        L->gne0 = gne[0];
        if (slab) {
            const int64_t div = int64_t(1) << l;
            const bool win = l < firstRep;
            if (l <= firstRep && (fine->slabBegin % div || fine->slabEnd % div)) throw std::runtime_error("slab boundaries must be multiples of 2^(first replicated level)");
            L->sb = fine->slabBegin / div; L->se = fine->slabEnd / div; L->windowed = win;
            if (win && L->se - L->sb < 1) throw std::runtime_error("slab too thin for the requested number of windowed levels");
            const int64_t wlo = win ? std::max<int64_t>(L->sb - 1, 0) : 0, whi = win ? std::min<int64_t>(L->se + 1, gne[0]) : gne[0];
            ne[0] = whi - wlo;
            L->g = make_grid(N, ne);
            GridDesc &g = L->g;
            g.xoff = (int)wlo;
            if (l <= firstRep) { g.ownLo = (int)(L->sb - wlo); g.ownHi = (int)((L->se == gne[0] ? L->se + 1 : L->se) - wlo); }
            if (win) { g.cmpLo = (int)(L->sb - wlo); g.cmpHi = (int)(L->se + 1 - wlo); g.oeLo = (int)(L->sb - wlo); g.oeHi = (int)(L->se - wlo); }
            if (l == 0 && (g.xoff != fine->g.xoff || g.nn[0] != fine->g.nn[0])) throw std::logic_error("slab window mismatch");
        } else {
            L->g = make_grid(N, ne);
        }
        for (int d = 0; d < 3; ++d) L->ne[d] = ne[d];
        L->stretchBD = (fine->dmax[1] - fine->dmin[1]) / double(ne[1]);
        const size_t len = (size_t)L->g.numNodes * N;
        L->x.alloc(len, true); L->b.alloc(len, true); L->r.alloc(len, true);
        if (l == 0) L->nodeMask = fine->nodeMask;
        else {
            coarsen_dirichlet_mask(N, *mg->lv.back(), mg->lv.back()->nodeMask, *L);
            L->dmask.alloc(L->g.numNodes, false); L->dmask.upload(L->nodeMask.data(), L->g.numNodes, fine->stream);
            static const bool usePosTab = [] { const char *e = std::getenv("VF_ST_POSTAB"); return !(e && e[0] == '0'); }();
            if (usePosTab && !stencil_level_streams(L->g)) {   // measured: helps the latency-bound levels, costs 2-6 % on the streaming one (profiles/r06f_ab.log)
                L->posTab.alloc((size_t)L->g.numPos, false);
                launch_fill_pos_table(fine->stream, L->g, L->posTab.p);
            }
            VF_CUDA(cudaStreamSynchronize(fine->stream));
        }
        mg->lv.push_back(std::move(L));
    }
    // coarsenedFineK0s[fi] = Phi_fi^T K0 Phi_fi (MultigridSolver.hh:116-120, 664-687, 711-722)
    const int npe = 1 << N, ke = N * npe;
    mg->cK0.assign((size_t)npe * ke * ke, 0.0);
    for (int fi = 0; fi < npe; ++fi) {
        std::vector<double> phi((size_t)npe * npe);
        for (int fn = 0; fn < npe; ++fn) for (int cn = 0; cn < npe; ++cn) {
            double v = 1;
            for (int d = 0; d < N; ++d) {
                const double pos = 0.5 * ((fn >> (N - 1 - d)) & 1) + 0.5 * ((fi >> (N - 1 - d)) & 1);
                v *= ((cn >> (N - 1 - d)) & 1) ? pos : (1.0 - pos);
            }
            phi[fn * npe + cn] = v;
        }
        std::vector<double> T((size_t)ke * ke, 0.0);
        for (int a = 0; a < ke; ++a) for (int j = 0; j < npe; ++j) for (int dc = 0; dc < N; ++dc) {
            double sacc = 0; for (int i = 0; i < npe; ++i) sacc += fine->K0[(size_t)a * ke + (N * i + dc)] * phi[i * npe + j];
            T[(size_t)a * ke + (N * j + dc)] = sacc;
        }
        double *Kc = &mg->cK0[(size_t)fi * ke * ke];
        for (int j = 0; j < npe; ++j) for (int c = 0; c < N; ++c) for (int bcol = 0; bcol < ke; ++bcol) {
            double sacc = 0; for (int i = 0; i < npe; ++i) sacc += phi[i * npe + j] * T[(size_t)(N * i + c) * ke + bcol];
            Kc[(size_t)(N * j + c) * ke + bcol] = sacc;
        }
    }
    mg->cK0dev.alloc(mg->cK0.size(), false); mg->cK0dev.upload(mg->cK0.data(), mg->cK0.size(), fine->stream);
    VF_CUDA(cudaStreamSynchronize(fine->stream));
    const size_t len0 = (size_t)fine->g.numNodes * N;
    mg->Ad.alloc(len0, true); mg->d.alloc(len0, true);
    mg->scalars.alloc(SC_COUNT, true); mg->scratch.alloc(reduce_scratch_doubles(), true); mg->gridBar.alloc(2, true);
    VF_CUDA(cudaMallocHost(&mg->hostScalars, SC_COUNT * sizeof(double)));
    mg_sync_level_masks(*mg);
    return mg.release();
}
int vf_mg_create(vf_sim *fine, int levels, vf_mg **out) {
    VF_TRY if (fine->window) throw std::runtime_error("use vf_mg_create_slab for a slab-window simulator");
    *out = mg_create_common(fine, levels, INT_MAX); VF_CATCH
}
int vf_mg_create_slab(vf_sim *fine, int levels, int first_replicated_level, vf_mg **out) {
    VF_TRY if (!fine->window) throw std::runtime_error("vf_mg_create_slab needs a simulator created with vf_sim_create_slab");
    *out = mg_create_common(fine, levels, first_replicated_level); VF_CATCH
}
// ---- slab groups ---------------------------------------------------------------------------
namespace {
// Dirichlet masks of the coarse levels must also be right on the ghost planes (they are applied pointwise to fields whose
// ghost copies have to stay identical to the owner's values) and on the replicated levels (every part only saw its own
// window).  Done once at group creation through the ordinary field exchanges: bit c of the mask travels as component c.
__global__ void k_mask_to_field(long long nn, int N, const uint8_t *__restrict__ m, double *__restrict__ f) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < nn) for (int c = 0; c < N; ++c) f[c * nn + n] = ((m[n] >> c) & 1) ? 1.0 : 0.0;
}
__global__ void k_field_to_mask(long long nn, int N, const double *__restrict__ f, uint8_t *__restrict__ m) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < nn) { unsigned v = 0; for (int c = 0; c < N; ++c) if (f[c * nn + n] > 0.5) v |= 1u << c; m[n] = (uint8_t)v; }
}
void group_complete_masks(vf_mg &lead) {
    std::vector<vf_mg *> &P = parts_of(lead);
    const int nl = lead.numLevels(), N = lead.N;
    for (int l = 1; l < nl; ++l) {
        for (vf_mg *m : P) {
            MGLevel &L = *m->lv[l];
            coarsen_dirichlet_mask(N, *m->lv[l - 1], m->lv[l - 1]->nodeMask, L);   // from the (now complete) finer mask
            L.dmask.upload(L.nodeMask.data(), L.g.numNodes, m->ctx.stream);
            k_mask_to_field<<<(unsigned)((L.g.numNodes + 255) / 256), 256, 0, m->ctx.stream>>>(L.g.numNodes, N, L.dmask.p, L.x.p);
            VF_KERNEL_CHECK();
        }
        if (l < lead.firstRep) grp_exchange(lead, l, fx(l));
        else if (l == lead.firstRep) {
            for (vf_mg *m : P) zero_unowned_planes(*m, l, lx(*m, l));
            grp_allreduce(lead, fx(l), (size_t)lead.grid(l).numNodes * N);
        }
        for (vf_mg *m : P) {
            MGLevel &L = *m->lv[l];
            k_field_to_mask<<<(unsigned)((L.g.numNodes + 255) / 256), 256, 0, m->ctx.stream>>>(L.g.numNodes, N, L.x.p, L.dmask.p);
            VF_KERNEL_CHECK();
            L.dmask.download(L.nodeMask.data(), L.g.numNodes, m->ctx.stream);
            VF_CUDA(cudaMemsetAsync(L.x.p, 0, sizeof(double) * L.g.numNodes * N, m->ctx.stream));
        }
    }
    for (vf_mg *m : P) VF_CUDA(cudaStreamSynchronize(m->ctx.stream));
}
void group_validate(vf_group &G) {
    if (G.parts.empty()) throw std::runtime_error("empty slab group");
    for (vf_mg *m : G.parts) {
        if (!m->sim->window) throw std::runtime_error("slab groups are made of solvers created with vf_mg_create_slab");
        if (m->numLevels() != G.parts[0]->numLevels() || m->firstRep != G.parts[0]->firstRep) throw std::runtime_error("all parts of a slab group need the same hierarchy");
        if (m->grp) throw std::runtime_error("solver already belongs to a slab group");
    }
}
} // namespace

int vf_group_create_local(int nparts, vf_mg **parts, vf_group **out) {
    VF_TRY
    auto G = std::make_unique<vf_group>();
    G->parts.assign(parts, parts + nparts); G->world = nparts;
    group_validate(*G);
    for (int i = 0; i < nparts; ++i) {
        if (parts[i]->ctx.stream != parts[0]->ctx.stream) throw std::runtime_error("the parts of a local slab group must share one stream (vf_sim_create_slab share_stream_with)");
        if (i > 0 && parts[i]->sim->slabBegin != parts[i - 1]->sim->slabEnd) throw std::runtime_error("parts must be consecutive slabs");
    }
    if (parts[0]->sim->slabBegin != 0 || parts[nparts - 1]->sim->slabEnd != parts[0]->sim->gne0) throw std::runtime_error("the slabs must cover the grid");
    for (int i = 0; i < nparts; ++i) parts[i]->grp = G.get();
    group_complete_masks(*parts[0]);
    *out = G.release();
    VF_CATCH
}
int vf_nccl_unique_id(void *out128) { VF_TRY NcclApi &A = NcclApi::get(); A.check(A.GetUniqueId(reinterpret_cast<NcclUniqueId *>(out128)), "ncclGetUniqueId"); VF_CATCH }
int vf_group_create_nccl(vf_mg *part, int rank, int world, const void *unique_id128, vf_group **out) {
    VF_TRY
    auto G = std::make_unique<vf_group>();
    G->parts.assign(1, part); G->rank = rank; G->world = world;
    group_validate(*G);
    if ((rank == 0) != (part->sim->slabBegin == 0) || (rank == world - 1) != (part->sim->slabEnd == part->sim->gne0)) throw std::runtime_error("slab does not match the rank's position");
    NcclApi &A = NcclApi::get();
    NcclUniqueId id; std::memcpy(&id, unique_id128, sizeof(id));
    A.check(A.CommInitRank(&G->comm, world, id, rank), "ncclCommInitRank");
    part->grp = G.get();
    group_complete_masks(*part);
    p2p_setup(*G, *part);
    *out = G.release();
    VF_CATCH
}
int vf_group_destroy(vf_group *g) {
    VF_TRY if (g) {
        for (vf_mg *m : g->parts) {
            cudaStreamSynchronize(m->ctx.stream);
            // a captured preconditioner of an NCCL rank holds the communicator's all-reduce nodes: ncclCommDestroy waits for them
            if (m->pg.exec) { cudaGraphExecDestroy(m->pg.exec); m->pg.exec = nullptr; }
            if (m->sg.exec) { cudaGraphExecDestroy(m->sg.exec); m->sg.exec = nullptr; }
            m->grp = nullptr;
        }
        delete g;
    } VF_CATCH
}
int vf_group_pcg_dev(vf_group *g, double *const *x_dev, const double *const *b_dev, int maxIter, double tol, int mgIt, int mgSmooth, int fmg,
                     int dirichletOK, int *iters, double *residualNorms, vf_pcg_callback cb, void *user) {
    VF_TRY vf_mg &lead = *g->parts[0];
    mg_pcg(lead, x_dev, b_dev, maxIter, tol, mgIt, mgSmooth, fmg != 0, dirichletOK != 0, cb, user);
    p2p_check(*g, lead.ctx.stream);
    if (iters) *iters = lead.lastIters;
    if (residualNorms) std::copy(lead.lastResiduals.begin(), lead.lastResiduals.end(), residualNorms);
    VF_CATCH
}

int vf_mg_destroy(vf_mg *mg) { VF_TRY if (mg) { cudaStreamSynchronize(mg->ctx.stream); delete mg; } VF_CATCH }
int vf_mg_num_levels(const vf_mg *mg) { return mg->numLevels(); }
int64_t vf_mg_level_num_nodes(const vf_mg *mg, int l) { return mg->grid(l).numNodes; }
int vf_mg_level_grid(const vf_mg *mg, int l, int64_t *ne) { VF_TRY for (int d = 0; d < mg->N; ++d) ne[d] = mg->lv.at(l)->ne[d]; VF_CATCH }
int vf_mg_level_dirichlet_mask(const vf_mg *mg, int l, uint8_t *m) {
    VF_TRY const std::vector<uint8_t> &src = (l == 0) ? mg->sim->nodeMask : mg->lv.at(l)->nodeMask; std::copy(src.begin(), src.end(), m); VF_CATCH
}
int vf_mg_get_coarsened_fine_K0(const vf_mg *mg, int fi, double *out) {
    VF_TRY const int ke = mg->N * (1 << mg->N); std::copy(mg->cK0.begin() + (size_t)fi * ke * ke, mg->cK0.begin() + (size_t)(fi + 1) * ke * ke, out); VF_CATCH
}
int vf_mg_update_stiffness_matrices(vf_mg *mg) { VF_TRY mg_update_stiffness(*mg, true); VF_CUDA(cudaStreamSynchronize(mg->ctx.stream)); VF_CATCH }
int vf_mg_apply_K(vf_mg *mg, int l, const double *u, double *out) {
    VF_TRY const size_t len = (size_t)mg->grid(l).numNodes * mg->N;
    double *du = mg_tmp(mg, 0, len), *dout = mg_tmp(mg, 1, len);
    h2d(du, u, len, mg->ctx.stream); mg_sync_level_masks(*mg);
    mg_apply_K(*mg, l, fu(du), fu(du), fu(dout), APPLY_SET, false);
    d2h(out, dout, len, mg->ctx.stream); VF_CATCH
}
int vf_mg_compute_residual(vf_mg *mg, int l, const double *u, const double *b, double *r) {
    VF_TRY const size_t len = (size_t)mg->grid(l).numNodes * mg->N;
    double *du = mg_tmp(mg, 0, len), *db = mg_tmp(mg, 1, len), *dr = mg_tmp(mg, 2, len);
    h2d(du, u, len, mg->ctx.stream); h2d(db, b, len, mg->ctx.stream); mg_sync_level_masks(*mg);
    VF_CUDA(cudaMemsetAsync(dr, 0, len * sizeof(double), mg->ctx.stream));
    mg_residual(*mg, l, fu(du), fu(db), fu(dr));
    d2h(r, dr, len, mg->ctx.stream); VF_CATCH
}
int vf_mg_smooth(vf_mg *mg, int l, double *u, const double *b, int forward) {
    VF_TRY const size_t len = (size_t)mg->grid(l).numNodes * mg->N;
    double *du = mg_tmp(mg, 0, len), *db = mg_tmp(mg, 1, len);
    h2d(du, u, len, mg->ctx.stream); h2d(db, b, len, mg->ctx.stream); mg_sync_level_masks(*mg);
    mg_smooth(*mg, l, fu(du), fu(db), forward != 0);
    d2h(u, du, len, mg->ctx.stream); VF_CATCH
}
int vf_mg_smooth_residual(vf_mg *mg, int l, double *u, const double *b, int forward, double *r) {
    VF_TRY
    mg_sync_level_masks(*mg);
    if (!(l > 0 && l < mg->numLevels() && !mg->grp && gs_residual_fusable(mg->grid(l)))) throw std::runtime_error("vf_mg_smooth_residual: the residual-emitting sweep is not available on this level");
    const size_t len = (size_t)mg->grid(l).numNodes * mg->N;
    double *du = mg_tmp(mg, 0, len), *db = mg_tmp(mg, 1, len), *dr = mg_tmp(mg, 2, len);
    h2d(du, u, len, mg->ctx.stream); h2d(db, b, len, mg->ctx.stream);
    VF_CUDA(cudaMemsetAsync(dr, 0xff, len * sizeof(double), mg->ctx.stream));   // NaN pattern: every entry must be written by the sweep
    const Field rf = fu(dr);
    mg_smooth(*mg, l, fu(du), fu(db), forward != 0, &rf);
    launch_zero_dirichlet(mg->ctx, mg->grid(l), mg->dmask(l), dr);
    d2h(u, du, len, mg->ctx.stream); d2h(r, dr, len, mg->ctx.stream);
    VF_CATCH
}
int vf_mg_restrict(vf_mg *mg, int lf, const double *fine, double *coarse) {
    VF_TRY const size_t lenF = (size_t)mg->grid(lf).numNodes * mg->N, lenC = (size_t)mg->grid(lf + 1).numNodes * mg->N;
    double *df = mg_tmp(mg, 0, lenF), *dc = mg_tmp(mg, 1, lenC);
    h2d(df, fine, lenF, mg->ctx.stream); h2d(dc, coarse, lenC, mg->ctx.stream); mg_sync_level_masks(*mg);
    launch_restrict(mg->ctx, mg->grid(lf), mg->grid(lf + 1), df, dc);
    d2h(coarse, dc, lenC, mg->ctx.stream); VF_CATCH
}
int vf_mg_interpolate(vf_mg *mg, int lf, const double *coarse, double *fine, int accumulate) {
    VF_TRY const size_t lenF = (size_t)mg->grid(lf).numNodes * mg->N, lenC = (size_t)mg->grid(lf + 1).numNodes * mg->N;
    double *df = mg_tmp(mg, 0, lenF), *dc = mg_tmp(mg, 1, lenC);
    h2d(dc, coarse, lenC, mg->ctx.stream); h2d(df, fine, lenF, mg->ctx.stream); mg_sync_level_masks(*mg);
    launch_prolong(mg->ctx, mg->grid(lf), mg->grid(lf + 1), dc, df, accumulate != 0);
    d2h(fine, df, lenF, mg->ctx.stream); VF_CATCH
}
int vf_mg_get_stencil(vf_mg *mg, int l, double *out) {
    VF_TRY if (l < 1 || l >= mg->numLevels()) throw std::runtime_error("stencils are stored for levels 1..numLevels-1");
    mg_update_stiffness(*mg);
    const GridDesc &g = mg->grid(l); const int ns = mg->N == 3 ? 27 : 9, NN = mg->N * mg->N;
    std::vector<double> h((size_t)g.numPos * ns * NN);
    d2h(h.data(), mg->lv[l]->S.p, h.size(), mg->ctx.stream);
    for (int c0 = 0; c0 < g.nn[0]; ++c0) for (int c1 = 0; c1 < g.nn[1]; ++c1) for (int c2 = 0; c2 < g.nn[2]; ++c2) {
        const long long n = (long long)c0 * g.ns[0] + (long long)c1 * g.ns[1] + c2, p = stencil_pos(g, c0, c1, c2);
        for (int s = 0; s < ns; ++s) for (int i = 0; i < NN; ++i) out[((size_t)n * ns + s) * NN + i] = h[(size_t)stencil_addr(p, s * NN + i, ns * NN)];
    }
    VF_CATCH
}
int vf_mg_coarse_solve(vf_mg *mg, const double *f, double *x) {
    VF_TRY const int l = mg->numLevels() - 1; const size_t len = (size_t)mg->grid(l).numNodes * mg->N;
    double *df = mg_tmp(mg, 0, len), *dx = mg_tmp(mg, 1, len);
    h2d(df, f, len, mg->ctx.stream); mg_sync_level_masks(*mg);
    if (l == 0) sim_direct_solve(*mg->sim, df, dx); else mg_coarse_solve(*mg, fu(df), fu(dx));
    d2h(x, dx, len, mg->ctx.stream); VF_CATCH
}
int vf_mg_solve(vf_mg *mg, const double *u, const double *f, int numSteps, int nsmooth, int stiffnessUpdated, int zeroDirichlet, int fmg, double *out) {
    VF_TRY (void)stiffnessUpdated; // coarse operators are version-tracked; stale ones are always rebuilt
    const size_t len = (size_t)mg->grid(0).numNodes * mg->N;
    mg_sync_level_masks(*mg);
    h2d(lx(*mg, 0), u, len, mg->ctx.stream);
    if (numSteps > 0) {
        h2d(lb(*mg, 0), f, len, mg->ctx.stream);
        if (mg->numLevels() == 1) sim_direct_solve(*mg->sim, lb(*mg, 0), lx(*mg, 0));
        else { mg_update_stiffness(*mg); mg_solve_inplace(*mg, numSteps, nsmooth, zeroDirichlet != 0, fmg != 0); }
    }
    d2h(out, lx(*mg, 0), len, mg->ctx.stream); VF_CATCH
}
int vf_mg_pcg_dev(vf_mg *mg, double *x, const double *b, int maxIter, double tol, int mgIt, int mgSmooth, int fmg, int dirichletOK,
                  int *outIters, double *resNorms, vf_pcg_callback cb, void *user) {
    VF_TRY mg_pcg(*mg, x, b, maxIter, tol, mgIt, mgSmooth, fmg != 0, dirichletOK != 0, cb, user);
    VF_CUDA(cudaStreamSynchronize(mg->ctx.stream));
    if (outIters) *outIters = mg->lastIters;
    if (resNorms) std::copy(mg->lastResiduals.begin(), mg->lastResiduals.end(), resNorms);
    VF_CATCH
}
int vf_mg_pcg(vf_mg *mg, double *x, const double *b, int maxIter, double tol, int mgIt, int mgSmooth, int fmg, int dirichletOK,
              int *outIters, double *resNorms, vf_pcg_callback cb, void *user) {
    VF_TRY const size_t len = (size_t)mg->grid(0).numNodes * mg->N;
    double *dx = mg_tmp(mg, 0, len), *db = mg_tmp(mg, 1, len);
    h2d(dx, x, len, mg->ctx.stream); h2d(db, b, len, mg->ctx.stream);
    mg_pcg(*mg, dx, db, maxIter, tol, mgIt, mgSmooth, fmg != 0, dirichletOK != 0, cb, user);
    d2h(x, dx, len, mg->ctx.stream);
    if (outIters) *outIters = mg->lastIters;
    if (resNorms) std::copy(mg->lastResiduals.begin(), mg->lastResiduals.end(), resNorms);
    VF_CATCH
}
// Out-of-place form, as the reference's Python binding calls it (VoxelFEM.cc:174-186: `VField x = u; pcg(x, ...); return x`):
// the initial guess u0 is read, the solution is written to x_out; neither host buffer needs preparing between solves.
int vf_mg_pcg_io(vf_mg *mg, const double *u0, const double *b, double *x_out, int maxIter, double tol, int mgIt, int mgSmooth, int fmg, int dirichletOK,
                 int *outIters, double *resNorms, vf_pcg_callback cb, void *user) {
    VF_TRY const size_t len = (size_t)mg->grid(0).numNodes * mg->N;
    double *dx = mg_tmp(mg, 0, len), *db = mg_tmp(mg, 1, len);
    // The copies run on their own stream so that the hierarchy rebuild (which depends on the moduli only) overlaps them: at 256^3 the
    // 6.6 ms rebuild hides behind 15 ms of PCIe traffic.
    if (!mg->ioStream) {
        VF_CUDA(cudaStreamCreateWithFlags(&mg->ioStream, cudaStreamNonBlocking));
        VF_CUDA(cudaEventCreateWithFlags(&mg->ioEvA, cudaEventDisableTiming)); VF_CUDA(cudaEventCreateWithFlags(&mg->ioEvB, cudaEventDisableTiming));
    }
    VF_CUDA(cudaEventRecord(mg->ioEvA, mg->ctx.stream)); VF_CUDA(cudaStreamWaitEvent(mg->ioStream, mg->ioEvA, 0));   // earlier work on the buffers is done
    VF_CUDA(cudaMemcpyAsync(dx, u0, len * sizeof(double), cudaMemcpyHostToDevice, mg->ioStream));
    VF_CUDA(cudaMemcpyAsync(db, b, len * sizeof(double), cudaMemcpyHostToDevice, mg->ioStream));
    VF_CUDA(cudaEventRecord(mg->ioEvB, mg->ioStream));
    if (mgIt > 0 && mgSmooth > 0 && mg->numLevels() > 1 && !mg->grp) {
        mg_sync_level_masks(*mg);
        mg_update_stiffness(*mg, mg->rebuildEverySolve);
        mg->rebuiltForThisSolve = true;
    }
    VF_CUDA(cudaStreamWaitEvent(mg->ctx.stream, mg->ioEvB, 0));
    mg_pcg(*mg, dx, db, maxIter, tol, mgIt, mgSmooth, fmg != 0, dirichletOK != 0, cb, user);
    d2h(x_out, dx, len, mg->ctx.stream);
    if (outIters) *outIters = mg->lastIters;
    if (resNorms) std::copy(mg->lastResiduals.begin(), mg->lastResiduals.end(), resNorms);
    VF_CATCH
}
int vf_mg_get_pcg_residual(vf_mg *mg, double *r) { VF_TRY d2h(r, lb(*mg, 0), (size_t)mg->grid(0).numNodes * mg->N, mg->ctx.stream); VF_CATCH }
int vf_mg_get_pcg_iterate(vf_mg *mg, double *x) {
    VF_TRY if (!mg->pcgX) throw std::runtime_error("no PCG solve has run on this solver");
    d2h(x, mg->pcgX, (size_t)mg->grid(0).numNodes * mg->N, mg->ctx.stream); VF_CATCH
}
int vf_mg_set_symmetric_gauss_seidel(vf_mg *mg, int s) { mg->symmetricGS = s != 0; return 0; }
int vf_mg_set_rebuild_every_solve(vf_mg *mg, int on) { mg->rebuildEverySolve = on != 0; return 0; }
int vf_mg_set_mask_layer(vf_mg *mg, int64_t layer) { if (int rc = vf_sim_set_mask_layer(mg->sim, layer)) return rc; VF_TRY mg_sync_level_masks(*mg); VF_CATCH }
int vf_mg_decrement_mask(vf_mg *mg, int inc) {
    VF_TRY // decrementFabricationMaskHeightByLayer (TensorProductSimulator.hh:311-324; MultigridSolver.hh:1030-1036)
    vf_sim &s = *mg->sim;
    if ((int64_t)s.firstMasked > s.ne[1]) throw std::runtime_error("Mask must already be applied");
    if (s.firstMasked < inc) throw std::runtime_error("Mask decrement of bounds");
    const bool pendingBand = mg->bandActive && mg->bandVersion == s.version;
    const bool upToDate = mg->stiffnessVersion == s.version || pendingBand;
    s.maskHeight -= inc * s.spacing[1];
    s.firstMasked -= inc; s.firstDetached = s.firstMasked + 1;
    s.refreshGridMask();
    launch_zero_moduli_layers(s.ctx, s.g, s.E.p, s.firstMasked, s.firstMasked + inc);
    s.touch(); mg_sync_level_masks(*mg);
    if (upToDate && !mg->grp && mg->numLevels() > 1 && mg->lv[1]->S.n) {   // only these element layers changed since the hierarchy was built
        mg->bandLo = pendingBand ? std::min(mg->bandLo, s.firstMasked) : s.firstMasked;
        mg->bandHi = pendingBand ? std::max(mg->bandHi, s.firstMasked + inc) : s.firstMasked + inc;
        mg->bandActive = true; mg->bandVersion = s.version;
    } else mg->bandActive = false;
    VF_CATCH
}
int vf_mg_debug_get(vf_mg *mg, int which, int l, double *out) {
    VF_TRY MGLevel &L = *mg->lv.at(l); const double *p = which == 0 ? L.x.p : (which == 1 ? L.b.p : L.r.p);
    d2h(out, p, (size_t)L.g.numNodes * mg->N, mg->ctx.stream); VF_CATCH
}
int vf_mg_debug_multicolor_visit(vf_mg *mg, int32_t *order) {
    VF_TRY // host restatement of the visit order for the debug exposer (:444-450); colour-major, row-major within a colour
    const GridDesc &g = mg->sim->g; int32_t i = 0;
    for (int color = 0; color < (1 << mg->N); ++color) {
        ColorDesc col; if (!make_color(g, color, col)) continue;
        for (int i0 = 0; i0 < col.cnt[0]; ++i0) for (int i1 = 0; i1 < col.cnt[1]; ++i1) for (int i2 = 0; i2 < col.cnt[2]; ++i2)
            order[(long long)(col.off[0] + 2 * i0) * g.ns[0] + (long long)(col.off[1] + 2 * i1) * g.ns[1] + (col.off[2] + 2 * i2)] = i++;
    }
    VF_CATCH
}

// Device-buffer helpers: complete on return (they run on the legacy default stream, which the non-blocking solver streams are not ordered against).
int vf_dev_alloc(size_t n, double **out) { VF_TRY ensure_device(); VF_CUDA(cudaMalloc(out, n * sizeof(double))); VF_CUDA(cudaMemset(*out, 0, n * sizeof(double))); VF_CUDA(cudaStreamSynchronize(0)); VF_CATCH }
int vf_dev_free(double *p) { VF_TRY VF_CUDA(cudaFree(p)); VF_CATCH }
int vf_dev_upload(double *dev, const double *host, size_t n) { VF_TRY VF_CUDA(cudaMemcpy(dev, host, n * sizeof(double), cudaMemcpyHostToDevice)); VF_CUDA(cudaStreamSynchronize(0)); VF_CATCH }
int vf_dev_download(double *host, const double *dev, size_t n) { VF_TRY VF_CUDA(cudaDeviceSynchronize()); VF_CUDA(cudaMemcpy(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost)); VF_CATCH }
int vf_dev_memset_zero(double *dev, size_t n) { VF_TRY VF_CUDA(cudaMemset(dev, 0, n * sizeof(double))); VF_CUDA(cudaStreamSynchronize(0)); VF_CATCH }
void *vf_mg_stream(vf_mg *mg) { return (void *)mg->ctx.stream; }
int vf_mg_synchronize(vf_mg *mg) { VF_TRY VF_CUDA(cudaStreamSynchronize(mg->ctx.stream)); VF_CATCH }

// Device time of `reps` back-to-back repetitions of one multigrid operation on the level's own x/b/r fields
// (CUDA events on the solver's stream).  op: 0 one smoothing sweep (2^N colour passes), 1 residual, 2 applyK,
// 3 restrict (level -> level+1), 4 prolong-add (level+1 -> level), 5 coarse solve, 6 V-cycle from `level`, 7 FMG cycle.
int vf_mg_time_op(vf_mg *mg, int op, int level, int reps, int nsmooth, double *ms_per_rep) {
    VF_TRY
    mg_sync_level_masks(*mg); mg_update_stiffness(*mg);
    cudaEvent_t e0, e1; VF_CUDA(cudaEventCreate(&e0)); VF_CUDA(cudaEventCreate(&e1));
    const bool profWas = mg->sim->prof.enabled; mg->sim->prof.enabled = false;
    auto body = [&]() {
        switch (op) {
            case 0: mg_smooth(*mg, level, fx(level), fb(level), true); break;
            case 1: mg_residual(*mg, level, fx(level), fb(level), fr(level)); break;
            case 2: mg_apply_K(*mg, level, fx(level), fx(level), fr(level), APPLY_SET, true); break;
            case 3: launch_restrict(mg->ctx, mg->grid(level), mg->grid(level + 1), lr(*mg, level), lb(*mg, level + 1)); break;
            case 4: launch_prolong(mg->ctx, mg->grid(level), mg->grid(level + 1), lx(*mg, level + 1), lx(*mg, level), true); break;
            case 5: mg_coarse_solve(*mg, fb(mg->numLevels() - 1), fx(mg->numLevels() - 1)); break;
            case 6: mg_vcycle(*mg, level, nsmooth, true); break;
            case 7: mg_fmg(*mg, 0, nsmooth, true); break;
            case 8: {   // forward sweep that also emits the residual (stored-stencil levels of an undivided solver)
                if (!(level > 0 && !mg->grp && gs_residual_fusable(mg->grid(level)))) throw std::runtime_error("smooth_residual: not available on this level");
                const Field rl = fr(level);
                mg_smooth(*mg, level, fx(level), fb(level), true, &rl);
                launch_zero_dirichlet(mg->ctx, mg->grid(level), mg->dmask(level), lr(*mg, level));
                break;
            }
            default: throw std::runtime_error("unknown op");
        }
    };
    body(); // warm-up
    VF_CUDA(cudaEventRecord(e0, mg->ctx.stream));
    for (int i = 0; i < reps; ++i) body();
    VF_CUDA(cudaEventRecord(e1, mg->ctx.stream));
    VF_CUDA(cudaEventSynchronize(e1));
    float ms = 0; VF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_rep = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    mg->sim->prof.enabled = profWas;
    VF_CATCH
}
int vf_prof_enable(vf_mg *mg, int enable) { mg->sim->prof.enabled = enable != 0; return 0; }
int vf_prof_reset(vf_mg *mg) { VF_TRY mg->sim->prof.reset(); VF_CATCH }
int vf_prof_num_categories(void) { return PC_COUNT; }
const char *vf_prof_name(int c) { return (c >= 0 && c < PC_COUNT) ? kProfNames[c] : ""; }
int vf_prof_get(vf_mg *mg, int c, int64_t *launches, double *ms, double *units) {
    VF_TRY Profiler &p = mg->sim->prof; VF_CUDA(cudaStreamSynchronize(mg->ctx.stream)); p.resolve();
    if (c < 0 || c >= PC_COUNT) throw std::runtime_error("bad profiling category");
    if (launches) *launches = p.launches[c]; if (ms) *ms = p.ms[c]; if (units) *units = p.units[c];
    VF_CATCH
}

// ---- stand-alone filters --------------------------------------------------------------------
static LaunchCtx default_ctx() { LaunchCtx c; c.stream = nullptr; c.prof = nullptr; return c; }
int vf_filter_smooth(int dim, const int64_t *sizes, int radius, int type, const double *in, double *out) {
    VF_TRY ensure_device(); long long n = 1; int sz[3] = {1, 1, 1}; for (int d = 0; d < dim; ++d) { sz[d] = (int)sizes[d]; n *= sizes[d]; }
    DevBuf<double> a, b; a.alloc(n, false); b.alloc(n, false); LaunchCtx c = default_ctx();
    h2d(a.p, in, n, c.stream); launch_filter_smooth(c, dim, sz, radius, type, a.p, b.p); d2h(out, b.p, n, c.stream); VF_CATCH
}
int vf_filter_project(int64_t n, double beta, const double *in, double *out) {
    VF_TRY ensure_device(); DevBuf<double> a, b; a.alloc(n, false); b.alloc(n, false); LaunchCtx c = default_ctx();
    h2d(a.p, in, n, c.stream); launch_filter_project(c, n, beta, a.p, b.p); d2h(out, b.p, n, c.stream); VF_CATCH
}
int vf_filter_project_backprop(int64_t n, double beta, const double *in, const double *vars, double *out) {
    VF_TRY ensure_device(); DevBuf<double> a, v, b; a.alloc(n, false); v.alloc(n, false); b.alloc(n, false); LaunchCtx c = default_ctx();
    h2d(a.p, in, n, c.stream); h2d(v.p, vars, n, c.stream); launch_filter_project_backprop(c, n, beta, a.p, v.p, b.p); d2h(out, b.p, n, c.stream); VF_CATCH
}

static long long prod_sizes(int dim, const int *sz) { long long n = 1; for (int d = 0; d < dim; ++d) n *= sz[d]; return n; }
// UpsampleFilter (TopologyOptimizationFilter.hh:418-523): coarse_sizes -> (coarse_sizes - 1) * factor + 1
int vf_filter_upsample(int dim, const int64_t *coarse_sizes, int factor, const double *in, double *out) {
    VF_TRY ensure_device(); int cs[3] = {1, 1, 1}, fs[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { cs[d] = (int)coarse_sizes[d]; fs[d] = (cs[d] - 1) * factor + 1; if (cs[d] < 2) throw std::runtime_error("Interpolation can only be applied to a 2^d grid or larger"); }
    const long long nc = prod_sizes(dim, cs), nf = prod_sizes(dim, fs);
    DevBuf<double> a, b; a.alloc(nc, false); b.alloc(nf, false); LaunchCtx c = default_ctx();
    h2d(a.p, in, nc, c.stream); launch_filter_upsample(c, dim, cs, factor, a.p, b.p); d2h(out, b.p, nf, c.stream); VF_CATCH
}
int vf_filter_upsample_backprop(int dim, const int64_t *coarse_sizes, int factor, const double *d_dout, double *d_din) {
    VF_TRY ensure_device(); int cs[3] = {1, 1, 1}, fs[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { cs[d] = (int)coarse_sizes[d]; fs[d] = (cs[d] - 1) * factor + 1; }
    const long long nc = prod_sizes(dim, cs), nf = prod_sizes(dim, fs);
    DevBuf<double> a, b; a.alloc(nf, false); b.alloc(nc, false); LaunchCtx c = default_ctx();
    h2d(a.p, d_dout, nf, c.stream); launch_filter_upsample_backprop(c, dim, cs, factor, a.p, b.p); d2h(d_din, b.p, nc, c.stream); VF_CATCH
}
// VertexToCellFilter (:528-598): vertex_sizes -> vertex_sizes - 1
int vf_filter_vertex_to_cell(int dim, const int64_t *vertex_sizes, const double *in, double *out) {
    VF_TRY ensure_device(); int vs[3] = {1, 1, 1}, es[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { vs[d] = (int)vertex_sizes[d]; es[d] = vs[d] - 1; if (vs[d] < 2) throw std::runtime_error("Input grid must be 2^d or larger."); }
    const long long nv = prod_sizes(dim, vs), ne = prod_sizes(dim, es);
    DevBuf<double> a, b; a.alloc(nv, false); b.alloc(ne, false); LaunchCtx c = default_ctx();
    h2d(a.p, in, nv, c.stream); launch_filter_v2c(c, dim, vs, a.p, b.p); d2h(out, b.p, ne, c.stream); VF_CATCH
}
int vf_filter_vertex_to_cell_backprop(int dim, const int64_t *vertex_sizes, const double *d_dout, double *d_din) {
    VF_TRY ensure_device(); int vs[3] = {1, 1, 1}, es[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { vs[d] = (int)vertex_sizes[d]; es[d] = vs[d] - 1; }
    const long long nv = prod_sizes(dim, vs), ne = prod_sizes(dim, es);
    DevBuf<double> a, b; a.alloc(ne, false); b.alloc(nv, false); LaunchCtx c = default_ctx();
    h2d(a.p, d_dout, ne, c.stream); launch_filter_v2c_backprop(c, dim, vs, a.p, b.p); d2h(d_din, b.p, nv, c.stream); VF_CATCH
}
// LangelaarFilter (:601-712).  `out` is IN/OUT: in 3D the support of a voxel holds the voxel itself (NDVector.hh:211-229 walks the first
// N - 1 axes), so the previous content of the output array is read -- pass the array the reference would be writing into.  smax
// receives m_cachedSmax; the output itself is m_cachedFiltered.
int vf_filter_langelaar(int dim, const int64_t *sizes, const double *in, double *out, double *smax) {
    VF_TRY ensure_device(); int sz[3] = {1, 1, 1}; for (int d = 0; d < dim; ++d) sz[d] = (int)sizes[d];
    const long long n = prod_sizes(dim, sz);
    DevBuf<double> a, b, m; a.alloc(n, false); b.alloc(n, false); m.alloc(n, true); LaunchCtx c = default_ctx();
    h2d(a.p, in, n, c.stream); h2d(b.p, out, n, c.stream); launch_filter_langelaar(c, dim, sz, a.p, b.p, m.p);
    d2h(out, b.p, n, c.stream); if (smax) d2h(smax, m.p, n, c.stream); VF_CATCH
}
int vf_filter_langelaar_backprop(int dim, const int64_t *sizes, const double *d_dout, const double *vars, const double *filtered, const double *smax, double *d_din) {
    VF_TRY ensure_device(); int sz[3] = {1, 1, 1}; for (int d = 0; d < dim; ++d) sz[d] = (int)sizes[d];
    const long long n = prod_sizes(dim, sz);
    DevBuf<double> g, v, f, m, scr, o; g.alloc(n, false); v.alloc(n, false); f.alloc(n, false); m.alloc(n, false); scr.alloc(3 * n, true); o.alloc(n, false);
    LaunchCtx c = default_ctx();
    h2d(g.p, d_dout, n, c.stream); h2d(v.p, vars, n, c.stream); h2d(f.p, filtered, n, c.stream); h2d(m.p, smax, n, c.stream);
    launch_filter_langelaar_backprop(c, dim, sz, g.p, v.p, f.p, m.p, scr.p, o.p); d2h(d_din, o.p, n, c.stream); VF_CATCH
}

// ---- device-pointer variants (element arrays stay in HBM; kernels run on the simulator's stream) ------------------------
// Building blocks of a topology-optimization iteration whose element arrays are partitioned into slabs (one per GPU): the
// host side exchanges filter halos and all-reduces scalars between these calls (voxelfem_b200/capi.py: SlabProblem).
int vf_dev_filter_smooth(vf_sim *s, int dim, const int64_t *sizes, int radius, int type, const double *in_dev, double *out_dev) {
    VF_TRY int sz[3] = {1, 1, 1}; for (int d = 0; d < dim; ++d) sz[d] = (int)sizes[d];
    launch_filter_smooth(s->ctx, dim, sz, radius, type, in_dev, out_dev); VF_CATCH
}
int vf_dev_filter_project(vf_sim *s, int64_t n, double beta, const double *in_dev, double *out_dev) {
    VF_TRY launch_filter_project(s->ctx, n, beta, in_dev, out_dev); VF_CATCH
}
int vf_dev_filter_project_backprop(vf_sim *s, int64_t n, double beta, const double *g_dev, const double *vars_dev, double *out_dev) {
    VF_TRY launch_filter_project_backprop(s->ctx, n, beta, g_dev, vars_dev, out_dev); VF_CATCH
}
int vf_dev_oc_update(vf_sim *s, int64_t n, const double *x0_dev, const double *dJ_dev, const double *dc_dev, double lambda, double m, double p, double *out_dev) {
    VF_TRY launch_oc_update(s->ctx, n, x0_dev, dJ_dev, dc_dev, lambda, m, p, out_dev); VF_CATCH
}
int vf_dev_sum(vf_sim *s, int64_t n, const double *x_dev, double *result) {
    VF_TRY
    if (s->scratch.n < reduce_scratch_doubles()) s->scratch.alloc(reduce_scratch_doubles(), true);
    if (s->scalars.n < 8) s->scalars.alloc(8, true);
    launch_sum(s->ctx, n, x_dev, s->scalars.p, s->scratch.p);
    d2h(result, s->scalars.p, 1, s->stream);
    VF_CATCH
}
int vf_sim_set_densities_dev(vf_sim *s, const double *rho_dev) {
    VF_TRY if (rho_dev != s->rho.p) VF_CUDA(cudaMemcpyAsync(s->rho.p, rho_dev, sizeof(double) * s->g.numElems, cudaMemcpyDeviceToDevice, s->stream));
    s->updateModuli(); VF_CATCH
}
int vf_sim_compliance_gradient_dev(vf_sim *s, const double *u_dev, double *g_dev, int accumulate) {
    VF_TRY launch_compliance_gradient(s->ctx, s->g, s->K0p, u_dev, s->rho.p, g_dev, s->law, s->E0, s->Emin, s->gamma, s->q, s->gravity, s->elemVolume(), accumulate != 0); VF_CATCH
}
void *vf_sim_stream(vf_sim *s) { return (void *)s->stream; }
int vf_sim_synchronize(vf_sim *s) { VF_TRY VF_CUDA(cudaStreamSynchronize(s->stream)); VF_CATCH }

} // extern "C"

// ---------------------------------------------------------------------------
// Topology optimization problem
// ---------------------------------------------------------------------------
struct FilterSpec {
    int kind, radius, type; double beta;       // spec quadruple {kind, radius | factor, type, beta}
    int in[3] = {1, 1, 1}, out[3] = {1, 1, 1}; // grid sizes (first N entries), set by vf_top_create from the physical grid backwards (FilterChain::setOutputDimensions, TopologyOptimizationFilter.hh:117-132)
    long long nIn = 0, nOut = 0;
    // LangelaarFilter state: smax of every voxel's support and the filtered values of the LAST application (m_cachedSmax, m_cachedFiltered, :697-701)
    std::unique_ptr<DevBuf<double>> smaxCache, filteredCache, lgScratch;
    // PythonFilter (:247-275): host callbacks
    vf_filter_apply_cb applyCb = nullptr; vf_filter_backprop_cb backpropCb = nullptr; void *user = nullptr;
};
struct vf_top {
    vf_mg *mg; vf_sim *sim;
    std::vector<FilterSpec> filters; double volFrac;
    std::vector<std::unique_ptr<DevBuf<double>>> vars; // m_vars of FilterChain (TopologyOptimizationFilter.hh:111-132)
    DevBuf<double> u, f, dJ, dc, stepped, xv, tmp, scalar, scratch;
    std::vector<double> hostA, hostB, hostC;           // staging for PythonFilter callbacks
    vf_pcg_callback residualCb = nullptr; void *residualUser = nullptr;   // MultigridComplianceObjective::residual_cb
    int cgIter = 100; double tol = 1e-5; int mgIt = 1, mgSmooth = 2; bool fmg = true, zeroInit = false; // TopologyOptimizationObjective.hh:99-103
    double lamMin = 1, lamMax = 2; // OptimalityCriterion.hh:46-49
    int lastPcgIters = 0;
    double *hostScalar = nullptr;
    ~vf_top() { if (hostScalar) cudaFreeHost(hostScalar); }
    long long ne() const { return sim->g.numElems; }                              // physical variables (one per element)
    long long nDesign() const { return filters.empty() ? ne() : filters.front().nIn; }
    long long nMax() const { long long m = ne(); for (const auto &fs : filters) m = std::max(m, std::max(fs.nIn, fs.nOut)); return m; }
    // Filter::apply; `out` is in/out for the Langelaar filter (in 3D a voxel's support holds the voxel itself: its previous value is read)
    void applyFilter(FilterSpec &fs, const double *in, double *out) {
        const int N = sim->N;
        switch (fs.kind) {
            case VF_FILTER_SMOOTH:  launch_filter_smooth(mg->ctx, N, fs.in, fs.radius, fs.type, in, out); break;
            case VF_FILTER_PROJECT: launch_filter_project(mg->ctx, fs.nIn, fs.beta, in, out); break;
            case VF_FILTER_UPSAMPLE: launch_filter_upsample(mg->ctx, N, fs.in, fs.radius, in, out); break;
            case VF_FILTER_VERTEX_TO_CELL: launch_filter_v2c(mg->ctx, N, fs.in, in, out); break;
            case VF_FILTER_LANGELAAR:
                launch_filter_langelaar(mg->ctx, N, fs.in, in, out, fs.smaxCache->p);
                VF_CUDA(cudaMemcpyAsync(fs.filteredCache->p, out, sizeof(double) * fs.nOut, cudaMemcpyDeviceToDevice, mg->ctx.stream));
                break;
            case VF_FILTER_PYTHON: {
                if (!fs.applyCb) throw std::runtime_error("Apply callback must be configured");
                hostA.resize(fs.nIn); hostB.assign(fs.nOut, 0.0);
                d2h(hostA.data(), in, fs.nIn, mg->ctx.stream);
                if (fs.applyCb(hostA.data(), fs.nIn, hostB.data(), fs.nOut, fs.user)) throw std::runtime_error("PythonFilter apply callback failed");
                h2d(out, hostB.data(), fs.nOut, mg->ctx.stream); VF_CUDA(cudaStreamSynchronize(mg->ctx.stream));
                break;
            }
            default: throw std::runtime_error("unknown filter kind");
        }
    }
    // Filter::backprop(in = dJ/d(out), vars = the filter's input, out = dJ/d(in))
    void backpropFilter(FilterSpec &fs, const double *g, const double *vars_, double *out) {
        const int N = sim->N;
        switch (fs.kind) {
            case VF_FILTER_SMOOTH:  launch_filter_smooth(mg->ctx, N, fs.in, fs.radius, fs.type, g, out); break;     // symmetric operator (:297-310)
            case VF_FILTER_PROJECT: launch_filter_project_backprop(mg->ctx, fs.nIn, fs.beta, g, vars_, out); break;
            case VF_FILTER_UPSAMPLE: launch_filter_upsample_backprop(mg->ctx, N, fs.in, fs.radius, g, out); break;
            case VF_FILTER_VERTEX_TO_CELL: launch_filter_v2c_backprop(mg->ctx, N, fs.in, g, out); break;
            case VF_FILTER_LANGELAAR: launch_filter_langelaar_backprop(mg->ctx, N, fs.in, g, vars_, fs.filteredCache->p, fs.smaxCache->p, fs.lgScratch->p, out); break;
            case VF_FILTER_PYTHON: {
                if (!fs.backpropCb) throw std::runtime_error("Backprop callback must be configured");
                hostA.resize(fs.nOut); hostB.resize(fs.nIn); hostC.assign(fs.nIn, 0.0);
                d2h(hostA.data(), g, fs.nOut, mg->ctx.stream); d2h(hostB.data(), vars_, fs.nIn, mg->ctx.stream);
                if (fs.backpropCb(hostA.data(), fs.nOut, hostB.data(), fs.nIn, hostC.data(), fs.user)) throw std::runtime_error("PythonFilter backprop callback failed");
                h2d(out, hostC.data(), fs.nIn, mg->ctx.stream); VF_CUDA(cudaStreamSynchronize(mg->ctx.stream));
                break;
            }
            default: throw std::runtime_error("unknown filter kind");
        }
    }
    double readScalar() {
        VF_CUDA(cudaMemcpyAsync(hostScalar, scalar.p, sizeof(double), cudaMemcpyDeviceToHost, mg->ctx.stream));
        VF_CUDA(cudaStreamSynchronize(mg->ctx.stream));
        return *hostScalar;
    }
    // MultigridComplianceObjective::updateCache (TopologyOptimizationObjective.hh:88-96)
    void updateCache(const double *xPhysDev) {
        if (xPhysDev != sim->rho.p) VF_CUDA(cudaMemcpyAsync(sim->rho.p, xPhysDev, sizeof(double) * ne(), cudaMemcpyDeviceToDevice, mg->ctx.stream));
        sim->updateModuli();
        if (zeroInit) VF_CUDA(cudaMemsetAsync(u.p, 0, sizeof(double) * u.n, mg->ctx.stream));
        mg_pcg(*mg, u.p, f.p, cgIter, tol, mgIt, mgSmooth, fmg, false, residualCb, residualUser);   // residual_cb (TopologyOptimizationObjective.hh:93, 104)
        lastPcgIters = mg->lastIters;
    }
    void setVarsDev() { // FilterChain::setDesignVars (:142-152) + updateCache
        TraceScope ts("setVars");                              // TopologyOptimizationProblem.hh:42
        for (size_t i = 0; i < filters.size(); ++i) applyFilter(filters[i], vars[i]->p, vars[i + 1]->p);
        updateCache(vars.back()->p);
    }
    void backprop(DevBuf<double> &g, DevBuf<double> &scr) { // FilterChain::backprop (:162-170); g: physical size in, design size out
        double *a = g.p, *b = scr.p;
        for (size_t i = filters.size(); i-- > 0;) { backpropFilter(filters[i], a, vars[i]->p, b); std::swap(a, b); }
        if (a != g.p) VF_CUDA(cudaMemcpyAsync(g.p, a, sizeof(double) * nDesign(), cudaMemcpyDeviceToDevice, mg->ctx.stream));
    }
    void objectiveGradient() { // into dJ
        TraceScope ts("evaluateObjectiveGradient");            // TopologyOptimizationProblem.hh:78
        launch_compliance_gradient(mg->ctx, sim->g, sim->K0p, u.p, sim->rho.p, dJ.p, sim->law, sim->E0, sim->Emin, sim->gamma, sim->q, sim->gravity, sim->elemVolume(), false);
        backprop(dJ, tmp);
    }
    void constraintJacobian() { // into dc (TopologyOptimizationConstraint.hh:34-36)
        TraceScope ts("evaluateConstraintsJacobian");          // TopologyOptimizationProblem.hh:107
        launch_fill(mg->ctx, ne(), -1.0 / (volFrac * double(ne())), dc.p);
        backprop(dc, tmp);
    }
    double constraintOf(const double *xPhys) { // :30-32
        launch_sum(mg->ctx, ne(), xPhys, scalar.p, scratch.p);
        return 1.0 - (readScalar() / double(ne())) / volFrac;
    }
    double ceval(double lambda, double m, double p) { // OptimalityCriterion.hh:64-83 + TopologyOptimizationProblem.hh:58-66
        launch_oc_update(mg->ctx, nDesign(), vars[0]->p, dJ.p, dc.p, lambda, m, p, stepped.p);
        const double *cur = stepped.p; double *a = xv.p, *b = tmp.p;
        for (auto &fs : filters) { applyFilter(fs, cur, a); cur = a; std::swap(a, b); }   // FilterChain::applyInPlace (:154-160)
        return constraintOf(cur);
    }
};

extern "C" {
int vf_top_create(vf_mg *mg, int nf, const double *spec, double volFrac, vf_top **out) {
    VF_TRY
    auto t = std::make_unique<vf_top>();
    t->mg = mg; t->sim = mg->sim; t->volFrac = volFrac;
    const int N = t->sim->N;
    for (int i = 0; i < nf; ++i) { FilterSpec fs; fs.kind = (int)spec[4 * i]; fs.radius = (int)spec[4 * i + 1]; fs.type = (int)spec[4 * i + 2]; fs.beta = spec[4 * i + 3]; t->filters.push_back(std::move(fs)); }
    // grid sizes from the physical grid backwards (FilterChain::setOutputDimensions, TopologyOptimizationFilter.hh:117-132)
    int dims[3] = {1, 1, 1}; for (int a = 0; a < N; ++a) dims[a] = (int)t->sim->ne[a];
    for (int i = nf - 1; i >= 0; --i) {
        FilterSpec &fs = t->filters[i];
        for (int a = 0; a < 3; ++a) fs.out[a] = fs.in[a] = dims[a];
        if (fs.kind == VF_FILTER_UPSAMPLE) {                       // m_setOutputDimensions (:448-454)
            const int f = fs.radius;
            if (f < 1) throw std::runtime_error("UpsampleFilter factor must be positive");
            for (int a = 0; a < N; ++a) {
                if (dims[a] < 2) throw std::runtime_error("Interpolation can only be applied to a 2^d grid or larger");
                fs.in[a] = (dims[a] - 1) / f + 1;
                if ((fs.in[a] - 1) * f + 1 != dims[a]) throw std::runtime_error("Output size is not divisible by factor");
            }
        } else if (fs.kind == VF_FILTER_VERTEX_TO_CELL) {          // (:593-596)
            for (int a = 0; a < N; ++a) fs.in[a] = dims[a] + 1;
        } else if (fs.kind < 0 || fs.kind > VF_FILTER_PYTHON) throw std::runtime_error("unknown filter kind");
        fs.nIn = fs.nOut = 1;
        for (int a = 0; a < N; ++a) { fs.nIn *= fs.in[a]; fs.nOut *= fs.out[a]; dims[a] = fs.in[a]; }
        if (fs.kind == VF_FILTER_LANGELAAR) {
            fs.smaxCache = std::make_unique<DevBuf<double>>(); fs.smaxCache->alloc(fs.nIn, true);
            fs.filteredCache = std::make_unique<DevBuf<double>>(); fs.filteredCache->alloc(fs.nIn, true);
            fs.lgScratch = std::make_unique<DevBuf<double>>(); fs.lgScratch->alloc(3 * fs.nIn, true);
        }
    }
    const long long ne = t->ne(), nmax = t->nMax(); const size_t len = (size_t)t->sim->g.numNodes * t->sim->N;
    for (int i = 0; i <= nf; ++i) { t->vars.push_back(std::make_unique<DevBuf<double>>()); t->vars.back()->alloc(i < nf ? t->filters[i].nIn : ne, true); }
    t->u.alloc(len, true); t->f.alloc(len, true);
    t->dJ.alloc(nmax, true); t->dc.alloc(nmax, true); t->stepped.alloc(t->nDesign(), true); t->xv.alloc(nmax, true); t->tmp.alloc(nmax, true);
    t->scalar.alloc(1, true); t->scratch.alloc(reduce_scratch_doubles(), true);
    VF_CUDA(cudaMallocHost(&t->hostScalar, sizeof(double)));
    sim_build_load_dev(*t->sim, t->f.p);      // ComplianceObjective ctor (TopologyOptimizationObjective.hh:32-35)
    t->updateCache(t->sim->rho.p);            // MultigridComplianceObjective ctor (:82-86)
    VF_CUDA(cudaStreamSynchronize(mg->ctx.stream));
    *out = t.release();
    VF_CATCH
}
int64_t vf_top_num_vars(const vf_top *t) { return t->nDesign(); }                 /* FilterChain::numVars (:134) */
int64_t vf_top_num_physical_vars(const vf_top *t) { return t->ne(); }             /* numPhysicalVars (:138) */
int vf_top_get_grid_dims(const vf_top *t, int physical, int64_t *dims) {          /* gridDims / physicalGridDims (:136, 139) */
    const int N = t->sim->N;
    for (int a = 0; a < N; ++a) dims[a] = (physical || t->filters.empty()) ? (int64_t)t->sim->ne[a] : t->filters.front().in[a];
    return 0;
}
int vf_top_set_residual_callback(vf_top *t, vf_pcg_callback cb, void *user) { t->residualCb = cb; t->residualUser = user; return 0; }
int vf_top_set_python_filter(vf_top *t, int index, vf_filter_apply_cb apply_cb, vf_filter_backprop_cb backprop_cb, void *user) {
    VF_TRY
    if (index < 0 || index >= (int)t->filters.size() || t->filters[index].kind != VF_FILTER_PYTHON) throw std::runtime_error("filter " + std::to_string(index) + " is not a PythonFilter");
    t->filters[index].applyCb = apply_cb; t->filters[index].backpropCb = backprop_cb; t->filters[index].user = user;
    VF_CATCH
}
int vf_top_destroy(vf_top *t) { VF_TRY if (t) { cudaStreamSynchronize(t->mg->ctx.stream); delete t; } VF_CATCH }
int vf_top_set_solver(vf_top *t, int cgIter, double tol, int mgIt, int mgSmooth, int fmg, int zeroInit) { t->cgIter = cgIter; t->tol = tol; t->mgIt = mgIt; t->mgSmooth = mgSmooth; t->fmg = fmg != 0; t->zeroInit = zeroInit != 0; return 0; }
int vf_top_set_vars(vf_top *t, const double *x) { VF_TRY h2d(t->vars[0]->p, x, t->nDesign(), t->mg->ctx.stream); t->setVarsDev(); VF_CUDA(cudaStreamSynchronize(t->mg->ctx.stream)); VF_CATCH }
int vf_top_get_vars(vf_top *t, int which, double *out) { VF_TRY d2h(out, which == 0 ? t->vars.front()->p : t->vars.back()->p, which == 0 ? t->nDesign() : t->ne(), t->mg->ctx.stream); VF_CATCH }
int vf_top_compliance(vf_top *t, double *out) {
    VF_TRY launch_dot_plain(t->mg->ctx, (long long)t->u.n, t->f.p, t->u.p, t->scalar.p, t->scratch.p); *out = 0.5 * t->readScalar(); VF_CATCH
}
int vf_top_constraint(vf_top *t, double *out) { VF_TRY *out = t->constraintOf(t->vars.back()->p); VF_CATCH }
int vf_top_objective_gradient(vf_top *t, double *g) { VF_TRY t->objectiveGradient(); d2h(g, t->dJ.p, t->nDesign(), t->mg->ctx.stream); VF_CATCH }
int vf_top_constraint_jacobian(vf_top *t, double *g) { VF_TRY t->constraintJacobian(); d2h(g, t->dc.p, t->nDesign(), t->mg->ctx.stream); VF_CATCH }
int vf_top_get_u(vf_top *t, double *u) { VF_TRY d2h(u, t->u.p, t->u.n, t->mg->ctx.stream); VF_CATCH }
int vf_top_last_pcg_iterations(vf_top *t) { return t->lastPcgIters; }
int vf_top_get_lambda_bracket(vf_top *t, double *lo, double *hi) { *lo = t->lamMin; *hi = t->lamMax; return 0; }
// OCOptimizer::step (OptimalityCriterion.hh:51-134)
static int oc_bisect(vf_top *t, double m, double p, double ctol) {   // bracket + bisection of OCOptimizer::step (:95-129) on t->dJ, t->dc; result in t->stepped
    TraceScope ts("Bisection");                                // OptimalityCriterion.hh:91
    int nevals = 0;
    auto ceval = [&](double lam) { ++nevals; return t->ceval(lam, m, p); };
    const double dilation = 32;
    double mid = 0.5 * (t->lamMin + t->lamMax);
    t->lamMax = dilation * t->lamMax + (1 - dilation) * mid;
    t->lamMin = std::max(dilation * t->lamMin + (1 - dilation) * mid, 0.01);
    const int guard = 100; int nit = 0;
    for (; nit < guard; ++nit) { if (ceval(t->lamMin) < 0) break; t->lamMax = t->lamMin; t->lamMin /= 2; }
    if (nit == guard) throw std::runtime_error("Bracketing constraint(lambda_min) < 0 failed (100 times).");
    if (nit == 0) for (; nit < guard; ++nit) { if (ceval(t->lamMax) > 0) break; t->lamMin = t->lamMax; t->lamMax *= 2; }
    if (nit == guard) throw std::runtime_error("Bracketing constraint(lambda_max) > 0 failed (100 times).");
    double violation;
    do {
        mid = 0.5 * (t->lamMin + t->lamMax);
        violation = ceval(mid);
        if (std::abs(violation) <= ctol) break;
        ++nit;
        if (violation < 0) t->lamMin = mid;
        if (violation > 0) t->lamMax = mid;
    } while (true);
    return nevals;
}
int vf_top_oc_step(vf_top *t, double m, double p, double ctol, int *nevalsOut) {
    VF_TRY
    TraceScope ts("OC step");                                  // OptimalityCriterion.hh:52
    t->objectiveGradient(); t->constraintJacobian();
    const int nevals = oc_bisect(t, m, p, ctol);
    // m_p.setVars(m_steppedVars) (:133)
    VF_CUDA(cudaMemcpyAsync(t->vars[0]->p, t->stepped.p, sizeof(double) * t->nDesign(), cudaMemcpyDeviceToDevice, t->mg->ctx.stream));
    t->setVarsDev();
    VF_CUDA(cudaStreamSynchronize(t->mg->ctx.stream));
    if (nevalsOut) *nevalsOut = nevals;
    VF_CATCH
}
// The search half of OCOptimizer::step for problems whose virtual methods are overridden on the host (the trampoline of
// python_bindings/VoxelFEM.cc:58-66; step(inplace = false), OptimalityCriterion.hh:57-60): the objective gradient comes from the
// caller (evaluateObjectiveGradientAndReturn), the constraint is the problem's own (evaluateOCConstraintAtVars is not virtual), and
// the stepped variables are returned instead of being set -- the caller then invokes its (possibly overridden) setVars (:133).
int vf_top_oc_search(vf_top *t, const double *dJ, double m, double p, double ctol, double *stepped, int *nevalsOut) {
    VF_TRY
    TraceScope ts("OC step");
    if (dJ) h2d(t->dJ.p, dJ, (size_t)t->nDesign(), t->mg->ctx.stream); else t->objectiveGradient();
    t->constraintJacobian();
    const int nevals = oc_bisect(t, m, p, ctol);
    d2h(stepped, t->stepped.p, (size_t)t->nDesign(), t->mg->ctx.stream);
    if (nevalsOut) *nevalsOut = nevals;
    VF_CATCH
}

} // extern "C"

// ---------------------------------------------------------------------------
// Compliance topology optimization on a slab group (BASELINE.json configs[3]): TopologyOptimizationProblem +
// MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer (TopologyOptimizationProblem.hh:17-155,
// OptimalityCriterion.hh:38-149) with the element arrays partitioned like the solver.  Every part owns the design variables of
// its element layers [sb, se).  Per evaluation of the filter chain the parts exchange R = max(2 sum(radii), sum(radii) + 1) element
// layers with their neighbours and run the chain on slab + halo with the kernels of the undivided problem: the filter's reflecting
// boundary then applies at true domain ends only, and what the halo's far end contaminates reaches neither the owned layers, nor
// the ghost layer of the stiffness window, nor the layers back-propagation needs (R >= 2 rho_i for the nonlinear filter i).
// Volume and compliance are all-reduced; every rank runs the same bracket / bisection on the same all-reduced constraint values.
// Transport: device copies inside a local group, ncclSend/Recv + ncclAllReduce between NCCL ranks.  Filters: Smoothing, Projection.
// ---------------------------------------------------------------------------
struct GTopPart {
    vf_mg *mg = nullptr; vf_sim *sim = nullptr;
    int64_t sb = 0, se = 0, elo = 0, ehi = 0;       // owned element layers, owned + halo layers
    long long nOwn = 0, nExt = 0, layer = 0;        // entries of an owned / extended array, of one element layer
    DevBuf<double> x, stepped, dJ, dc, extA, extB, win, u, f, scalar, scratch, gwin;
    std::vector<std::unique_ptr<DevBuf<double>>> vars;   // chain variables on slab + halo (vars[0] = design variables)
    double *own(double *ext) const { return ext + (sb - elo) * layer; }
};
struct vf_gtop {
    vf_group *grp = nullptr;
    std::vector<FilterSpec> filters; double volFrac = 0;
    std::vector<std::unique_ptr<GTopPart>> parts;
    int64_t gne[3] = {0, 0, 0}, R = 0; long long neGlobal = 0;
    int cgIter = 100; double tol = 1e-5; int mgIt = 1, mgSmooth = 2; bool fmg = true, zeroInit = false;
    double lamMin = 1, lamMax = 2;
    int lastPcgIters = 0;
    DevBuf<double> gather;                           // whole-grid staging for the get_* entry points
    vf_mg &lead() { return *grp->parts[0]; }
    cudaStream_t stream() { return lead().ctx.stream; }
    // sum of one device scalar per part over all parts / ranks
    double sumScalars() {
        vf_mg &L = lead();
        if (grp->comm) {
            NcclApi &A = NcclApi::get(); double *p = parts[0]->scalar.p;
            A.check(A.AllReduce(p, p, 1, kNcclDouble, kNcclSum, grp->comm, L.ctx.stream), "ncclAllReduce");
        }
        double tot = 0;
        for (auto &pp : parts) { double v = 0; d2h(&v, pp->scalar.p, 1, L.ctx.stream); tot += v; if (grp->comm) break; }
        return tot;
    }
    // ext <- owned layers of `ownedSel` + halo layers from the neighbouring slabs
    template<class Sel> void withHalo(Sel ownedSel, std::vector<double *> &ext) {
        cudaStream_t st = stream();
        for (size_t i = 0; i < parts.size(); ++i) VF_CUDA(cudaMemcpyAsync(parts[i]->own(ext[i]), ownedSel(*parts[i]), sizeof(double) * parts[i]->nOwn, cudaMemcpyDeviceToDevice, st));
        if (grp->comm) {
            GTopPart &P = *parts[0]; NcclApi &A = NcclApi::get();
            const int64_t kl = P.sb - P.elo, kr = P.ehi - P.se;
            const double *o = ownedSel(P);
            A.check(A.GroupStart(), "ncclGroupStart");
            if (grp->rank > 0) {
                A.check(A.Send(o, (size_t)(R * P.layer), kNcclDouble, grp->rank - 1, grp->comm, st), "ncclSend");
                A.check(A.Recv(ext[0], (size_t)(kl * P.layer), kNcclDouble, grp->rank - 1, grp->comm, st), "ncclRecv");
            }
            if (grp->rank + 1 < grp->world) {
                A.check(A.Send(o + (P.se - P.sb - R) * P.layer, (size_t)(R * P.layer), kNcclDouble, grp->rank + 1, grp->comm, st), "ncclSend");
                A.check(A.Recv(ext[0] + (P.se - P.elo) * P.layer, (size_t)(kr * P.layer), kNcclDouble, grp->rank + 1, grp->comm, st), "ncclRecv");
            }
            A.check(A.GroupEnd(), "ncclGroupEnd");
            return;
        }
        for (size_t i = 0; i < parts.size(); ++i) {
            GTopPart &P = *parts[i];
            if (i > 0) { GTopPart &Q = *parts[i - 1]; const int64_t k = P.sb - P.elo;
                VF_CUDA(cudaMemcpyAsync(ext[i], ownedSel(Q) + (Q.se - Q.sb - k) * Q.layer, sizeof(double) * k * P.layer, cudaMemcpyDeviceToDevice, st)); }
            if (i + 1 < parts.size()) { GTopPart &Q = *parts[i + 1]; const int64_t k = P.ehi - P.se;
                VF_CUDA(cudaMemcpyAsync(ext[i] + (P.se - P.elo) * P.layer, ownedSel(Q), sizeof(double) * k * P.layer, cudaMemcpyDeviceToDevice, st)); }
        }
    }
    void applyFilter(GTopPart &P, const FilterSpec &fs, const double *in, double *out) {
        int sz[3] = {(int)(P.ehi - P.elo), (int)gne[1], (int)gne[2]};
        if (fs.kind == VF_FILTER_SMOOTH) launch_filter_smooth(P.mg->ctx, 3, sz, fs.radius, fs.type, in, out);
        else launch_filter_project(P.mg->ctx, P.nExt, fs.beta, in, out);
    }
    // FilterChain::setDesignVars (:142-152) into the parts' chain variables; returns nothing (vars.back() = physical densities on slab + halo)
    template<class Sel> void forwardInto(Sel ownedSel, bool keep, std::vector<double *> &last) {
        std::vector<double *> ext(parts.size());
        for (size_t i = 0; i < parts.size(); ++i) ext[i] = keep ? parts[i]->vars[0]->p : parts[i]->extA.p;
        withHalo(ownedSel, ext);
        last = ext;
        for (size_t k = 0; k < filters.size(); ++k)
            for (size_t i = 0; i < parts.size(); ++i) {
                GTopPart &P = *parts[i];
                double *out = keep ? P.vars[k + 1]->p : (last[i] == P.extA.p ? P.extB.p : P.extA.p);
                applyFilter(P, filters[k], last[i], out); last[i] = out;
            }
    }
    double volumeConstraint(const std::vector<double *> &physExt) {   // TopologyOptimizationConstraint.hh:30-32
        for (size_t i = 0; i < parts.size(); ++i) { GTopPart &P = *parts[i]; launch_sum(P.mg->ctx, P.nOwn, P.own(physExt[i]), P.scalar.p, P.scratch.p); }
        return 1.0 - (sumScalars() / double(neGlobal)) / volFrac;
    }
    // FilterChain::backprop (:162-170) of owned gradients g (in place: result in the owned arrays selected by outSel)
    template<class SelIn, class SelOut> void backprop(SelIn gSel, SelOut outSel) {
        std::vector<double *> ext(parts.size());
        for (size_t i = 0; i < parts.size(); ++i) ext[i] = parts[i]->extA.p;
        withHalo(gSel, ext);
        for (size_t k = filters.size(); k-- > 0;)
            for (size_t i = 0; i < parts.size(); ++i) {
                GTopPart &P = *parts[i];
                double *out = ext[i] == P.extA.p ? P.extB.p : P.extA.p;
                if (filters[k].kind == VF_FILTER_SMOOTH) applyFilter(P, filters[k], ext[i], out);
                else launch_filter_project_backprop(P.mg->ctx, P.nExt, filters[k].beta, ext[i], P.vars[k]->p, out);
                ext[i] = out;
            }
        for (size_t i = 0; i < parts.size(); ++i) VF_CUDA(cudaMemcpyAsync(outSel(*parts[i]), parts[i]->own(ext[i]), sizeof(double) * parts[i]->nOwn, cudaMemcpyDeviceToDevice, stream()));
    }
    // setVars (TopologyOptimizationProblem.hh:41-50): chain, densities of the stiffness window, MultigridComplianceObjective::updateCache (:88-96)
    void update() {
        TraceScope ts("setVars");
        std::vector<double *> last;
        forwardInto([](GTopPart &P) { return P.x.p; }, true, last);
        std::vector<double *> xs, bs_;
        for (auto &pp : parts) {
            GTopPart &P = *pp;
            VF_CUDA(cudaMemcpyAsync(P.sim->rho.p, last[&pp - &parts[0]] + (P.sim->xoff - P.elo) * P.layer, sizeof(double) * P.sim->g.numElems, cudaMemcpyDeviceToDevice, stream()));
            P.sim->updateModuli();
            if (zeroInit) VF_CUDA(cudaMemsetAsync(P.u.p, 0, sizeof(double) * P.u.n, stream()));
            xs.push_back(P.u.p); bs_.push_back(P.f.p);
        }
        std::vector<const double *> bs(bs_.begin(), bs_.end());
        mg_pcg(lead(), xs.data(), bs.data(), cgIter, tol, mgIt, mgSmooth, fmg, false, nullptr, nullptr);
        lastPcgIters = lead().lastIters;
    }
    double compliance() {
        for (auto &pp : parts) launch_masked_dot(pp->mg->ctx, pp->sim->g, pp->f.p, pp->u.p, pp->scalar.p, pp->scratch.p);   // owned node planes only
        return 0.5 * sumScalars();
    }
    void gradients() {   // dJ and dc with respect to the owned design variables
        for (auto &pp : parts) {
            GTopPart &P = *pp; vf_sim &sm = *P.sim;
            launch_compliance_gradient(P.mg->ctx, sm.g, sm.K0p, P.u.p, sm.rho.p, P.gwin.p, sm.law, sm.E0, sm.Emin, sm.gamma, sm.q, sm.gravity, sm.elemVolume(), false);
        }
        backprop([](GTopPart &P) { return P.gwin.p + (P.sb - P.sim->xoff) * P.layer; }, [](GTopPart &P) { return P.dJ.p; });
        for (auto &pp : parts) launch_fill(pp->mg->ctx, pp->nOwn, -1.0 / (volFrac * double(neGlobal)), pp->stepped.p);   // :34-36 (stepped as scratch)
        backprop([](GTopPart &P) { return P.stepped.p; }, [](GTopPart &P) { return P.dc.p; });
    }
    double ceval(double lambda, double m, double p) {
        for (auto &pp : parts) launch_oc_update(pp->mg->ctx, pp->nOwn, pp->x.p, pp->dJ.p, pp->dc.p, lambda, m, p, pp->stepped.p);
        std::vector<double *> last;
        forwardInto([](GTopPart &P) { return P.stepped.p; }, false, last);
        return volumeConstraint(last);
    }
    // owned arrays of all parts / ranks -> whole-grid host array
    template<class Sel> void gatherTo(Sel sel, double *outHost) {
        cudaStream_t st = stream();
        if (gather.n != (size_t)neGlobal) gather.alloc(neGlobal, false);
        VF_CUDA(cudaMemsetAsync(gather.p, 0, sizeof(double) * neGlobal, st));
        for (auto &pp : parts) VF_CUDA(cudaMemcpyAsync(gather.p + pp->sb * pp->layer, sel(*pp), sizeof(double) * pp->nOwn, cudaMemcpyDeviceToDevice, st));
        if (grp->comm) { NcclApi &A = NcclApi::get(); A.check(A.AllReduce(gather.p, gather.p, (size_t)neGlobal, kNcclDouble, kNcclSum, grp->comm, st), "ncclAllReduce"); }
        d2h(outHost, gather.p, (size_t)neGlobal, st);
    }
};

extern "C" {
int vf_group_top_create(vf_group *g, int nf, const double *spec, double volFrac, vf_gtop **out) {
    VF_TRY
    auto t = std::make_unique<vf_gtop>();
    t->grp = g; t->volFrac = volFrac;
    vf_mg &lead = *g->parts[0];
    if (lead.N != 3) throw std::runtime_error("slab-partitioned topology optimization is 3D");
    int64_t sr = 0;
    for (int i = 0; i < nf; ++i) {
        FilterSpec fs; fs.kind = (int)spec[4 * i]; fs.radius = (int)spec[4 * i + 1]; fs.type = (int)spec[4 * i + 2]; fs.beta = spec[4 * i + 3];
        if (fs.kind != VF_FILTER_SMOOTH && fs.kind != VF_FILTER_PROJECT) throw std::runtime_error("slab-partitioned problems support the Smoothing and Projection filters");
        if (fs.kind == VF_FILTER_SMOOTH) sr += fs.radius;
        t->filters.push_back(std::move(fs));
    }
    t->R = std::max<int64_t>(2 * sr, sr + 1);
    t->gne[0] = lead.sim->gne0; t->gne[1] = lead.sim->ne[1]; t->gne[2] = lead.sim->ne[2];
    t->neGlobal = (long long)t->gne[0] * t->gne[1] * t->gne[2];
    const bool multi = g->comm ? g->world > 1 : g->parts.size() > 1;
    for (vf_mg *m : g->parts) {
        auto P = std::make_unique<GTopPart>();
        P->mg = m; P->sim = m->sim; P->sb = m->sim->slabBegin; P->se = m->sim->slabEnd;
        P->elo = std::max<int64_t>(0, P->sb - t->R); P->ehi = std::min<int64_t>(t->gne[0], P->se + t->R);
        if (multi && P->se - P->sb < t->R) throw std::runtime_error("a slab must hold at least R = " + std::to_string(t->R) + " element layers");
        P->layer = (long long)t->gne[1] * t->gne[2]; P->nOwn = (P->se - P->sb) * P->layer; P->nExt = (P->ehi - P->elo) * P->layer;
        const size_t len = (size_t)m->sim->g.numNodes * 3;
        P->x.alloc(P->nOwn, true); P->stepped.alloc(P->nOwn, true); P->dJ.alloc(P->nOwn, true); P->dc.alloc(P->nOwn, true);
        P->extA.alloc(P->nExt, true); P->extB.alloc(P->nExt, true); P->gwin.alloc(m->sim->g.numElems, true);
        P->u.alloc(len, true); P->f.alloc(len, true); P->scalar.alloc(1, true); P->scratch.alloc(reduce_scratch_doubles(), true);
        for (int i = 0; i <= nf; ++i) { P->vars.push_back(std::make_unique<DevBuf<double>>()); P->vars.back()->alloc(P->nExt, true); }
        sim_build_load_dev(*m->sim, P->f.p);
        t->parts.push_back(std::move(P));
    }
    VF_CUDA(cudaStreamSynchronize(lead.ctx.stream));
    *out = t.release();
    VF_CATCH
}
int vf_group_top_destroy(vf_gtop *t) { VF_TRY if (t) { cudaStreamSynchronize(t->stream()); delete t; } VF_CATCH }
int64_t vf_group_top_halo_layers(const vf_gtop *t) { return t->R; }
int vf_group_top_set_solver(vf_gtop *t, int cgIter, double tol, int mgIt, int mgSmooth, int fmg, int zeroInit) { t->cgIter = cgIter; t->tol = tol; t->mgIt = mgIt; t->mgSmooth = mgSmooth; t->fmg = fmg != 0; t->zeroInit = zeroInit != 0; return 0; }
// x: design variables of the WHOLE grid (host, numElements of the global grid); every part keeps the layers it owns
int vf_group_top_set_vars(vf_gtop *t, const double *x) {
    VF_TRY for (auto &pp : t->parts) h2d(pp->x.p, x + pp->sb * pp->layer, (size_t)pp->nOwn, t->stream());
    t->update(); VF_CUDA(cudaStreamSynchronize(t->stream())); VF_CATCH
}
// which: 0 design variables, 1 physical densities; out: whole grid (host); every rank receives the whole array
int vf_group_top_get_vars(vf_gtop *t, int which, double *out) {
    VF_TRY if (which == 0) t->gatherTo([](GTopPart &P) { return P.x.p; }, out);
    else t->gatherTo([](GTopPart &P) { return P.own(P.vars.back()->p); }, out); VF_CATCH
}
int vf_group_top_compliance(vf_gtop *t, double *out) { VF_TRY *out = t->compliance(); VF_CATCH }
int vf_group_top_constraint(vf_gtop *t, double *out) {
    VF_TRY std::vector<double *> last; for (auto &pp : t->parts) last.push_back(pp->vars.back()->p); *out = t->volumeConstraint(last); VF_CATCH
}
int vf_group_top_objective_gradient(vf_gtop *t, double *g) { VF_TRY t->gradients(); t->gatherTo([](GTopPart &P) { return P.dJ.p; }, g); VF_CATCH }
int vf_group_top_constraint_jacobian(vf_gtop *t, double *g) { VF_TRY t->gradients(); t->gatherTo([](GTopPart &P) { return P.dc.p; }, g); VF_CATCH }
int vf_group_top_last_pcg_iterations(vf_gtop *t) { return t->lastPcgIters; }
// this rank's (first local part's) displacement window, component-major
int vf_group_top_get_u(vf_gtop *t, int part, double *u) { VF_TRY GTopPart &P = *t->parts.at(part); d2h(u, P.u.p, P.u.n, t->stream()); VF_CATCH }
// OCOptimizer::step (OptimalityCriterion.hh:51-134) on the partitioned problem
int vf_group_top_oc_step(vf_gtop *t, double m, double p, double ctol, int *nevalsOut) {
    VF_TRY
    TraceScope ts("OC step");
    t->gradients();
    int nevals = 0;
    auto ceval = [&](double lam) { ++nevals; return t->ceval(lam, m, p); };
    {
        TraceScope tb("Bisection");
        const double dilation = 32;
        double mid = 0.5 * (t->lamMin + t->lamMax);
        t->lamMax = dilation * t->lamMax + (1 - dilation) * mid;
        t->lamMin = std::max(dilation * t->lamMin + (1 - dilation) * mid, 0.01);
        const int guard = 100; int nit = 0;
        for (; nit < guard; ++nit) { if (ceval(t->lamMin) < 0) break; t->lamMax = t->lamMin; t->lamMin /= 2; }
        if (nit == guard) throw std::runtime_error("Bracketing constraint(lambda_min) < 0 failed (100 times).");
        if (nit == 0) for (; nit < guard; ++nit) { if (ceval(t->lamMax) > 0) break; t->lamMin = t->lamMax; t->lamMax *= 2; }
        if (nit == guard) throw std::runtime_error("Bracketing constraint(lambda_max) > 0 failed (100 times).");
        double violation;
        do {
            mid = 0.5 * (t->lamMin + t->lamMax);
            violation = ceval(mid);
            if (std::abs(violation) <= ctol) break;
            if (violation < 0) t->lamMin = mid;
            if (violation > 0) t->lamMax = mid;
        } while (true);
    }
    for (auto &pp : t->parts) VF_CUDA(cudaMemcpyAsync(pp->x.p, pp->stepped.p, sizeof(double) * pp->nOwn, cudaMemcpyDeviceToDevice, t->stream()));
    t->update();
    VF_CUDA(cudaStreamSynchronize(t->stream()));
    if (nevalsOut) *nevalsOut = nevals;
    VF_CATCH
}
} // extern "C"

// ---------------------------------------------------------------------------
// Layer-by-layer evaluator (LayerByLayer.hh:25-309)
// ---------------------------------------------------------------------------
// Initial-guess generators: Zero (:56-61), FD with history 1 ("constant") or 2 ("fd") (:89-101) and
// Subspace N=k (:103-209).  The reference maintains the k x k subspace system  A = U^T K U, b = U^T f  with
// band-limited recurrences (:149-202, TensorProductSimulator.hh:1309-1406) that are exact identities for
// A and b *of the next layer's system*; here A and b are evaluated directly for the current (next) system at
// constructGuess time with k stiffness applies and k(k+3)/2 masked dot products -- same values up to rounding.
struct vf_lbl {
    vf_mg *mg; vf_sim *sim;
    int method = 2; // 0 zero, 1 FD, 2 subspace
    size_t maxHist = 3;
    std::vector<std::unique_ptr<DevBuf<double>>> hist; // front = most recent
    DevBuf<double> f, u, uFull, w, totalGrad, layerGrad, scalar, scratch;
    bool uValid = false, uFullValid = false;
    double totalCompliance = 0; long long layersAccumulated = 0;
    double *hostScalar = nullptr;
    std::vector<int> layerIters;
    ~vf_lbl() { if (hostScalar) cudaFreeHost(hostScalar); }
    size_t len() const { return (size_t)sim->g.numNodes * sim->N; }
    double maskedDot(const double *a, const double *b) {
        launch_masked_dot(mg->ctx, sim->g, a, b, scalar.p, scratch.p);
        VF_CUDA(cudaMemcpyAsync(hostScalar, scalar.p, sizeof(double), cudaMemcpyDeviceToHost, mg->ctx.stream));
        VF_CUDA(cudaStreamSynchronize(mg->ctx.stream));
        return *hostScalar;
    }
    void addToHistory() { // m_addToHistory (:72-84): the solved field moves to the front, the stalest buffer is recycled as `u`
        if (maxHist == 0) return;
        std::unique_ptr<DevBuf<double>> slot;
        if (hist.size() == maxHist) { slot = std::move(hist.back()); hist.pop_back(); }
        else { slot = std::make_unique<DevBuf<double>>(); slot->alloc(len(), true); }
        std::swap(slot->p, u.p); std::swap(slot->n, u.n);
        hist.insert(hist.begin(), std::move(slot));
        uValid = false;
    }
    // pseudo-inverse solve of the symmetric k x k system (Eigen::JacobiSVD::solve semantics, :120-121)
    static std::vector<double> solveSymmetricPinv(std::vector<double> A, const std::vector<double> &b, int k) {
        std::vector<double> V((size_t)k * k, 0.0);
        for (int i = 0; i < k; ++i) V[i * k + i] = 1;
        for (int sweep = 0; sweep < 64; ++sweep) {
            double off = 0; for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j) off += A[i * k + j] * A[i * k + j];
            if (off < 1e-300) break;
            for (int p = 0; p < k; ++p) for (int q = p + 1; q < k; ++q) {
                if (A[p * k + q] == 0) continue;
                const double theta = (A[q * k + q] - A[p * k + p]) / (2 * A[p * k + q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), sn = t * c;
                for (int r = 0; r < k; ++r) { const double arp = A[r * k + p], arq = A[r * k + q]; A[r * k + p] = c * arp - sn * arq; A[r * k + q] = sn * arp + c * arq; }
                for (int r = 0; r < k; ++r) { const double apr = A[p * k + r], aqr = A[q * k + r]; A[p * k + r] = c * apr - sn * aqr; A[q * k + r] = sn * apr + c * aqr; }
                for (int r = 0; r < k; ++r) { const double vrp = V[r * k + p], vrq = V[r * k + q]; V[r * k + p] = c * vrp - sn * vrq; V[r * k + q] = sn * vrp + c * vrq; }
            }
        }
        double smax = 0; for (int i = 0; i < k; ++i) smax = std::max(smax, std::abs(A[i * k + i]));
        const double thresh = std::numeric_limits<double>::epsilon() * k * smax;
        std::vector<double> x(k, 0.0);
        for (int i = 0; i < k; ++i) {
            const double lam = A[i * k + i];
            if (std::abs(lam) <= thresh) continue;
            double proj = 0; for (int r = 0; r < k; ++r) proj += V[r * k + i] * b[r];
            for (int r = 0; r < k; ++r) x[r] += V[r * k + i] * proj / lam;
        }
        return x;
    }
    void constructGuess() {
        const size_t s = hist.size();
        if (u.n != len()) u.alloc(len(), true);
        if (method == 0 || s == 0) { VF_CUDA(cudaMemsetAsync(u.p, 0, sizeof(double) * len(), mg->ctx.stream)); uValid = true; return; }
        if (method == 1) { // InitGenFD (:95-100)
            VF_CUDA(cudaMemcpyAsync(u.p, hist[0]->p, sizeof(double) * len(), cudaMemcpyDeviceToDevice, mg->ctx.stream));
            if (s == 2) { launch_scale(mg->ctx, (long long)len(), 2.0, u.p); launch_axpy(mg->ctx, (long long)len(), -1.0, hist[1]->p, u.p); }
            if (s > 2) throw std::runtime_error("Unimplemented");
            uValid = true; return;
        }
        // InitGenSubspace (:110-147): solve A c = b, g = sum c_i u_i on the attached nodes, zero above
        const int k = (int)s;
        std::vector<double> A((size_t)k * k, 0.0), b(k, 0.0);
        if (w.n != len()) w.alloc(len(), true);
        for (int j = 0; j < k; ++j) {
            b[j] = maskedDot(hist[j]->p, f.p);
            launch_apply_l0(mg->ctx, sim->g, sim->K0p, hist[j]->p, sim->E.p, nullptr, nullptr, w.p, APPLY_SET, nullptr, nullptr);
            for (int i = j; i < k; ++i) A[i * k + j] = A[j * k + i] = maskedDot(hist[i]->p, w.p);
        }
        const std::vector<double> c = solveSymmetricPinv(A, b, k);
        VF_CUDA(cudaMemsetAsync(u.p, 0, sizeof(double) * len(), mg->ctx.stream));
        for (int i = 0; i < k; ++i) launch_axpy(mg->ctx, (long long)len(), c[i], hist[i]->p, u.p);
        launch_detached_zero(mg->ctx, sim->g, u.p);
        uValid = true;
    }
};

extern "C" {
int vf_lbl_create(vf_mg *mg, vf_lbl **out) {
    VF_TRY auto l = std::make_unique<vf_lbl>(); l->mg = mg; l->sim = mg->sim;
    l->scalar.alloc(1, true); l->scratch.alloc(reduce_scratch_doubles(), true);
    VF_CUDA(cudaMallocHost(&l->hostScalar, sizeof(double)));
    *out = l.release(); VF_CATCH
}
int vf_lbl_destroy(vf_lbl *l) { VF_TRY if (l) { cudaStreamSynchronize(l->mg->ctx.stream); delete l; } VF_CATCH }
int vf_lbl_select_init_method(vf_lbl *l, const char *method) { // selectInitMethod (:214-220)
    VF_TRY const std::string m(method);
    if (m == "zero") { l->method = 0; l->maxHist = 0; }
    else if (m == "constant") { l->method = 1; l->maxHist = 1; }
    else if (m == "fd") { l->method = 1; l->maxHist = 2; }
    else if (m.substr(0, 2) == "N=") { l->method = 2; l->maxHist = (size_t)std::stoi(m.substr(2)); }
    else throw std::runtime_error("Unrecognized method " + m);
    l->hist.clear(); VF_CATCH
}
// LayerByLayerEvaluator::run (:223-296)
int vf_lbl_run(vf_lbl *l, int zeroInit, int64_t layerIncrement, int maxIter, double tol, int mgIt, int mgSmooth, int fmg, vf_lbl_callback cb, void *user,
               vf_pcg_callback pcgCb, void *pcgUser) {
    VF_TRY
    vf_mg &mg = *l->mg; vf_sim &sim = *l->sim;
    const int64_t numLayers = sim.ne[1];
    const size_t len = l->len();
    if (layerIncrement < 1) throw std::runtime_error("layerIncrement must be positive");
    l->uValid = false;
    if (!zeroInit && l->uFullValid) { if (l->u.n != len) l->u.alloc(len, false); VF_CUDA(cudaMemcpyAsync(l->u.p, l->uFull.p, sizeof(double) * len, cudaMemcpyDeviceToDevice, mg.ctx.stream)); l->uValid = true; }
    l->hist.clear();                                           // m_initGen->reset() (:237)
    l->layersAccumulated = 0; l->totalCompliance = 0; l->layerIters.clear();
    l->totalGrad.alloc(sim.g.numElems, true);
    trace_push("Build load");                                  // (:243)
    if (int rc = vf_mg_set_mask_layer(&mg, numLayers)) { trace_pop("Build load"); return rc; } // (:244)
    if (l->f.n != len) l->f.alloc(len, true);
    sim_build_load_dev(sim, l->f.p);                           // (:245)
    trace_pop("Build load");
    for (int64_t layer = numLayers; layer > 0; layer -= std::min(layerIncrement, layer)) {
        if (layer < numLayers) {
            if (int rc = vf_mg_decrement_mask(&mg, (int)layerIncrement)) return rc;               // (:250)
            const double g2 = sim.gravity[0] * sim.gravity[0] + sim.gravity[1] * sim.gravity[1] + sim.gravity[2] * sim.gravity[2];
            if (g2 == 0 || std::abs(g2 - sim.gravity[1] * sim.gravity[1]) > 1e-10) throw std::runtime_error("Unexpected gravity vector");
            TraceScope tl("Update load");                                                        // (:251)
            launch_self_weight_load(mg.ctx, sim.g, sim.rho.p, sim.gravity, sim.elemVolume(), l->f.p, (int)layer, (int)(layer + layerIncrement), -1.0); // (:252)
        }
        if (layer < numLayers || !l->uValid) { TraceScope tg("Construct initial guess"); l->constructGuess(); } // (:258-260)
        try {
            mg_pcg(mg, l->u.p, l->f.p, maxIter, tol, mgIt, mgSmooth, fmg != 0, /* dirichletAlreadySatisfied */ true, pcgCb, pcgUser); // (:265)
        } catch (const std::exception &e) { throw std::runtime_error(std::string("PCG exception ") + e.what() + " at l = " + std::to_string(layer)); }
        l->layerIters.push_back(mg.lastIters);
        trace_push("Compute compliance");                                                          // (:272)
        const double compliance = l->maskedDot(l->f.p, l->u.p);                                    // (:273)
        trace_pop("Compute compliance");
        if (cb) cb(layer, compliance, mg.lastIters, user);
        l->totalCompliance += compliance;
        TraceScope tgr("Compute gradient");                                                        // (:282)
        launch_compliance_gradient(mg.ctx, sim.g, sim.K0p, l->u.p, sim.rho.p, l->totalGrad.p, sim.law, sim.E0, sim.Emin, sim.gamma, sim.q, sim.gravity, sim.elemVolume(), true); // (:283)
        ++l->layersAccumulated;
        if (layer == numLayers) { if (l->uFull.n != len) l->uFull.alloc(len, false); VF_CUDA(cudaMemcpyAsync(l->uFull.p, l->u.p, sizeof(double) * len, cudaMemcpyDeviceToDevice, mg.ctx.stream)); l->uFullValid = true; }
        if (layer >= layerIncrement) l->addToHistory();                                            // finalizeLayer (:290-294)
    }
    VF_CUDA(cudaStreamSynchronize(mg.ctx.stream));
    VF_CATCH
}
// Inside a vf_lbl_callback: the layer's displacement and its compliance gradient, the last two arguments of the reference's
// lblCallback(l, compliance, grad_compliance, u) (LayerByLayer.hh:222, 277-279) -- fetched only when the callback asks for them.
int vf_lbl_get_layer_u(vf_lbl *l, double *u) { VF_TRY d2h(u, l->u.p, l->len(), l->mg->ctx.stream); VF_CATCH }
int vf_lbl_get_layer_gradient(vf_lbl *l, double *g) {
    VF_TRY vf_sim &sim = *l->sim; vf_mg &mg = *l->mg;
    if (l->layerGrad.n != (size_t)sim.g.numElems) l->layerGrad.alloc(sim.g.numElems, false);
    launch_compliance_gradient(mg.ctx, sim.g, sim.K0p, l->u.p, sim.rho.p, l->layerGrad.p, sim.law, sim.E0, sim.Emin, sim.gamma, sim.q, sim.gravity, sim.elemVolume(), false); // complianceGradientFlattened(u)
    d2h(g, l->layerGrad.p, (size_t)sim.g.numElems, mg.ctx.stream); VF_CATCH
}
int vf_lbl_objective(vf_lbl *l, double *out) { *out = 0.5 * l->totalCompliance / double(l->layersAccumulated); return 0; } // (:299)
int vf_lbl_gradient(vf_lbl *l, double *g) { // (:300)
    VF_TRY std::vector<double> h(l->sim->g.numElems); d2h(h.data(), l->totalGrad.p, h.size(), l->mg->ctx.stream);
    for (size_t i = 0; i < h.size(); ++i) g[i] = h[i] / double(l->layersAccumulated); VF_CATCH
}
}

// ---------------------------------------------------------------------------
// Layer-by-layer evaluator on a slab group (LayerByLayer.hh:25-309 with the grid partitioned along axis 0).  The build direction is
// axis 1, so every slab holds a piece of every layer: the fabrication mask, the self-weight load update, the initial-guess history
// and the compliance gradient are per-part operations on the part's window; what couples the parts is the partitioned MG-PCG
// (vf_group_pcg), the ghost planes of the history fields before the stiffness applies of the subspace guess, and the scalar sums
// (owned node planes only, summed over parts / all-reduced over ranks).  The coarse hierarchy is rebuilt in full for every layer
// (the banded update of the undivided solver is not partitioned).
// ---------------------------------------------------------------------------
struct GLblPart {
    vf_mg *mg = nullptr; vf_sim *sim = nullptr;
    std::vector<std::unique_ptr<DevBuf<double>>> hist;
    DevBuf<double> f, u, uFull, w, totalGrad, scalar, scratch;
    size_t len() const { return (size_t)sim->g.numNodes * sim->N; }
};
struct vf_glbl {
    vf_group *grp = nullptr; std::vector<std::unique_ptr<GLblPart>> parts;
    int method = 2; size_t maxHist = 3;
    bool uValid = false, uFullValid = false;
    double totalCompliance = 0; long long layersAccumulated = 0; std::vector<int> layerIters;
    DevBuf<double> gather;
    vf_mg &lead() { return *grp->parts[0]; }
    cudaStream_t stream() { return lead().ctx.stream; }
    // sum over all parts / ranks of one masked dot product per part (owned node planes only)
    template<class SelA, class SelB> double dot(SelA a, SelB b) {
        for (auto &pp : parts) launch_masked_dot(pp->mg->ctx, pp->sim->g, a(*pp), b(*pp), pp->scalar.p, pp->scratch.p);
        if (grp->comm) { NcclApi &A = NcclApi::get(); double *p = parts[0]->scalar.p; A.check(A.AllReduce(p, p, 1, kNcclDouble, kNcclSum, grp->comm, stream()), "ncclAllReduce"); }
        double tot = 0;
        for (auto &pp : parts) { double v = 0; d2h(&v, pp->scalar.p, 1, stream()); tot += v; if (grp->comm) break; }
        return tot;
    }
    void addToHistory() {   // m_addToHistory (:72-84)
        if (maxHist == 0) return;
        for (auto &pp : parts) {
            GLblPart &P = *pp; std::unique_ptr<DevBuf<double>> slot;
            if (P.hist.size() == maxHist) { slot = std::move(P.hist.back()); P.hist.pop_back(); }
            else { slot = std::make_unique<DevBuf<double>>(); slot->alloc(P.len(), true); }
            std::swap(slot->p, P.u.p); std::swap(slot->n, P.u.n);
            P.hist.insert(P.hist.begin(), std::move(slot));
        }
        uValid = false;
    }
    void constructGuess() {
        const size_t s = parts[0]->hist.size();
        for (auto &pp : parts) if (pp->u.n != pp->len()) pp->u.alloc(pp->len(), true);
        if (method == 0 || s == 0) { for (auto &pp : parts) VF_CUDA(cudaMemsetAsync(pp->u.p, 0, sizeof(double) * pp->len(), stream())); uValid = true; return; }
        if (method == 1) {   // InitGenFD (:95-100)
            if (s > 2) throw std::runtime_error("Unimplemented");
            for (auto &pp : parts) {
                GLblPart &P = *pp;
                VF_CUDA(cudaMemcpyAsync(P.u.p, P.hist[0]->p, sizeof(double) * P.len(), cudaMemcpyDeviceToDevice, stream()));
                if (s == 2) { launch_scale(P.mg->ctx, (long long)P.len(), 2.0, P.u.p); launch_axpy(P.mg->ctx, (long long)P.len(), -1.0, P.hist[1]->p, P.u.p); }
            }
            uValid = true; return;
        }
        // InitGenSubspace (:110-147): A c = b with A = U^T K U, b = U^T f over the whole grid
        const int k = (int)s;
        std::vector<double> A((size_t)k * k, 0.0), b(k, 0.0);
        for (auto &pp : parts) if (pp->w.n != pp->len()) pp->w.alloc(pp->len(), true);
        for (int j = 0; j < k; ++j) {
            b[j] = dot([&](GLblPart &P) { return P.hist[j]->p; }, [](GLblPart &P) { return P.f.p; });
            // the stiffness apply reads the ghost planes of the history field: the solve left them current for the latest field only
            std::vector<double *> hp; for (auto &pp : parts) hp.push_back(pp->hist[j]->p);
            grp_exchange(lead(), 0, [&](vf_mg &m) { for (size_t q = 0; q < parts.size(); ++q) if (parts[q]->mg == &m) return hp[q]; return (double *)nullptr; });
            for (auto &pp : parts) part_apply_K(*pp->mg, 0, pp->hist[j]->p, nullptr, pp->w.p, APPLY_SET, false);
            for (int i = j; i < k; ++i) A[i * k + j] = A[j * k + i] = dot([&](GLblPart &P) { return P.hist[i]->p; }, [](GLblPart &P) { return P.w.p; });
        }
        const std::vector<double> c = vf_lbl::solveSymmetricPinv(A, b, k);
        for (auto &pp : parts) {
            GLblPart &P = *pp;
            VF_CUDA(cudaMemsetAsync(P.u.p, 0, sizeof(double) * P.len(), stream()));
            for (int i = 0; i < k; ++i) launch_axpy(P.mg->ctx, (long long)P.len(), c[i], P.hist[i]->p, P.u.p);
            launch_detached_zero(P.mg->ctx, P.sim->g, P.u.p);
        }
        uValid = true;
    }
};

extern "C" {
int vf_group_lbl_create(vf_group *g, vf_glbl **out) {
    VF_TRY auto l = std::make_unique<vf_glbl>(); l->grp = g;
    if (g->parts[0]->N != 3) throw std::runtime_error("the slab-partitioned layer-by-layer evaluator is 3D");
    for (vf_mg *m : g->parts) {
        auto P = std::make_unique<GLblPart>(); P->mg = m; P->sim = m->sim;
        P->scalar.alloc(1, true); P->scratch.alloc(reduce_scratch_doubles(), true);
        l->parts.push_back(std::move(P));
    }
    *out = l.release(); VF_CATCH
}
int vf_group_lbl_destroy(vf_glbl *l) { VF_TRY if (l) { cudaStreamSynchronize(l->stream()); delete l; } VF_CATCH }
int vf_group_lbl_select_init_method(vf_glbl *l, const char *method) {   // selectInitMethod (:214-220)
    VF_TRY const std::string m(method);
    if (m == "zero") { l->method = 0; l->maxHist = 0; }
    else if (m == "constant") { l->method = 1; l->maxHist = 1; }
    else if (m == "fd") { l->method = 1; l->maxHist = 2; }
    else if (m.substr(0, 2) == "N=") { l->method = 2; l->maxHist = (size_t)std::stoi(m.substr(2)); }
    else throw std::runtime_error("Unrecognized method " + m);
    for (auto &pp : l->parts) pp->hist.clear(); VF_CATCH
}
// LayerByLayerEvaluator::run (:223-296) on the partitioned grid; every rank calls it with the same arguments
int vf_group_lbl_run(vf_glbl *l, int zeroInit, int64_t layerIncrement, int maxIter, double tol, int mgIt, int mgSmooth, int fmg, vf_lbl_callback cb, void *user) {
    VF_TRY
    vf_mg &lead = l->lead(); vf_sim &sim0 = *lead.sim;
    const int64_t numLayers = sim0.ne[1];
    if (layerIncrement < 1) throw std::runtime_error("layerIncrement must be positive");
    cudaStream_t st = l->stream();
    l->uValid = false;
    if (!zeroInit && l->uFullValid) {
        for (auto &pp : l->parts) { GLblPart &P = *pp; if (P.u.n != P.len()) P.u.alloc(P.len(), false); VF_CUDA(cudaMemcpyAsync(P.u.p, P.uFull.p, sizeof(double) * P.len(), cudaMemcpyDeviceToDevice, st)); }
        l->uValid = true;
    }
    l->layersAccumulated = 0; l->totalCompliance = 0; l->layerIters.clear();
    for (auto &pp : l->parts) {
        GLblPart &P = *pp;
        P.hist.clear(); P.totalGrad.alloc(P.sim->g.numElems, true);
        if (int rc = vf_mg_set_mask_layer(P.mg, numLayers)) return rc;
        if (P.f.n != P.len()) P.f.alloc(P.len(), true);
        sim_build_load_dev(*P.sim, P.f.p);
    }
    const double g2 = sim0.gravity[0] * sim0.gravity[0] + sim0.gravity[1] * sim0.gravity[1] + sim0.gravity[2] * sim0.gravity[2];
    for (int64_t layer = numLayers; layer > 0; layer -= std::min(layerIncrement, layer)) {
        if (layer < numLayers) {
            if (g2 == 0 || std::abs(g2 - sim0.gravity[1] * sim0.gravity[1]) > 1e-10) throw std::runtime_error("Unexpected gravity vector");
            for (auto &pp : l->parts) {
                GLblPart &P = *pp;
                if (int rc = vf_mg_decrement_mask(P.mg, (int)layerIncrement)) return rc;
                launch_self_weight_load(P.mg->ctx, P.sim->g, P.sim->rho.p, P.sim->gravity, P.sim->elemVolume(), P.f.p, (int)layer, (int)(layer + layerIncrement), -1.0);
            }
        }
        if (layer < numLayers || !l->uValid) l->constructGuess();
        std::vector<double *> xs; std::vector<const double *> bs;
        for (auto &pp : l->parts) { xs.push_back(pp->u.p); bs.push_back(pp->f.p); }
        try { mg_pcg(lead, xs.data(), bs.data(), maxIter, tol, mgIt, mgSmooth, fmg != 0, /* dirichletAlreadySatisfied */ true, nullptr, nullptr); }
        catch (const std::exception &e) { throw std::runtime_error(std::string("PCG exception ") + e.what() + " at l = " + std::to_string(layer)); }
        l->layerIters.push_back(lead.lastIters);
        const double compliance = l->dot([](GLblPart &P) { return P.f.p; }, [](GLblPart &P) { return P.u.p; });
        if (cb) cb(layer, compliance, lead.lastIters, user);
        l->totalCompliance += compliance;
        for (auto &pp : l->parts) {
            GLblPart &P = *pp; vf_sim &sm = *P.sim;
            launch_compliance_gradient(P.mg->ctx, sm.g, sm.K0p, P.u.p, sm.rho.p, P.totalGrad.p, sm.law, sm.E0, sm.Emin, sm.gamma, sm.q, sm.gravity, sm.elemVolume(), true);
            if (layer == numLayers) { if (P.uFull.n != P.len()) P.uFull.alloc(P.len(), false); VF_CUDA(cudaMemcpyAsync(P.uFull.p, P.u.p, sizeof(double) * P.len(), cudaMemcpyDeviceToDevice, st)); }
        }
        if (layer == numLayers) l->uFullValid = true;
        ++l->layersAccumulated;
        if (layer >= layerIncrement) l->addToHistory();
    }
    p2p_check(*l->grp, st);
    VF_CUDA(cudaStreamSynchronize(st));
    VF_CATCH
}
int vf_group_lbl_objective(vf_glbl *l, double *out) { *out = 0.5 * l->totalCompliance / double(l->layersAccumulated); return 0; }   // (:299)
// gradient (:300) over the WHOLE grid (host array, every rank receives all of it)
int vf_group_lbl_gradient(vf_glbl *l, double *g) {
    VF_TRY
    vf_sim &s0 = *l->lead().sim; cudaStream_t st = l->stream();
    const long long layer = (long long)s0.ne[1] * s0.ne[2], neGlobal = (long long)s0.gne0 * layer;
    if (l->gather.n != (size_t)neGlobal) l->gather.alloc(neGlobal, false);
    VF_CUDA(cudaMemsetAsync(l->gather.p, 0, sizeof(double) * neGlobal, st));
    for (auto &pp : l->parts) {
        vf_sim &sm = *pp->sim;
        VF_CUDA(cudaMemcpyAsync(l->gather.p + sm.slabBegin * layer, pp->totalGrad.p + (sm.slabBegin - sm.xoff) * layer, sizeof(double) * (sm.slabEnd - sm.slabBegin) * layer, cudaMemcpyDeviceToDevice, st));
    }
    if (l->grp->comm) { NcclApi &A = NcclApi::get(); A.check(A.AllReduce(l->gather.p, l->gather.p, (size_t)neGlobal, kNcclDouble, kNcclSum, l->grp->comm, st), "ncclAllReduce"); }
    std::vector<double> h(neGlobal); d2h(h.data(), l->gather.p, (size_t)neGlobal, st);
    for (long long i = 0; i < neGlobal; ++i) g[i] = h[i] / double(l->layersAccumulated);
    VF_CATCH
}
int vf_group_lbl_num_layer_iterations(vf_glbl *l, int *iters /* one per simulated layer, may be NULL */) {
    if (iters) std::copy(l->layerIters.begin(), l->layerIters.end(), iters);
    return (int)l->layerIters.size();
}
}
