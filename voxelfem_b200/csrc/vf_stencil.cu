// vf_stencil.cu -- coarse-level operators stored as 3^N-point block stencils (sm_100a).
//
// The reference keeps levels >= 2 as a block-CSC matrix ("blockK", TensorProductSimulator.hh:
// 885-966; applied by CSCMatrix::applyTransposeParallel, MeshFEM SparseMatrices.hh:1613-1677,
// smoothed with NodeSmoothStencilBlockK, MultigridSolver.hh:323-334), rebuilds the level-1 operator
// on the fly from the fine moduli (MultigridSolver.hh:294-321, 475-499) and caches per-element
// matrices at the coarsest level (:815-817).  On a structured grid all three are the same object: a
// 3^N-point stencil of N x N blocks per node.  Here every level >= 1 stores that stencil in a
// slot-major SoA layout  S[(slot*N*N + a*N + b) * numNodes + pos(node)]  where pos() is the colour-major node
// numbering of GridDesc (stencil_pos): a colour pass of the smoother streams its rows contiguously.
//
// Galerkin coarsening (MultigridSolver.hh:711-819) is done directly on stencils: level 1 from the
// fine moduli and the 2^N matrices coarsenedFineK0s[fi]; level l >= 2 as P^T A_{l-1} P.  Both are
// algebraically identical to the reference's per-element recursion (sum over coarse elements of
// Phi^T Ke Phi), differing only in floating-point summation order.
#include "vf_internal.cuh"
#include "vf_reduce.cuh"

namespace vf {

// Slot-parallel stencil row evaluation.  A thread block handles 32 consecutive nodes (in colour-major order)
// with one warp-row per stencil slot: thread (tx, s) loads the N x N block S[s](node tx) (coalesced across tx)
// and the neighbour displacement, and contributes S u to a shared-memory reduction over the 3^N slots.  Every
// thread issues a single batch of independent loads, so the latency chain of a row is one memory round trip
// instead of 3^N dependent ones -- this is what matters on the small coarse levels that sit on the critical
// path of every V-cycle, while the large level 1 streams its 1944 B/node of stencil at HBM rate.
constexpr int kSlotNodes = 16; // nodes per block (16 consecutive doubles = 128 B = 4 full sectors per load)
template<int N>
struct SlotShared {
    double red[N][Dims<N>::NS][kSlotNodes];
    double Md[N * N][kSlotNodes];
    double us[N][kSlotNodes];
    double bs[N][kSlotNodes];
    unsigned dm[kSlotNodes];
};

// coordinates of the q-th node: either of one colour pass (col != nullptr) or of the whole grid in colour-major order
template<int N>
__device__ __forceinline__ bool slot_node_coords(const GridDesc &g, const ColorDesc *col, long long q, int (&c)[3]) {
    if (col) {
        const long long tot = (long long)col->cnt[0] * col->cnt[1] * col->cnt[2];
        if (q >= tot) return false;
        const int i2 = (int)(q % col->cnt[2]); q /= col->cnt[2];
        const int i1 = (int)(q % col->cnt[1]); const int i0 = (int)(q / col->cnt[1]);
        c[0] = col->off[0] + 2 * i0; c[1] = col->off[1] + 2 * i1; c[2] = col->off[2] + 2 * i2;
        return true;
    }
    if (q >= g.numNodes) return false;
    int cc = 0;
    #pragma unroll
    for (int k = 1; k < 8; ++k) cc += (q >= g.cbase[k]) ? 1 : 0;   // cbase is non-decreasing
    long long idx = q - g.cbase[cc];
    const int i2 = (int)(idx % g.ccnt[cc][2]); idx /= g.ccnt[cc][2];
    const int i1 = (int)(idx % g.ccnt[cc][1]); const int i0 = (int)(idx / g.ccnt[cc][1]);
    c[0] = 2 * i0 + ((cc >> 2) & 1); c[1] = 2 * i1 + ((cc >> 1) & 1); c[2] = 2 * i2 + (cc & 1);
    return true;
}

template<int N, bool GS, int MODE>
__global__ void __launch_bounds__(kSlotNodes * Dims<N>::NS)
k_stencil_slots(const __grid_constant__ GridDesc g, const __grid_constant__ ColorDesc col, const double *__restrict__ S,
                const double *uin, const double *__restrict__ b, const uint8_t *__restrict__ dmask,
                double *out, int forward) {
    constexpr int NS = Dims<N>::NS, A0 = Dims<N>::A0, NN = N * N;
    __shared__ SlotShared<N> sh;
    const int tx = threadIdx.x, s = threadIdx.y;
    const long long q = (long long)blockIdx.x * kSlotNodes + tx;
    int c[3] = {0, 0, 0};
    const bool inRange = slot_node_coords<N>(g, GS ? &col : nullptr, q, c);
    const long long n = (long long)c[0] * g.ns[0] + (long long)c[1] * g.ns[1] + c[2];
    const bool detached = inRange && (((g.bd == 1) ? c[1] : c[2]) >= g.nActive);
    double acc[N];
    #pragma unroll
    for (int a = 0; a < N; ++a) acc[a] = 0.0;
    if (inRange && !detached) {
        // stage b and the Dirichlet mask alongside the stencil loads so that the finalising warp has no global loads left
        if (s < N && (GS || MODE == APPLY_RESIDUAL)) sh.bs[s][tx] = b[s * g.numNodes + n];
        if (s == N) sh.dm[tx] = dmask ? dmask[n] : 0u;
        int d[3] = {0, 0, 0};
        { int r = s;
          #pragma unroll
          for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
        bool valid = true; long long off = 0;
        #pragma unroll
        for (int a = A0; a < 3; ++a) { const int qq = c[a] + d[a]; valid = valid && qq >= 0 && qq < g.nn[a]; off += (long long)d[a] * g.ns[a]; }
        if (valid) {
            const long long p = stencil_pos(g, c[0], c[1], c[2]);
            double un[N], sv[NN];
            #pragma unroll
            for (int k = 0; k < N; ++k) un[k] = uin[k * g.numNodes + n + off];
            #pragma unroll
            for (int k = 0; k < NN; ++k) sv[k] = __ldg(S + (long long)(s * NN + k) * g.numNodes + p);
            #pragma unroll
            for (int a = 0; a < N; ++a) {
                #pragma unroll
                for (int k = 0; k < N; ++k) acc[a] = fma(sv[a * N + k], un[k], acc[a]);
            }
            if (s == NS / 2) {
                #pragma unroll
                for (int k = 0; k < NN; ++k) sh.Md[k][tx] = sv[k];
                #pragma unroll
                for (int k = 0; k < N; ++k) sh.us[k][tx] = un[k];
            }
        }
    }
    #pragma unroll
    for (int a = 0; a < N; ++a) sh.red[a][s][tx] = acc[a];
    __syncthreads();
    if (s < N) { // warp-row a = s sums the slot contributions of component a (fixed order -> deterministic)
        double t = 0.0;
        #pragma unroll
        for (int k = 0; k < NS; ++k) t += sh.red[s][k][tx];
        sh.red[s][0][tx] = t;
    }
    __syncthreads();
    if (s != 0 || !inRange) return;
    if (detached) {
        if (!GS && MODE == APPLY_SET) {
            #pragma unroll
            for (int a = 0; a < N; ++a) out[a * g.numNodes + n] = 0.0;
        }
        return;
    }
    const unsigned dm = sh.dm[tx];
    if (GS) {
        if (dm == (unsigned)((1 << N) - 1)) return; // hasFullDirichlet (MultigridSolver.hh:350)
        double rhs[N], M[N][N], du[N];
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            rhs[a] = sh.bs[a][tx] - sh.red[a][0][tx];
            #pragma unroll
            for (int k = 0; k < N; ++k) M[a][k] = sh.Md[a * N + k][tx];
        }
        gs_node_update<N>(M, rhs, dm, forward != 0, du);
        #pragma unroll
        for (int a = 0; a < N; ++a) out[a * g.numNodes + n] = sh.us[a][tx] + du[a];
    } else {
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            const double av = sh.red[a][0][tx];
            double res;
            if (MODE == APPLY_SET) res = av;
            else if (MODE == APPLY_ADD) res = out[a * g.numNodes + n] + av;
            else if (MODE == APPLY_SUB) res = out[a * g.numNodes + n] - av;
            else res = sh.bs[a][tx] - av;
            if ((dm >> a) & 1u) res = 0.0;
            out[a * g.numNodes + n] = res;
        }
    }
}

void launch_apply_stencil(const LaunchCtx &ctx, const GridDesc &g, const double *S, const double *u, const double *b,
                          const uint8_t *dmask, double *out, int mode) {
    ProfScope ps(ctx, mode == APPLY_RESIDUAL ? PC_RESIDUAL_ST : PC_APPLY_ST, (double)g.numNodes);
    ColorDesc col; make_color(g, 0, col);
    dim3 block(kSlotNodes, g.N == 3 ? 27 : 9), grid((unsigned)((g.numNodes + kSlotNodes - 1) / kSlotNodes));
#define VF_CASE(NN_, M) if (g.N == NN_ && mode == M) k_stencil_slots<NN_, false, M><<<grid, block, 0, ctx.stream>>>(g, col, S, u, b, dmask, out, 1);
    VF_CASE(3, APPLY_SET) VF_CASE(3, APPLY_ADD) VF_CASE(3, APPLY_SUB) VF_CASE(3, APPLY_RESIDUAL)
    VF_CASE(2, APPLY_SET) VF_CASE(2, APPLY_ADD) VF_CASE(2, APPLY_SUB) VF_CASE(2, APPLY_RESIDUAL)
#undef VF_CASE
    VF_KERNEL_CHECK();
}

void launch_gs_stencil(const LaunchCtx &ctx, const GridDesc &g, const double *S, double *u, const double *b,
                       const uint8_t *dmask, int color, bool forward) {
    ColorDesc col;
    if (!make_color(g, color, col)) return;
    const long long tot = (long long)col.cnt[0] * col.cnt[1] * col.cnt[2];
    ProfScope ps(ctx, PC_GS_ST, (double)tot);
    dim3 block(kSlotNodes, g.N == 3 ? 27 : 9), grid((unsigned)((tot + kSlotNodes - 1) / kSlotNodes));
    if (g.N == 3) k_stencil_slots<3, true, APPLY_SET><<<grid, block, 0, ctx.stream>>>(g, col, S, u, b, dmask, u, forward ? 1 : 0);
    else          k_stencil_slots<2, true, APPLY_SET><<<grid, block, 0, ctx.stream>>>(g, col, S, u, b, dmask, u, forward ? 1 : 0);
    VF_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------
// Galerkin coarsening
// ---------------------------------------------------------------------------
// Level-1 stencil from the fine moduli:  A_delta(n) = sum_{coarse e containing n and n+delta}
//   sum_{fi} E_{child(e, fi)} * cK0[fi][ln_e(n) rows, ln_e(n+delta) cols]
// (m_firstLevelCoarsenedStiffnessMatrix, MultigridSolver.hh:724-732, assembled as in :782-814).
// One thread per (coarse node, slot); blockIdx.y = slot so (e, ln, m) are warp-uniform.
template<int N>
__global__ void __launch_bounds__(128)
k_coarsen_from_moduli(const __grid_constant__ GridDesc gc, const __grid_constant__ GridDesc gf,
                      const double *__restrict__ E, const double *__restrict__ cK0, double *__restrict__ Sc) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE, A0 = Dims<N>::A0, NN = N * N;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= gc.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % gc.nn[2]); r /= gc.nn[2]; c[1] = (int)(r % gc.nn[1]); c[0] = (int)(r / gc.nn[1]); }
    int d[3] = {0, 0, 0};
    { int r = s; for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
    double acc[NN];
    #pragma unroll
    for (int i = 0; i < NN; ++i) acc[i] = 0.0;
    bool nbValid = true;
    for (int a = A0; a < 3; ++a) { const int q = c[a] + d[a]; nbValid = nbValid && q >= 0 && q < gc.nn[a]; }
    if (nbValid) {
        for (int e = 0; e < NPE; ++e) { // incident coarse elements; node is local node ln == e
            bool ok = true; int m = 0; int ec[3] = {0, 0, 0};
            for (int a = A0; a < 3; ++a) {
                const int ob = (e >> (2 - a)) & 1;
                ec[a] = c[a] - ob;
                ok = ok && ec[a] >= 0 && ec[a] < gc.ne[a];
                const int mb = d[a] + ob;
                ok = ok && (mb == 0 || mb == 1);
                m |= (mb & 1) << (2 - a);
            }
            if (!ok) continue;
            for (int fi = 0; fi < NPE; ++fi) {
                long long ef = 0;
                for (int a = A0; a < 3; ++a) ef += (long long)(2 * ec[a] + ((fi >> (2 - a)) & 1)) * gf.es[a];
                const double Ef = __ldg(E + ef);
                const double *Kb = cK0 + (size_t)fi * KE * KE;
                #pragma unroll
                for (int a = 0; a < N; ++a) {
                    #pragma unroll
                    for (int b = 0; b < N; ++b) acc[a * N + b] = fma(Ef, __ldg(Kb + (N * e + a) * KE + (N * m + b)), acc[a * N + b]);
                }
            }
        }
    }
    const long long p = stencil_pos(gc, c[0], c[1], c[2]);
    #pragma unroll
    for (int i = 0; i < NN; ++i) Sc[(long long)(s * NN + i) * gc.numNodes + p] = acc[i];
}

void launch_coarsen_from_moduli(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *E, const double *cK0, double *Sc) {
    ProfScope ps(ctx, PC_COARSEN, (double)gc.numNodes);
    dim3 block(128), grid((unsigned)((gc.numNodes + 127) / 128), gc.N == 3 ? 27 : 9);
    if (gc.N == 3) k_coarsen_from_moduli<3><<<grid, block, 0, ctx.stream>>>(gc, gf, E, cK0, Sc);
    else           k_coarsen_from_moduli<2><<<grid, block, 0, ctx.stream>>>(gc, gf, E, cK0, Sc);
    VF_KERNEL_CHECK();
}

// A^c_delta(n) = sum_{a, a' in {-1,0,1}^N} w(a) w(a') A^f_{eps}(2n + a),  eps = 2 delta + a' - a in {-1,0,1}^N,
// w(a) = prod_d (1 - |a_d| / 2)   (= P^T A^f P with the multilinear P of MultigridSolver.hh:130-176).
template<int N>
__global__ void __launch_bounds__(128)
k_coarsen_stencil(const __grid_constant__ GridDesc gc, const __grid_constant__ GridDesc gf,
                  const double *__restrict__ Sf, double *__restrict__ Sc) {
    constexpr int A0 = Dims<N>::A0, NN = N * N, NS = Dims<N>::NS;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= gc.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % gc.nn[2]); r /= gc.nn[2]; c[1] = (int)(r % gc.nn[1]); c[0] = (int)(r / gc.nn[1]); }
    int d[3] = {0, 0, 0};
    { int r = s; for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
    double acc[NN];
    #pragma unroll
    for (int i = 0; i < NN; ++i) acc[i] = 0.0;
    bool nbValid = true;
    for (int a = A0; a < 3; ++a) { const int q = c[a] + d[a]; nbValid = nbValid && q >= 0 && q < gc.nn[a]; }
    if (nbValid) {
        for (int sa = 0; sa < NS; ++sa) {          // a: fine node i = 2n + a
            int av[3] = {0, 0, 0};
            { int r = sa; for (int a = 2; a >= A0; --a) { av[a] = r % 3 - 1; r /= 3; } }
            bool ok = true; double wa = 1.0; int fq[3] = {0, 0, 0};
            for (int a = A0; a < 3; ++a) {
                const int q = 2 * c[a] + av[a];
                ok = ok && q >= 0 && q < gf.nn[a];
                fq[a] = q;
                wa *= av[a] == 0 ? 1.0 : 0.5;
            }
            if (!ok) continue;
            const long long fi = stencil_pos(gf, fq[0], fq[1], fq[2]);
            for (int sb = 0; sb < NS; ++sb) {      // a': fine node j = 2(n + delta) + a'
                int bv[3] = {0, 0, 0};
                { int r = sb; for (int a = 2; a >= A0; --a) { bv[a] = r % 3 - 1; r /= 3; } }
                bool ok2 = true; int se = 0; double w = wa;
                for (int a = A0; a < 3; ++a) {
                    const int eps = 2 * d[a] + bv[a] - av[a];
                    ok2 = ok2 && eps >= -1 && eps <= 1;
                    const int qj = 2 * (c[a] + d[a]) + bv[a];
                    ok2 = ok2 && qj >= 0 && qj < gf.nn[a];
                    se = se * 3 + (eps + 1);
                    w *= bv[a] == 0 ? 1.0 : 0.5;
                }
                if (!ok2) continue;
                #pragma unroll
                for (int i = 0; i < NN; ++i) acc[i] = fma(w, __ldg(Sf + (long long)(se * NN + i) * gf.numNodes + fi), acc[i]);
            }
        }
    }
    const long long p = stencil_pos(gc, c[0], c[1], c[2]);
    #pragma unroll
    for (int i = 0; i < NN; ++i) Sc[(long long)(s * NN + i) * gc.numNodes + p] = acc[i];
}

void launch_coarsen_stencil(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *Sf, double *Sc) {
    ProfScope ps(ctx, PC_COARSEN, (double)gc.numNodes);
    dim3 block(128), grid((unsigned)((gc.numNodes + 127) / 128), gc.N == 3 ? 27 : 9);
    if (gc.N == 3) k_coarsen_stencil<3><<<grid, block, 0, ctx.stream>>>(gc, gf, Sf, Sc);
    else           k_coarsen_stencil<2><<<grid, block, 0, ctx.stream>>>(gc, gf, Sf, Sc);
    VF_KERNEL_CHECK();
}

// Level-0 stencil S = sum_e E_e * K0 blocks (assembled K in stencil form; used by the single-level
// direct solve TPS::solve, TensorProductSimulator.hh:1198-1230).
template<int N>
__global__ void __launch_bounds__(128)
k_stencil_from_moduli_l0(const __grid_constant__ GridDesc g, const double *__restrict__ E, const double *__restrict__ K0, double *__restrict__ S) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE, A0 = Dims<N>::A0, NN = N * N;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= g.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % g.nn[2]); r /= g.nn[2]; c[1] = (int)(r % g.nn[1]); c[0] = (int)(r / g.nn[1]); }
    int d[3] = {0, 0, 0};
    { int r = s; for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
    double acc[NN];
    #pragma unroll
    for (int i = 0; i < NN; ++i) acc[i] = 0.0;
    for (int e = 0; e < NPE; ++e) {
        bool ok = true; int m = 0; long long ei = 0;
        for (int a = A0; a < 3; ++a) {
            const int ob = (e >> (2 - a)) & 1;
            const int ec = c[a] - ob;
            ok = ok && ec >= 0 && ec < g.ne[a];
            ei += (long long)ec * g.es[a];
            const int mb = d[a] + ob;
            ok = ok && (mb == 0 || mb == 1);
            m |= (mb & 1) << (2 - a);
        }
        if (!ok) continue;
        const double Ee = __ldg(E + ei);
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            #pragma unroll
            for (int b = 0; b < N; ++b) acc[a * N + b] = fma(Ee, __ldg(K0 + (N * e + a) * KE + (N * m + b)), acc[a * N + b]);
        }
    }
    const long long p = stencil_pos(g, c[0], c[1], c[2]);
    #pragma unroll
    for (int i = 0; i < NN; ++i) S[(long long)(s * NN + i) * g.numNodes + p] = acc[i];
}

void launch_stencil_from_moduli_l0(const LaunchCtx &ctx, const GridDesc &g, const double *E, const double *K0dev, double *S) {
    ProfScope ps(ctx, PC_COARSEN, (double)g.numNodes);
    dim3 block(128), grid((unsigned)((g.numNodes + 127) / 128), g.N == 3 ? 27 : 9);
    if (g.N == 3) k_stencil_from_moduli_l0<3><<<grid, block, 0, ctx.stream>>>(g, E, K0dev, S);
    else          k_stencil_from_moduli_l0<2><<<grid, block, 0, ctx.stream>>>(g, E, K0dev, S);
    VF_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------
// Coarsest-level dense system (replaces m_assembleStiffnessMatrix + rowColRemoval + CHOLMOD,
// TensorProductSimulator.hh:834-865, 1198-1230)
// ---------------------------------------------------------------------------
__global__ void k_stencil_to_dense(const __grid_constant__ GridDesc g, const double *__restrict__ S, const int *__restrict__ red, int nfree, double *__restrict__ A) {
    const int N = g.N, NN = N * N, NS = (N == 3) ? 27 : 9, A0 = 3 - N;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= g.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % g.nn[2]); r /= g.nn[2]; c[1] = (int)(r % g.nn[1]); c[0] = (int)(r / g.nn[1]); }
    int d[3] = {0, 0, 0};
    { int r = s; for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
    long long m = 0;
    for (int a = 0; a < 3; ++a) { const int q = c[a] + d[a]; if (q < 0 || q >= g.nn[a]) return; m += (long long)q * g.ns[a]; }
    (void)NS;
    for (int a = 0; a < N; ++a) {
        const int ri = red[n * N + a]; if (ri < 0) continue;
        for (int b = 0; b < N; ++b) {
            const int rj = red[m * N + b]; if (rj < 0) continue;
            A[(size_t)ri * nfree + rj] = S[(long long)(s * NN + a * N + b) * g.numNodes + stencil_pos(g, c[0], c[1], c[2])];
        }
    }
}
void launch_stencil_to_dense(const LaunchCtx &ctx, const GridDesc &g, const double *S, const int *redIdx, int nfree, double *A) {
    ProfScope ps(ctx, PC_COARSEN, (double)g.numNodes);
    dim3 block(128), grid((unsigned)((g.numNodes + 127) / 128), g.N == 3 ? 27 : 9);
    k_stencil_to_dense<<<grid, block, 0, ctx.stream>>>(g, S, redIdx, nfree, A);
    VF_KERNEL_CHECK();
}

__global__ void k_symmetrize_lower(double *A, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < n && j < n && j > i) A[(size_t)i * n + j] = A[(size_t)j * n + i];
}
void launch_symmetrize_lower(const LaunchCtx &ctx, double *A, int n) {
    ProfScope ps(ctx, PC_COARSEN, (double)n * n);
    dim3 block(32, 8), grid((n + 31) / 32, (n + 7) / 8);
    k_symmetrize_lower<<<grid, block, 0, ctx.stream>>>(A, n);
    VF_KERNEL_CHECK();
}

// y = A x for a dense symmetric row-major A: one warp per row, coalesced along the row.
__global__ void __launch_bounds__(256) k_dense_symv(const double *__restrict__ A, int n, const double *__restrict__ x, double *__restrict__ y) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double *row = A + (size_t)warp * n;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(__ldg(row + j), x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[warp] = s;
}
void launch_dense_symv(const LaunchCtx &ctx, const double *A, int n, const double *x, double *y) {
    ProfScope ps(ctx, PC_COARSE_SOLVE, (double)n * n);
    if (n == 0) return;
    dim3 block(256), grid((unsigned)(((size_t)n * 32 + 255) / 256));
    k_dense_symv<<<grid, block, 0, ctx.stream>>>(A, n, x, y);
    VF_KERNEL_CHECK();
}

__global__ void k_gather_free(const double *__restrict__ f, const int *__restrict__ freeDofs, int nfree, long long numNodes, int N, double *__restrict__ rhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfree) return;
    const int dof = freeDofs[i];
    rhs[i] = f[(long long)(dof % N) * numNodes + dof / N];
}
void launch_gather_free(const LaunchCtx &ctx, const double *f, const int *freeDofs, int nfree, long long numNodes, int N, double *rhs) {
    ProfScope ps(ctx, PC_COARSE_SOLVE, (double)nfree);
    if (nfree == 0) return;
    k_gather_free<<<(nfree + 255) / 256, 256, 0, ctx.stream>>>(f, freeDofs, nfree, numNodes, N, rhs);
    VF_KERNEL_CHECK();
}
__global__ void k_scatter_free(const double *__restrict__ y, const int *__restrict__ freeDofs, int nfree, long long numNodes, int N, double *__restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfree) return;
    const int dof = freeDofs[i];
    x[(long long)(dof % N) * numNodes + dof / N] = y[i];
}
void launch_scatter_free(const LaunchCtx &ctx, const double *y, const int *freeDofs, int nfree, long long numNodes, int N, double *x) {
    ProfScope ps(ctx, PC_COARSE_SOLVE, (double)nfree);
    if (nfree == 0) return;
    k_scatter_free<<<(nfree + 255) / 256, 256, 0, ctx.stream>>>(y, freeDofs, nfree, numNodes, N, x);
    VF_KERNEL_CHECK();
}

} // namespace vf
