// vf_l0.cu -- matrix-free level-0 operator kernels (sm_100a).
//
// Replaces SpecializedTPSStencils<Real,1,1[,1]>::applyK (TPSStencils.hh:231-396, 431-728),
// computeResidual (MultigridSolver.hh:527-541) and the level-0 multicoloured block
// Gauss-Seidel pass (NodeSmoothStencilFinest + m_smoothNode, MultigridSolver.hh:277-292,
// 347-378, 408-442).
//
// Formulation: one thread per node, lanes along the fastest grid axis (coalesced SoA
// loads).  The 3^N neighbour displacements are loaded once each and scattered into 2^N
// per-incident-element partial products t_e = K0[ln_e rows, :] u_e, which are then
// combined with the element moduli: (K u)_n = sum_e E_e t_e.  K0 is a __grid_constant__
// kernel parameter, so every K0 entry is a constant-bank operand of a DFMA.
#include "vf_internal.cuh"
#include "vf_reduce.cuh"
#include <cstdlib>

namespace vf {

template<int N>
struct NodeCtx {
    int c[3];          // embedded coordinates
    long long n;       // flat node index
    bool lo[3], hi[3]; // neighbour at -1 / +1 exists along each embedded axis
};

template<int N>
__device__ __forceinline__ void make_ctx(const GridDesc &g, int c0, int c1, int c2, NodeCtx<N> &x) {
    x.c[0] = c0; x.c[1] = c1; x.c[2] = c2;
    x.n = (long long)c0 * g.ns[0] + (long long)c1 * g.ns[1] + c2;
    #pragma unroll
    for (int a = 0; a < 3; ++a) { x.lo[a] = x.c[a] >= 1; x.hi[a] = x.c[a] + 1 < g.nn[a]; }
}

// Gather the per-element partial products t[e][c] and moduli Ee[e] around a node.
// Element e (bit (2-a) of e set <=> the element lies at offset -1 along embedded axis a) has the
// node as its local node ln == e (TPSStencils.hh:47-70).
template<int N>
__device__ __forceinline__ void gather_l0(const GridDesc &g, const K0Param &K, const double *__restrict__ u,
                                          const double *__restrict__ E, const NodeCtx<N> &x,
                                          double (&t)[1 << N][N], double (&Ee)[1 << N], double (&uself)[N]) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE, NS = Dims<N>::NS, A0 = Dims<N>::A0;
    const long long pe = (long long)x.c[0] * g.es[0] + (long long)x.c[1] * g.es[1] + x.c[2]; // "primary" element (may be out of range)
    #pragma unroll
    for (int e = 0; e < NPE; ++e) {
        bool valid = true; long long off = 0;
        #pragma unroll
        for (int a = A0; a < 3; ++a) {
            const int ob = (e >> (2 - a)) & 1;
            valid = valid && (ob ? x.lo[a] : x.hi[a]);
            off += ob ? g.es[a] : 0;
        }
        Ee[e] = valid ? __ldg(E + (pe - off)) : 0.0;
        #pragma unroll
        for (int c = 0; c < N; ++c) t[e][c] = 0.0;
    }
    #pragma unroll
    for (int s = 0; s < NS; ++s) {
        int d[3] = {0, 0, 0};
        {   // decode slot -> offsets in {-1,0,1} over the active axes (row-major, axis A0 slowest)
            int r = s;
            #pragma unroll
            for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; }
        }
        bool valid = true; long long off = 0;
        #pragma unroll
        for (int a = A0; a < 3; ++a) {
            valid = valid && (d[a] == 0 || (d[a] < 0 ? x.lo[a] : x.hi[a]));
            off += (long long)d[a] * g.ns[a];
        }
        double un[N];
        #pragma unroll
        for (int c = 0; c < N; ++c) un[c] = valid ? u[c * g.numNodes + x.n + off] : 0.0;
        if (s == NS / 2) {
            #pragma unroll
            for (int c = 0; c < N; ++c) uself[c] = un[c];
        }
        #pragma unroll
        for (int e = 0; e < NPE; ++e) {
            // local index m of the neighbour inside element e: bit_a = d_a + ob_a must be 0 or 1
            bool inElem = true; int m = 0;
            #pragma unroll
            for (int a = A0; a < 3; ++a) {
                const int ob = (e >> (2 - a)) & 1;
                const int mb = d[a] + ob;
                inElem = inElem && (mb == 0 || mb == 1);
                m |= (mb & 1) << (2 - a);
            }
            if (inElem) {
                #pragma unroll
                for (int c = 0; c < N; ++c) {
                    #pragma unroll
                    for (int dc = 0; dc < N; ++dc) t[e][c] = fma(K.v[(N * e + c) * KE + (N * m + dc)], un[dc], t[e][c]);
                }
            }
        }
    }
}

template<int N, int MODE, bool DOT>
__global__ void __launch_bounds__(256)
k_apply_l0(const __grid_constant__ GridDesc g, const __grid_constant__ K0Param K, const double *__restrict__ u,
           const double *__restrict__ E, const double *__restrict__ b, const uint8_t *__restrict__ dmask,
           double *__restrict__ out, double *dotOut, double *scratch) {
    constexpr int NPE = Dims<N>::NPE;
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int c1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = blockIdx.z * blockDim.z + threadIdx.z;
    double dotv = 0.0;
    if (c2 < g.nn[2] && c1 < g.nn[1] && c0 < g.nn[0]) {
        NodeCtx<N> x; make_ctx<N>(g, c0, c1, c2, x);
        const bool detached = ((g.bd == 1) ? c1 : c2) >= g.nActive;
        if (detached) {
            // applyK<ZeroInit = true> zero-fills the detached margin (TPSStencils.hh:385-395, 717-727)
            if (MODE == APPLY_SET) {
                #pragma unroll
                for (int c = 0; c < N; ++c) out[c * g.numNodes + x.n] = 0.0;
            }
        } else {
            double t[NPE][N], Ee[NPE], uself[N];
            gather_l0<N>(g, K, u, E, x, t, Ee, uself);
            const unsigned dm = dmask ? dmask[x.n] : 0u;
            #pragma unroll
            for (int c = 0; c < N; ++c) {
                double acc = 0.0;
                #pragma unroll
                for (int e = 0; e < NPE; ++e) acc = fma(Ee[e], t[e][c], acc);
                double res;
                if (MODE == APPLY_SET) res = acc;
                else if (MODE == APPLY_ADD) res = out[c * g.numNodes + x.n] + acc;
                else if (MODE == APPLY_SUB) res = out[c * g.numNodes + x.n] - acc;
                else res = b[c * g.numNodes + x.n] - acc;
                if ((dm >> c) & 1u) res = 0.0;
                out[c * g.numNodes + x.n] = res;
                if (DOT && c0 >= g.ownLo && c0 < g.ownHi) dotv = fma(uself[c], res, dotv);
            }
        }
    }
    if (DOT) grid_sum(dotv, scratch, dotOut);
}

// ---------------------------------------------------------------------------
// 3D apply, two nodes per thread.
//
// Every K0 entry is used exactly once per node (8 incident elements x 3 rows x 24 columns = 24 x 24), so with one
// node per thread each DFMA needs its own constant fetch (LDCU) and the kernel is issue-bound at ~40% of the FP64
// pipe.  Here a thread owns the two nodes (x, y, z), (x, y+1, z) -- lanes along the fastest axis -- so every constant
// feeds two DFMAs and the 4 x 3 neighbour rows of an x-plane are loaded once for both nodes.  There are no
// boundary branches: neighbour addresses are clamped into the grid and the moduli of elements outside the grid are
// zero, so whatever a clamped load returns is multiplied by zero (an element inside the grid has all 8 nodes inside).
// ---------------------------------------------------------------------------
template<int MODE, bool DOT>
__global__ void __launch_bounds__(128, 3)
k_apply3_l0(const __grid_constant__ GridDesc g, const __grid_constant__ K0Param K, const double *__restrict__ u,
            const double *__restrict__ E, const double *__restrict__ b, const uint8_t *__restrict__ dmask,
            double *__restrict__ out, double *dotOut, double *scratch) {
    const int c2 = blockIdx.x * 32 + threadIdx.x;
    const int c1 = 2 * (blockIdx.y * 4 + threadIdx.y);
    const int c0 = blockIdx.z;
    double dotv = 0.0;
    const bool inGrid = c2 < g.nn[2] && c1 < g.nn[1];
    if (inGrid) {
        const int nx = g.nn[0], ny = g.nn[1], nz = g.nn[2];
        // clamped neighbour offsets
        long long xo[3], yo[4]; int zo[3];
        #pragma unroll
        for (int i = 0; i < 3; ++i) { xo[i] = (long long)min(max(c0 + i - 1, 0), nx - 1) * g.ns[0]; zo[i] = min(max(c2 + i - 1, 0), nz - 1); }
        #pragma unroll
        for (int r = 0; r < 4; ++r) yo[r] = (long long)min(max(c1 + r - 1, 0), ny - 1) * g.ns[1];
        // moduli of the 2 x 3 x 2 elements around the node pair (zero outside the grid)
        double Ee[2][3][2];
        #pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            #pragma unroll
            for (int iy = 0; iy < 3; ++iy) {
                #pragma unroll
                for (int iz = 0; iz < 2; ++iz) {
                    const int ex = c0 - 1 + ix, ey = c1 - 1 + iy, ez = c2 - 1 + iz;
                    const bool ok = ex >= 0 && ex < g.ne[0] && ey >= 0 && ey < g.ne[1] && ez >= 0 && ez < g.ne[2];
                    Ee[ix][iy][iz] = ok ? __ldg(E + ((long long)ex * g.es[0] + (long long)ey * g.es[1] + ez)) : 0.0;
                }
            }
        }
        double t[2][8][3];
        #pragma unroll
        for (int j = 0; j < 2; ++j) {
            #pragma unroll
            for (int e = 0; e < 8; ++e) {
                #pragma unroll
                for (int c = 0; c < 3; ++c) t[j][e][c] = 0.0;
            }
        }
        double uself[2][3];
        #pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            #pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                double uu[4][3];
                #pragma unroll
                for (int r = 0; r < 4; ++r) {
                    #pragma unroll
                    for (int dc = 0; dc < 3; ++dc) uu[r][dc] = u[dc * g.numNodes + xo[dx] + yo[r] + zo[dz]];
                }
                if (dx == 1 && dz == 1) {
                    #pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        #pragma unroll
                        for (int dc = 0; dc < 3; ++dc) uself[j][dc] = uu[j + 1][dc];
                    }
                }
                #pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    #pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // local index m of neighbour (dx-1, dy-1, dz-1) in incident element e (bit (2-a) of e: element at offset -1 along axis a)
                        const int m0 = (dx - 1) + ((e >> 2) & 1), m1 = (dy - 1) + ((e >> 1) & 1), m2 = (dz - 1) + (e & 1);
                        if (m0 < 0 || m0 > 1 || m1 < 0 || m1 > 1 || m2 < 0 || m2 > 1) continue;
                        const int m = (m0 << 2) | (m1 << 1) | m2;
                        #pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            #pragma unroll
                            for (int dc = 0; dc < 3; ++dc) {
                                const double k = K.v[(3 * e + c) * 24 + (3 * m + dc)];
                                t[0][e][c] = fma(k, uu[dy][dc], t[0][e][c]);
                                t[1][e][c] = fma(k, uu[dy + 1][dc], t[1][e][c]);
                            }
                        }
                    }
                }
            }
        }
        #pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int y = c1 + j;
            if (y >= ny) break;
            const long long n = (long long)c0 * g.ns[0] + (long long)y * g.ns[1] + c2;
            if (y >= g.nActive) { // detached layer: applyK<ZeroInit = true> zero-fills it (TPSStencils.hh:717-727)
                if (MODE == APPLY_SET) {
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) out[c * g.numNodes + n] = 0.0;
                }
                continue;
            }
            const unsigned dm = dmask ? dmask[n] : 0u;
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                double acc = 0.0;
                #pragma unroll
                for (int e = 0; e < 8; ++e) acc = fma(Ee[1 - ((e >> 2) & 1)][j + 1 - ((e >> 1) & 1)][1 - (e & 1)], t[j][e][c], acc);
                double res;
                if (MODE == APPLY_SET) res = acc;
                else if (MODE == APPLY_ADD) res = out[c * g.numNodes + n] + acc;
                else if (MODE == APPLY_SUB) res = out[c * g.numNodes + n] - acc;
                else res = b[c * g.numNodes + n] - acc;
                if ((dm >> c) & 1u) res = 0.0;
                out[c * g.numNodes + n] = res;
                if (DOT && c0 >= g.ownLo && c0 < g.ownHi) dotv = fma(uself[j][c], res, dotv);
            }
        }
    }
    if (DOT) grid_sum(dotv, scratch, dotOut);
}

static void apply3_l0_dispatch(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                               const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    dim3 block(32, 4, 1);
    dim3 grid((g.nn[2] + 31) / 32, ((g.nn[1] + 1) / 2 + 3) / 4, g.nn[0]);
    // the fused u . out reduction keeps one partial per block; beyond the scratch capacity fall back to a separate dot
    const bool fused = dotOut && (size_t)grid.x * grid.y * grid.z <= (size_t)kReduceMaxBlocks;
#define VF_APPLY_CASE(M) \
    if (mode == M) { \
        if (fused) k_apply3_l0<M, true><<<grid, block, 0, ctx.stream>>>(g, K, u, E, b, dmask, out, dotOut, scratch); \
        else       k_apply3_l0<M, false><<<grid, block, 0, ctx.stream>>>(g, K, u, E, b, dmask, out, nullptr, nullptr); \
    }
    VF_APPLY_CASE(APPLY_SET) VF_APPLY_CASE(APPLY_ADD) VF_APPLY_CASE(APPLY_SUB) VF_APPLY_CASE(APPLY_RESIDUAL)
#undef VF_APPLY_CASE
    VF_KERNEL_CHECK();
    if (dotOut && !fused) launch_masked_dot(ctx, g, u, out, dotOut, scratch);
}

template<int N>
static void apply_l0_dispatch(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                              const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    dim3 block = (N == 3) ? dim3(32, 4, 2) : dim3(32, 8, 1);
    dim3 grid((g.nn[2] + block.x - 1) / block.x, (g.nn[1] + block.y - 1) / block.y, (g.nn[0] + block.z - 1) / block.z);
    const bool fused = dotOut && (size_t)grid.x * grid.y * grid.z <= (size_t)kReduceMaxBlocks;
#define VF_APPLY_CASE(M) \
    if (mode == M) { \
        if (fused) k_apply_l0<N, M, true><<<grid, block, 0, ctx.stream>>>(g, K, u, E, b, dmask, out, dotOut, scratch); \
        else       k_apply_l0<N, M, false><<<grid, block, 0, ctx.stream>>>(g, K, u, E, b, dmask, out, nullptr, nullptr); \
    }
    VF_APPLY_CASE(APPLY_SET) VF_APPLY_CASE(APPLY_ADD) VF_APPLY_CASE(APPLY_SUB) VF_APPLY_CASE(APPLY_RESIDUAL)
#undef VF_APPLY_CASE
    VF_KERNEL_CHECK();
    if (dotOut && !fused) launch_masked_dot(ctx, g, u, out, dotOut, scratch);
}

void launch_apply_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                     const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    ProfScope ps(ctx, mode == APPLY_RESIDUAL ? PC_RESIDUAL_L0 : PC_APPLY_L0, (double)g.numNodes);
    if (g.N == 3) apply3_l0_dispatch(ctx, g, K, u, E, b, dmask, out, mode, dotOut, scratch);
    else          apply_l0_dispatch<2>(ctx, g, K, u, E, b, dmask, out, mode, dotOut, scratch);
}

// ---------------------------------------------------------------------------
// Multicoloured block Gauss-Seidel, level 0
// ---------------------------------------------------------------------------


template<int N>
__global__ void __launch_bounds__(256)
k_gs_l0(const __grid_constant__ GridDesc g, const __grid_constant__ K0Param K, const __grid_constant__ ColorDesc col,
        double *__restrict__ u, const double *__restrict__ b, const double *__restrict__ E,
        const uint8_t *__restrict__ dmask, int forward) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE;
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int i0 = blockIdx.z * blockDim.z + threadIdx.z;
    if (i2 >= col.cnt[2] || i1 >= col.cnt[1] || i0 >= col.cnt[0]) return;
    NodeCtx<N> x; make_ctx<N>(g, col.off[0] + 2 * i0, col.off[1] + 2 * i1, col.off[2] + 2 * i2, x);
    if (x.c[0] < g.cmpLo || x.c[0] >= g.cmpHi) return;   // ghost planes of a slab window are received, not computed
    const unsigned dm = dmask[x.n];
    if (dm == (unsigned)((1 << N) - 1)) return; // hasFullDirichlet (:350)
    double t[NPE][N], Ee[NPE], uself[N];
    gather_l0<N>(g, K, u, E, x, t, Ee, uself);
    double rhs[N], M[N][N];
    #pragma unroll
    for (int c = 0; c < N; ++c) {
        double acc = 0.0;
        #pragma unroll
        for (int e = 0; e < NPE; ++e) acc = fma(Ee[e], t[e][c], acc);
        rhs[c] = b[c * g.numNodes + x.n] - acc;
        #pragma unroll
        for (int c2 = 0; c2 < N; ++c2) {
            double m = 0.0;
            #pragma unroll
            for (int e = 0; e < NPE; ++e) m = fma(Ee[e], K.v[(N * e + c) * KE + (N * e + c2)], m);
            M[c][c2] = m;
        }
    }
    double du[N];
    gs_node_update<N>(M, rhs, dm, forward != 0, du);
    #pragma unroll
    for (int c = 0; c < N; ++c) u[c * g.numNodes + x.n] = uself[c] + du[c];
}

// ---------------------------------------------------------------------------
// 3D single-colour Gauss-Seidel pass, two same-colour nodes (y, y+2) per thread: every K0 constant feeds two DFMAs and
// there are no boundary branches (clamped addresses, zero moduli outside the grid), as in k_apply3_l0.
// ---------------------------------------------------------------------------
template<bool FWD>
__global__ void __launch_bounds__(128, 3)
k_gs3_color(const __grid_constant__ GridDesc g, const __grid_constant__ K0Param K, const __grid_constant__ ColorDesc col,
            double *u, const double *__restrict__ b, const double *__restrict__ E, const uint8_t *__restrict__ dmask) {
    const int i2 = blockIdx.x * 32 + threadIdx.x;
    const int i1 = 2 * (blockIdx.y * 4 + threadIdx.y);   // first of the two colour-local y indices
    const int i0 = blockIdx.z;
    if (i2 >= col.cnt[2] || i1 >= col.cnt[1]) return;
    const int c0 = col.off[0] + 2 * i0, c1 = col.off[1] + 2 * i1, c2 = col.off[2] + 2 * i2;
    if (c0 < g.cmpLo || c0 >= g.cmpHi) return;           // ghost planes of a slab window are received, not computed
    const int nx = g.nn[0], ny = g.nn[1], nz = g.nn[2];
    const long long NN = g.numNodes;
    long long xo[3], yo[5]; int zo[3];
    #pragma unroll
    for (int i = 0; i < 3; ++i) { xo[i] = (long long)min(max(c0 + i - 1, 0), nx - 1) * g.ns[0]; zo[i] = min(max(c2 + i - 1, 0), nz - 1); }
    #pragma unroll
    for (int r = 0; r < 5; ++r) yo[r] = (long long)min(max(c1 + r - 1, 0), ny - 1) * g.ns[1];
    double t[2][8][3];
    #pragma unroll
    for (int j = 0; j < 2; ++j) {
        #pragma unroll
        for (int e = 0; e < 8; ++e) {
            #pragma unroll
            for (int c = 0; c < 3; ++c) t[j][e][c] = 0.0;
        }
    }
    double uself[2][3];
    #pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
        #pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
            double uu[5][3];
            #pragma unroll
            for (int r = 0; r < 5; ++r) {
                #pragma unroll
                for (int dc = 0; dc < 3; ++dc) uu[r][dc] = u[dc * NN + xo[dx] + yo[r] + zo[dz]];
            }
            if (dx == 1 && dz == 1) {
                #pragma unroll
                for (int j = 0; j < 2; ++j) {
                    #pragma unroll
                    for (int dc = 0; dc < 3; ++dc) uself[j][dc] = uu[2 * j + 1][dc];
                }
            }
            #pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                #pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int m0 = (dx - 1) + ((e >> 2) & 1), m1 = (dy - 1) + ((e >> 1) & 1), m2 = (dz - 1) + (e & 1);
                    if (m0 < 0 || m0 > 1 || m1 < 0 || m1 > 1 || m2 < 0 || m2 > 1) continue;
                    const int m = (m0 << 2) | (m1 << 1) | m2;
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        #pragma unroll
                        for (int dc = 0; dc < 3; ++dc) {
                            const double k = K.v[(3 * e + c) * 24 + (3 * m + dc)];
                            t[0][e][c] = fma(k, uu[dy][dc], t[0][e][c]);
                            t[1][e][c] = fma(k, uu[dy + 2][dc], t[1][e][c]);
                        }
                    }
                }
            }
        }
    }
    #pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (i1 + j >= col.cnt[1]) break;
        const int y = c1 + 2 * j;
        const long long n = (long long)c0 * g.ns[0] + (long long)y * g.ns[1] + c2;
        const unsigned dm = dmask[n];
        if (dm == 7u) continue; // hasFullDirichlet (MultigridSolver.hh:350)
        double Ee[8];
        #pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ex = c0 - ((e >> 2) & 1), ey = y - ((e >> 1) & 1), ez = c2 - (e & 1);
            const bool ok = ex >= 0 && ex < g.ne[0] && ey >= 0 && ey < g.ne[1] && ez >= 0 && ez < g.ne[2];
            Ee[e] = ok ? __ldg(E + ((long long)ex * g.es[0] + (long long)ey * g.es[1] + ez)) : 0.0;
        }
        double rhs[3], M[3][3], du[3];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            #pragma unroll
            for (int e = 0; e < 8; ++e) acc = fma(Ee[e], t[j][e][c], acc);
            rhs[c] = b[c * NN + n] - acc;
            #pragma unroll
            for (int c2_ = 0; c2_ < 3; ++c2_) {
                double mm = 0.0;
                #pragma unroll
                for (int e = 0; e < 8; ++e) mm = fma(Ee[e], K.v[(3 * e + c) * 24 + (3 * e + c2_)], mm);
                M[c][c2_] = mm;
            }
        }
        gs_node_update<3>(M, rhs, dm, FWD, du);
        #pragma unroll
        for (int c = 0; c < 3; ++c) u[c * NN + n] = uself[j][c] + du[c];
    }
}

void launch_gs3_color_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                         const uint8_t *dmask, int color, bool forward) {
    ColorDesc col;
    if (!make_color(g, color, col)) return;
    ProfScope ps(ctx, PC_GS_L0, (double)col.cnt[0] * col.cnt[1] * col.cnt[2]);
    dim3 block(32, 4, 1), grid((col.cnt[2] + 31) / 32, ((col.cnt[1] + 1) / 2 + 3) / 4, col.cnt[0]);
    if (forward) k_gs3_color<true><<<grid, block, 0, ctx.stream>>>(g, K, col, u, b, E, dmask);
    else         k_gs3_color<false><<<grid, block, 0, ctx.stream>>>(g, K, col, u, b, E, dmask);
    VF_KERNEL_CHECK();
}

void launch_gs_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                  const uint8_t *dmask, int color, bool forward) {
    ColorDesc col;
    if (!make_color(g, color, col)) return;
    ProfScope ps(ctx, PC_GS_L0, (double)col.cnt[0] * col.cnt[1] * col.cnt[2]);
    dim3 block = (g.N == 3) ? dim3(32, 4, 2) : dim3(32, 8, 1);
    dim3 grid((col.cnt[2] + block.x - 1) / block.x, (col.cnt[1] + block.y - 1) / block.y, (col.cnt[0] + block.z - 1) / block.z);
    if (g.N == 3) k_gs_l0<3><<<grid, block, 0, ctx.stream>>>(g, K, col, u, b, E, dmask, forward ? 1 : 0);
    else          k_gs_l0<2><<<grid, block, 0, ctx.stream>>>(g, K, col, u, b, E, dmask, forward ? 1 : 0);
    VF_KERNEL_CHECK();
}

} // namespace vf
