// vf_stencil.cu -- coarse-level operators stored as 3^N-point block stencils (sm_100a).
//
// The reference keeps levels >= 2 as a block-CSC matrix ("blockK", TensorProductSimulator.hh:
// 885-966; applied by CSCMatrix::applyTransposeParallel, MeshFEM SparseMatrices.hh:1613-1677,
// smoothed with NodeSmoothStencilBlockK, MultigridSolver.hh:323-334), rebuilds the level-1 operator
// on the fly from the fine moduli (MultigridSolver.hh:294-321, 475-499) and caches per-element
// matrices at the coarsest level (:815-817).  On a structured grid all three are the same object: a
// 3^N-point stencil of N x N blocks per node.  Here every level >= 1 stores that stencil in a
// tiled layout  S[tile][slot*N*N + a*N + b][lane]  (stencil_addr) over the colour-major node numbering of GridDesc
// (stencil_pos): the rows of 16 consecutive same-colour nodes are one contiguous block, fetched by one TMA bulk copy.
//
// Galerkin coarsening (MultigridSolver.hh:711-819) is done directly on stencils: level 1 from the
// fine moduli and the 2^N matrices coarsenedFineK0s[fi]; level l >= 2 as P^T A_{l-1} P.  Both are
// algebraically identical to the reference's per-element recursion (sum over coarse elements of
// Phi^T Ke Phi), differing only in floating-point summation order.
#include "vf_internal.cuh"
#include "vf_reduce.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace vf {

// Tile kernel for the stored-stencil levels.  A thread block handles one tile of kStencilTile consecutive positions
// (same colour, colour-major order) with one thread row per stencil slot.  Thread 0 stages the tile's contiguous
// NE x 16 block of stencil entries into shared memory with a single TMA bulk copy (cp.async.bulk + mbarrier) while all
// threads resolve their node coordinates and gather the neighbour displacements; thread (tx, s) then multiplies the
// N x N block of slot s with u(neighbour s) and the 3^N slot contributions are reduced in shared memory in a fixed
// order (deterministic).  The level-1 stencil (1944 B per node in 3D) therefore streams from HBM as contiguous 31 KB
// blocks; the small coarse levels cost one memory round trip per colour pass.
template<int N, int SPT>
struct TileShared {
    alignas(128) double S[Dims<N>::NE * kStencilTile];
    double red[N][Dims<N>::NS / SPT][kStencilTile];   // k_stencil_tile uses the first (NS / SPT + 1) / 2 rows (pairs of thread rows are pre-added by shuffle)
    double bs[N][kStencilTile];
    double us[N][kStencilTile];
    double G[N * N][kStencilTile];   // Gauss-Seidel passes: the node's update matrix (k_stencil_tile)
    unsigned dm[kStencilTile];
    alignas(8) unsigned long long mbar;
};

// coordinates of position pos (colour-major numbering); false for the padding lanes of a colour's last tile
template<int N>
__device__ __forceinline__ bool pos_coords(const GridDesc &g, long long pos, int (&c)[3]) {
    int cc = 0;
    #pragma unroll
    for (int k = 1; k < 8; ++k) cc += (pos >= g.cbase[k]) ? 1 : 0;   // cbase is non-decreasing
    const unsigned idx = (unsigned)(pos - g.cbase[cc]);
    const unsigned n2 = (unsigned)g.ccnt[cc][2], n1 = (unsigned)g.ccnt[cc][1], n0 = (unsigned)g.ccnt[cc][0];
    if (idx >= n0 * n1 * n2) return false;
    const unsigned i2 = idx % n2, r = idx / n2, i1 = r % n1, i0 = r / n1;
    c[0] = 2 * (int)i0 + ((cc >> 2) & 1); c[1] = 2 * (int)i1 + ((cc >> 1) & 1); c[2] = 2 * (int)i2 + (cc & 1);
    return true;
}

// SPT = stencil slots per thread: 1 (one thread row per slot, 432 threads, 4 blocks per SM in 3D) or 3 (one thread row per
// (dx, dy) pair handling its three z-neighbours, 144 threads, 6 blocks per SM: more tiles in flight per SM).
//
// RES (Gauss-Seidel passes only): the sweep also leaves the residual  r = b - K u  of its FINAL iterate in rout, so that the V-cycle
// needs no separate residual kernel (which would stream the whole stencil a second time: 4.2 GB at level 1 of a 256^3 grid).  Right
// after its update a node's residual is  rhs - M du  (zero for a block solve, non-zero after the point sweep of a partially
// constrained node); every LATER update du_j of a neighbour j changes it by  -K_ij du_j = -(K_ji)^T du_j  (the Galerkin operator is
// symmetric), and K_ji is a slot of the row of j that the pass updating j holds in shared memory anyway.  So the thread of slot
// (j -> i) adds  -(S_slot)^T du_j  to r_i for the neighbours i of colours visited EARLIER in this sweep (red.global.add.f64; the
// summation order of those <= 26 contributions is not fixed, results agree with the direct residual to rounding).  Dirichlet
// components of r are zeroed by the caller afterwards.  Only on fully attached grids (the host checks); in a slab window the shared
// planes next to the ghost planes miss the contributions of the neighbouring part's updates and are recomputed by the caller
// (launch_residual_stencil_plane).
//
// Instruction economy (ncu: the tile kernels issue ~300 instructions per warp for 9 DFMA per thread and sit at 60 % issue-slot
// utilisation with 32 registers per thread): on the levels that do not stream from HBM node coordinates come from the level's
// position table (posTab, one 8-byte load instead of the divisions of pos_coords; on the streaming level the dependent
// load costs more than the divisions), all node indices are 32-bit (3 * numNodes < 2^31 on every stored-stencil level, checked by
// the host), the slot contributions of a warp's two thread rows are added by one shuffle before the shared-memory reduction (14
// partial sums per component instead of 27), and a Gauss-Seidel pass has no serial 3 x 3 solve after the reduction (see G below).
// Same-box A/B of these pieces: profiles/r06f_ab.log.
using sidx = int;   // node indices of the stored-stencil levels (check_index_range)
template<int N, bool GS, int MODE, int SPT, bool RES = false>
__global__ void __launch_bounds__(kStencilTile * Dims<N>::NS / SPT, (N == 3 ? 4 : 8) * (SPT == 3 ? 3 : 2) / 2)
k_stencil_tile(const __grid_constant__ GridDesc g, long long tile0, const double *__restrict__ S,
               const double *uin, const double *__restrict__ b, const uint8_t *__restrict__ dmask,
               double *out, int flags, double *rout, const unsigned long long *__restrict__ posTab, int planeSel) {
    // planeSel >= 0 (apply / residual modes): only the nodes of local node plane planeSel are computed and written
    static_assert(!RES || (GS && SPT == 1), "the residual-emitting variant is a Gauss-Seidel pass with one slot per thread");
    // flags bit 0: forward sweep; bit 1: the previous kernel on the stream does not write S (a colour pass of the same sweep),
    // so the stencil tile may be requested BEFORE waiting for it -- the HBM round trip of this kernel's first wave then
    // overlaps the tail of the previous colour pass.
    pdl_trigger();
    const int forward = flags & 1;
    const bool earlyTile = (flags & 2) != 0;
    constexpr int NS = Dims<N>::NS, A0 = Dims<N>::A0, NN = N * N, NE = Dims<N>::NE, ROWS = NS / SPT, RP = (ROWS + 1) / 2;
    __shared__ TileShared<N, SPT> sh;
    const int tx = threadIdx.x, s = threadIdx.y;   // s: thread row, slots s * SPT .. s * SPT + SPT - 1
    const long long tile = tile0 + blockIdx.x;
    // flags bit 4 (off by default, see stencil_stream_hint): thread 0 requests the stencil tile first of all, before any coordinate
    // work.  (Padding tiles are zero-filled, so the request is always in bounds; a block without active nodes drains the copy before
    // it exits; grids with tiles that are skipped as a whole -- ghost planes of a slab window, detached layers -- never do this.)
    const bool requestFirst = (flags & 16) && !(GS && (g.cmpLo > 0 || g.cmpHi < g.nn[0])) && g.nActive >= g.nn[g.bd];
    if (tx == 0 && s == 0) {
        mbar_init(&sh.mbar, 1);
        if (requestFirst) {
            if (!earlyTile) pdl_wait();
            if (flags & 4) tma_load_1d_stream(sh.S, S + tile * (long long)(NE * kStencilTile), NE * kStencilTile * sizeof(double), &sh.mbar);
            else           tma_load_1d(sh.S, S + tile * (long long)(NE * kStencilTile), NE * kStencilTile * sizeof(double), &sh.mbar);
        }
    }
    int c[3] = {0, 0, 0};
    bool inRange;
    if (posTab) {
        const unsigned long long e = __ldg(posTab + tile * kStencilTile + tx);
        inRange = (e >> 63) == 0ull;
        c[2] = (int)(e & 0xffffull); c[1] = (int)((e >> 16) & 0xffffull); c[0] = (int)((e >> 32) & 0xffffull);
    } else inRange = pos_coords<N>(g, tile * kStencilTile + tx, c);
    const sidx nnodes = (sidx)g.numNodes, ns0 = (sidx)g.ns[0], ns1 = (sidx)g.ns[1];
    const sidx n = c[0] * ns0 + c[1] * ns1 + c[2];
    const bool detached = inRange && (((g.bd == 1) ? c[1] : c[2]) >= g.nActive);
    const bool active = inRange && !detached && !(GS && (c[0] < g.cmpLo || c[0] >= g.cmpHi)) // ghost planes of a slab window are not smoothed
                        && !(!GS && planeSel >= 0 && c[0] != planeSel);
    const int anyActive = __syncthreads_or(active ? 1 : 0);     // also publishes the barrier initialisation
    if (!requestFirst) {
        if (!earlyTile) pdl_wait();
        if (anyActive && tx == 0 && s == 0) {
            if (flags & 4) tma_load_1d_stream(sh.S, S + tile * (long long)(NE * kStencilTile), NE * kStencilTile * sizeof(double), &sh.mbar);
            else           tma_load_1d(sh.S, S + tile * (long long)(NE * kStencilTile), NE * kStencilTile * sizeof(double), &sh.mbar);
        }
    }
    pdl_wait();
    sidx pushTo = -1;   // RES: node index of this slot's neighbour if it receives this pass's contribution, else -1
    if (anyActive) {
        double acc[N], un[SPT][N];
        #pragma unroll
        for (int a = 0; a < N; ++a) acc[a] = 0.0;
        bool valid[SPT];
        #pragma unroll
        for (int j = 0; j < SPT; ++j) {
            valid[j] = false;
            #pragma unroll
            for (int k = 0; k < N; ++k) un[j][k] = 0.0;
        }
        if (active) {
            // stage b and the Dirichlet mask alongside so that the finalising threads have no global loads left
            if (s < N && (GS || MODE == APPLY_RESIDUAL)) sh.bs[s][tx] = b[s * nnodes + n];
            if (s == N) sh.dm[tx] = dmask ? dmask[n] : 0u;
            #pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const int slot = s * SPT + j;
                int d[3] = {0, 0, 0};
                { int r = slot;
                  #pragma unroll
                  for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
                bool v = true;
                #pragma unroll
                for (int a = A0; a < 3; ++a) { const int qq = c[a] + d[a]; v = v && qq >= 0 && qq < g.nn[a]; }
                const sidx nb = n + d[0] * ns0 + d[1] * ns1 + d[2];
                valid[j] = v;
                if (v) {
                    if (flags & 8) {
                        const unsigned long long keep = l2_policy_evict_last();
                        #pragma unroll
                        for (int k = 0; k < N; ++k) un[j][k] = ld_l2_hint(uin + (k * nnodes + nb), keep);
                    } else {
                        #pragma unroll
                        for (int k = 0; k < N; ++k) un[j][k] = uin[k * nnodes + nb];
                    }
                    if (GS && slot == NS / 2) {
                        #pragma unroll
                        for (int k = 0; k < N; ++k) sh.us[k][tx] = un[j][k];
                    }
                    if (RES) {   // neighbours whose colour was visited earlier in this sweep receive -(S_slot)^T du
                        const int cc = ((((c[0] + g.xoff) & 1) << 2) | ((c[1] & 1) << 1) | (c[2] & 1));   // colour = parity class of the GLOBAL index
                        const int m = ((d[0] != 0) << 2) | ((d[1] != 0) << 1) | (d[2] != 0);
                        const bool visited = forward ? ((cc ^ m) < cc) : ((cc ^ m) > cc);
                        if (visited) pushTo = nb;
                    }
                }
            }
        }
        mbar_wait(&sh.mbar, 0);
        #pragma unroll
        for (int j = 0; j < SPT; ++j) {
            if (valid[j]) {
                const int slot = s * SPT + j;
                #pragma unroll
                for (int a = 0; a < N; ++a) {
                    #pragma unroll
                    for (int k = 0; k < N; ++k) acc[a] = fma(sh.S[(slot * NN + a * N + k) * kStencilTile + tx], un[j][k], acc[a]);
                }
            }
        }
        // thread rows 2w and 2w + 1 are the two halves of warp w (the last row has no partner): add them by shuffle
        if (s != ROWS - 1) {
            #pragma unroll
            for (int a = 0; a < N; ++a) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], 16);
        }
        if (!(s & 1)) {
            #pragma unroll
            for (int a = 0; a < N; ++a) sh.red[a][s >> 1][tx] = acc[a];
        }

    }
    __syncthreads();
    if (!anyActive) {
        if (!GS && MODE == APPLY_SET && s == 0 && inRange) {   // applyK<ZeroInit> zero-fills the detached margin
            #pragma unroll
            for (int a = 0; a < N; ++a) out[a * nnodes + n] = 0.0;
        }
        if (requestFirst && tx == 0 && s == 0) mbar_wait(&sh.mbar, 0);   // the requested tile must have landed before the block's shared memory is released
        return;
    }
    if (s < N) { // thread row a = s sums the partial sums of component a (fixed order -> deterministic)
        double t = 0.0;
        #pragma unroll
        for (int k = 0; k < RP; ++k) t += sh.red[s][k][tx];
        sh.red[s][0][tx] = t;
    }
    if (GS && s == ROWS - 1) {
        // Meanwhile the last thread row (idle during the reduction) prepares the node's update matrix G:  du = G (b - K u).
        // G = M^-1 for a free node (cofactor inverse, as Eigen's fixed-size inverse(), MultigridSolver.hh:370), the linear map of the
        // point Gauss-Seidel sweep over the free components for a partially constrained node (:358-365), 0 for skipped nodes (:350).
        // The serial tail of the block after the reduction is then 9 multiply-adds instead of a 3 x 3 solve with its reciprocal.
        double G[N][N];
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            #pragma unroll
            for (int k = 0; k < N; ++k) G[a][k] = 0.0;
        }
        const unsigned dm = active ? sh.dm[tx] : (unsigned)((1 << N) - 1);
        if (dm != (unsigned)((1 << N) - 1)) {
            double M[N][N];
            #pragma unroll
            for (int a = 0; a < N; ++a) {
                #pragma unroll
                for (int k = 0; k < N; ++k) M[a][k] = sh.S[((NS / 2) * NN + a * N + k) * kStencilTile + tx];
            }
            if (dm == 0u) block_inverse<N>(M, G);
            else {
                #pragma unroll
                for (int k = 0; k < N; ++k) {
                    double e[N], col[N];
                    #pragma unroll
                    for (int a = 0; a < N; ++a) e[a] = (a == k) ? 1.0 : 0.0;
                    #pragma unroll
                    for (int a = 0; a < N; ++a) col[a] = 0.0;
                    if (forward) gs_point_sweep<N, true>(M, e, dm, col);
                    else         gs_point_sweep<N, false>(M, e, dm, col);
                    #pragma unroll
                    for (int a = 0; a < N; ++a) G[a][k] = col[a];
                }
            }
        }
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            #pragma unroll
            for (int k = 0; k < N; ++k) sh.G[a * N + k][tx] = G[a][k];
        }
    }
    __syncthreads();
    if (GS) {
        // every thread that needs du forms it itself from G and the reduced sums: no further barrier
        if (!(s == 0 ? active : (RES && pushTo >= 0))) return;
        double rhs[N], du[N];
        #pragma unroll
        for (int a = 0; a < N; ++a) rhs[a] = sh.bs[a][tx] - sh.red[a][0][tx];
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            double t = 0.0;
            #pragma unroll
            for (int k = 0; k < N; ++k) t = fma(sh.G[a * N + k][tx], rhs[k], t);
            du[a] = t;
        }
        if (s == 0) {
            const unsigned dm = sh.dm[tx];
            if (dm != (unsigned)((1 << N) - 1)) {   // hasFullDirichlet nodes are skipped (:350)
                #pragma unroll
                for (int a = 0; a < N; ++a) out[a * nnodes + n] = sh.us[a][tx] + du[a];
            }
            if (RES) {   // the node's own residual after its update: the first write of r_n in this sweep, later passes add to it
                #pragma unroll
                for (int a = 0; a < N; ++a) {
                    double t = rhs[a];
                    #pragma unroll
                    for (int k = 0; k < N; ++k) t = fma(-sh.S[((NS / 2) * NN + a * N + k) * kStencilTile + tx], du[k], t);
                    rout[a * nnodes + n] = ((dm >> a) & 1u) ? 0.0 : t;
                }
            }
        }
        if (RES && pushTo >= 0) {
            #pragma unroll
            for (int k = 0; k < N; ++k) {
                double t = 0.0;
                #pragma unroll
                for (int a = 0; a < N; ++a) t = fma(sh.S[(s * NN + a * N + k) * kStencilTile + tx], du[a], t);
                atomicAdd(rout + (k * nnodes + pushTo), -t);
            }
        }
        return;
    }
    if (s != 0 || !inRange) return;
    if (planeSel >= 0 && !active) return;
    if (detached) {
        if (!GS && MODE == APPLY_SET) {
            #pragma unroll
            for (int a = 0; a < N; ++a) out[a * nnodes + n] = 0.0;
        }
        return;
    }
    const unsigned dm = sh.dm[tx];
    {
        #pragma unroll
        for (int a = 0; a < N; ++a) {
            const double av = sh.red[a][0][tx];
            double res;
            if (MODE == APPLY_SET) res = av;
            else if (MODE == APPLY_ADD) res = out[a * nnodes + n] + av;
            else if (MODE == APPLY_SUB) res = out[a * nnodes + n] - av;
            else res = sh.bs[a][tx] - av;
            if ((dm >> a) & 1u) res = 0.0;
            out[a * nnodes + n] = res;
        }
    }
}

// Position table of a stored-stencil level: entry pos = c2 | c1 << 16 | c0 << 32 of the node at position pos of the colour-major
// numbering, bit 63 set for the padding positions of a colour's last tile.
template<int N>
__global__ void __launch_bounds__(256) k_fill_pos_table(const __grid_constant__ GridDesc g, unsigned long long *tab) {
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.numPos) return;
    int c[3] = {0, 0, 0};
    const bool ok = pos_coords<N>(g, pos, c);
    tab[pos] = ok ? ((unsigned long long)c[2] | ((unsigned long long)c[1] << 16) | ((unsigned long long)c[0] << 32)) : (1ull << 63);
}
void launch_fill_pos_table(cudaStream_t stream, const GridDesc &g, unsigned long long *tab) {
    if (g.nn[0] > 65535 || g.nn[1] > 65535 || g.nn[2] > 65535) throw std::runtime_error("stored-stencil level too large for the position table");
    const unsigned blocks = (unsigned)((g.numPos + 255) / 256);
    if (g.N == 3) k_fill_pos_table<3><<<blocks, 256, 0, stream>>>(g, tab);
    else          k_fill_pos_table<2><<<blocks, 256, 0, stream>>>(g, tab);
    VF_KERNEL_CHECK();
}
static void check_index_range(const GridDesc &g) {
    if (3.0 * (double)g.numNodes >= 2147483648.0) throw std::runtime_error("stored-stencil level exceeds the 32-bit node index range of the tile kernels");
}

// flags bit 2: fetch the stencil tile with an L2 evict-first hint (levels whose stencil streams from HBM; VF_ST_EVICT_FIRST=0 disables)
static int stencil_stream_hint(const GridDesc &g) {
    static const bool on = [] { const char *e = std::getenv("VF_ST_EVICT_FIRST"); return !(e && e[0] == '0'); }();
    const double bytes = (double)g.numNodes * (g.N == 3 ? 1944.0 : 288.0);
    static const bool keepU = [] { const char *e = std::getenv("VF_ST_KEEP_U"); return e && e[0] == '1'; }();
    // flags bit 4: request the tile before any coordinate work.  Measured and rejected (profiles/r06e_time_ops_early*.log: level-1 apply
    // 0.69 -> 0.77 ms, sweep 0.93 -> 0.96 ms); VF_ST_EARLY_TMA=1 enables it.
    static const bool early = [] { const char *e = std::getenv("VF_ST_EARLY_TMA"); return e && e[0] == '1'; }();
    return ((on && bytes > 64.0 * 1048576.0) ? (keepU ? 12 : 4) : 0) | (early ? 16 : 0);   // larger than half the L2: the stencil cannot stay resident between passes anyway
}
static int stencil_slots_per_thread() {
    static const int v = [] { const char *e = std::getenv("VF_ST_SPT"); return (e && std::atoi(e) == 3) ? 3 : 1; }();
    return v;
}
// the tile kernels want the largest shared-memory carve-out (4 resident blocks x 42 KB in 3D)
template<typename Kern> static void prefer_shared(Kern k) {
    VF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
}
static void stencil_kernel_attributes() {
    static PerDeviceFlags done;
    if (!first_use_on_device(done)) return;
#define VF_ATTR(NN_, M) prefer_shared(k_stencil_tile<NN_, false, M, 1>); prefer_shared(k_stencil_tile<NN_, false, M, 3>);
    VF_ATTR(3, APPLY_SET) VF_ATTR(3, APPLY_ADD) VF_ATTR(3, APPLY_SUB) VF_ATTR(3, APPLY_RESIDUAL)
    VF_ATTR(2, APPLY_SET) VF_ATTR(2, APPLY_ADD) VF_ATTR(2, APPLY_SUB) VF_ATTR(2, APPLY_RESIDUAL)
#undef VF_ATTR
    prefer_shared(k_stencil_tile<3, true, APPLY_SET, 1>); prefer_shared(k_stencil_tile<2, true, APPLY_SET, 1>);
    prefer_shared(k_stencil_tile<3, true, APPLY_SET, 3>); prefer_shared(k_stencil_tile<2, true, APPLY_SET, 3>);
    prefer_shared(k_stencil_tile<3, true, APPLY_SET, 1, true>); prefer_shared(k_stencil_tile<2, true, APPLY_SET, 1, true>);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// One smoothing sweep of a SMALL stored-stencil level in one launch.  A colour pass of a level that fits the L2 (level >= 3 of the
// 256^3 hierarchy: <= 36k nodes) is a 2-3 us kernel behind ~8 us of launch / dependency latency, and a sweep is 2^N of them
// (13 % of the solve in round 1).  Here a persistent grid (every block resident) runs the 2^N colour passes back to back with a
// grid-wide barrier in between (sense-reversing counter in global memory; release / acquire at gpu scope) and loops over the tiles
// of a colour.  Displacements are read with ld.global.cg: other blocks wrote them one colour earlier in the same launch, and L1 is
// not coherent.  Same tile arithmetic, same visiting order (MultigridSolver.hh:408-442) as k_stencil_tile<GS>.
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_gpu_u32(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// bar[0]: arrivals of the current generation, bar[1]: generation.  All blocks of the grid must be resident.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned gen = ld_acquire_gpu_u32(bar + 1);
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) { bar[0] = 0u; __threadfence(); st_release_gpu_u32(bar + 1, gen + 1); }
        else while (ld_acquire_gpu_u32(bar + 1) == gen) { }
    }
    __syncthreads();
}

template<int N>
__global__ void __launch_bounds__(kStencilTile * Dims<N>::NS, N == 3 ? 4 : 8)
k_stencil_sweep(const __grid_constant__ GridDesc g, const double *__restrict__ S, double *u, const double *__restrict__ b,
                const uint8_t *__restrict__ dmask, int forward, int xparity, unsigned *bar) {
    pdl_prologue();
    constexpr int NS = Dims<N>::NS, A0 = Dims<N>::A0, NN = N * N, NE = Dims<N>::NE, NC = 1 << N;
    __shared__ TileShared<N, 1> sh;
    const int tx = threadIdx.x, s = threadIdx.y;
    if (tx == 0 && s == 0) mbar_init(&sh.mbar, 1);
    __syncthreads();
    unsigned phase = 0;
    int d[3] = {0, 0, 0};
    { int r = s;
      #pragma unroll
      for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
    #pragma unroll 1
    for (int ci = 0; ci < NC; ++ci) {
        const int gcol = forward ? ci : NC - 1 - ci;
        const int color = N == 3 ? (gcol ^ (xparity << 2)) : gcol;     // local parity class of this global colour (slab windows)
        const long long tile0 = g.cbase[color] / kStencilTile;
        const long long tot = (long long)g.ccnt[color][0] * g.ccnt[color][1] * g.ccnt[color][2];
        const int ntile = (int)((tot + kStencilTile - 1) / kStencilTile);
        #pragma unroll 1
        for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
            const long long tile = tile0 + t;
            const long long pos = tile * kStencilTile + tx;
            int c[3] = {0, 0, 0};
            const bool inRange = pos_coords<N>(g, pos, c);
            const long long n = (long long)c[0] * g.ns[0] + (long long)c[1] * g.ns[1] + c[2];
            const bool detached = inRange && (((g.bd == 1) ? c[1] : c[2]) >= g.nActive);
            const bool active = inRange && !detached && !(c[0] < g.cmpLo || c[0] >= g.cmpHi);
            const int anyActive = __syncthreads_or(active ? 1 : 0);   // also: everybody is done with the previous tile's shared data
            if (!anyActive) continue;
            if (tx == 0 && s == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the previous tile was read through the generic proxy
                tma_load_1d(sh.S, S + tile * (long long)(NE * kStencilTile), NE * kStencilTile * sizeof(double), &sh.mbar);
            }
            double acc[N], un[N];
            #pragma unroll
            for (int a = 0; a < N; ++a) { acc[a] = 0.0; un[a] = 0.0; }
            bool valid = false;
            if (active) {
                if (s < N) sh.bs[s][tx] = b[s * g.numNodes + n];
                if (s == N) sh.dm[tx] = dmask ? dmask[n] : 0u;
                bool v = true; long long off = 0;
                #pragma unroll
                for (int a = A0; a < 3; ++a) { const int qq = c[a] + d[a]; v = v && qq >= 0 && qq < g.nn[a]; off += (long long)d[a] * g.ns[a]; }
                valid = v;
                if (v) {
                    #pragma unroll
                    for (int k = 0; k < N; ++k) un[k] = __ldcg(u + k * g.numNodes + n + off);
                    if (s == NS / 2) {
                        #pragma unroll
                        for (int k = 0; k < N; ++k) sh.us[k][tx] = un[k];
                    }
                }
            }
            mbar_wait(&sh.mbar, phase); phase ^= 1u;
            if (valid) {
                #pragma unroll
                for (int a = 0; a < N; ++a) {
                    #pragma unroll
                    for (int k = 0; k < N; ++k) acc[a] = fma(sh.S[(s * NN + a * N + k) * kStencilTile + tx], un[k], acc[a]);
                }
            }
            #pragma unroll
            for (int a = 0; a < N; ++a) sh.red[a][s][tx] = acc[a];
            __syncthreads();
            if (s < N) {   // thread row a = s sums the slot contributions of component a (fixed order -> deterministic)
                double tsum = 0.0;
                #pragma unroll
                for (int k = 0; k < NS; ++k) tsum += sh.red[s][k][tx];
                sh.red[s][0][tx] = tsum;
            }
            __syncthreads();
            if (s == 0 && active) {
                const unsigned dm = sh.dm[tx];
                if (dm != (unsigned)((1 << N) - 1)) {      // hasFullDirichlet nodes are skipped (MultigridSolver.hh:350)
                    double rhs[N], M[N][N], du[N];
                    #pragma unroll
                    for (int a = 0; a < N; ++a) {
                        rhs[a] = sh.bs[a][tx] - sh.red[a][0][tx];
                        #pragma unroll
                        for (int k = 0; k < N; ++k) M[a][k] = sh.S[((NS / 2) * NN + a * N + k) * kStencilTile + tx];
                    }
                    gs_node_update<N>(M, rhs, dm, forward != 0, du);
                    #pragma unroll
                    for (int a = 0; a < N; ++a) u[a * g.numNodes + n] = sh.us[a][tx] + du[a];
                }
            }
        }
        if (ci + 1 < NC) grid_barrier(bar, gridDim.x);
    }
}

// whether launch_gs_stencil_sweep() handles this level.  OFF by default: measured on B200 (profiles/r04f_bench_fused_sweep.log) the
// persistent sweep LOSES to the 2^N programmatic-dependent-launch chained colour passes replayed from the CUDA graph -- 256^3
// solve 175.4 ms (chained) vs 177.9 ms (levels <= 5k nodes fused) vs 180.7 ms (levels <= 70k nodes fused): inside a graph a chained
// colour pass costs ~4 us, less than a grid barrier plus an exposed TMA round trip.  VF_SWEEP_FUSED_NODES=<n> enables it for levels
// of up to n nodes (the parity tests run with it on and off).
bool stencil_sweep_fused(const GridDesc &g) {
    static const long long limit = [] { const char *e = std::getenv("VF_SWEEP_FUSED_NODES"); return e ? std::atoll(e) : 0LL; }();
    return g.numNodes <= limit;
}
void launch_gs_stencil_sweep(const LaunchCtx &ctx, const GridDesc &g, const double *S, double *u, const double *b, const uint8_t *dmask,
                             bool forward, int xparity, unsigned *bar) {
    stencil_kernel_attributes();
    static PerDeviceFlags once; static int maxBlocks[2] = {0, 0};
    if (first_use_on_device(once)) {
        int dev = 0, sms = 0, occ3 = 0, occ2 = 0;
        VF_CUDA(cudaGetDevice(&dev)); VF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        prefer_shared(k_stencil_sweep<3>); prefer_shared(k_stencil_sweep<2>);
        VF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, k_stencil_sweep<3>, kStencilTile * 27, 0));
        VF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_stencil_sweep<2>, kStencilTile * 9, 0));
        maxBlocks[1] = sms * std::max(occ3, 1); maxBlocks[0] = sms * std::max(occ2, 1);
    }
    long long ntileMax = 1, nodes = 0;
    for (int c = 0; c < (1 << g.N); ++c) {
        const long long tot = (long long)g.ccnt[c][0] * g.ccnt[c][1] * g.ccnt[c][2];
        ntileMax = std::max(ntileMax, (tot + kStencilTile - 1) / kStencilTile); nodes += tot;
    }
    ProfScope ps(ctx, PC_GS_ST_SMALL, (double)nodes);
    // every block must be resident (grid barrier): never more blocks than fit, and not the whole machine for a handful of tiles
    const int blocks = (int)std::min<long long>(ntileMax, maxBlocks[g.N == 3 ? 1 : 0]);
    dim3 block(kStencilTile, g.N == 3 ? 27 : 9), grid((unsigned)blocks);
    if (g.N == 3) VF_LAUNCH((k_stencil_sweep<3>), grid, block, 0, ctx.stream, g, S, u, b, dmask, forward ? 1 : 0, xparity, bar);
    else          VF_LAUNCH((k_stencil_sweep<2>), grid, block, 0, ctx.stream, g, S, u, b, dmask, forward ? 1 : 0, xparity, bar);
    VF_KERNEL_CHECK();
}

void launch_apply_stencil(const LaunchCtx &ctx, const GridDesc &g, const double *S, const double *u, const double *b,
                          const uint8_t *dmask, double *out, int mode, const unsigned long long *posTab) {
    stencil_kernel_attributes(); check_index_range(g);
    const bool big = stencil_level_streams(g);
    ProfScope ps(ctx, mode == APPLY_RESIDUAL ? (big ? PC_RESIDUAL_ST : PC_RESIDUAL_ST_SMALL) : (big ? PC_APPLY_ST : PC_APPLY_ST_SMALL), (double)g.numNodes);
    const int spt = stencil_slots_per_thread();
    const int hint = stencil_stream_hint(g);
    dim3 block(kStencilTile, (g.N == 3 ? 27 : 9) / spt), grid((unsigned)(g.numPos / kStencilTile));
#define VF_CASE(NN_, M) if (g.N == NN_ && mode == M) { \
        if (spt == 3) VF_LAUNCH((k_stencil_tile<NN_, false, M, 3>), grid, block, 0, ctx.stream, g, 0, S, u, b, dmask, out, 1 | hint, nullptr, posTab, -1); \
        else          VF_LAUNCH((k_stencil_tile<NN_, false, M, 1>), grid, block, 0, ctx.stream, g, 0, S, u, b, dmask, out, 1 | hint, nullptr, posTab, -1); }
    VF_CASE(3, APPLY_SET) VF_CASE(3, APPLY_ADD) VF_CASE(3, APPLY_SUB) VF_CASE(3, APPLY_RESIDUAL)
    VF_CASE(2, APPLY_SET) VF_CASE(2, APPLY_ADD) VF_CASE(2, APPLY_SUB) VF_CASE(2, APPLY_RESIDUAL)
#undef VF_CASE
    VF_KERNEL_CHECK();
}

// VF_GS_RESIDUAL=0 keeps the separate residual kernel; VF_GS_RESIDUAL_MIN_NODES=<n> fuses only on levels of at least n nodes
bool gs_residual_fusable(const GridDesc &g) {
    static const bool on = [] { const char *e = std::getenv("VF_GS_RESIDUAL"); return !(e && e[0] == '0'); }();
    static const long long minNodes = [] { const char *e = std::getenv("VF_GS_RESIDUAL_MIN_NODES"); return e ? std::atoll(e) : 0LL; }();
    return on && g.nActive >= g.nn[g.bd] && g.numNodes >= minNodes && !stencil_sweep_fused(g);
}
// out = b - K u (Dirichlet components zeroed) on the nodes of local node plane `plane` only: one launch per colour of that x parity
// over the tiles that hold the plane's nodes.  Completes the residual of a slab window's shared planes after a residual-emitting
// sweep (their neighbours in the ghost planes are updated by the neighbouring part, so the pushed contributions are incomplete).
void launch_residual_stencil_plane(const LaunchCtx &ctx, const GridDesc &g, const double *S, const double *u, const double *b,
                                   const uint8_t *dmask, double *out, int plane, const unsigned long long *posTab) {
    stencil_kernel_attributes(); check_index_range(g);
    const int hint = stencil_stream_hint(g);
    const int nc = 1 << g.N;
    for (int c = 0; c < nc; ++c) {
        if (g.N == 3 && ((c >> 2) & 1) != (plane & 1)) continue;
        const long long perPlane = (long long)g.ccnt[c][1] * g.ccnt[c][2];
        if (perPlane == 0 || (plane >> 1) >= g.ccnt[c][0]) continue;
        const long long first = g.cbase[c] + (long long)(plane >> 1) * perPlane, last = first + perPlane - 1;
        const long long t0 = first / kStencilTile, t1 = last / kStencilTile;
        ProfScope ps(ctx, stencil_level_streams(g) ? PC_RESIDUAL_ST : PC_RESIDUAL_ST_SMALL, (double)perPlane);
        dim3 block(kStencilTile, g.N == 3 ? 27 : 9), grid((unsigned)(t1 - t0 + 1));
        if (g.N == 3) VF_LAUNCH((k_stencil_tile<3, false, APPLY_RESIDUAL, 1>), grid, block, 0, ctx.stream, g, t0, S, u, b, dmask, out, 1 | hint, nullptr, posTab, plane);
        else          VF_LAUNCH((k_stencil_tile<2, false, APPLY_RESIDUAL, 1>), grid, block, 0, ctx.stream, g, t0, S, u, b, dmask, out, 1 | hint, nullptr, posTab, plane);
        VF_KERNEL_CHECK();
    }
}
void launch_gs_stencil(const LaunchCtx &ctx, const GridDesc &g, const double *S, double *u, const double *b,
                       const uint8_t *dmask, int color, bool forward, bool chained, double *resOut, const unsigned long long *posTab) {
    ColorDesc col;
    if (!make_color(g, color, col)) return;
    stencil_kernel_attributes(); check_index_range(g);
    const long long tot = (long long)g.ccnt[color][0] * g.ccnt[color][1] * g.ccnt[color][2];
    ProfScope ps(ctx, stencil_level_streams(g) ? PC_GS_ST : PC_GS_ST_SMALL, (double)col.cnt[0] * col.cnt[1] * col.cnt[2]);
    const int spt = stencil_slots_per_thread();
    dim3 block(kStencilTile, (g.N == 3 ? 27 : 9) / spt), grid((unsigned)((tot + kStencilTile - 1) / kStencilTile));
    const long long tile0 = g.cbase[color] / kStencilTile;
    const int fl = (forward ? 1 : 0) | (chained ? 2 : 0) | stencil_stream_hint(g);
    double *const noRes = nullptr;
    if (resOut) {   // residual-emitting pass: one slot per thread
        block = dim3(kStencilTile, g.N == 3 ? 27 : 9);
        if (g.N == 3) VF_LAUNCH_PDL(chained, (k_stencil_tile<3, true, APPLY_SET, 1, true>), grid, block, 0, ctx.stream, g, tile0, S, u, b, dmask, u, fl, resOut, posTab, -1);
        else          VF_LAUNCH_PDL(chained, (k_stencil_tile<2, true, APPLY_SET, 1, true>), grid, block, 0, ctx.stream, g, tile0, S, u, b, dmask, u, fl, resOut, posTab, -1);
    }
    else if (g.N == 3 && spt == 3) VF_LAUNCH_PDL(chained, (k_stencil_tile<3, true, APPLY_SET, 3>), grid, block, 0, ctx.stream, g, tile0, S, u, b, dmask, u, fl, noRes, posTab, -1);
    else if (g.N == 3)        VF_LAUNCH_PDL(chained, (k_stencil_tile<3, true, APPLY_SET, 1>), grid, block, 0, ctx.stream, g, tile0, S, u, b, dmask, u, fl, noRes, posTab, -1);
    else if (spt == 3)        VF_LAUNCH_PDL(chained, (k_stencil_tile<2, true, APPLY_SET, 3>), grid, block, 0, ctx.stream, g, tile0, S, u, b, dmask, u, fl, noRes, posTab, -1);
    else                      VF_LAUNCH_PDL(chained, (k_stencil_tile<2, true, APPLY_SET, 1>), grid, block, 0, ctx.stream, g, tile0, S, u, b, dmask, u, fl, noRes, posTab, -1);
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
                      const double *__restrict__ E, const double *__restrict__ cK0, double *__restrict__ Sc, int bandLo, int bandHi) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE, A0 = Dims<N>::A0, NN = N * N;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= gc.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % gc.nn[2]); r /= gc.nn[2]; c[1] = (int)(r % gc.nn[1]); c[0] = (int)(r / gc.nn[1]); }
    { const int cb = (gc.bd == 1) ? c[1] : c[2]; if (cb < bandLo || cb > bandHi) return; }   // banded update: rows outside keep their values
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
                long long ef = 0; bool owned = true;
                for (int a = A0; a < 3; ++a) {
                    const int q = 2 * ec[a] + ((fi >> (2 - a)) & 1) + (a == 0 ? xshift(gf, gc) : 0);
                    if (a == 0) owned = q >= gf.oeLo && q < gf.oeHi;   // sub-assembly over the fine element layers this part owns
                    ef += (long long)q * gf.es[a];
                }
                if (!owned) continue;
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
    for (int i = 0; i < NN; ++i) Sc[stencil_addr(p, s * NN + i, Dims<N>::NE)] = acc[i];
}

// The same sum with the coarsened full-density blocks as KERNEL PARAMETERS: one launch per stencil slot, and for a given slot the
// blocks K[e][fi] = cK0[fi][rows of local node e, columns of local node m(e, slot)] (<= 4.6 KB) are warp-uniform compile-time
// offsets into the constant bank, so the multiply-adds take them as operands directly.  The first kernel issues one global load per
// multiply-add (the table is fetched through the load/store path although every lane reads the same entry) and is bound by that path:
// 3.35 ms for level 1 of a 256^3 grid, whose 9.9 G multiply-adds are 0.6 ms of FP64 pipe time.  Same (e, fi, entry) summation order
// as above: bit-identical stencils.
template<int N> struct SlotK {
    double k[1 << N][1 << N][N * N];   // [incident coarse element e][child fi][a * N + b]
    int m[1 << N];                     // local node of n + delta in element e, or -1 if the element does not contain it
};
template<int N>
__global__ void __launch_bounds__(128)
k_coarsen_from_moduli_slot(const __grid_constant__ GridDesc gc, const __grid_constant__ GridDesc gf, const __grid_constant__ SlotK<N> K,
                           const double *__restrict__ E, double *__restrict__ Sc, int s, int bandLo, int bandHi) {
    constexpr int NPE = Dims<N>::NPE, A0 = Dims<N>::A0, NN = N * N;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= gc.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % gc.nn[2]); r /= gc.nn[2]; c[1] = (int)(r % gc.nn[1]); c[0] = (int)(r / gc.nn[1]); }
    { const int cb = (gc.bd == 1) ? c[1] : c[2]; if (cb < bandLo || cb > bandHi) return; }   // banded update: rows outside keep their values
    int d[3] = {0, 0, 0};
    { int r = s; for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
    double acc[NN];
    #pragma unroll
    for (int i = 0; i < NN; ++i) acc[i] = 0.0;
    bool nbValid = true;
    for (int a = A0; a < 3; ++a) { const int q = c[a] + d[a]; nbValid = nbValid && q >= 0 && q < gc.nn[a]; }
    if (nbValid) {
        const int xs = xshift(gf, gc);
        #pragma unroll
        for (int e = 0; e < NPE; ++e) { // incident coarse elements; node is local node ln == e
            if (K.m[e] < 0) continue;   // warp-uniform
            bool ok = true; int ec[3] = {0, 0, 0};
            #pragma unroll
            for (int a = A0; a < 3; ++a) {
                ec[a] = c[a] - ((e >> (2 - a)) & 1);
                ok = ok && ec[a] >= 0 && ec[a] < gc.ne[a];
            }
            if (!ok) continue;
            // first child of the element; child fi adds its bits times the element strides (compile-time selection of uniform operands)
            long long base = 0;
            #pragma unroll
            for (int a = A0; a < 3; ++a) base += (long long)(2 * ec[a] + (a == 0 ? xs : 0)) * gf.es[a];
            const int q0 = 2 * ec[0] + xs;
            #pragma unroll
            for (int fi = 0; fi < NPE; ++fi) {
                if (A0 == 0) { const int q = q0 + ((fi >> 2) & 1); if (q < gf.oeLo || q >= gf.oeHi) continue; }   // sub-assembly over the fine element layers this part owns
                long long ef = base;
                #pragma unroll
                for (int a = A0; a < 3; ++a) if ((fi >> (2 - a)) & 1) ef += gf.es[a];
                const double Ef = __ldg(E + ef);
                #pragma unroll
                for (int i = 0; i < NN; ++i) acc[i] = fma(Ef, K.k[e][fi][i], acc[i]);
            }
        }
    }
    const long long p = stencil_pos(gc, c[0], c[1], c[2]);
    #pragma unroll
    for (int i = 0; i < NN; ++i) Sc[stencil_addr(p, s * NN + i, Dims<N>::NE)] = acc[i];
}
template<int N>
static void coarsen_from_moduli_slots(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *E, const double *cK0host, double *Sc, int bandLo, int bandHi) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE, A0 = Dims<N>::A0, NS = Dims<N>::NS;
    const unsigned blocks = (unsigned)((gc.numNodes + 127) / 128);
    for (int s = 0; s < NS; ++s) {
        int d[3] = {0, 0, 0};
        { int r = s; for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
        SlotK<N> K; std::memset(&K, 0, sizeof(K));
        for (int e = 0; e < NPE; ++e) {
            bool ok = true; int m = 0;
            for (int a = A0; a < 3; ++a) {
                const int mb = d[a] + ((e >> (2 - a)) & 1);
                ok = ok && (mb == 0 || mb == 1);
                m |= (mb & 1) << (2 - a);
            }
            K.m[e] = ok ? m : -1;
            if (!ok) continue;
            for (int fi = 0; fi < NPE; ++fi)
                for (int a = 0; a < N; ++a)
                    for (int b = 0; b < N; ++b) K.k[e][fi][a * N + b] = cK0host[(size_t)fi * KE * KE + (size_t)(N * e + a) * KE + (N * m + b)];
        }
        count_launch();
        k_coarsen_from_moduli_slot<N><<<blocks, 128, 0, ctx.stream>>>(gc, gf, K, E, Sc, s, bandLo, bandHi);
        VF_KERNEL_CHECK();
    }
}

// cK0host: the host copy of cK0 (the per-slot parameter form needs it); VF_COARSEN_PARAMK=0 selects the first kernel
void launch_coarsen_from_moduli(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *E, const double *cK0, double *Sc, int bandLo, int bandHi,
                                const double *cK0host) {
    ProfScope ps(ctx, PC_COARSEN, (double)gc.numNodes);
    static const bool paramK = [] { const char *e = std::getenv("VF_COARSEN_PARAMK"); return !(e && e[0] == '0'); }();
    if (cK0host && paramK) {
        if (gc.N == 3) coarsen_from_moduli_slots<3>(ctx, gc, gf, E, cK0host, Sc, bandLo, bandHi);
        else           coarsen_from_moduli_slots<2>(ctx, gc, gf, E, cK0host, Sc, bandLo, bandHi);
        return;
    }
    dim3 block(128), grid((unsigned)((gc.numNodes + 127) / 128), gc.N == 3 ? 27 : 9);
    if (gc.N == 3) k_coarsen_from_moduli<3><<<grid, block, 0, ctx.stream>>>(gc, gf, E, cK0, Sc, bandLo, bandHi);
    else           k_coarsen_from_moduli<2><<<grid, block, 0, ctx.stream>>>(gc, gf, E, cK0, Sc, bandLo, bandHi);
    VF_KERNEL_CHECK();
}

// A^c_delta(n) = sum_{a, a' in {-1,0,1}^N} w(a) w(a') A^f_{eps}(2n + a),  eps = 2 delta + a' - a in {-1,0,1}^N,
// w(a) = prod_d (1 - |a_d| / 2)   (= P^T A^f P with the multilinear P of MultigridSolver.hh:130-176).
template<int N>
__global__ void __launch_bounds__(128)
k_coarsen_stencil(const __grid_constant__ GridDesc gc, const __grid_constant__ GridDesc gf,
                  const double *__restrict__ Sf, double *__restrict__ Sc, int bandLo, int bandHi) {
    constexpr int A0 = Dims<N>::A0, NN = N * N, NS = Dims<N>::NS;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= gc.numNodes) return;
    int c[3]; { long long r = n; c[2] = (int)(r % gc.nn[2]); r /= gc.nn[2]; c[1] = (int)(r % gc.nn[1]); c[0] = (int)(r / gc.nn[1]); }
    { const int cb = (gc.bd == 1) ? c[1] : c[2]; if (cb < bandLo || cb > bandHi) return; }   // banded update: rows outside keep their values
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
                const int q = 2 * c[a] + av[a] + (a == 0 ? xshift(gf, gc) : 0);
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
                    const int qj = 2 * (c[a] + d[a]) + bv[a] + (a == 0 ? xshift(gf, gc) : 0);
                    ok2 = ok2 && qj >= 0 && qj < gf.nn[a];
                    se = se * 3 + (eps + 1);
                    w *= bv[a] == 0 ? 1.0 : 0.5;
                }
                if (!ok2) continue;
                #pragma unroll
                for (int i = 0; i < NN; ++i) acc[i] = fma(w, __ldg(Sf + stencil_addr(fi, se * NN + i, Dims<N>::NE)), acc[i]);
            }
        }
    }
    const long long p = stencil_pos(gc, c[0], c[1], c[2]);
    #pragma unroll
    for (int i = 0; i < NN; ++i) Sc[stencil_addr(p, s * NN + i, Dims<N>::NE)] = acc[i];
}

void launch_coarsen_stencil(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *Sf, double *Sc, int bandLo, int bandHi) {
    ProfScope ps(ctx, PC_COARSEN, (double)gc.numNodes);
    dim3 block(128), grid((unsigned)((gc.numNodes + 127) / 128), gc.N == 3 ? 27 : 9);
    if (gc.N == 3) k_coarsen_stencil<3><<<grid, block, 0, ctx.stream>>>(gc, gf, Sf, Sc, bandLo, bandHi);
    else           k_coarsen_stencil<2><<<grid, block, 0, ctx.stream>>>(gc, gf, Sf, Sc, bandLo, bandHi);
    VF_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------
// Separable Galerkin coarsening (3D, full rebuilds).  The multilinear prolongation is a tensor product P = Px (x) Py (x) Pz,
// so P^T A P collapses one axis at a time:  B_(delta, o)(n') = sum_{alpha, alpha'} w(alpha) w(alpha') A_(eps, o)(2 n'_a + alpha),
// eps = 2 delta + alpha' - alpha, with o the (unchanged) offsets along the other two axes.  Every intermediate operator is again
// a 27-point block stencil on a mixed grid, each pass reads its input once and writes an output of half the size: 11 GB of
// traffic for level 2 of a 256^3 grid where the one-shot kernel above issues 43 GB of cache-resident re-reads (14.7 ms).
// Intermediates are SoA (T[entry][node]); the first pass reads and the last pass writes the colour-tiled layout.
// ---------------------------------------------------------------------------
struct AxisPass { int inNN[3], outNN[3]; int axis; long long inNodes, outNodes; int shift; };   // shift: fine index = 2 * coarse index + al + shift (slab windows, axis 0)

template<bool IN_TILED, bool OUT_TILED>
__global__ void __launch_bounds__(128)
k_coarsen_axis(const __grid_constant__ AxisPass P, const __grid_constant__ GridDesc gIn, const __grid_constant__ GridDesc gOut,
               const double *__restrict__ In, double *__restrict__ Out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.outNodes) return;
    const int q = blockIdx.y, o = q / 9, i = q % 9;     // o: offsets along the two untouched axes, i: entry of the 3 x 3 block
    int c[3]; { long long r = n; c[2] = (int)(r % P.outNN[2]); r /= P.outNN[2]; c[1] = (int)(r % P.outNN[1]); c[0] = (int)(r / P.outNN[1]); }
    const int a = P.axis, b0 = a == 0 ? 1 : 0, b1 = a == 2 ? 1 : 2;   // b0 < b1: the other axes
    int d[3]; d[b0] = o / 3 - 1; d[b1] = o % 3 - 1;
    double acc[3] = {0.0, 0.0, 0.0};
    #pragma unroll
    for (int al = -1; al <= 1; ++al) {
        const int f = 2 * c[a] + al + P.shift;
        if (f < 0 || f >= P.inNN[a]) continue;
        const double wa = al == 0 ? 1.0 : 0.5;
        int fc[3] = {c[0], c[1], c[2]}; fc[a] = f;
        const long long inLin = ((long long)fc[0] * P.inNN[1] + fc[1]) * P.inNN[2] + fc[2];
        const long long inPos = IN_TILED ? stencil_pos(gIn, fc[0], fc[1], fc[2]) : 0;
        #pragma unroll
        for (int ep = -1; ep <= 1; ++ep) {
            d[a] = ep;
            const int slot = ((d[0] + 1) * 3 + (d[1] + 1)) * 3 + (d[2] + 1);
            const double v = IN_TILED ? __ldg(In + stencil_addr(inPos, slot * 9 + i, 243)) : __ldg(In + (long long)(slot * 9 + i) * P.inNodes + inLin);
            #pragma unroll
            for (int de = -1; de <= 1; ++de) {
                const int ap = ep + al - 2 * de;              // alpha': fine neighbour 2 (n' + delta) + alpha' = f + eps
                if (ap < -1 || ap > 1) continue;
                const int cn = c[a] + de;
                if (cn < 0 || cn >= P.outNN[a]) continue;
                acc[de + 1] = fma(wa * (ap == 0 ? 1.0 : 0.5), v, acc[de + 1]);
            }
        }
    }
    const long long outPos = OUT_TILED ? stencil_pos(gOut, c[0], c[1], c[2]) : 0;
    #pragma unroll
    for (int de = -1; de <= 1; ++de) {
        d[a] = de;
        const int slot = ((d[0] + 1) * 3 + (d[1] + 1)) * 3 + (d[2] + 1);
        if (OUT_TILED) Out[stencil_addr(outPos, slot * 9 + i, 243)] = acc[de + 1];
        else           Out[(long long)(slot * 9 + i) * P.outNodes + n] = acc[de + 1];
    }
}

// scratch: at least coarsen_separable_scratch(gc, gf) doubles
size_t coarsen_separable_scratch(const GridDesc &gc, const GridDesc &gf) {
    return (size_t)243 * ((size_t)gc.nn[0] * gf.nn[1] * gf.nn[2] + (size_t)gc.nn[0] * gc.nn[1] * gf.nn[2]);
}
void launch_coarsen_stencil_separable(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *Sf, double *Sc, double *scratch) {
    ProfScope ps(ctx, PC_COARSEN, (double)gc.numNodes);
    AxisPass P[3];
    int cur[3] = {gf.nn[0], gf.nn[1], gf.nn[2]};
    for (int a = 0; a < 3; ++a) {
        P[a].axis = a; P[a].shift = a == 0 ? xshift(gf, gc) : 0;
        for (int k = 0; k < 3; ++k) { P[a].inNN[k] = cur[k]; }
        cur[a] = gc.nn[a];
        for (int k = 0; k < 3; ++k) { P[a].outNN[k] = cur[k]; }
        P[a].inNodes = (long long)P[a].inNN[0] * P[a].inNN[1] * P[a].inNN[2];
        P[a].outNodes = (long long)P[a].outNN[0] * P[a].outNN[1] * P[a].outNN[2];
    }
    double *T1 = scratch, *T2 = scratch + (size_t)243 * P[0].outNodes;
    auto grid = [](const AxisPass &p) { return dim3((unsigned)((p.outNodes + 127) / 128), 81); };
    k_coarsen_axis<true, false><<<grid(P[0]), 128, 0, ctx.stream>>>(P[0], gf, gc, Sf, T1);
    k_coarsen_axis<false, false><<<grid(P[1]), 128, 0, ctx.stream>>>(P[1], gf, gc, T1, T2);
    k_coarsen_axis<false, true><<<grid(P[2]), 128, 0, ctx.stream>>>(P[2], gf, gc, T2, Sc);
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
    for (int i = 0; i < NN; ++i) S[stencil_addr(p, s * NN + i, Dims<N>::NE)] = acc[i];
}

void launch_stencil_from_moduli_l0(const LaunchCtx &ctx, const GridDesc &g, const double *E, const double *K0dev, double *S) {
    ProfScope ps(ctx, PC_COARSEN, (double)g.numNodes);
    dim3 block(128), grid((unsigned)((g.numNodes + 127) / 128), g.N == 3 ? 27 : 9);
    if (g.N == 3) k_stencil_from_moduli_l0<3><<<grid, block, 0, ctx.stream>>>(g, E, K0dev, S);
    else          k_stencil_from_moduli_l0<2><<<grid, block, 0, ctx.stream>>>(g, E, K0dev, S);
    VF_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------
// Slab completion: rows of one node plane <-> contiguous buffer [node in plane (row-major y, z)][entry]
// ---------------------------------------------------------------------------
long long stencil_plane_rows(const GridDesc &g, int plane) { (void)plane; return (long long)g.nn[1] * g.nn[2]; }
template<bool ADD>
__global__ void __launch_bounds__(256) k_stencil_plane(const __grid_constant__ GridDesc g, double *S, int plane, double *buf, int NE) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rows = (long long)g.nn[1] * g.nn[2];
    if (i >= rows * NE) return;
    const int entry = (int)(i % NE); const long long r = i / NE;
    const int c2 = (int)(r % g.nn[2]), c1 = (int)(r / g.nn[2]);
    const long long a = stencil_addr(stencil_pos(g, plane, c1, c2), entry, NE);
    if (ADD) S[a] += buf[i]; else buf[i] = S[a];
}
void launch_stencil_plane_pack(const LaunchCtx &ctx, const GridDesc &g, const double *S, int plane, double *buf) {
    const int NE = (g.N == 3 ? 27 : 9) * g.N * g.N; const long long tot = stencil_plane_rows(g, plane) * NE;
    ProfScope ps(ctx, PC_COARSEN, (double)tot);
    k_stencil_plane<false><<<(unsigned)((tot + 255) / 256), 256, 0, ctx.stream>>>(g, const_cast<double *>(S), plane, buf, NE);
    VF_KERNEL_CHECK();
}
void launch_stencil_plane_add(const LaunchCtx &ctx, const GridDesc &g, double *S, int plane, const double *buf) {
    const int NE = (g.N == 3 ? 27 : 9) * g.N * g.N; const long long tot = stencil_plane_rows(g, plane) * NE;
    ProfScope ps(ctx, PC_COARSEN, (double)tot);
    k_stencil_plane<true><<<(unsigned)((tot + 255) / 256), 256, 0, ctx.stream>>>(g, S, plane, const_cast<double *>(buf), NE);
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
    for (int a = 0; a < N; ++a) {
        const int ri = red[n * N + a]; if (ri < 0) continue;
        for (int b = 0; b < N; ++b) {
            const int rj = red[m * N + b]; if (rj < 0) continue;
            A[(size_t)ri * nfree + rj] = S[stencil_addr(stencil_pos(g, c[0], c[1], c[2]), s * NN + a * N + b, NS * NN)];
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
    pdl_prologue();
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
    VF_LAUNCH((k_dense_symv), grid, block, 0, ctx.stream, A, n, x, y);
    VF_KERNEL_CHECK();
}

// y = tril(A) x or triu(A) x for a row-major matrix: one warp per row over the row's triangular part only
template<bool LOWER>
__global__ void __launch_bounds__(256) k_dense_trmv(const double *__restrict__ A, int n, const double *__restrict__ x, double *__restrict__ y) {
    pdl_prologue();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *a = A + (size_t)row * n;
    const int j0 = LOWER ? 0 : row, j1 = LOWER ? row + 1 : n;
    double s = 0.0;
    for (int j = (j0 & ~31) + lane; j < j1; j += 32) if (j >= j0) s = fma(__ldg(a + j), x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[row] = s;
}
void launch_dense_trmv(const LaunchCtx &ctx, const double *A, int n, const double *x, double *y, bool lower) {
    ProfScope ps(ctx, PC_COARSE_SOLVE, (double)n * n / 2);
    if (n == 0) return;
    dim3 block(256), grid((unsigned)(((size_t)n * 32 + 255) / 256));
    if (lower) VF_LAUNCH((k_dense_trmv<true>), grid, block, 0, ctx.stream, A, n, x, y);
    else       VF_LAUNCH((k_dense_trmv<false>), grid, block, 0, ctx.stream, A, n, x, y);
    VF_KERNEL_CHECK();
}

__global__ void k_gather_free(const double *__restrict__ f, const int *__restrict__ freeDofs, int nfree, long long numNodes, int N, double *__restrict__ rhs) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfree) return;
    const int dof = freeDofs[i];
    rhs[i] = f[(long long)(dof % N) * numNodes + dof / N];
}
void launch_gather_free(const LaunchCtx &ctx, const double *f, const int *freeDofs, int nfree, long long numNodes, int N, double *rhs) {
    ProfScope ps(ctx, PC_COARSE_SOLVE, (double)nfree);
    if (nfree == 0) return;
    VF_LAUNCH((k_gather_free), (nfree + 255) / 256, 256, 0, ctx.stream, f, freeDofs, nfree, numNodes, N, rhs);
    VF_KERNEL_CHECK();
}
__global__ void k_scatter_free(const double *__restrict__ y, const int *__restrict__ freeDofs, int nfree, long long numNodes, int N, double *__restrict__ x) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfree) return;
    const int dof = freeDofs[i];
    x[(long long)(dof % N) * numNodes + dof / N] = y[i];
}
void launch_scatter_free(const LaunchCtx &ctx, const double *y, const int *freeDofs, int nfree, long long numNodes, int N, double *x) {
    ProfScope ps(ctx, PC_COARSE_SOLVE, (double)nfree);
    if (nfree == 0) return;
    VF_LAUNCH((k_scatter_free), (nfree + 255) / 256, 256, 0, ctx.stream, y, freeDofs, nfree, numNodes, N, x);
    VF_KERNEL_CHECK();
}

} // namespace vf
