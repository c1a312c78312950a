// vf_internal.cuh -- shared declarations for the CUDA side of libvoxelfem_b200.
// Internal to the library; the public surface is include/voxelfem_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <mutex>

namespace vf {

// ---------------------------------------------------------------------------
// Grid descriptor, passed to kernels by value.
// 2D grids are embedded with a leading dummy axis: nn = {1, nx, ny}, so node (x, y) has
// flat index x*ny + y exactly as in the reference (NDVector.hh:256-264) and the fastest
// axis is always embedded axis 2.
// ---------------------------------------------------------------------------
struct GridDesc {
    int N;              // 2 or 3
    int nn[3];          // nodes per embedded axis
    int ne[3];          // elements per embedded axis (1 along the dummy axis in 2D)
    long long ns[3];    // node strides
    long long es[3];    // element strides
    long long numNodes, numElems;
    int bd;             // embedded build-direction axis (reference axis 1): 1 in 3D, 2 in 2D
    int nActive;        // non-detached node layers along bd (nondetachedNodesPerDim, TensorProductSimulator.hh:371-375)
    int neActive;       // non-masked element layers along bd (nonmaskedElementsPerDim, :377-381)
    // Colour-major node numbering used by the stored stencils of the coarse levels: the nodes of parity class
    // (colour) c occupy positions cbase[c] .. cbase[c] + prod(ccnt[c]) in row-major order of (i_a >> 1).  A colour
    // pass of the smoother then streams its stencil rows from contiguous memory.
    // Colour bases are padded to multiples of kStencilTile so that every tile of kStencilTile consecutive positions belongs
    // to one colour; numPos = padded number of positions (>= numNodes).
    long long cbase[8];
    int ccnt[8][3];
    long long numPos;
    // Slab windows (multi-GPU).  A level of a slab-partitioned solver stores a window of the global grid:
    // local plane i along embedded axis 0 is global plane xoff + i.  Defaults describe an undivided grid.
    int xoff;           // global index of local node plane 0 (and of local element layer 0)
    int ownLo, ownHi;   // local node planes [ownLo, ownHi) this part owns: reductions count exactly these
    int cmpLo, cmpHi;   // local node planes [cmpLo, cmpHi) the smoother updates (owned + shared planes; ghost planes are received)
    int oeLo, oeHi;     // local element layers [oeLo, oeHi) this part owns: Galerkin coarsening sub-assembles exactly these
    // NB: the level-0 kernels are compiled at their register limit with this struct ahead of their constant-bank tables; changing its
    // size shifted those operands and cost k_apply3w_l0 48 B of extra spills and 30 % of its speed (profiles/r06c_bench.log).
};

// Stored stencils are tiled: the NE = 3^N * N * N entries of kStencilTile consecutive positions form one contiguous
// block  S[tile][entry][lane]  (31,104 B in 3D), which one cp.async.bulk (TMA) copy stages into shared memory.
constexpr int kStencilTile = 16;

#if defined(__CUDACC__)
#define VF_HD __host__ __device__ __forceinline__
#else
#define VF_HD inline
#endif
// position of node (c0, c1, c2) in the colour-major numbering
VF_HD long long stencil_pos(const GridDesc &g, int c0, int c1, int c2) {
    const int col = ((c0 & 1) << 2) | ((c1 & 1) << 1) | (c2 & 1);
    return g.cbase[col] + ((long long)(c0 >> 1) * g.ccnt[col][1] + (c1 >> 1)) * g.ccnt[col][2] + (c2 >> 1);
}

// offset of stencil entry `entry` (= slot * N*N + a*N + b) of position pos; NE = entries per node
VF_HD long long stencil_addr(long long pos, int entry, int NE) {
    return ((pos / kStencilTile) * NE + entry) * kStencilTile + (pos % kStencilTile);
}

template<int N> struct Dims {
    static constexpr int NPE = 1 << N;       // nodes per element
    static constexpr int KE  = N * NPE;      // element matrix size
    static constexpr int NS  = (N == 3) ? 27 : 9; // stencil slots
    static constexpr int NE  = NS * N * N;        // stencil entries per node
    static constexpr int A0  = 3 - N;        // first active embedded axis
};

// Full-density element stiffness matrix, passed by value (__grid_constant__) so that its
// entries become constant-bank operands of the DFMAs.
struct K0Param {
    double v[24 * 24];
    // 3D only: K0 in the symmetry-adapted (Walsh) basis of the voxel's mirror group, filled by finalize_k0_param().
    // kh[p][c][c'] couples the modes (component c, sign pattern s = p ^ (4 >> c)) of irreducible representation p.
    double kh[8][3][3];
    int walsh;      // 1 if K0 is block diagonal in that basis (orthotropic material aligned with the grid)
    int sparse7;    // 1 if additionally the trilinear mode (s = 7) of a component only couples with itself
    // The same mirror symmetry in the nodal basis: K0[(m,c),(m^D,d)] = s(m,c,d) vt[D][3c+d] with s = +1 for c == d and
    // (-1)^(m_c + m_d) otherwise (m_a = coordinate of local node m along axis a): 72 numbers instead of 576, each shared by the
    // 8 incident elements of a node.  Rows padded to 10 doubles so that constant-bank pairs are 16-byte aligned.
    double vt[8][10];
};
// Passed by value to the Walsh-basis kernels (constant-bank operands).
struct KhatParam { double v[8][3][3]; };
void finalize_k0_param(K0Param &K, int N);   // vf_l0.cu


// Colour (parity class) description for one pass of the multicoloured smoother
// (visitNodesMulticolored, MultigridSolver.hh:408-442): nodes off + 2*i, i < cnt, per embedded axis.
struct ColorDesc { int off[3]; int cnt[3]; };
inline bool make_color(const GridDesc &g, int color, ColorDesc &col) {
    const int A0 = 3 - g.N;
    for (int a = 0; a < 3; ++a) { col.off[a] = 0; col.cnt[a] = 1; }
    for (int a = A0; a < 3; ++a) {
        col.off[a] = (color >> (2 - a)) & 1;
        const int lim = (a == g.bd) ? g.nActive : g.nn[a];
        if (lim - 1 - col.off[a] < 0) return false;
        col.cnt[a] = (lim - 1 - col.off[a]) / 2 + 1;
    }
    return true;
}

#ifdef __CUDACC__
template<int N> __device__ __forceinline__ void solve_block(const double (&M)[N][N], const double (&rhs)[N], double (&du)[N]);
// Closed-form cofactor inverse, as Eigen's fixed-size Matrix::inverse() (MultigridSolver.hh:370)
template<> __device__ __forceinline__ void solve_block<3>(const double (&M)[3][3], const double (&r)[3], double (&du)[3]) {
    const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
    const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const double det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
    const double id = 1.0 / det;
    du[0] = (c00 * r[0] + (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * r[1] + (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * r[2]) * id;
    du[1] = (c01 * r[0] + (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * r[1] + (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * r[2]) * id;
    du[2] = (c02 * r[0] + (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * r[1] + (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * r[2]) * id;
}
template<> __device__ __forceinline__ void solve_block<2>(const double (&M)[2][2], const double (&r)[2], double (&du)[2]) {
    const double id = 1.0 / (M[0][0] * M[1][1] - M[0][1] * M[1][0]);
    du[0] = (M[1][1] * r[0] - M[0][1] * r[1]) * id;
    du[1] = (M[0][0] * r[1] - M[1][0] * r[0]) * id;
}
// the same cofactor inverse as a matrix (one reciprocal): the stored-stencil smoother applies it to the reduced right-hand side
template<int N> __device__ __forceinline__ void block_inverse(const double (&M)[N][N], double (&G)[N][N]);
template<> __device__ __forceinline__ void block_inverse<3>(const double (&M)[3][3], double (&G)[3][3]) {
    const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
    const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const double id = 1.0 / (M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02);
    G[0][0] = c00 * id; G[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id; G[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
    G[1][0] = c01 * id; G[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id; G[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
    G[2][0] = c02 * id; G[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id; G[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
}
template<> __device__ __forceinline__ void block_inverse<2>(const double (&M)[2][2], double (&G)[2][2]) {
    const double id = 1.0 / (M[0][0] * M[1][1] - M[0][1] * M[1][0]);
    G[0][0] = M[1][1] * id; G[0][1] = -M[0][1] * id; G[1][0] = -M[1][0] * id; G[1][1] = M[0][0] * id;
}
// u_n += M^-1 (b - S') for free nodes; point Gauss-Seidel on the free components of partially
// constrained nodes, direction following the sweep (MultigridSolver.hh:358-365).
template<int N, bool FWD>
__device__ __forceinline__ void gs_point_sweep(const double (&M)[N][N], const double (&rhs)[N], unsigned dm, double (&du)[N]) {
    #pragma unroll
    for (int k = 0; k < N; ++k) {
        constexpr int dummy = 0; (void)dummy;
        const int i = FWD ? k : (N - 1 - k);   // compile-time after unrolling: M, rhs, du stay in registers
        double s = rhs[i];
        #pragma unroll
        for (int j = 0; j < N; ++j) s -= M[i][j] * du[j];
        du[i] = s * ((((dm >> i) & 1u) ? 0.0 : 1.0) / M[i][i]);
    }
}
template<int N>
__device__ __forceinline__ void gs_node_update(const double (&M)[N][N], const double (&rhs)[N], unsigned dm, bool forward, double (&du)[N]) {
    if (dm == 0u) { solve_block<N>(M, rhs, du); return; }
    #pragma unroll
    for (int c = 0; c < N; ++c) du[c] = 0.0;
    if (forward) gs_point_sweep<N, true>(M, rhs, dm, du);
    else         gs_point_sweep<N, false>(M, rhs, dm, du);
}
#endif

// ---------------------------------------------------------------------------
// Profiling categories (device time per kernel family; vf_prof_* in the C ABI)
// ---------------------------------------------------------------------------
enum ProfCat {
    PC_APPLY_L0 = 0, PC_RESIDUAL_L0, PC_GS_L0, PC_APPLY_ST, PC_RESIDUAL_ST, PC_GS_ST, PC_RESTRICT, PC_PROLONG,
    PC_COARSE_SOLVE, PC_VEC, PC_COARSEN, PC_TOPOPT, PC_OTHER,
    // stored-stencil levels whose stencil fits the 126 MB L2 several times over are launch-latency bound, not HBM bound: own families
    PC_APPLY_ST_SMALL, PC_RESIDUAL_ST_SMALL, PC_GS_ST_SMALL, PC_COUNT
};

// a stencil level streams from HBM when its stencil (1944 B per node in 3D) is larger than 1 GiB (level 1 of the 256^3 grid: 4.2 GB; level 2: 0.5 GB, a few L2 sizes, latency-dominated)
inline bool stencil_level_streams(const GridDesc &g) { return (double)g.numNodes * (g.N == 3 ? 1944.0 : 288.0) > 1024.0 * 1048576.0; }
struct Profiler;
struct LaunchCtx {
    cudaStream_t stream = nullptr;
    Profiler *prof = nullptr;
};
void prof_begin(const LaunchCtx &ctx, int cat, double units);
void prof_end(const LaunchCtx &ctx, int cat);
void count_launch();
void set_last_error(const std::string &msg);   // message returned by vf_last_error() on this thread

struct ProfScope {
    const LaunchCtx &c; int cat;
    ProfScope(const LaunchCtx &ctx, int cat_, double units) : c(ctx), cat(cat_) { prof_begin(c, cat, units); count_launch(); }
    ~ProfScope() { prof_end(c, cat); }
};

// vf_trace.cu: section timers + NVTX ranges under the reference's section names (GlobalBenchmark.hh, Timer.hh)
void trace_push(const char *name);
void trace_pop(const char *name);
bool trace_enabled();
struct TraceScope {   // BENCHMARK_SCOPED_TIMER_SECTION
    std::string name;
    explicit TraceScope(std::string n) : name(std::move(n)) { trace_push(name.c_str()); }
    ~TraceScope() { trace_pop(name.c_str()); }
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
    if (e != cudaSuccess) {
        throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")");
    }
}
#define VF_CUDA(x) ::vf::cuda_check((x), #x, __FILE__, __LINE__)
#define VF_KERNEL_CHECK() ::vf::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// Per-device one-time setup (function attributes, lookup tables): kernels' attributes and device allocations belong to ONE device,
// and a process may drive several (vf_set_device).  first_use_on_device(flags) is true exactly once per device and flag array.
constexpr int kMaxDevices = 64;
struct PerDeviceFlags { bool done[kMaxDevices] = {}; std::mutex m; };
inline int current_device() { int d = 0; cuda_check(cudaGetDevice(&d), "cudaGetDevice", __FILE__, __LINE__); return d; }
inline bool first_use_on_device(PerDeviceFlags &f) {
    const int d = current_device();
    std::lock_guard<std::mutex> lock(f.m);
    if (d < 0 || d >= kMaxDevices) return true;
    if (f.done[d]) return false;
    f.done[d] = true;
    return true;
}

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Kernels of the solve path start with pdl_prologue(): griddepcontrol.launch_dependents
// lets a NEXT kernel launched with the programmatic-stream-serialization attribute be scheduled as soon as this grid's blocks
// have all started, griddepcontrol.wait holds its blocks until the PREVIOUS grid has completed and its writes are visible.
// The attribute is set on the colour passes 2..8 of a stored-stencil smoothing sweep only, whose blocks request their stencil
// tile (TMA) before the wait: the HBM round trip of a pass's first wave overlaps the tail of the previous pass (level-1
// sweep 0.914 -> 0.891 ms, level-2 sweep 0.156 -> 0.132 ms).  Setting it on every launch of the captured preconditioner
// graph made the FMG cycle slower (14.4 vs 13.5 ms, profiles/r02b_time_ab.log).  VF_PDL=0 disables the attribute.
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
// mbarrier + TMA bulk copy (cp.async.bulk) helpers shared by the stored-stencil and the level-0 smoother kernels
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same copy with an L2 evict-first hint: for data that is streamed once per pass (stored stencils), so that it does not push
// the re-used nodal fields out of the L2
__device__ __forceinline__ void tma_load_1d_stream(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// a global load that asks the L2 to keep the line (evict-last): the nodal fields of a level whose stencil streams past them
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ double ld_l2_hint(const double *p, unsigned long long pol) {
    double v; asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }
bool pdl_enabled();   // vf_vec.cu
template<class... KArgs, class... Args>
inline void launch_pdl(bool programmatic, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (programmatic && pdl_enabled()) ? 1 : 0;
    cuda_check(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...), "cudaLaunchKernelEx", __FILE__, __LINE__);
}
// VF_LAUNCH: ordinary stream-ordered launch; VF_LAUNCH_PDL: may start before its predecessor has drained (see above)
#define VF_LAUNCH(kernel, grid, block, smem, stream, ...) ::vf::launch_pdl(false, kernel, grid, block, smem, stream, __VA_ARGS__)
#define VF_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) ::vf::launch_pdl(pdl, kernel, grid, block, smem, stream, __VA_ARGS__)
#endif

// ---------------------------------------------------------------------------
// Kernel launchers (defined in the .cu files named in the comments)
// ---------------------------------------------------------------------------
// --- vf_l0.cu: matrix-free level-0 operator (TPSStencils.hh:231-396, 431-728; MultigridSolver.hh:277-292, 347-378)
enum ApplyMode { APPLY_SET = 0, APPLY_ADD = 1, APPLY_SUB = 2, APPLY_RESIDUAL = 3 };
// out (=, +=, -=) K u   or   out = b - K u (APPLY_RESIDUAL);  dmask != nullptr zeroes Dirichlet components of out.
// dotOut != nullptr additionally accumulates sum(u . out) over the non-detached nodes (deterministic two-stage reduction).
void launch_apply_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                     const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch);
// One colour pass of the block Gauss-Seidel smoother at level 0.
void launch_gs_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                  const uint8_t *dmask, int color, bool forward);

// --- vf_gs0.cu: level-0 smoother on row units (3D, mirror-symmetric K0)
// One smoothing sweep (all 8 colours in the reference order) as 4 launches: one per (x, y) parity class of node rows; a thread block
// stages the 9 neighbouring u rows, b and the 4 adjacent modulus rows of one z-row in shared memory (cp.async, parity-split) and
// runs both z-colours of the row.  gs_rows_supported(): whether the kernel covers this grid / material.
bool gs_rows_supported(const GridDesc &g, const K0Param &K);
// cls: 0..3, the (x parity, y parity) = (cls >> 1, cls & 1) class in LOCAL parities of the window
void launch_gs_rows_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                       const uint8_t *dmask, int cls, bool forward);
// measured FP64 FMA throughput (T DFMA/s) of the device: independent register-resident DFMA chains on every SM
double measure_dfma_peak(cudaStream_t stream);

// One colour pass, two nodes per thread (3D only).
void launch_gs3_color_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                         const uint8_t *dmask, int color, bool forward);

// --- vf_stencil.cu: 3^N-point block-stencil levels (MultigridSolver.hh:323-334; TensorProductSimulator.hh:1500-1504)
// Stencil layout: S[stencil_addr(stencil_pos(node), slot * N*N + a*N + b, NE)], numPos * NE doubles per level
void launch_apply_stencil(const LaunchCtx &ctx, const GridDesc &g, const double *S, const double *u, const double *b,
                          const uint8_t *dmask, double *out, int mode, const unsigned long long *posTab = nullptr);
// chained: the previous kernel on the stream is a colour pass of the same sweep (nothing in flight writes S)
bool stencil_sweep_fused(const GridDesc &g);   // small level: all colour passes of a sweep in one persistent launch
void launch_gs_stencil_sweep(const LaunchCtx &ctx, const GridDesc &g, const double *S, double *u, const double *b, const uint8_t *dmask,
                             bool forward, int xparity, unsigned *bar);
void launch_gs_stencil(const LaunchCtx &ctx, const GridDesc &g, const double *S, double *u, const double *b,
                       const uint8_t *dmask, int color, bool forward, bool chained = false, double *resOut = nullptr,
                       const unsigned long long *posTab = nullptr);
// posTab: the level's device table  position -> packed node coordinates  (launch_fill_pos_table), or nullptr (the tile kernels
// then derive the coordinates arithmetically)
// resOut != nullptr: the pass also accumulates the residual of the sweep's final iterate (k_stencil_tile<RES>); all 2^N passes of
// the sweep must be given the same resOut, the grid must be fully attached, Dirichlet components are left unmasked and the shared
// planes of a slab window incomplete (launch_residual_stencil_plane)
bool gs_residual_fusable(const GridDesc &g);
void launch_residual_stencil_plane(const LaunchCtx &ctx, const GridDesc &g, const double *S, const double *u, const double *b,
                                   const uint8_t *dmask, double *out, int plane, const unsigned long long *posTab = nullptr);
void launch_fill_pos_table(cudaStream_t stream, const GridDesc &g, unsigned long long *tab);   // tab: g.numPos entries
// Galerkin coarsening (MultigridSolver.hh:711-819): level-1 stencil from the fine Young's moduli and the 2^N
// coarsened full-density matrices cK0[fi] (device, [fi][KE][KE]); level l >= 2 stencil as P^T A_{l-1} P.
// bandLo..bandHi (inclusive, coarse node layers along the build direction): only those rows are recomputed -- the banded update of
// updateStiffnessMatrices after a change of the fabrication mask (MultigridSolver.hh:907-1017); default: every row.
void launch_coarsen_from_moduli(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *E, const double *cK0, double *Sc, int bandLo = 0, int bandHi = 0x7fffffff,
                                const double *cK0host = nullptr);   // cK0host: host copy of cK0, enables the per-slot kernel with the blocks as kernel parameters
void launch_coarsen_stencil(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *Sf, double *Sc, int bandLo = 0, int bandHi = 0x7fffffff);
// P^T A P one axis at a time (3D, undivided grids, full rebuilds); scratch holds the two intermediate operators
size_t coarsen_separable_scratch(const GridDesc &gc, const GridDesc &gf);
void launch_coarsen_stencil_separable(const LaunchCtx &ctx, const GridDesc &gc, const GridDesc &gf, const double *Sf, double *Sc, double *scratch);
// Level-0 stencil straight from moduli (used for single-level direct solves): S = sum_e E_e K0 blocks.
void launch_stencil_from_moduli_l0(const LaunchCtx &ctx, const GridDesc &g, const double *E, const double *K0dev, double *S);
// Dense matrix of the free DOFs from a stencil: A[red(i)][red(j)], row-major n x n; redIdx[dof] = -1 for fixed DOFs.
// Slab completion of a sub-assembled stencil: rows of local node plane `plane` <-> contiguous buffer [position in plane][entry]
long long stencil_plane_rows(const GridDesc &g, int plane);   // number of nodes in a plane
void launch_stencil_plane_pack(const LaunchCtx &ctx, const GridDesc &g, const double *S, int plane, double *buf);
void launch_stencil_plane_add(const LaunchCtx &ctx, const GridDesc &g, double *S, int plane, const double *buf);
void launch_stencil_to_dense(const LaunchCtx &ctx, const GridDesc &g, const double *S, const int *redIdx, int nfree, double *A);
void launch_symmetrize_lower(const LaunchCtx &ctx, double *A, int n);   // copy lower (row-major) triangle to upper
void launch_dense_symv(const LaunchCtx &ctx, const double *A, int n, const double *x, double *y);
// y = tril(A) x (lower = true) or y = triu(A) x (lower = false), diagonal included; row-major A
void launch_dense_trmv(const LaunchCtx &ctx, const double *A, int n, const double *x, double *y, bool lower);
void launch_gather_free(const LaunchCtx &ctx, const double *f, const int *freeDofs, int nfree, long long numNodes, int N, double *rhs);
void launch_scatter_free(const LaunchCtx &ctx, const double *y, const int *freeDofs, int nfree, long long numNodes, int N, double *x);

// --- vf_dense.cu: dense FP64 kernels of the coarsest-level direct solver (column-major, leading dimensions)
constexpr int kDiagBlockHost = 64;   // panel width of the blocked factorization (== kDiagBlock of vf_dense.cu)
// C (m x n) = alpha * A (m x k) * op(B) + beta * C;  transB: B is n x k;  lowerOnly: skip the tiles strictly above the diagonal
void launch_dgemm(const LaunchCtx &ctx, bool transB, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
                  double beta, double *C, int ldc, bool lowerOnly);
// M (m x m SPD, lower triangle read, trailing part overwritten) -> Lb = L (scratch), X = L^-1 (zeros above the diagonal must be there);
// tmp: kDiagBlockHost * m doubles; *info (zero-initialised) receives infoBase + 1 + index of the first non-positive pivot
void potrf_inv_blocked(const LaunchCtx &ctx, double *M, int ld, int m, double *Lb, int ldl, double *X, int ldx, double *tmp, int *info, int infoBase);

// --- vf_vec.cu: transfers and PCG vector kernels
// Grid transfers between two windows: fine plane index = 2 * coarse plane index + xshift(gf, gc) along embedded axis 0.
VF_HD int xshift(const GridDesc &gf, const GridDesc &gc) { return 2 * gc.xoff - gf.xoff; }
void launch_restrict(const LaunchCtx &ctx, const GridDesc &gf, const GridDesc &gc, const double *fine, double *coarse);
void launch_prolong(const LaunchCtx &ctx, const GridDesc &gf, const GridDesc &gc, const double *coarse, double *fine, bool accumulate);
void launch_zero_dirichlet(const LaunchCtx &ctx, const GridDesc &g, const uint8_t *dmask, double *u);
void launch_enforce_dirichlet(const LaunchCtx &ctx, long long numNodes, int N, int ndir, const long long *nodes, const uint8_t *masks, const double *vals, double *u);
void launch_masked_zero(const LaunchCtx &ctx, const GridDesc &g, double *u, int margin);
void launch_detached_zero(const LaunchCtx &ctx, const GridDesc &g, double *u); // zero the detached node layers
void launch_masked_copy(const LaunchCtx &ctx, const GridDesc &g, const double *in, double *out, int margin);
// result[slot] = sum over non-detached nodes of a . b   (deterministic)
void launch_masked_dot(const LaunchCtx &ctx, const GridDesc &g, const double *a, const double *b, double *result, double *scratch);
// d = s + (num/den) d   (first == true: d = s)           (scaleAndAddInPlace, ParallelVectorOps.hh:76-84)
void launch_cg_direction(const LaunchCtx &ctx, const GridDesc &g, const double *s, double *d, const double *num, const double *den, bool first);
// alpha = rMr / dAd;  x += alpha d;  r -= alpha Ad;  rsq = ||r||^2      (MultigridSolver.hh:1134-1143)
void launch_cg_update(const LaunchCtx &ctx, const GridDesc &g, double *x, const double *d, double *r, const double *Ad,
                      const double *rMr, const double *dAd, double *rsq, double *scratch);
size_t reduce_scratch_doubles();

// --- vf_top.cu: optimization-layer kernels
void launch_update_moduli(const LaunchCtx &ctx, const GridDesc &g, const double *rho, double *E, int law, double E0, double Emin, double gamma, double q, bool maskActive);
void launch_zero_moduli_layers(const LaunchCtx &ctx, const GridDesc &g, double *E, int layerBegin, int layerEnd);
void launch_compliance_gradient(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *rho, double *out,
                                int law, double E0, double Emin, double gamma, double q, const double *gravity, double elemVol, bool accumulate);
void launch_energy_density(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E, double *out);
void launch_self_weight_load(const LaunchCtx &ctx, const GridDesc &g, const double *rho, const double *gravity, double elemVol, double *f, int layerBegin, int layerEnd, double sign);
void launch_filter_smooth(const LaunchCtx &ctx, int N, const int *sizes, int radius, int type, const double *in, double *out);
void launch_filter_project(const LaunchCtx &ctx, long long n, double beta, const double *in, double *out);
void launch_filter_project_backprop(const LaunchCtx &ctx, long long n, double beta, const double *in, const double *vars, double *out);
// --- vf_filters.cu: UpsampleFilter, VertexToCellFilter, LangelaarFilter (TopologyOptimizationFilter.hh:418-712); sizes: first N entries
void launch_filter_upsample(const LaunchCtx &ctx, int N, const int *coarseSizes, int factor, const double *in, double *out);
void launch_filter_upsample_backprop(const LaunchCtx &ctx, int N, const int *coarseSizes, int factor, const double *dout, double *din);
void launch_filter_v2c(const LaunchCtx &ctx, int N, const int *vertexSizes, const double *in, double *out);
void launch_filter_v2c_backprop(const LaunchCtx &ctx, int N, const int *vertexSizes, const double *dout, double *din);
void launch_filter_langelaar(const LaunchCtx &ctx, int N, const int *sizes, const double *in, double *out, double *smaxCache);
void launch_filter_langelaar_backprop(const LaunchCtx &ctx, int N, const int *sizes, const double *g_in, const double *vars, const double *filtered,
                                      const double *smaxCache, double *scratch, double *out);
// OC update (OptimalityCriterion.hh:64-83): out = clamp(x0 * (dJ / (dc*lambda))^p, x0 -+ m, [0,1]); non-finite -> x0
void launch_oc_update(const LaunchCtx &ctx, long long n, const double *x0, const double *dJ, const double *dc, double lambda, double m, double p, double *out);
void launch_sum(const LaunchCtx &ctx, long long n, const double *x, double *result, double *scratch);
void launch_fill(const LaunchCtx &ctx, long long n, double v, double *x);
void launch_dot_plain(const LaunchCtx &ctx, long long n, const double *a, const double *b, double *result, double *scratch);
void launch_axpy(const LaunchCtx &ctx, long long n, double a, const double *x, double *y); // y += a x
void launch_scale(const LaunchCtx &ctx, long long n, double a, double *x);

} // namespace vf
