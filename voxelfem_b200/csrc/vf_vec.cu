// vf_vec.cu -- grid-transfer and PCG vector kernels (sm_100a).  All HBM-bound streams.
//
// restriction / interpolation: MultigridSolver.hh:178-262 with the weights of
// TPSStencils::fineNodesInSupport (TPSStencils.hh:82-128).
// Masked vector ops: TensorProductSimulator.hh:413-459; PCG updates: MultigridSolver.hh:1108-1143,
// ParallelVectorOps.hh:28-84.
#include "vf_internal.cuh"
#include <cstdlib>
#include "vf_reduce.cuh"

namespace vf {

bool pdl_enabled() {
    static const bool v = [] { const char *e = std::getenv("VF_PDL"); return !(e && e[0] == '0'); }();
    return v;
}

static inline dim3 node_block(const GridDesc &g) { return (g.N == 3) ? dim3(32, 4, 2) : dim3(32, 8, 1); }
static inline dim3 node_grid(const GridDesc &g, dim3 b) { return dim3((g.nn[2] + b.x - 1) / b.x, (g.nn[1] + b.y - 1) / b.y, (g.nn[0] + b.z - 1) / b.z); }

// coarse_n = sum_{i in [-1,1]^N, fine node in range and non-detached} prod_d (1 - |i_d|/2) * fine[2n + i]
template<int N>
__global__ void __launch_bounds__(256)
k_restrict(const __grid_constant__ GridDesc gf, const __grid_constant__ GridDesc gc, const double *__restrict__ fine, double *__restrict__ coarse) {
    pdl_prologue();
    constexpr int A0 = Dims<N>::A0, NS = Dims<N>::NS;
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int c1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = blockIdx.z * blockDim.z + threadIdx.z;
    if (c2 >= gc.nn[2] || c1 >= gc.nn[1] || c0 >= gc.nn[0]) return;
    const int cc[3] = {c0, c1, c2};
    const long long n = (long long)c0 * gc.ns[0] + (long long)c1 * gc.ns[1] + c2;
    const int cbd = (gc.bd == 1) ? c1 : c2;
    if (cbd > gc.nActive) return;                     // deeper detached layers are left untouched
    double acc[N];
    #pragma unroll
    for (int c = 0; c < N; ++c) acc[c] = 0.0;
    if (cbd < gc.nActive) {                           // cbd == nActive: first detached layer is zeroed (MultigridSolver.hh:251-262)
        #pragma unroll
        for (int s = 0; s < NS; ++s) {
            int d[3] = {0, 0, 0};
            { int r = s;
              #pragma unroll
              for (int a = 2; a >= A0; --a) { d[a] = r % 3 - 1; r /= 3; } }
            // branch-free: an out-of-range neighbour is read at a clamped (valid, non-detached) index with weight 0, so all
            // 3^N * N loads are independent and in flight together (the branchy version serialised on load latency: 11 us for
            // a 17^3 coarse grid)
            bool valid = true; long long fi = 0; double w = 1.0;
            #pragma unroll
            for (int a = A0; a < 3; ++a) {
                const int q = 2 * cc[a] + d[a] + (a == 0 ? xshift(gf, gc) : 0);
                const int lim = (a == gf.bd) ? gf.nActive : gf.nn[a];
                valid = valid && q >= 0 && q < lim;
                fi += (long long)min(max(q, 0), lim - 1) * gf.ns[a];
                w *= (d[a] == 0) ? 1.0 : 0.5;
            }
            w = valid ? w : 0.0;
            #pragma unroll
            for (int c = 0; c < N; ++c) acc[c] = fma(w, fine[c * gf.numNodes + fi], acc[c]);
        }
    }
    #pragma unroll
    for (int c = 0; c < N; ++c) coarse[c * gc.numNodes + n] = acc[c];
}

void launch_restrict(const LaunchCtx &ctx, const GridDesc &gf, const GridDesc &gc, const double *fine, double *coarse) {
    ProfScope ps(ctx, PC_RESTRICT, (double)gf.numNodes);
    dim3 b = node_block(gc), g = node_grid(gc, b);
    if (gc.N == 3) VF_LAUNCH((k_restrict<3>), g, b, 0, ctx.stream, gf, gc, fine, coarse);
    else           VF_LAUNCH((k_restrict<2>), g, b, 0, ctx.stream, gf, gc, fine, coarse);
    VF_KERNEL_CHECK();
}

// fine_n (=, +=) sum over the <= 2^N coarse nodes around it of the multilinear weights.
// The weight prod_a (odd_a ? 1/2 : 1) does not depend on which of the surrounding coarse nodes is visited and is a power of two, so
// the coarse values are summed (same order as the reference's row visit, MultigridSolver.hh:162-176) and scaled once: bit-identical
// to accumulating w * value, with a third of the instructions (the first version spent 357 instructions per fine node and was
// issue-bound at 0.29 / 0.46 ms for the 256^3 grid).  x and y parities are warp-uniform, so even rows skip their absent neighbours.
template<int N, bool ACC>
__global__ void __launch_bounds__(256)
k_prolong(const __grid_constant__ GridDesc gf, const __grid_constant__ GridDesc gc, const double *__restrict__ coarse, double *__restrict__ fine) {
    pdl_prologue();
    constexpr int A0 = Dims<N>::A0;
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int c1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = blockIdx.z * blockDim.z + threadIdx.z;
    if (c2 >= gf.nn[2] || c1 >= gf.nn[1] || c0 >= gf.nn[0]) return;
    const int cc[3] = {c0, c1, c2};
    if (ACC && ((gf.bd == 1) ? c1 : c2) >= gf.nActive) return; // accum_interpolation visits non-detached nodes only (:192)
    const long long n = (long long)c0 * gf.ns[0] + (long long)c1 * gf.ns[1] + c2;
    long long base = 0; bool ok0[3] = {true, true, true}, ok1[3] = {false, false, false}; double w = 1.0;
    #pragma unroll
    for (int a = A0; a < 3; ++a) {
        const int fa = cc[a] - (a == 0 ? xshift(gf, gc) : 0);   // fine index relative to coarse plane 0 (may be negative in a window)
        const int q = fa >> 1;                                  // arithmetic shift = floor division
        const bool odd = fa & 1;
        ok0[a] = q >= 0 && q < gc.nn[a];                        // outside the coarse window: only for ghost planes, which are received
        ok1[a] = odd && q + 1 >= 0 && q + 1 < gc.nn[a];         // even fine index: single coarse node
        base += (long long)q * gc.ns[a];
        if (odd) w *= 0.5;
    }
    double acc[N];
    #pragma unroll
    for (int c = 0; c < N; ++c) acc[c] = 0.0;
    const long long NC = gc.numNodes;
    #pragma unroll
    for (int b0 = 0; b0 < (N == 3 ? 2 : 1); ++b0) {
        if (!(b0 ? ok1[0] : ok0[0])) continue;
        #pragma unroll
        for (int b1 = 0; b1 < 2; ++b1) {
            if (!(b1 ? ok1[1] : ok0[1])) continue;
            const long long row = base + (b0 ? gc.ns[0] : 0) + (b1 ? gc.ns[1] : 0);
            #pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                if (!(b2 ? ok1[2] : ok0[2])) continue;
                #pragma unroll
                for (int c = 0; c < N; ++c) acc[c] += coarse[c * NC + row + b2];
            }
        }
    }
    #pragma unroll
    for (int c = 0; c < N; ++c) {
        if (ACC) fine[c * gf.numNodes + n] += w * acc[c]; else fine[c * gf.numNodes + n] = w * acc[c];
    }
}

void launch_prolong(const LaunchCtx &ctx, const GridDesc &gf, const GridDesc &gc, const double *coarse, double *fine, bool accumulate) {
    ProfScope ps(ctx, PC_PROLONG, (double)gf.numNodes);
    dim3 b = node_block(gf), g = node_grid(gf, b);
    if (gf.N == 3) { if (accumulate) VF_LAUNCH((k_prolong<3, true>), g, b, 0, ctx.stream, gf, gc, coarse, fine); else VF_LAUNCH((k_prolong<3, false>), g, b, 0, ctx.stream, gf, gc, coarse, fine); }
    else           { if (accumulate) VF_LAUNCH((k_prolong<2, true>), g, b, 0, ctx.stream, gf, gc, coarse, fine); else VF_LAUNCH((k_prolong<2, false>), g, b, 0, ctx.stream, gf, gc, coarse, fine); }
    VF_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------
// Flat masked loops: index i over all nodes; active iff coordinate along bd < limit
// ---------------------------------------------------------------------------
// node belongs to the planes this part owns (slab windows; always true for an undivided grid)
__device__ __forceinline__ bool node_owned(const GridDesc &g, long long n) {
    if (g.ownLo <= 0 && g.ownHi >= g.nn[0]) return true;
    const int p = (int)(n / g.ns[0]);
    return p >= g.ownLo && p < g.ownHi;
}
__device__ __forceinline__ bool node_active(const GridDesc &g, long long n, int limit) {
    if (limit >= g.nn[g.bd]) return true;
    const int cbd = (g.bd == 2) ? (int)(n % g.nn[2]) : (int)((n / g.nn[2]) % g.nn[1]);
    return cbd < limit;
}
static inline int flat_blocks(long long n) { long long b = (n + 255) / 256; const long long cap = 148LL * 16; return (int)(b < cap ? (b > 0 ? b : 1) : cap); }

__global__ void __launch_bounds__(256) k_zero_dirichlet(const __grid_constant__ GridDesc g, const uint8_t *__restrict__ dmask, double *__restrict__ u) {
    pdl_prologue();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x) {
        const unsigned dm = dmask[n];
        if (dm) { for (int c = 0; c < g.N; ++c) if ((dm >> c) & 1u) u[c * g.numNodes + n] = 0.0; }
    }
}
void launch_zero_dirichlet(const LaunchCtx &ctx, const GridDesc &g, const uint8_t *dmask, double *u) {
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_zero_dirichlet), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, dmask, u);
    VF_KERNEL_CHECK();
}

__global__ void k_enforce_dirichlet(long long numNodes, int N, int ndir, const long long *__restrict__ nodes, const uint8_t *__restrict__ masks, const double *__restrict__ vals, double *__restrict__ u) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndir) return;
    const long long n = nodes[i]; const unsigned dm = masks[i];
    for (int c = 0; c < N; ++c) if ((dm >> c) & 1u) u[c * numNodes + n] = vals[(long long)i * N + c];
}
void launch_enforce_dirichlet(const LaunchCtx &ctx, long long numNodes, int N, int ndir, const long long *nodes, const uint8_t *masks, const double *vals, double *u) {
    if (ndir == 0) return;
    ProfScope ps(ctx, PC_VEC, (double)ndir);
    VF_LAUNCH((k_enforce_dirichlet), (ndir + 255) / 256, 256, 0, ctx.stream, numNodes, N, ndir, nodes, masks, vals, u);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256) k_masked_zero(const __grid_constant__ GridDesc g, double *__restrict__ u, int limit) {
    pdl_prologue();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x)
        if (node_active(g, n, limit)) for (int c = 0; c < g.N; ++c) u[c * g.numNodes + n] = 0.0;
}
void launch_masked_zero(const LaunchCtx &ctx, const GridDesc &g, double *u, int margin) {
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_masked_zero), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, u, g.nActive + margin);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_detached_zero(const __grid_constant__ GridDesc g, double *__restrict__ u) {
    pdl_prologue();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x)
        if (!node_active(g, n, g.nActive)) for (int c = 0; c < g.N; ++c) u[c * g.numNodes + n] = 0.0;
}
void launch_detached_zero(const LaunchCtx &ctx, const GridDesc &g, double *u) {
    if (g.nActive >= g.nn[g.bd]) return;
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_detached_zero), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, u);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_masked_copy(const __grid_constant__ GridDesc g, const double *__restrict__ in, double *__restrict__ out, int limit) {
    pdl_prologue();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x)
        if (node_active(g, n, limit)) for (int c = 0; c < g.N; ++c) out[c * g.numNodes + n] = in[c * g.numNodes + n];
}
void launch_masked_copy(const LaunchCtx &ctx, const GridDesc &g, const double *in, double *out, int margin) {
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_masked_copy), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, in, out, g.nActive + margin);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256) k_masked_dot(const __grid_constant__ GridDesc g, const double *__restrict__ a, const double *__restrict__ b, double *result, double *scratch) {
    pdl_prologue();
    double s = 0.0;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x)
        if (node_active(g, n, g.nActive) && node_owned(g, n)) for (int c = 0; c < g.N; ++c) s = fma(a[c * g.numNodes + n], b[c * g.numNodes + n], s);
    grid_sum(s, scratch, result);
}
void launch_masked_dot(const LaunchCtx &ctx, const GridDesc &g, const double *a, const double *b, double *result, double *scratch) {
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_masked_dot), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, a, b, result, scratch);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256) k_cg_direction(const __grid_constant__ GridDesc g, const double *__restrict__ s, double *__restrict__ d, const double *num, const double *den, int first) {
    pdl_prologue();
    const double beta = first ? 0.0 : (*num / *den);
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x) {
        if (!node_active(g, n, g.nActive)) continue;
        for (int c = 0; c < g.N; ++c) {
            const long long i = c * g.numNodes + n;
            d[i] = first ? s[i] : fma(beta, d[i], s[i]);
        }
    }
}
void launch_cg_direction(const LaunchCtx &ctx, const GridDesc &g, const double *s, double *d, const double *num, const double *den, bool first) {
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_cg_direction), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, s, d, num, den, first ? 1 : 0);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256) k_cg_update(const __grid_constant__ GridDesc g, double *__restrict__ x, const double *__restrict__ d, double *__restrict__ r, const double *__restrict__ Ad,
                                                   const double *rMr, const double *dAd, double *rsq, double *scratch) {
    pdl_prologue();
    const double alpha = *rMr / *dAd;
    double s = 0.0;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < g.numNodes; n += (long long)gridDim.x * blockDim.x) {
        if (!node_active(g, n, g.nActive)) continue;
        const bool own = node_owned(g, n);
        for (int c = 0; c < g.N; ++c) {
            const long long i = c * g.numNodes + n;
            x[i] = fma(alpha, d[i], x[i]);
            const double rn = fma(-alpha, Ad[i], r[i]);
            r[i] = rn;
            if (own) s = fma(rn, rn, s);
        }
    }
    grid_sum(s, scratch, rsq);
}
void launch_cg_update(const LaunchCtx &ctx, const GridDesc &g, double *x, const double *d, double *r, const double *Ad,
                      const double *rMr, const double *dAd, double *rsq, double *scratch) {
    ProfScope ps(ctx, PC_VEC, (double)g.numNodes);
    VF_LAUNCH((k_cg_update), flat_blocks(g.numNodes), 256, 0, ctx.stream, g, x, d, r, Ad, rMr, dAd, rsq, scratch);
    VF_KERNEL_CHECK();
}

size_t reduce_scratch_doubles() { return (size_t)kReduceMaxBlocks + 2; }

// ---------------------------------------------------------------------------
// Plain (unmasked) flat helpers used by the optimization layer
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sum(long long n, const double *__restrict__ x, double *result, double *scratch) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i];
    grid_sum(s, scratch, result);
}
void launch_sum(const LaunchCtx &ctx, long long n, const double *x, double *result, double *scratch) {
    ProfScope ps(ctx, PC_TOPOPT, (double)n);
    k_sum<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, x, result, scratch);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_dot_plain(long long n, const double *__restrict__ a, const double *__restrict__ b, double *result, double *scratch) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fma(a[i], b[i], s);
    grid_sum(s, scratch, result);
}
void launch_dot_plain(const LaunchCtx &ctx, long long n, const double *a, const double *b, double *result, double *scratch) {
    ProfScope ps(ctx, PC_VEC, (double)n);
    k_dot_plain<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, a, b, result, scratch);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_fill(long long n, double v, double *__restrict__ x) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
void launch_fill(const LaunchCtx &ctx, long long n, double v, double *x) {
    ProfScope ps(ctx, PC_OTHER, (double)n);
    k_fill<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, v, x);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_axpy(long long n, double a, const double *__restrict__ x, double *__restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fma(a, x[i], y[i]);
}
void launch_axpy(const LaunchCtx &ctx, long long n, double a, const double *x, double *y) {
    ProfScope ps(ctx, PC_VEC, (double)n);
    k_axpy<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, a, x, y);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_scale(long long n, double a, double *__restrict__ x) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= a;
}
void launch_scale(const LaunchCtx &ctx, long long n, double a, double *x) {
    ProfScope ps(ctx, PC_VEC, (double)n);
    k_scale<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, a, x);
    VF_KERNEL_CHECK();
}

} // namespace vf
