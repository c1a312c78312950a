// vf_filters.cu -- the dimension-changing and layer-sequential density filters (sm_100a).
//
//   UpsampleFilter      TopologyOptimizationFilter.hh:418-523   multilinear upsampling of a vertex grid s -> (s - 1) * factor + 1
//   VertexToCellFilter  TopologyOptimizationFilter.hh:528-598   vertex values -> cell values (average of the 2^N corners)
//   LangelaarFilter     TopologyOptimizationFilter.hh:601-712   self-supporting (overhang) filter, layers along the build direction
// (SmoothingFilter / ProjectionFilter live in vf_top.cu.)  Arrays are flat, row-major over the grid with the last axis fastest, as
// NDVector (NDVector.hh:256-264); 2D grids are embedded as 1 x d0 x d1.
//
// LangelaarFilter notes.  smax / smin constants P = 40, Q = 40 - 1.58, epsilon = 1e-4 (:703-711).  The support of a voxel is what
// NDVector::visitSupportingRegion (NDVector.hh:211-229) visits: the voxel below and its two neighbours along axis 0 -- and, because
// that loop runs over the first N - 1 axes rather than over the axes orthogonal to the build direction, in 3D also the voxel TWO
// layers below and the voxel ITSELF (whose output value is whatever the output array held before this application).  The z
// neighbours are never visited.  This is restated as is: results have to match the reference, not the paper.
#include "vf_internal.cuh"
#include <cmath>

namespace vf {

struct FGrid { int d[3]; long long n; };     // embedded sizes, number of entries
static FGrid fgrid(int N, const int *sizes) {
    FGrid g; g.d[0] = (N == 3) ? sizes[0] : 1; g.d[1] = sizes[N - 2]; g.d[2] = sizes[N - 1];
    g.n = (long long)g.d[0] * g.d[1] * g.d[2];
    return g;
}
__device__ __forceinline__ void funflat(const FGrid &g, long long i, int (&c)[3]) {
    c[2] = (int)(i % g.d[2]); i /= g.d[2]; c[1] = (int)(i % g.d[1]); c[0] = (int)(i / g.d[1]);
}
__device__ __forceinline__ long long fflat(const FGrid &g, int c0, int c1, int c2) { return ((long long)c0 * g.d[1] + c1) * g.d[2] + c2; }

// ---------------------------------------------------------------------------------------------------------------------------------
// UpsampleFilter
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void k_upsample(const __grid_constant__ FGrid gc, const __grid_constant__ FGrid gf, int N, int factor, const double *__restrict__ in, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gf.n) return;
    int nf[3]; funflat(gf, i, nf);
    int c[3] = {0, 0, 0}; double t[3] = {0.0, 0.0, 0.0};
    for (int a = 3 - N; a < 3; ++a) {   // coarse cell holding the fine node (the last cell also owns its upper boundary), local coordinate in [0, 1]
        c[a] = min(nf[a] / factor, gc.d[a] - 2);
        t[a] = double(nf[a] - c[a] * factor) / double(factor);
    }
    double v = 0.0;
    for (int b = 0; b < (1 << N); ++b) {
        double w = 1.0; int q[3] = {c[0], c[1], c[2]};
        for (int a = 3 - N; a < 3; ++a) { const int bit = (b >> (2 - a)) & 1; w *= bit ? t[a] : 1.0 - t[a]; q[a] += bit; }
        v += w * in[fflat(gc, q[0], q[1], q[2])];
    }
    out[i] = v;
}
__global__ void k_upsample_backprop(const __grid_constant__ FGrid gc, const __grid_constant__ FGrid gf, int N, int factor, const double *__restrict__ dout, double *__restrict__ din) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gc.n) return;
    int nc[3]; funflat(gc, i, nc);
    int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, ctr[3] = {0, 0, 0};
    for (int a = 3 - N; a < 3; ++a) {   // fine nodes inside the support of the coarse hat function (:492-500)
        ctr[a] = factor * nc[a]; lo[a] = ctr[a]; hi[a] = ctr[a] + 1;
        if (nc[a] > 0) lo[a] -= factor - 1;
        if (nc[a] < gc.d[a] - 1) hi[a] += factor - 1;
    }
    double s = 0.0;
    for (int x = lo[0]; x < hi[0]; ++x) for (int y = lo[1]; y < hi[1]; ++y) for (int z = lo[2]; z < hi[2]; ++z) {
        double phi = 1.0; const int q[3] = {x, y, z};
        for (int a = 3 - N; a < 3; ++a) phi *= 1.0 - double(abs(q[a] - ctr[a])) / double(factor);
        s += phi * dout[fflat(gf, x, y, z)];
    }
    din[i] = s;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// VertexToCellFilter
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void k_v2c(const __grid_constant__ FGrid gv, const __grid_constant__ FGrid ge, int N, const double *__restrict__ in, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ge.n) return;
    int e[3]; funflat(ge, i, e);
    double s = 0.0;
    for (int b = 0; b < (1 << N); ++b) s += in[fflat(gv, e[0] + ((N == 3) ? ((b >> 2) & 1) : 0), e[1] + ((b >> 1) & 1), e[2] + (b & 1))];
    out[i] = s * ((N == 3) ? 0.125 : 0.25);
}
__global__ void k_v2c_backprop(const __grid_constant__ FGrid gv, const __grid_constant__ FGrid ge, int N, const double *__restrict__ dout, double *__restrict__ din) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gv.n) return;
    int v[3]; funflat(gv, i, v);
    double s = 0.0;
    for (int b = 0; b < (1 << N); ++b) {
        const int e0 = v[0] - ((N == 3) ? ((b >> 2) & 1) : 0), e1 = v[1] - ((b >> 1) & 1), e2 = v[2] - (b & 1);
        if (e0 >= 0 && e0 < ge.d[0] && e1 >= 0 && e1 < ge.d[1] && e2 >= 0 && e2 < ge.d[2]) s += dout[fflat(ge, e0, e1, e2)];
    }
    din[i] = s * ((N == 3) ? 0.125 : 0.25);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// LangelaarFilter
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr double kLgP = 40.0, kLgQ = 40.0 - 1.58, kLgEps = 1e-4;
__device__ __forceinline__ double lg_smin(double x1, double x2) { return 0.5 * (x1 + x2 - sqrt((x1 - x2) * (x1 - x2) + kLgEps) + sqrt(kLgEps)); }
__device__ __forceinline__ double lg_dsmin_dx1(double x1, double x2) { return 0.5 * (1.0 - (x1 - x2) * pow((x1 - x2) * (x1 - x2) + kLgEps, -0.5)); }
__device__ __forceinline__ double lg_dsmin_dx2(double x1, double x2) { return 0.5 * (1.0 + (x1 - x2) * pow((x1 - x2) * (x1 - x2) + kLgEps, -0.5)); }

// Support of voxel (c0, layer, c2) [3D] / (layer-axis = embedded axis 2 in 2D] in visiting order; returns the number of entries.
// bd = embedded build axis (1 in 3D, 2 in 2D); side = embedded axis of the reference's axis 0 (0 in 3D, 1 in 2D).
__device__ __forceinline__ int lg_support(const FGrid &g, int N, const int (&c)[3], long long (&idx)[5]) {
    const int bd = (N == 3) ? 1 : 2, side = (N == 3) ? 0 : 1;
    int q[3] = {c[0], c[1], c[2]};
    q[bd] -= 1;
    int n = 0;
    idx[n++] = fflat(g, q[0], q[1], q[2]);                                                  // voxel below
    q[side] -= 1; if (q[side] >= 0) idx[n++] = fflat(g, q[0], q[1], q[2]);
    q[side] += 2; if (q[side] < g.d[side]) idx[n++] = fflat(g, q[0], q[1], q[2]);
    q[side] -= 1;
    if (N == 3) {                                                                           // d = 1 of the reference's loop is the build axis itself
        q[bd] -= 1; if (q[bd] >= 0) idx[n++] = fflat(g, q[0], q[1], q[2]);                  // two layers below
        q[bd] += 2; if (q[bd] < g.d[bd]) idx[n++] = fflat(g, q[0], q[1], q[2]);             // the voxel itself
    }
    return n;
}
__device__ __forceinline__ void lg_layer_voxel(const FGrid &g, int N, int layer, long long t, int (&c)[3]) {
    if (N == 3) { c[0] = (int)(t / g.d[2]); c[1] = layer; c[2] = (int)(t % g.d[2]); }
    else { c[0] = 0; c[1] = (int)t; c[2] = layer; }
}
__global__ void k_langelaar_layer(const __grid_constant__ FGrid g, int N, int layer, const double *__restrict__ in, double *out, double *__restrict__ smaxCache) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (N == 3) ? (long long)g.d[0] * g.d[2] : g.d[1];
    if (t >= per) return;
    int c[3]; lg_layer_voxel(g, N, layer, t, c);
    const long long i = fflat(g, c[0], c[1], c[2]);
    if (layer == 0) { out[i] = in[i]; return; }                                             // attached to the build platform (:614)
    long long sup[5]; const int ns = lg_support(g, N, c, sup);
    double sum = 0.0;
    for (int k = 0; k < ns; ++k) sum += pow(out[sup[k]], kLgP);                             // smax (:673-677); in 3D sup includes i itself: its previous value
    const double sm = pow(sum, 1.0 / kLgQ);
    smaxCache[i] = sm;
    out[i] = lg_smin(in[i], sm);
}
// S_i = sum over the support of i of filtered^P (the sum inside smaxDerivative, :680-684), for every voxel above layer 0
__global__ void k_langelaar_supsum(const __grid_constant__ FGrid g, int N, const double *__restrict__ filtered, double *__restrict__ S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int c[3]; funflat(g, i, c);
    const int bd = (N == 3) ? 1 : 2;
    if (c[bd] == 0) { S[i] = 0.0; return; }
    long long sup[5]; const int ns = lg_support(g, N, c, sup);
    double sum = 0.0;
    for (int k = 0; k < ns; ++k) sum += pow(filtered[sup[k]], kLgP);
    S[i] = sum;
}
// sminDerivative(vars, i, k) (:687-690) = dsmin_dx2(vars_i, smax_i) * P filtered_k^(P-1) / Q * S_i^(1/Q - 1)
__device__ __forceinline__ double lg_D(long long i, long long k, const double *vars, const double *smaxCache, const double *filtered, const double *S) {
    return lg_dsmin_dx2(vars[i], smaxCache[i]) * (kLgP * pow(filtered[k], kLgP - 1.0) / kLgQ * pow(S[i], 1.0 / kLgQ - 1.0));
}
// Lagrange multipliers of one layer (computeLagrangeMultipliers, :643-661), gather form: lambda_k = in_k + sum over the voxels i of the
// layer above whose support holds k of lambda_i D(i, k).  lam holds the multipliers BEFORE a voxel's own self-term (3D), lamOut after.
__global__ void k_langelaar_lambda_layer(const __grid_constant__ FGrid g, int N, int layer, const double *__restrict__ in, const double *__restrict__ vars,
                                         const double *__restrict__ smaxCache, const double *__restrict__ filtered, const double *__restrict__ S,
                                         double *lam, double *lamOut) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (N == 3) ? (long long)g.d[0] * g.d[2] : g.d[1];
    if (t >= per) return;
    int c[3]; lg_layer_voxel(g, N, layer, t, c);
    const int bd = (N == 3) ? 1 : 2, side = (N == 3) ? 0 : 1;
    const long long k = fflat(g, c[0], c[1], c[2]);
    double l = in[k];
    if (layer < g.d[bd] - 1) {
        // the reference walks the voxels i of the layer above in index order and adds to their supports: i = above-left, above, above-right
        for (int ds = -1; ds <= 1; ++ds) {
            int q[3] = {c[0], c[1], c[2]}; q[bd] += 1; q[side] += ds;
            if (q[side] < 0 || q[side] >= g.d[side]) continue;
            const long long i = fflat(g, q[0], q[1], q[2]);
            l += lam[i] * lg_D(i, k, vars, smaxCache, filtered, S);
        }
    }
    lam[k] = l;
    // 3D: a voxel is in its own support, so after its contributions went out it adds lambda_i D(i, i) to itself (visiting order of
    // NDVector::visitSupportingRegion: itself comes last)
    lamOut[k] = (N == 3 && layer > 0) ? l + l * lg_D(k, k, vars, smaxCache, filtered, S) : l;
}
__global__ void k_langelaar_backprop_out(const __grid_constant__ FGrid g, int N, const double *__restrict__ lamOut, const double *__restrict__ vars,
                                         const double *__restrict__ smaxCache, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int c[3]; funflat(g, i, c);
    const int bd = (N == 3) ? 1 : 2;
    out[i] = (c[bd] == 0) ? lamOut[i] : lamOut[i] * lg_dsmin_dx1(vars[i], smaxCache[i]);   // (:631-633)
}

// ---------------------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

void launch_filter_upsample(const LaunchCtx &ctx, int N, const int *coarseSizes, int factor, const double *in, double *out) {
    int fs[3]; for (int a = 0; a < N; ++a) fs[a] = (coarseSizes[a] - 1) * factor + 1;
    const FGrid gc = fgrid(N, coarseSizes), gf = fgrid(N, fs);
    ProfScope ps(ctx, PC_TOPOPT, (double)gf.n);
    k_upsample<<<blocks_for(gf.n), 256, 0, ctx.stream>>>(gc, gf, N, factor, in, out);
    VF_KERNEL_CHECK();
}
void launch_filter_upsample_backprop(const LaunchCtx &ctx, int N, const int *coarseSizes, int factor, const double *dout, double *din) {
    int fs[3]; for (int a = 0; a < N; ++a) fs[a] = (coarseSizes[a] - 1) * factor + 1;
    const FGrid gc = fgrid(N, coarseSizes), gf = fgrid(N, fs);
    ProfScope ps(ctx, PC_TOPOPT, (double)gf.n);
    k_upsample_backprop<<<blocks_for(gc.n), 256, 0, ctx.stream>>>(gc, gf, N, factor, dout, din);
    VF_KERNEL_CHECK();
}
void launch_filter_v2c(const LaunchCtx &ctx, int N, const int *vertexSizes, const double *in, double *out) {
    int es[3]; for (int a = 0; a < N; ++a) es[a] = vertexSizes[a] - 1;
    const FGrid gv = fgrid(N, vertexSizes), ge = fgrid(N, es);
    ProfScope ps(ctx, PC_TOPOPT, (double)ge.n);
    k_v2c<<<blocks_for(ge.n), 256, 0, ctx.stream>>>(gv, ge, N, in, out);
    VF_KERNEL_CHECK();
}
void launch_filter_v2c_backprop(const LaunchCtx &ctx, int N, const int *vertexSizes, const double *dout, double *din) {
    int es[3]; for (int a = 0; a < N; ++a) es[a] = vertexSizes[a] - 1;
    const FGrid gv = fgrid(N, vertexSizes), ge = fgrid(N, es);
    ProfScope ps(ctx, PC_TOPOPT, (double)gv.n);
    k_v2c_backprop<<<blocks_for(gv.n), 256, 0, ctx.stream>>>(gv, ge, N, dout, din);
    VF_KERNEL_CHECK();
}
// out is in/out (see the header note); smaxCache: one double per voxel
void launch_filter_langelaar(const LaunchCtx &ctx, int N, const int *sizes, const double *in, double *out, double *smaxCache) {
    TraceScope ts("applyLangelaarFilter");                        // TopologyOptimizationFilter.hh:610
    const FGrid g = fgrid(N, sizes);
    const int layers = sizes[1];
    const long long per = g.n / layers;
    ProfScope ps(ctx, PC_TOPOPT, (double)g.n);
    for (int l = 0; l < layers; ++l) {                               // a layer depends on the one below: one launch per layer
        k_langelaar_layer<<<blocks_for(per), 256, 0, ctx.stream>>>(g, N, l, in, out, smaxCache);
        if (l) count_launch();
    }
    VF_KERNEL_CHECK();
}
// scratch: 3 doubles per voxel (S, lambda before / after the self-term); filtered = the output of the last application (m_cachedFiltered)
void launch_filter_langelaar_backprop(const LaunchCtx &ctx, int N, const int *sizes, const double *g_in, const double *vars, const double *filtered,
                                      const double *smaxCache, double *scratch, double *out) {
    TraceScope ts("backpropLangelaarFilter");                     // (:626)
    const FGrid g = fgrid(N, sizes);
    const int layers = sizes[1];
    const long long per = g.n / layers;
    double *S = scratch, *lam = scratch + g.n, *lamOut = scratch + 2 * g.n;
    ProfScope ps(ctx, PC_TOPOPT, (double)g.n);
    k_langelaar_supsum<<<blocks_for(g.n), 256, 0, ctx.stream>>>(g, N, filtered, S);
    for (int l = layers - 1; l >= 0; --l) {
        k_langelaar_lambda_layer<<<blocks_for(per), 256, 0, ctx.stream>>>(g, N, l, g_in, vars, smaxCache, filtered, S, lam, lamOut);
        count_launch();
    }
    k_langelaar_backprop_out<<<blocks_for(g.n), 256, 0, ctx.stream>>>(g, N, lamOut, vars, smaxCache, out);
    count_launch();
    VF_KERNEL_CHECK();
}

} // namespace vf
