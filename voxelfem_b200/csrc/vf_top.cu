// vf_top.cu -- element-wise kernels of the optimization layer (sm_100a).
//
// m_updateYoungModuli (TensorProductSimulator.hh:2055-2060, 2088-2102), compliance gradient
// (:972-1040), elementEnergyDensity (:1057-1073), self-weight loads (:1275-1305),
// SmoothingFilter / ProjectionFilter (TopologyOptimizationFilter.hh:199-225, 328-395),
// OC update (OptimalityCriterion.hh:64-83).
#include "vf_internal.cuh"
#include "vf_reduce.cuh"
#include <vector>
#include <cmath>

namespace vf {

static inline int flat_blocks(long long n) { long long b = (n + 255) / 256; const long long cap = 148LL * 16; return (int)(b < cap ? (b > 0 ? b : 1) : cap); }

__device__ __forceinline__ int elem_layer(const GridDesc &g, long long e) {
    return (g.bd == 2) ? (int)(e % g.ne[2]) : (int)((e / g.ne[2]) % g.ne[1]);
}

__global__ void __launch_bounds__(256)
k_update_moduli(const __grid_constant__ GridDesc g, const double *__restrict__ rho, double *__restrict__ E,
                int law, double E0, double Emin, double gamma, double q, int maskActive) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < g.numElems; e += (long long)gridDim.x * blockDim.x) {
        double v;
        if (maskActive && elem_layer(g, e) >= g.neActive) v = 0.0;
        else {
            const double r = rho[e];
            v = (law == 0) ? (Emin + pow(r, gamma) * (E0 - Emin)) : (Emin + r * (E0 - Emin) / (1.0 + q * (1.0 - r)));
        }
        E[e] = v;
    }
}
void launch_update_moduli(const LaunchCtx &ctx, const GridDesc &g, const double *rho, double *E, int law, double E0, double Emin, double gamma, double q, bool maskActive) {
    ProfScope ps(ctx, PC_TOPOPT, (double)g.numElems);
    k_update_moduli<<<flat_blocks(g.numElems), 256, 0, ctx.stream>>>(g, rho, E, law, E0, Emin, gamma, q, maskActive ? 1 : 0);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256)
k_zero_moduli_layers(const __grid_constant__ GridDesc g, double *__restrict__ E, int layerBegin, int layerEnd) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < g.numElems; e += (long long)gridDim.x * blockDim.x) {
        const int l = elem_layer(g, e);
        if (l >= layerBegin && l < layerEnd) E[e] = 0.0;
    }
}
void launch_zero_moduli_layers(const LaunchCtx &ctx, const GridDesc &g, double *E, int layerBegin, int layerEnd) {
    ProfScope ps(ctx, PC_TOPOPT, (double)g.numElems);
    k_zero_moduli_layers<<<flat_blocks(g.numElems), 256, 0, ctx.stream>>>(g, E, layerBegin, layerEnd);
    VF_KERNEL_CHECK();
}

// u_e^T K0 u_e for one element (thread-local gather of its 2^N nodes)
template<int N>
__device__ __forceinline__ double elem_uKu(const GridDesc &g, const K0Param &K, const double *__restrict__ u, long long e, double (&ue)[Dims<N>::KE]) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE, A0 = Dims<N>::A0;
    int ec[3]; { long long r = e; ec[2] = (int)(r % g.ne[2]); r /= g.ne[2]; ec[1] = (int)(r % g.ne[1]); ec[0] = (int)(r / g.ne[1]); }
    long long n0 = 0;
    #pragma unroll
    for (int a = A0; a < 3; ++a) n0 += (long long)ec[a] * g.ns[a];
    #pragma unroll
    for (int m = 0; m < NPE; ++m) {
        long long off = 0;
        #pragma unroll
        for (int a = A0; a < 3; ++a) off += ((m >> (2 - a)) & 1) ? g.ns[a] : 0;
        #pragma unroll
        for (int c = 0; c < N; ++c) ue[N * m + c] = u[c * g.numNodes + n0 + off];
    }
    double uKu = 0.0;
    #pragma unroll
    for (int a = 0; a < KE; ++a) {
        double s = 0.0;
        #pragma unroll
        for (int b = 0; b < KE; ++b) s = fma(K.v[a * KE + b], ue[b], s);
        uKu = fma(ue[a], s, uKu);
    }
    return uKu;
}

template<int N>
__global__ void __launch_bounds__(128)
k_compliance_gradient(const __grid_constant__ GridDesc g, const __grid_constant__ K0Param K, const double *__restrict__ u,
                      const double *__restrict__ rho, double *__restrict__ out, int law, double E0, double Emin, double gamma, double q,
                      double g0, double g1, double g2, double elemVol, int accumulate) {
    constexpr int NPE = Dims<N>::NPE, KE = Dims<N>::KE;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.numElems) return;
    if (elem_layer(g, e) >= g.neActive) { if (!accumulate) out[e] = 0.0; return; } // masked elements do not contribute (:984-987, :1016)
    double ue[KE];
    const double uKu = elem_uKu<N>(g, K, u, e, ue);
    const double r = rho[e];
    double val;
    if (law == 0) val = -0.5 * gamma * pow(r, gamma - 1.0) * (E0 - Emin) * uKu;
    else { const double den = 1.0 + q * (1.0 - r); val = -0.5 * (1.0 + q) * (E0 - Emin) / (den * den) * uKu; }
    const double grav[3] = {g0, g1, g2};
    if (g0 != 0.0 || g1 != 0.0 || g2 != 0.0) {
        const double intPhi = 1.0 / NPE;
        #pragma unroll
        for (int m = 0; m < NPE; ++m) {
            double gd = 0.0;
            #pragma unroll
            for (int c = 0; c < N; ++c) gd = fma(grav[c], ue[N * m + c], gd);
            val += intPhi * gd * elemVol;
        }
    }
    if (accumulate) out[e] += val; else out[e] = val;
}
void launch_compliance_gradient(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *rho, double *out,
                                int law, double E0, double Emin, double gamma, double q, const double *gravity, double elemVol, bool accumulate) {
    ProfScope ps(ctx, PC_TOPOPT, (double)g.numElems);
    const unsigned blocks = (unsigned)((g.numElems + 127) / 128);
    if (g.N == 3) k_compliance_gradient<3><<<blocks, 128, 0, ctx.stream>>>(g, K, u, rho, out, law, E0, Emin, gamma, q, gravity[0], gravity[1], gravity[2], elemVol, accumulate ? 1 : 0);
    else          k_compliance_gradient<2><<<blocks, 128, 0, ctx.stream>>>(g, K, u, rho, out, law, E0, Emin, gamma, q, gravity[0], gravity[1], 0.0, elemVol, accumulate ? 1 : 0);
    VF_KERNEL_CHECK();
}

template<int N>
__global__ void __launch_bounds__(128)
k_energy_density(const __grid_constant__ GridDesc g, const __grid_constant__ K0Param K, const double *__restrict__ u, const double *__restrict__ E, double *__restrict__ out) {
    constexpr int KE = Dims<N>::KE;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.numElems) return;
    double ue[KE];
    out[e] = 0.5 * E[e] * elem_uKu<N>(g, K, u, e, ue);
}
void launch_energy_density(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E, double *out) {
    ProfScope ps(ctx, PC_TOPOPT, (double)g.numElems);
    const unsigned blocks = (unsigned)((g.numElems + 127) / 128);
    if (g.N == 3) k_energy_density<3><<<blocks, 128, 0, ctx.stream>>>(g, K, u, E, out);
    else          k_energy_density<2><<<blocks, 128, 0, ctx.stream>>>(g, K, u, E, out);
    VF_KERNEL_CHECK();
}

// f_n += sign * gravity * (1/2^N) * elemVol * sum_{incident e, layer(e) in [layerBegin, layerEnd)} rho_e
// (buildLoadVector :1275-1286; addLayerRemovalDeltaLoadVector :1292-1305 with sign = -1)
template<int N>
__global__ void __launch_bounds__(256)
k_self_weight(const __grid_constant__ GridDesc g, const double *__restrict__ rho, double g0, double g1, double g2, double elemVol,
              double *__restrict__ f, int layerBegin, int layerEnd, double sign) {
    constexpr int NPE = Dims<N>::NPE, A0 = Dims<N>::A0;
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int c1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = blockIdx.z * blockDim.z + threadIdx.z;
    if (c2 >= g.nn[2] || c1 >= g.nn[1] || c0 >= g.nn[0]) return;
    const int cc[3] = {c0, c1, c2};
    const long long n = (long long)c0 * g.ns[0] + (long long)c1 * g.ns[1] + c2;
    double s = 0.0;
    #pragma unroll
    for (int e = 0; e < NPE; ++e) {
        bool ok = true; long long ei = 0;
        #pragma unroll
        for (int a = A0; a < 3; ++a) {
            const int ec = cc[a] - ((e >> (2 - a)) & 1);
            ok = ok && ec >= 0 && ec < g.ne[a];
            if (a == g.bd) ok = ok && ec >= layerBegin && ec < layerEnd;
            ei += (long long)ec * g.es[a];
        }
        if (ok) s += rho[ei];
    }
    if (s != 0.0) {
        const double w = sign * s * elemVol / NPE;
        const double grav[3] = {g0, g1, g2};
        #pragma unroll
        for (int c = 0; c < N; ++c) if (grav[c] != 0.0) f[c * g.numNodes + n] += grav[c] * w;
    }
}
void launch_self_weight_load(const LaunchCtx &ctx, const GridDesc &g, const double *rho, const double *gravity, double elemVol, double *f, int layerBegin, int layerEnd, double sign) {
    ProfScope ps(ctx, PC_TOPOPT, (double)g.numNodes);
    dim3 b = (g.N == 3) ? dim3(32, 4, 2) : dim3(32, 8, 1);
    dim3 gr((g.nn[2] + b.x - 1) / b.x, (g.nn[1] + b.y - 1) / b.y, (g.nn[0] + b.z - 1) / b.z);
    if (g.N == 3) k_self_weight<3><<<gr, b, 0, ctx.stream>>>(g, rho, gravity[0], gravity[1], gravity[2], elemVol, f, layerBegin, layerEnd, sign);
    else          k_self_weight<2><<<gr, b, 0, ctx.stream>>>(g, rho, gravity[0], gravity[1], 0.0, elemVol, f, layerBegin, layerEnd, sign);
    VF_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------
// Density filters
// ---------------------------------------------------------------------------
struct FilterDesc { int N; int sz[3]; int radius; int type; double invTotalWeight; };

__device__ __forceinline__ int reflect_index(int i, int s) { // -2,-1,0,1 -> 1,0,0,1  (TopologyOptimizationFilter.hh:339-345)
    while (i < 0 || i >= s) { if (i >= s) i = 2 * s - i - 1; if (i < 0) i = -i - 1; }
    return i;
}

// out_e = sum_{offset in [-r,r]^N, w > 0} w(offset) in[reflect(e + offset)] / sum w
// One thread per element, lanes along the fastest axis; weights staged in shared memory.
__global__ void __launch_bounds__(256)
k_filter_smooth(const __grid_constant__ FilterDesc fd, const double *__restrict__ in, double *__restrict__ out) {
    extern __shared__ double s_w[];
    const int w1 = 2 * fd.radius + 1;
    const int nOff = (fd.N == 3) ? w1 * w1 * w1 : w1 * w1;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nth = blockDim.x * blockDim.y * blockDim.z;
    for (int o = tid; o < nOff; o += nth) {
        int r = o; double sq = 0.0;
        for (int a = 0; a < fd.N; ++a) { const int q = r % w1 - fd.radius; r /= w1; sq += (double)q * q; }
        s_w[o] = (fd.type == 1) ? ((double)(fd.radius + 1) - sqrt(sq)) : 1.0;
    }
    __syncthreads();
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int c1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = blockIdx.z * blockDim.z + threadIdx.z;
    if (c2 >= fd.sz[2] || c1 >= fd.sz[1] || c0 >= fd.sz[0]) return;
    double acc = 0.0;
    const int r0 = (fd.N == 3) ? fd.radius : 0;
    for (int d0 = -r0; d0 <= r0; ++d0) {
        const int q0 = (fd.N == 3) ? reflect_index(c0 + d0, fd.sz[0]) : 0;
        for (int d1 = -fd.radius; d1 <= fd.radius; ++d1) {
            const int q1 = reflect_index(c1 + d1, fd.sz[1]);
            const long long rowBase = ((long long)q0 * fd.sz[1] + q1) * fd.sz[2];
            const int wBase = ((fd.N == 3 ? (d0 + fd.radius) * w1 : 0) + (d1 + fd.radius)) * w1 + fd.radius;
            for (int d2 = -fd.radius; d2 <= fd.radius; ++d2) {
                const double w = s_w[wBase + d2];
                if (w <= 0.0) continue;
                acc = fma(w, in[rowBase + reflect_index(c2 + d2, fd.sz[2])], acc);
            }
        }
    }
    out[((long long)c0 * fd.sz[1] + c1) * fd.sz[2] + c2] = acc * fd.invTotalWeight;
}
// 3D tiled variant: a block stages the (8 + 2r) x (4 + 2r) x (32 + 2r) neighbourhood of its 8 x 4 x 32 element tile in
// shared memory (reflection applied while loading, TopologyOptimizationFilter.hh:339-345) and every thread then sums the
// non-zero taps for 4 elements with immediate tile offsets: one LDS + one FMA per tap instead of three reflections,
// a weight test and a global load.  taps: [offset in tile | weight] pairs in the reference's row-major offset order.
constexpr int kFtX = 8, kFtY = 4, kFtZ = 32;
struct FilterTap { int off; int pad; double w; };
__global__ void __launch_bounds__(256)
k_filter_smooth3_tiled(const __grid_constant__ FilterDesc fd, const FilterTap *__restrict__ taps, int ntaps,
                       const double *__restrict__ in, double *__restrict__ out) {
    extern __shared__ double s_t[];
    const int r = fd.radius, uy = kFtY + 2 * r, uz = kFtZ + 2 * r, ux = kFtX + 2 * r;
    const int tid = threadIdx.x + 32 * (threadIdx.y + 4 * threadIdx.z);
    const int x0 = blockIdx.z * kFtX, y0 = blockIdx.y * kFtY, z0 = blockIdx.x * kFtZ;
    for (int i = tid; i < ux * uy * uz; i += 256) {
        const int zi = i % uz; const int q = i / uz; const int yi = q % uy, xi = q / uy;
        const int gx = reflect_index(x0 - r + xi, fd.sz[0]), gy = reflect_index(y0 - r + yi, fd.sz[1]), gz = reflect_index(z0 - r + zi, fd.sz[2]);
        s_t[i] = in[((long long)gx * fd.sz[1] + gy) * fd.sz[2] + gz];
    }
    __syncthreads();
    const int cz = z0 + threadIdx.x, cy = y0 + threadIdx.y;
    if (cz >= fd.sz[2] || cy >= fd.sz[1]) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // element (x0 + tz + 2k, cy, cz), k = 0..3 -> tile centre ((tz + 2k + r) * uy + ty + r) * uz + tx + r
    const double *base = s_t + ((threadIdx.z + r) * uy + threadIdx.y + r) * uz + threadIdx.x + r;
    const int kstride = 2 * uy * uz;
    for (int t = 0; t < ntaps; ++t) {
        const int off = taps[t].off; const double w = taps[t].w;
        #pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fma(w, base[off + k * kstride], acc[k]);
    }
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cx = x0 + threadIdx.z + 2 * k;
        if (cx < fd.sz[0]) out[((long long)cx * fd.sz[1] + cy) * fd.sz[2] + cz] = acc[k] * fd.invTotalWeight;
    }
}
struct FilterTapCache { int radius = -1, type = -1, n = 0, device = -1; FilterTap *dev = nullptr; };
static const FilterTap *filter_taps3(int radius, int type, int &ntaps, cudaStream_t stream) {
    static FilterTapCache cache[32];          // keyed by (device, radius, type): the table lives in that device's memory
    static std::mutex cacheMutex;
    std::lock_guard<std::mutex> lock(cacheMutex);
    const int device = current_device();
    for (auto &c : cache) if (c.radius == radius && c.type == type && c.device == device) { ntaps = c.n; return c.dev; }
    FilterTapCache *slot = nullptr;
    for (auto &c : cache) if (c.radius < 0) { slot = &c; break; }
    if (!slot) return nullptr;
    const int w1 = 2 * radius + 1, uy = kFtY + 2 * radius, uz = kFtZ + 2 * radius;
    std::vector<FilterTap> h;
    for (int d0 = -radius; d0 <= radius; ++d0) for (int d1 = -radius; d1 <= radius; ++d1) for (int d2 = -radius; d2 <= radius; ++d2) {
        const double w = (type == 1) ? ((double)(radius + 1) - std::sqrt((double)(d0 * d0 + d1 * d1 + d2 * d2))) : 1.0;
        if (w > 0.0) h.push_back(FilterTap{(d0 * uy + d1) * uz + d2, 0, w});
    }
    (void)w1;
    VF_CUDA(cudaMalloc(&slot->dev, h.size() * sizeof(FilterTap)));
    VF_CUDA(cudaMemcpyAsync(slot->dev, h.data(), h.size() * sizeof(FilterTap), cudaMemcpyHostToDevice, stream));
    VF_CUDA(cudaStreamSynchronize(stream));
    slot->radius = radius; slot->type = type; slot->n = (int)h.size(); slot->device = device;
    ntaps = slot->n;
    return slot->dev;
}

void launch_filter_smooth(const LaunchCtx &ctx, int N, const int *sizes, int radius, int type, const double *in, double *out) {
    FilterDesc fd; fd.N = N; fd.radius = radius; fd.type = type;
    fd.sz[0] = (N == 3) ? sizes[0] : 1; fd.sz[1] = sizes[N - 2]; fd.sz[2] = sizes[N - 1];
    const int w1 = 2 * radius + 1;
    double tot = 0.0;
    const int nOff = (N == 3) ? w1 * w1 * w1 : w1 * w1;
    for (int o = 0; o < nOff; ++o) {
        int r = o; double sq = 0.0;
        for (int a = 0; a < N; ++a) { const int q = r % w1 - radius; r /= w1; sq += (double)q * q; }
        const double w = (type == 1) ? ((double)(radius + 1) - std::sqrt(sq)) : 1.0;
        if (w > 0) tot += w;
    }
    // NOTE: the reference divides by the accumulated weight; multiplying by its reciprocal differs by <= 1 ulp.
    fd.invTotalWeight = 1.0 / tot;
    ProfScope ps(ctx, PC_TOPOPT, (double)fd.sz[0] * fd.sz[1] * fd.sz[2]);
    if (N == 3 && radius >= 1 && radius <= 4) {
        int ntaps = 0;
        const FilterTap *taps = filter_taps3(radius, type, ntaps, ctx.stream);
        if (taps) {
            const size_t tile = (size_t)(kFtX + 2 * radius) * (kFtY + 2 * radius) * (kFtZ + 2 * radius) * sizeof(double);
            static PerDeviceFlags attr;
            if (first_use_on_device(attr)) VF_CUDA(cudaFuncSetAttribute(k_filter_smooth3_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            dim3 bt(32, 4, 2), gt((fd.sz[2] + kFtZ - 1) / kFtZ, (fd.sz[1] + kFtY - 1) / kFtY, (fd.sz[0] + kFtX - 1) / kFtX);
            k_filter_smooth3_tiled<<<gt, bt, tile, ctx.stream>>>(fd, taps, ntaps, in, out);
            VF_KERNEL_CHECK();
            return;
        }
    }
    dim3 b = (N == 3) ? dim3(32, 4, 2) : dim3(32, 8, 1);
    dim3 gr((fd.sz[2] + b.x - 1) / b.x, (fd.sz[1] + b.y - 1) / b.y, (fd.sz[0] + b.z - 1) / b.z);
    const size_t smem = (size_t)nOff * sizeof(double);
    if (smem > 48 * 1024) throw std::runtime_error("SmoothingFilter radius too large");
    k_filter_smooth<<<gr, b, smem, ctx.stream>>>(fd, in, out);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256) k_filter_project(long long n, double beta, double th, const double *__restrict__ in, double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (th + tanh(beta * (in[i] - 0.5))) / (2.0 * th);
}
void launch_filter_project(const LaunchCtx &ctx, long long n, double beta, const double *in, double *out) {
    ProfScope ps(ctx, PC_TOPOPT, (double)n);
    k_filter_project<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, beta, std::tanh(0.5 * beta), in, out);
    VF_KERNEL_CHECK();
}
__global__ void __launch_bounds__(256) k_filter_project_backprop(long long n, double beta, double scale, const double *__restrict__ in, const double *__restrict__ vars, double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double t = tanh(beta * (vars[i] - 0.5));
        out[i] = in[i] * (1.0 - t * t) * scale;
    }
}
void launch_filter_project_backprop(const LaunchCtx &ctx, long long n, double beta, const double *in, const double *vars, double *out) {
    ProfScope ps(ctx, PC_TOPOPT, (double)n);
    k_filter_project_backprop<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, beta, 1.0 / (2.0 * std::tanh(0.5 * beta) / beta), in, vars, out);
    VF_KERNEL_CHECK();
}

__global__ void __launch_bounds__(256) k_oc_update(long long n, const double *__restrict__ x0, const double *__restrict__ dJ, const double *__restrict__ dc,
                                                   double lambda, double m, double p, double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = x0[i];
        // scalar-path semantics of OptimalityCriterion.hh:76-81: NaN candidates keep x0, +-inf are clamped
        const double raw = x * pow(dJ[i] / (dc[i] * lambda), p);
        const double res = isnan(raw) ? x : fmin(fmax(fmin(fmax(raw, x - m), x + m), 0.0), 1.0);
        out[i] = res;
    }
}
void launch_oc_update(const LaunchCtx &ctx, long long n, const double *x0, const double *dJ, const double *dc, double lambda, double m, double p, double *out) {
    ProfScope ps(ctx, PC_TOPOPT, (double)n);
    k_oc_update<<<flat_blocks(n), 256, 0, ctx.stream>>>(n, x0, dJ, dc, lambda, m, p, out);
    VF_KERNEL_CHECK();
}

} // namespace vf
