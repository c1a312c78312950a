// vf_mma.cu -- Method of Moving Asymptotes (MMA / GCMMA) with all O(n) work on the GPU.
//
// Restates MMA::step and MMA::Subproblem (MethodOfMovingAsymptotes.hh:28-469).  Every loop over the n design variables
// of the reference (tbb::parallel_for / parallel_reduce) is one fused CUDA kernel here; the (m+1)-sized algebra of the
// primal-dual interior-point solver (m = number of constraints, 1 in the reference's drivers) stays on the host.
// Where the reference stores derived arrays (u - x, (u - x)^2, p.lambda, q.lambda, dpsi/dx, G) these kernels recompute them
// from x, l, u, p, q in registers: the kernels are HBM-bound streams and the recomputation is free.
// Reductions are deterministic (fixed grid, per-block partials combined in index order).
#include "vf_internal.cuh"
#include "../../include/voxelfem_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace vf {
namespace {

constexpr int kMaxM = VF_MMA_MAX_CONSTRAINTS;
constexpr int kBlocks = 148 * 4, kThreads = 256;       // persistent grid: 4 CTAs per SM
constexpr int kMaxRed = kMaxM * (kMaxM + 1) / 2 + kMaxM + 2;

struct Ptrs {   // device arrays, all of length n unless noted
    long long n; int m;
    const double *xmin, *xmax;
    double *xh[3];                 // x^k, x^{k-1}, x^{k-2}
    double *l, *u, *alpha, *beta;
    double *df, *p, *q;            // (m+1) x n, row-major
    double *x, *xi, *eta, *dx, *dxi, *deta, *Dx;   // subproblem state
};
struct Small { double v[kMaxM + 1]; };
enum RedOp { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2 };
struct RedSpec { int count; unsigned char op[kMaxRed]; };

__device__ __forceinline__ double red_combine(double a, double b, int op) { return op == RED_SUM ? a + b : (op == RED_MIN ? fmin(a, b) : fmax(a, b)); }
__device__ __forceinline__ double red_identity(int op) { return op == RED_SUM ? 0.0 : (op == RED_MIN ? INFINITY : -INFINITY); }

// Block-level combine of R per-thread values; partials[r * gridDim.x + blockIdx.x] receives the block's value.
template<int R>
__device__ __forceinline__ void block_reduce_store(double (&v)[R], const RedSpec &spec, double *partials) {
    __shared__ double s_part[kThreads / 32][R];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r >= spec.count) break;
        double t = v[r];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = red_combine(t, __shfl_down_sync(0xffffffffu, t, o), spec.op[r]);
        if (lane == 0) s_part[warp][r] = t;
    }
    __syncthreads();
    if (threadIdx.x < spec.count) {
        const int r = threadIdx.x;
        double t = s_part[0][r];
        for (int w = 1; w < kThreads / 32; ++w) t = red_combine(t, s_part[w][r], spec.op[r]);
        partials[(size_t)r * gridDim.x + blockIdx.x] = t;
    }
}
// One warp per output: combine the per-block partials in index order.
__global__ void k_mma_finish(const double *partials, int nblocks, RedSpec spec, double *out) {
    const int r = blockIdx.x;
    const int op = spec.op[r];
    double t = red_identity(op);
    for (int b = threadIdx.x; b < nblocks; b += 32) t = red_combine(t, partials[(size_t)r * nblocks + b], op);
    // lanes hold strided partial results; combine lanes in a fixed order
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = red_combine(t, __shfl_down_sync(0xffffffffu, t, o), op);
    if (threadIdx.x == 0) out[r] = t;
}

// Asymptotes and move limits (MethodOfMovingAsymptotes.hh:66-91)
__global__ void k_mma_asymptotes(Ptrs P, int outerIter) {
    const double asyinit = 0.5, albefa = 0.1, move = 0.5;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double x = P.xh[0][j], xd = P.xmax[j] - P.xmin[j];
        double l, u;
        if (outerIter <= 2) { l = x - asyinit * xd; u = x + asyinit * xd; }
        else {
            const double x1 = P.xh[1][j], x2 = P.xh[2][j];
            const double gam = ((x - x1) * (x1 - x2) > 0) ? 1.2 : 0.7;
            l = fmax(fmin(x - gam * (x1 - P.l[j]), x - 0.01 * xd), x - 10 * xd);
            u = fmin(fmax(x + gam * (P.u[j] - x1), x + 0.01 * xd), x + 10 * xd);
        }
        P.l[j] = l; P.u[j] = u;
        P.alpha[j] = fmax(fmax(P.xmin[j], l + albefa * (x - l)), x - move * xd);
        P.beta[j]  = fmin(fmin(P.xmax[j], u - albefa * (u - x)), x + move * xd);
    }
}
// p, q from the split gradient (:108-114 / :123-129) and the row sums g_i(x^k) needed for r = f(x^k) - g(x^k) (:115, :130)
template<int M>
__global__ void k_mma_pq(Ptrs P, Small rho, RedSpec spec, double *partials) {
    double acc[M + 1];
    #pragma unroll
    for (int i = 0; i <= M; ++i) acc[i] = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double x = P.xh[0][j], xd = P.xmax[j] - P.xmin[j];
        const double umx = P.u[j] - x, xml = x - P.l[j], umx2 = umx * umx, xml2 = xml * xml;
        #pragma unroll
        for (int i = 0; i <= M; ++i) {
            const double g = P.df[(size_t)i * P.n + j], dp = fmax(g, 0.0), dm = fmax(-g, 0.0), rod = rho.v[i] / xd;
            const double p = umx2 * (1.001 * dp + 0.001 * dm + rod), q = xml2 * (0.001 * dp + 1.001 * dm + rod);
            P.p[(size_t)i * P.n + j] = p; P.q[(size_t)i * P.n + j] = q;
            acc[i] += p / umx + q / xml;
        }
    }
    block_reduce_store<M + 1>(acc, spec, partials);
}
// next_rho, first inner iteration (:138-145): rho_i = max(0.1/n * sum_j |df_ij| xdiff_j, 1e-6) (the max is applied on the host)
template<int M>
__global__ void k_mma_rho0(Ptrs P, RedSpec spec, double *partials) {
    double acc[M + 1];
    #pragma unroll
    for (int i = 0; i <= M; ++i) acc[i] = 0.0;
    const double w = 0.1 / double(P.n);
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double xd = P.xmax[j] - P.xmin[j];
        #pragma unroll
        for (int i = 0; i <= M; ++i) acc[i] += w * fabs(P.df[(size_t)i * P.n + j]) * xd;
    }
    block_reduce_store<M + 1>(acc, spec, partials);
}
// Row sums g_i(x) at the subproblem point x (sub_g_eval(true), :160-171) and the GCMMA distance d (:146-149):
// out[0..M] = g_i(x), out[M+1] = sum_j (u-l)(x - x^k)^2 / ((u-x)(x-l) xdiff)
template<int M>
__global__ void k_mma_geval(Ptrs P, RedSpec spec, double *partials) {
    double acc[M + 2];
    #pragma unroll
    for (int i = 0; i <= M + 1; ++i) acc[i] = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double x = P.x[j], u = P.u[j], l = P.l[j], umx = u - x, xml = x - l, dxk = x - P.xh[0][j];
        #pragma unroll
        for (int i = 0; i <= M; ++i) acc[i] += P.p[(size_t)i * P.n + j] / umx + P.q[(size_t)i * P.n + j] / xml;
        acc[M + 1] += (u - l) * dxk * dxk / (umx * xml * (P.xmax[j] - P.xmin[j]));
    }
    block_reduce_store<M + 2>(acc, spec, partials);
}
// Subproblem::init_vars (:281-313): x = (alpha+beta)/2, xi, eta
__global__ void k_mma_init(Ptrs P) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double a = P.alpha[j], b = P.beta[j], x = 0.5 * (a + b);
        P.x[j] = x; P.xi[j] = fmax(1.0 / (x - a), 1.0); P.eta[j] = fmax(1.0 / (b - x), 1.0);
    }
}
// solve_for_newton_direction, first half (:317-349): dx_pre, Dx and the reductions  G Dx^-1 G^T (upper triangle, row-major
// packed) followed by  G Dx^-1 dx_pre  (M values)
template<int M>
__global__ void k_mma_dir1(Ptrs P, double eps, Small lam, RedSpec spec, double *partials) {
    constexpr int R = M * (M + 1) / 2 + M;
    double acc[R];
    #pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double x = P.x[j], umx = P.u[j] - x, xml = x - P.l[j], umx2 = umx * umx, xml2 = xml * xml;
        const double xa = x - P.alpha[j], bx = P.beta[j] - x;
        double plam = P.p[j], qlam = P.q[j], G[M];
        #pragma unroll
        for (int i = 0; i < M; ++i) {
            const double pi = P.p[(size_t)(i + 1) * P.n + j], qi = P.q[(size_t)(i + 1) * P.n + j];
            plam += pi * lam.v[i]; qlam += qi * lam.v[i];
            G[i] = pi / umx2 - qi / xml2;
        }
        const double dpsi = plam / umx2 - qlam / xml2;
        const double dxp = dpsi - eps / xa + eps / bx;
        const double Dx = 2 * plam / (umx * umx2) + 2 * qlam / (xml * xml2) + P.xi[j] / xa + P.eta[j] / bx;
        P.dx[j] = dxp; P.Dx[j] = Dx;
        int k = 0;
        #pragma unroll
        for (int ci = 0; ci < M; ++ci) {
            #pragma unroll
            for (int cj = ci; cj < M; ++cj) acc[k++] += G[ci] * (G[cj] / Dx);
        }
        #pragma unroll
        for (int i = 0; i < M; ++i) acc[k++] += G[i] * (dxp / Dx);
    }
    block_reduce_store<R>(acc, spec, partials);
}
// solve_for_newton_direction, second half (:354-358) + the step-length ratio test over the n-sized variables (:385-394)
template<int M>
__global__ void k_mma_dir2(Ptrs P, double eps, Small dlam, RedSpec spec, double *partials) {
    double acc[1] = {INFINITY};
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        const double x = P.x[j], umx = P.u[j] - x, xml = x - P.l[j], umx2 = umx * umx, xml2 = xml * xml;
        const double xa = x - P.alpha[j], bx = P.beta[j] - x, xi = P.xi[j], eta = P.eta[j];
        double gl = 0.0;
        #pragma unroll
        for (int i = 0; i < M; ++i) gl += (P.p[(size_t)(i + 1) * P.n + j] / umx2 - P.q[(size_t)(i + 1) * P.n + j] / xml2) * dlam.v[i];
        const double dx = -(P.dx[j] + gl) / P.Dx[j];
        const double dxi = -xi + eps / xa - xi * dx / xa, deta = -eta + eps / bx + eta * dx / bx;
        P.dx[j] = dx; P.dxi[j] = dxi; P.deta[j] = deta;
        acc[0] = fmin(fmin(acc[0], dx / xa), fmin(dx / (-bx), fmin(dxi / xi, deta / eta)));
    }
    block_reduce_store<1>(acc, spec, partials);
}
// newton_step on (x, xi, eta) (:427-431) fused with the re-evaluation of squared_residual / KKT_inf_norm (:261-279, :399-424):
// out[0..M-1] = gvec (constraint rows of sub_g_eval), out[M] = |eq_a|^2 + |eq_e|^2 + |eq_f|^2, out[M+1] = their max-norm
template<int M>
__global__ void k_mma_step_eval(Ptrs P, double t, double eps, Small lam, RedSpec spec, double *partials) {
    double acc[M + 2];
    #pragma unroll
    for (int i = 0; i < M + 1; ++i) acc[i] = 0.0;
    acc[M + 1] = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += (long long)gridDim.x * blockDim.x) {
        double x = P.x[j], xi = P.xi[j], eta = P.eta[j];
        if (t != 0.0) { x += t * P.dx[j]; xi += t * P.dxi[j]; eta += t * P.deta[j]; P.x[j] = x; P.xi[j] = xi; P.eta[j] = eta; }
        const double umx = P.u[j] - x, xml = x - P.l[j], umx2 = umx * umx, xml2 = xml * xml;
        double plam = P.p[j], qlam = P.q[j];
        #pragma unroll
        for (int i = 0; i < M; ++i) {
            const double pi = P.p[(size_t)(i + 1) * P.n + j], qi = P.q[(size_t)(i + 1) * P.n + j];
            plam += pi * lam.v[i]; qlam += qi * lam.v[i];
            acc[i] += pi / umx + qi / xml;
        }
        const double ea = plam / umx2 - qlam / xml2 - xi + eta, ee = xi * (x - P.alpha[j]) - eps, ef = eta * (P.beta[j] - x) - eps;
        acc[M] += ea * ea + ee * ee + ef * ef;
        acc[M + 1] = fmax(acc[M + 1], fmax(fabs(ea), fmax(fabs(ee), fabs(ef))));
    }
    block_reduce_store<M + 2>(acc, spec, partials);
}

template<class T> struct Buf {
    T *p = nullptr; size_t n = 0;
    ~Buf() { if (p) cudaFree(p); }
    void alloc(size_t cnt) { if (p) cudaFree(p); p = nullptr; n = cnt; if (cnt) { VF_CUDA(cudaMalloc(&p, cnt * sizeof(T))); VF_CUDA(cudaMemset(p, 0, cnt * sizeof(T))); VF_CUDA(cudaStreamSynchronize(0)); } }   // complete before the (non-blocking) optimizer stream touches it
};

} // namespace
} // namespace vf

using namespace vf;

struct vf_mma {
    long long n = 0; int m = 0;
    Buf<double> xmin, xmax, xh[3], l, u, alpha, beta, df, p, q, x, xi, eta, dx, dxi, deta, Dx, partials, results;
    int nhist = 0, outerIter = 0, innerIter = 0; bool gcmma = false;
    std::vector<double> a, c, d, rho, r, fcur, diff;
    const double a0 = 1, raa0 = 1e-5;                       // MethodOfMovingAsymptotes.hh:196, 201
    struct Vars { double z = 1, zeta = 1; std::vector<double> y, lam, mu, s; } data, delta;
    std::vector<double> gvec;
    cudaStream_t stream = nullptr; LaunchCtx ctx;
    double *hostRes = nullptr, *hostX = nullptr, *hostDf = nullptr;   // pinned
    long long newtonIters = 0;
    ~vf_mma() { if (hostRes) cudaFreeHost(hostRes); if (hostX) cudaFreeHost(hostX); if (hostDf) cudaFreeHost(hostDf); if (stream) cudaStreamDestroy(stream); }

    Ptrs ptrs() {
        Ptrs P; P.n = n; P.m = m; P.xmin = xmin.p; P.xmax = xmax.p;
        for (int i = 0; i < 3; ++i) P.xh[i] = xh[i].p;
        P.l = l.p; P.u = u.p; P.alpha = alpha.p; P.beta = beta.p; P.df = df.p; P.p = p.p; P.q = q.p;
        P.x = x.p; P.xi = xi.p; P.eta = eta.p; P.dx = dx.p; P.dxi = dxi.p; P.deta = deta.p; P.Dx = Dx.p;
        return P;
    }
    static RedSpec spec(int count, int op = RED_SUM) { RedSpec s; s.count = count; for (int i = 0; i < kMaxRed; ++i) s.op[i] = (unsigned char)op; return s; }
    static Small small(const std::vector<double> &v) { Small s; std::memset(&s, 0, sizeof(s)); for (size_t i = 0; i < v.size(); ++i) s.v[i] = v[i]; return s; }
    // finish a reduction: combine the block partials and read the `count` results on the host
    const double *finish(const RedSpec &s) {
        ProfScope ps(ctx, PC_TOPOPT, 0);
        k_mma_finish<<<s.count, 32, 0, stream>>>(partials.p, kBlocks, s, results.p);
        VF_KERNEL_CHECK();
        VF_CUDA(cudaMemcpyAsync(hostRes, results.p, sizeof(double) * s.count, cudaMemcpyDeviceToHost, stream));
        VF_CUDA(cudaStreamSynchronize(stream));
        return hostRes;
    }
    void pushHistory(const double *srcDev) { // FixedSizeDeque<AXd>{3}::addToHistory: the stalest buffer becomes the new front
        std::swap(xh[2].p, xh[1].p); std::swap(xh[1].p, xh[0].p);
        VF_CUDA(cudaMemcpyAsync(xh[0].p, srcDev, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
        nhist = std::min(nhist + 1, 3);
    }
};

namespace {

#define VF_MMA_DISPATCH(M_RUNTIME, CALL) \
    switch (M_RUNTIME) { \
        case 1: { constexpr int M = 1; CALL; } break; case 2: { constexpr int M = 2; CALL; } break; \
        case 3: { constexpr int M = 3; CALL; } break; case 4: { constexpr int M = 4; CALL; } break; \
        case 5: { constexpr int M = 5; CALL; } break; case 6: { constexpr int M = 6; CALL; } break; \
        case 7: { constexpr int M = 7; CALL; } break; case 8: { constexpr int M = 8; CALL; } break; \
        default: throw std::runtime_error("MMA: unsupported number of constraints"); }

// Gaussian elimination with partial pivoting for the (m+1) x (m+1) Newton system (colPivHouseholderQr in the reference, :351)
std::vector<double> solve_dense(std::vector<double> A, std::vector<double> b, int k) {
    for (int c = 0; c < k; ++c) {
        int piv = c;
        for (int r = c + 1; r < k; ++r) if (std::abs(A[r * k + c]) > std::abs(A[piv * k + c])) piv = r;
        if (piv != c) { for (int j = 0; j < k; ++j) std::swap(A[c * k + j], A[piv * k + j]); std::swap(b[c], b[piv]); }
        for (int r = c + 1; r < k; ++r) {
            const double f = A[r * k + c] / A[c * k + c];
            for (int j = c; j < k; ++j) A[r * k + j] -= f * A[c * k + j];
            b[r] -= f * b[c];
        }
    }
    std::vector<double> x(k);
    for (int r = k - 1; r >= 0; --r) { double t = b[r]; for (int j = r + 1; j < k; ++j) t -= A[r * k + j] * x[j]; x[r] = t / A[r * k + r]; }
    return x;
}

// newton_step on (x, xi, eta) + residual re-evaluation; returns squared residual, sets kktMax
double step_eval(vf_mma &M_, double t, double eps, double &kktMax) {
    const int m = M_.m;
    RedSpec s = vf_mma::spec(m + 2); s.op[m + 1] = RED_MAX;
    { ProfScope ps(M_.ctx, PC_TOPOPT, (double)M_.n);
      VF_MMA_DISPATCH(m, (k_mma_step_eval<M><<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), t, eps, vf_mma::small(M_.data.lam), s, M_.partials.p)));
      VF_KERNEL_CHECK(); }
    const double *res = M_.finish(s);
    M_.gvec.assign(res, res + m);
    double sq = res[m], mx = res[m + 1];
    auto acc = [&](double v) { sq += v * v; mx = std::max(mx, std::abs(v)); };
    double la = 0;
    const auto &D = M_.data;
    for (int i = 0; i < m; ++i) {
        acc(M_.c[i] + M_.d[i] * D.y[i] - D.lam[i] - D.mu[i]);                          // eq_b (:443)
        acc(M_.gvec[i] - M_.a[i] * D.z - D.y[i] + D.s[i] + M_.r[i + 1]);               // eq_d (:445)
        acc(D.mu[i] * D.y[i] - eps); acc(D.lam[i] * D.s[i] - eps);                     // eq_g, eq_i
        la += D.lam[i] * M_.a[i];
    }
    acc(M_.a0 - D.zeta - la); acc(D.zeta * D.z - eps);                                 // eq_c, eq_h
    kktMax = mx;
    return sq;
}
void host_newton_step(vf_mma &M_, double t) { // (:432-437)
    auto &D = M_.data; const auto &d = M_.delta;
    for (int i = 0; i < M_.m; ++i) { D.y[i] += t * d.y[i]; D.lam[i] += t * d.lam[i]; D.mu[i] += t * d.mu[i]; D.s[i] += t * d.s[i]; }
    D.z += t * d.z; D.zeta += t * d.zeta;
}

// Subproblem::subsolve (:242-258); the solution is left in M_.x
void subsolve(vf_mma &M_) {
    const int m = M_.m; const long long n = M_.n;
    auto &D = M_.data; auto &dl = M_.delta;
    D.y.assign(m, 1.0); D.z = 1; D.zeta = 1; D.lam.assign(m, 1.0); D.s.assign(m, 1.0); D.mu.assign(m, 0.0);
    for (int i = 0; i < m; ++i) D.mu[i] = std::max(M_.c[i] / 2, 1.0);
    { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n); k_mma_init<<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs()); VF_KERNEL_CHECK(); }
    double kkt = 0, eps = 1;
    while (eps > 1e-7) {
        double resOld = 0;
        for (int it = 0; it < 10; ++it) {
            if (it == 0) resOld = step_eval(M_, 0.0, eps, kkt);                         // squared_residual(eps, init) (:251); also refreshes gvec
            // ---- solve_for_newton_direction (:315-363)
            const int R = m * (m + 1) / 2 + m;
            RedSpec s1 = vf_mma::spec(R);
            { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n);
              VF_MMA_DISPATCH(m, (k_mma_dir1<M><<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), eps, vf_mma::small(D.lam), s1, M_.partials.p)));
              VF_KERNEL_CHECK(); }
            const double *red = M_.finish(s1);
            std::vector<double> Dy(m), dy(m);
            for (int i = 0; i < m; ++i) { Dy[i] = M_.d[i] + D.mu[i] / D.y[i]; dy[i] = M_.c[i] + M_.d[i] * D.y[i] - D.lam[i] - eps / D.y[i]; }
            const int k = m + 1; std::vector<double> A((size_t)k * k, 0.0), rhs(k, 0.0);
            int idx = 0;
            for (int ci = 0; ci < m; ++ci) for (int cj = ci; cj < m; ++cj) A[ci * k + cj] = red[idx++];
            for (int i = 0; i < m; ++i) { A[i * k + i] += D.s[i] / D.lam[i] + 1 / Dy[i]; A[i * k + m] = M_.a[i]; }
            A[m * k + m] = -D.zeta / D.z;
            for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j) A[j * k + i] = A[i * k + j];
            double la = 0;
            for (int i = 0; i < m; ++i) {
                rhs[i] = M_.gvec[i] - M_.a[i] * D.z - D.y[i] + M_.r[i + 1] + eps / D.lam[i] + dy[i] / Dy[i] - red[idx + i];
                la += D.lam[i] * M_.a[i];
            }
            rhs[m] = M_.a0 - la - eps / D.z;
            const std::vector<double> sol = solve_dense(A, rhs, k);
            dl.lam.assign(sol.begin(), sol.begin() + m); dl.z = sol[m];
            RedSpec s2 = vf_mma::spec(1, RED_MIN);
            { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n);
              VF_MMA_DISPATCH(m, (k_mma_dir2<M><<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), eps, vf_mma::small(dl.lam), s2, M_.partials.p)));
              VF_KERNEL_CHECK(); }
            double mn = M_.finish(s2)[0];
            dl.y.assign(m, 0); dl.mu.assign(m, 0); dl.s.assign(m, 0);
            for (int i = 0; i < m; ++i) {
                dl.y[i] = dl.lam[i] / Dy[i] - dy[i] / Dy[i];
                dl.mu[i] = (eps - D.mu[i] * dl.y[i]) / D.y[i] - D.mu[i];
                dl.s[i] = (eps - D.s[i] * dl.lam[i]) / D.lam[i] - D.s[i];
            }
            dl.zeta = (eps - D.zeta * dl.z) / D.z - D.zeta;
            ++M_.newtonIters;
            // ---- newton_step_backtrack (:365-376) with step_satisfy_KKT (:384-397)
            for (int i = 0; i < m; ++i) mn = std::min({mn, dl.y[i] / D.y[i], dl.s[i] / D.s[i], dl.mu[i] / D.mu[i], dl.lam[i] / D.lam[i]});
            mn = std::min({mn, dl.z / D.z, dl.zeta / D.zeta});
            double t = 1 / std::max(-1.01 * mn, 1.0);
            host_newton_step(M_, t);
            double res = step_eval(M_, t, eps, kkt);
            while (res > resOld) { t /= 2; host_newton_step(M_, -t); res = step_eval(M_, -t, eps, kkt); }
            resOld = res;
            if (kkt <= 0.9 * eps) break;                                                // KKT_inf_norm (:253)
        }
        eps *= 0.1;
    }
}

// evaluate the user's functions at a device vector
void eval_f(vf_mma &M_, const double *xDev, vf_mma_f_callback f, void *user, int devPtrs, std::vector<double> &out) {
    out.assign(M_.m + 1, 0.0);
    if (devPtrs) { VF_CUDA(cudaStreamSynchronize(M_.stream)); if (f(xDev, out.data(), user)) throw std::runtime_error("MMA: objective/constraint callback failed"); return; }
    VF_CUDA(cudaMemcpyAsync(M_.hostX, xDev, sizeof(double) * M_.n, cudaMemcpyDeviceToHost, M_.stream));
    VF_CUDA(cudaStreamSynchronize(M_.stream));
    if (f(M_.hostX, out.data(), user)) throw std::runtime_error("MMA: objective/constraint callback failed");
}

} // namespace

extern "C" {

int vf_mma_create(int64_t num_vars, int num_constr, const double *xmin, const double *xmax, vf_mma **out) {
    try {
        if (num_vars < 1) throw std::runtime_error("MMA: numVars must be positive");
        if (num_constr < 1 || num_constr > kMaxM) throw std::runtime_error("MMA: numConstr must be between 1 and " + std::to_string(kMaxM));
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) throw std::runtime_error("voxelfem_b200: no usable CUDA device (this library has no CPU fallback)");
        auto M_ = std::make_unique<vf_mma>();
        const long long n = num_vars; const int m = num_constr;
        M_->n = n; M_->m = m;
        VF_CUDA(cudaStreamCreateWithFlags(&M_->stream, cudaStreamNonBlocking));
        M_->ctx.stream = M_->stream;
        for (Buf<double> *b : {&M_->xmin, &M_->xmax, &M_->xh[0], &M_->xh[1], &M_->xh[2], &M_->l, &M_->u, &M_->alpha, &M_->beta, &M_->x, &M_->xi, &M_->eta,
                               &M_->dx, &M_->dxi, &M_->deta, &M_->Dx}) b->alloc(n);
        for (Buf<double> *b : {&M_->df, &M_->p, &M_->q}) b->alloc((size_t)(m + 1) * n);
        M_->partials.alloc((size_t)kMaxRed * kBlocks); M_->results.alloc(kMaxRed);
        VF_CUDA(cudaMallocHost(&M_->hostRes, sizeof(double) * kMaxRed));
        VF_CUDA(cudaMallocHost(&M_->hostX, sizeof(double) * n));
        VF_CUDA(cudaMallocHost(&M_->hostDf, sizeof(double) * (size_t)(m + 1) * n));
        VF_CUDA(cudaMemcpyAsync(M_->xmin.p, xmin, sizeof(double) * n, cudaMemcpyHostToDevice, M_->stream));
        VF_CUDA(cudaMemcpyAsync(M_->xmax.p, xmax, sizeof(double) * n, cudaMemcpyHostToDevice, M_->stream));
        VF_CUDA(cudaStreamSynchronize(M_->stream));
        M_->a.assign(m, 0.0); M_->d.assign(m, 1.0); M_->c.assign(m, 1000.0);          // (:36)
        *out = M_.release();
    } catch (const std::exception &e) { vf::set_last_error(e.what()); return 1; }
    return 0;
}
int vf_mma_destroy(vf_mma *M_) { if (M_) { cudaStreamSynchronize(M_->stream); delete M_; } return 0; }
int vf_mma_enable_gcmma(vf_mma *M_, int enable) { M_->gcmma = enable != 0; return 0; }                 // enableGCMMA (:53)
int vf_mma_set_initial_var(vf_mma *M_, const double *x) {                                              // setInitialVar (:54-56)
    try {
        std::memcpy(M_->hostX, x, sizeof(double) * M_->n);
        VF_CUDA(cudaMemcpyAsync(M_->x.p, M_->hostX, sizeof(double) * M_->n, cudaMemcpyHostToDevice, M_->stream));
        M_->pushHistory(M_->x.p);
        VF_CUDA(cudaStreamSynchronize(M_->stream));
    } catch (const std::exception &e) { vf::set_last_error(e.what()); return 1; }
    return 0;
}
int vf_mma_get_optimal_var(vf_mma *M_, double *x) {                                                     // getOptimalVar (:134)
    try {
        if (M_->nhist == 0) throw std::runtime_error("Must specify an initial value");
        VF_CUDA(cudaMemcpyAsync(x, M_->xh[0].p, sizeof(double) * M_->n, cudaMemcpyDeviceToHost, M_->stream));
        VF_CUDA(cudaStreamSynchronize(M_->stream));
    } catch (const std::exception &e) { vf::set_last_error(e.what()); return 1; }
    return 0;
}
int vf_mma_get_optimal_var_dev(vf_mma *M_, const double **x_dev) { *x_dev = M_->xh[0].p; return M_->nhist ? 0 : 1; }
int64_t vf_mma_newton_iterations(const vf_mma *M_) { return M_->newtonIters; }

// MMA::step (:63-133)
int vf_mma_step(vf_mma *Mp, vf_mma_f_callback f, vf_mma_df_callback df, void *user, int callbacks_take_device_pointers) {
    try {
        vf_mma &M_ = *Mp; const int m = M_.m; const long long n = M_.n; const int dev = callbacks_take_device_pointers;
        if (M_.nhist == 0) throw std::runtime_error("Must specify an initial value");
        ++M_.outerIter;
        { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n); k_mma_asymptotes<<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), M_.outerIter); VF_KERNEL_CHECK(); }
        // f(x^k), df/dx(x^k) (:93-94)
        eval_f(M_, M_.xh[0].p, f, user, dev, M_.fcur);
        if (dev) { if (df(M_.xh[0].p, M_.df.p, user)) throw std::runtime_error("MMA: gradient callback failed"); }
        else {
            if (df(M_.hostX, M_.hostDf, user)) throw std::runtime_error("MMA: gradient callback failed");
            VF_CUDA(cudaMemcpyAsync(M_.df.p, M_.hostDf, sizeof(double) * (size_t)(m + 1) * n, cudaMemcpyHostToDevice, M_.stream));
        }
        auto build_pq = [&](const std::vector<double> &rho) { // p, q and r = f(x^k) - g(x^k)
            RedSpec s = vf_mma::spec(m + 1);
            { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n);
              VF_MMA_DISPATCH(m, (k_mma_pq<M><<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), vf_mma::small(rho), s, M_.partials.p)));
              VF_KERNEL_CHECK(); }
            const double *g = M_.finish(s);
            M_.r.assign(m + 1, 0.0);
            for (int i = 0; i <= m; ++i) M_.r[i] = M_.fcur[i] - g[i];
        };
        if (!M_.gcmma) {
            build_pq(std::vector<double>(m + 1, M_.raa0));
            subsolve(M_);
            M_.pushHistory(M_.x.p);
        } else {
            M_.innerIter = 0;
            std::vector<double> dd(1, 0.0);
            bool feasible = false;
            do {
                std::vector<double> rho(m + 1);
                if (M_.innerIter == 0) { // next_rho (:138-145)
                    RedSpec s = vf_mma::spec(m + 1);
                    { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n);
                      VF_MMA_DISPATCH(m, (k_mma_rho0<M><<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), s, M_.partials.p)));
                      VF_KERNEL_CHECK(); }
                    const double *res = M_.finish(s);
                    for (int i = 0; i <= m; ++i) rho[i] = std::max(res[i], 1e-6);
                } else {                  // (:146-155)
                    for (int i = 0; i <= m; ++i) {
                        const double del = M_.diff[i] / dd[0];
                        rho[i] = del < 0 ? M_.rho[i] : std::min(1.1 * (M_.rho[i] + del), 10 * M_.rho[i]);
                    }
                }
                M_.rho = rho;
                build_pq(rho);
                subsolve(M_);
                ++M_.innerIter;
                // isFeasible (:190-193): f(x^{k,l}) - (g(x^{k,l}) + r) < 0 for all rows
                std::vector<double> fx;
                eval_f(M_, M_.x.p, f, user, dev, fx);
                RedSpec s = vf_mma::spec(m + 2);
                { ProfScope ps(M_.ctx, PC_TOPOPT, (double)n);
                  VF_MMA_DISPATCH(m, (k_mma_geval<M><<<kBlocks, kThreads, 0, M_.stream>>>(M_.ptrs(), s, M_.partials.p)));
                  VF_KERNEL_CHECK(); }
                const double *g = M_.finish(s);
                M_.diff.assign(m + 1, 0.0);
                double mx = -std::numeric_limits<double>::infinity();
                for (int i = 0; i <= m; ++i) { M_.diff[i] = fx[i] - (g[i] + M_.r[i]); mx = std::max(mx, M_.diff[i]); }
                dd[0] = g[m + 1];
                feasible = mx < 0;
            } while (!feasible);
            M_.pushHistory(M_.x.p);
            M_.innerIter = 0;
        }
        VF_CUDA(cudaStreamSynchronize(M_.stream));
    } catch (const std::exception &e) { vf::set_last_error(e.what()); return 1; }
    return 0;
}

} // extern "C"
