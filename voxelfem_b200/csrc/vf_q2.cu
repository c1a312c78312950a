// vf_q2.cu -- degree-2 (Q2) tensor-product elements: TensorProductSimulator<double, 2, 2[, 2]> as far as the reference's GENERIC
// element path goes (sm_100a).  SURVEY.md section 8(f) rank 3; BASELINE.json north_star (a) names the Q2 stiffness apply.
//
// Reference: the generic SpecializedTPSStencils<Real, Degrees...>::applyK (TPSStencils.hh:163-185) -- a multicoloured element
// scatter  f[nodes(e)] (+,-)= E_e K0 u[nodes(e)]  over 3^N nodes per element -- with K0 from Element_T::Stiffness
// (TensorProductSimulator.hh:67-80: Gauss rule of degree 2 * deg per axis), the strains of TensorProductPolynomialInterpolant.hh:
// 204-231 and the Lagrange basis of LagrangePolynomial.hh:7-57.  Node grid (2 ne + 1)^N, flat indices row-major with the last axis
// fastest (NDVector.hh:249-275); local node (l0, l1[, l2]) of element e is global node 2 e + l, local index row-major over 3^N
// (TPSStencils.hh:139); K0 entry N * local + component.  The reference's python bindings never instantiate degree 2
// (python_bindings/VoxelFEM.cc:303-308), so there is no multigrid for it here either: the solver is a Jacobi-preconditioned CG on
// the same operator (the reference would use its CHOLMOD direct solve, TensorProductSimulator.hh:1198-1230).
//
// Kernel: the 2^N element colours (parity classes of the element index; same-colour elements share no node) run one after the other
// on the stream, so the scatter needs no atomics and is deterministic -- the reference's visitElementsMulticolored
// (TensorProductSimulator.hh:1444-1457).  A block owns a few elements; an element's N 3^N displacements are staged in shared
// memory and thread (element, row i) forms (K0 u_e)_i.  K0 is symmetric, so row i is read as column i: consecutive threads read
// consecutive addresses, and the 52 KB matrix (3D) stays in L1 / L2.
#include "vf_internal.cuh"
#include "vf_reduce.cuh"
#include "../../include/voxelfem_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

namespace vf {

struct Q2Grid {
    int N;
    int ne[3], nn[3];               // elements / nodes (2 ne + 1) per axis; unused axes: ne = 1, nn = 1 (leading, as in GridDesc)
    long long numNodes, numElems;
};

template<int N> struct Q2Dims { static constexpr int NPE = N == 3 ? 27 : 9, KE = N * NPE; };

template<int N>
__global__ void __launch_bounds__(N == 3 ? 4 * 81 : 8 * 18)
k_q2_apply(const __grid_constant__ Q2Grid g, int c0, int c1, int c2, int cnt0, int cnt1, int cnt2,
           const double *__restrict__ E, const double *__restrict__ K0, const double *__restrict__ u, double *out, double sign) {
    constexpr int NPE = Q2Dims<N>::NPE, KE = Q2Dims<N>::KE, EPB = N == 3 ? 4 : 8;
    __shared__ double us[EPB][KE];
    const int le = threadIdx.x / KE, i = threadIdx.x % KE;
    const long long idx = (long long)blockIdx.x * EPB + le;        // element within the colour
    const long long total = (long long)cnt0 * cnt1 * cnt2;
    const bool valid = idx < total;
    int e0 = 0, e1 = 0, e2 = 0;
    if (valid) { long long r = idx; e2 = c2 + 2 * (int)(r % cnt2); r /= cnt2; e1 = c1 + 2 * (int)(r % cnt1); e0 = c0 + 2 * (int)(r / cnt1); }
    const int m = i / N, c = i % N;                               // local node, component
    const int l2 = m % 3, l1 = (m / 3) % 3, l0 = N == 3 ? m / 9 : 0;
    // 2D grids are embedded with a leading dummy axis: (1, n1, n2)
    const long long node = ((long long)(N == 3 ? 2 * e0 + l0 : 0) * g.nn[1] + (2 * e1 + l1)) * g.nn[2] + (2 * e2 + l2);
    if (valid) us[le][i] = u[(long long)c * g.numNodes + node];
    __syncthreads();
    if (!valid) return;
    double acc = 0.0;
    #pragma unroll 9
    for (int j = 0; j < KE; ++j) acc = fma(__ldg(K0 + j * KE + i), us[le][j], acc);
    const long long e = ((long long)e0 * g.ne[1] + e1) * g.ne[2] + e2;
    double *o = out + (long long)c * g.numNodes + node;
    *o = fma(sign * E[e], acc, *o);
}

// u_e^T K0 u_e per element (compliance sensitivities, elementEnergyDensity): one block row per element, reduced over its KE threads
template<int N>
__global__ void __launch_bounds__(N == 3 ? 96 : 32)
k_q2_energy(const __grid_constant__ Q2Grid g, const double *__restrict__ K0, const double *__restrict__ u, double *__restrict__ energy) {
    constexpr int KE = Q2Dims<N>::KE;
    __shared__ double us[KE];
    const long long e = blockIdx.x;
    int e2 = (int)(e % g.ne[2]), e1 = (int)((e / g.ne[2]) % g.ne[1]), e0 = (int)(e / ((long long)g.ne[2] * g.ne[1]));
    const int i = threadIdx.x;
    double mine = 0.0;
    if (i < KE) {
        const int m = i / N, c = i % N, l2 = m % 3, l1 = (m / 3) % 3, l0 = N == 3 ? m / 9 : 0;
        const long long node = ((long long)(N == 3 ? 2 * e0 + l0 : 0) * g.nn[1] + (2 * e1 + l1)) * g.nn[2] + (2 * e2 + l2);
        mine = u[(long long)c * g.numNodes + node];
        us[i] = mine;
    }
    __syncthreads();
    double acc = 0.0;
    if (i < KE) { for (int j = 0; j < KE; ++j) acc = fma(__ldg(K0 + j * KE + i), us[j], acc); acc *= mine; }
    acc = block_sum(acc);
    if (i == 0) energy[e] = acc;
}

__global__ void __launch_bounds__(256) k_q2_moduli(long long n, const double *__restrict__ rho, double *__restrict__ E, int law, double E0, double Emin, double gamma, double q) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const double r = rho[e];
        E[e] = (law == 0) ? (Emin + pow(r, gamma) * (E0 - Emin)) : (Emin + r * (E0 - Emin) / (1.0 + q * (1.0 - r)));
    }
}
// y = m ? 0 : x  (Dirichlet components), z = x / d on the free components (Jacobi), and the fused CG vector updates
__global__ void __launch_bounds__(256) k_q2_mask(long long n, const uint8_t *__restrict__ fixed, double *x) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) if (fixed[i]) x[i] = 0.0;
}
__global__ void __launch_bounds__(256) k_q2_jacobi(long long n, const uint8_t *__restrict__ fixed, const double *__restrict__ d, const double *__restrict__ r, double *__restrict__ z) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) z[i] = fixed[i] ? 0.0 : r[i] / d[i];
}
__global__ void __launch_bounds__(256) k_q2_xpby(long long n, const double *__restrict__ x, double beta, double *y) {   // y = x + beta y
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fma(beta, y[i], x[i]);
}
static unsigned q2_blocks(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148 * 16); }

} // namespace vf

using namespace vf;

namespace {
template<class T> struct Buf {
    T *p = nullptr; size_t n = 0;
    ~Buf() { if (p) cudaFree(p); }
    void alloc(size_t count) { if (p) cudaFree(p); p = nullptr; n = count; if (count) { VF_CUDA(cudaMalloc(&p, count * sizeof(T))); VF_CUDA(cudaMemset(p, 0, count * sizeof(T))); VF_CUDA(cudaStreamSynchronize(0)); } }
};
}

struct vf_q2 {
    int N = 3; Q2Grid g;
    double dmin[3] = {0, 0, 0}, dmax[3] = {1, 1, 1}, h[3] = {1, 1, 1};
    double D[6][6];
    int law = VF_LAW_SIMP; double E0 = 1, Emin = 1e-4, gamma = 3, q = 3;
    std::vector<double> K0;
    Buf<double> K0dev, rho, E, tmpA, tmpB, diag, r, z, p, Ap, scalar, scratch, energy; Buf<uint8_t> fixed;
    cudaStream_t stream = nullptr; LaunchCtx ctx;
    ~vf_q2() { if (stream) cudaStreamDestroy(stream); }
    int ke() const { return N * (N == 3 ? 27 : 9); }
    size_t ndof() const { return (size_t)g.numNodes * N; }
    int symIdx(int i, int j) const { if (i == j) return i; if (N == 2) return 2; return 6 - i - j; }

    // 1D Lagrange basis of degree 2 on the nodes 0, 1/2, 1 (LagrangePolynomial.hh:7-57)
    static void lagrange2(double x, double (&v)[3], double (&d)[3]) {
        v[0] = 2.0 * (x - 0.5) * (x - 1.0); v[1] = -4.0 * x * (x - 1.0); v[2] = 2.0 * x * (x - 0.5);
        d[0] = 4.0 * x - 3.0; d[1] = -8.0 * x + 4.0; d[2] = 4.0 * x - 1.0;
    }
    // Element_T::Stiffness (TensorProductSimulator.hh:67-80) with the 3-point Gauss rule per axis (degree 2 * 2, :50)
    void updateK0() {
        const int npe = N == 3 ? 27 : 9, n = N * npe, fl = N == 3 ? 6 : 3;
        K0.assign((size_t)n * n, 0.0);
        const double s = std::sqrt(0.6), gp[3] = {0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s}, gw[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
        std::vector<double> B((size_t)n * fl);
        const int nq = N == 3 ? 27 : 9;
        for (int qi = 0; qi < nq; ++qi) {
            int qd[3] = {0, 0, 0}; { int r = qi; for (int d = N - 1; d >= 0; --d) { qd[d] = r % 3; r /= 3; } }
            double w = 1.0, val[3][3], der[3][3];
            for (int d = 0; d < N; ++d) { w *= gw[qd[d]]; lagrange2(gp[qd[d]], val[d], der[d]); }
            std::fill(B.begin(), B.end(), 0.0);
            for (int m = 0; m < npe; ++m) {
                int l[3] = {0, 0, 0}; { int r = m; for (int d = N - 1; d >= 0; --d) { l[d] = r % 3; r /= 3; } }
                double grad[3] = {0, 0, 0};
                for (int c = 0; c < N; ++c) { double v = 1.0; for (int d = 0; d < N; ++d) v *= d == c ? der[d][l[d]] / h[d] : val[d][l[d]]; grad[c] = v; }
                for (int c = 0; c < N; ++c) {
                    double *row = &B[(size_t)(N * m + c) * fl];
                    for (int i = 0; i < N; ++i) row[symIdx(c, i)] = 0.5 * grad[i];
                    row[symIdx(c, c)] = grad[c];
                }
            }
            for (int a = 0; a < n; ++a) {
                double sig[6];
                for (int i = 0; i < fl; ++i) { double t = 0; for (int j = 0; j < fl; ++j) t += D[i][j] * (j >= N ? 2.0 : 1.0) * B[(size_t)a * fl + j]; sig[i] = t * (i >= N ? 2.0 : 1.0); }
                for (int b = a; b < n; ++b) { double acc = 0; for (int i = 0; i < fl; ++i) acc += sig[i] * B[(size_t)b * fl + i]; K0[(size_t)a * n + b] += w * acc; }
            }
        }
        double vol = 1.0; for (int d = 0; d < N; ++d) vol *= h[d];
        for (int a = 0; a < n; ++a) for (int b = a; b < n; ++b) { K0[(size_t)a * n + b] *= vol; K0[(size_t)b * n + a] = K0[(size_t)a * n + b]; }
        K0dev.alloc(K0.size());
        VF_CUDA(cudaMemcpyAsync(K0dev.p, K0.data(), K0.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        VF_CUDA(cudaStreamSynchronize(stream));
    }
    void updateModuli() { k_q2_moduli<<<q2_blocks(g.numElems), 256, 0, stream>>>(g.numElems, rho.p, E.p, law, E0, Emin, gamma, q); VF_KERNEL_CHECK(); }
    // out (+,-)= K u on device arrays (component-major); the 2^N colours in HypercubeCornerVisitor order
    void apply(const double *u, double *out, bool zeroInit, bool negate) {
        if (zeroInit) VF_CUDA(cudaMemsetAsync(out, 0, ndof() * sizeof(double), stream));
        const int ncol = 1 << N;
        for (int col = 0; col < ncol; ++col) {
            int off[3] = {0, 0, 0}, cnt[3] = {1, 1, 1};
            bool empty = false;
            for (int d = 0; d < N; ++d) {
                const int a = 3 - N + d;                       // embedded axis
                off[a] = (col >> (N - 1 - d)) & 1;
                if (g.ne[a] - 1 - off[a] < 0) empty = true; else cnt[a] = (g.ne[a] - 1 - off[a]) / 2 + 1;
            }
            if (empty) continue;
            const long long total = (long long)cnt[0] * cnt[1] * cnt[2];
            count_launch();
            if (N == 3) k_q2_apply<3><<<(unsigned)((total + 3) / 4), 4 * 81, 0, stream>>>(g, off[0], off[1], off[2], cnt[0], cnt[1], cnt[2], E.p, K0dev.p, u, out, negate ? -1.0 : 1.0);
            else        k_q2_apply<2><<<(unsigned)((total + 7) / 8), 8 * 18, 0, stream>>>(g, off[0], off[1], off[2], cnt[0], cnt[1], cnt[2], E.p, K0dev.p, u, out, negate ? -1.0 : 1.0);
            VF_KERNEL_CHECK();
        }
    }
    double dot(const double *a, const double *b) {
        launch_dot_plain(ctx, (long long)ndof(), a, b, scalar.p, scratch.p);
        double v = 0; VF_CUDA(cudaMemcpyAsync(&v, scalar.p, sizeof(double), cudaMemcpyDeviceToHost, stream)); VF_CUDA(cudaStreamSynchronize(stream));
        return v;
    }
};

#define Q2_TRY try {
#define Q2_CATCH } catch (const std::logic_error &e) { vf::set_last_error(std::string("logic_error: ") + e.what()); return 2; } \
                   catch (const std::exception &e) { vf::set_last_error(e.what()); return 1; } return 0;

extern "C" {

int vf_q2_create(int dim, const int64_t *ne, const double *dmin, const double *dmax, vf_q2 **out) {
    Q2_TRY
    int cnt = 0; if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) throw std::runtime_error("voxelfem_b200: no usable CUDA device (this library has no CPU fallback)");
    if (dim != 2 && dim != 3) throw std::runtime_error("vf_q2_create: dimension must be 2 or 3");
    auto s = std::make_unique<vf_q2>();
    s->N = dim;
    std::memset(&s->g, 0, sizeof(s->g)); s->g.N = dim;
    for (int a = 0; a < 3; ++a) { s->g.ne[a] = 1; s->g.nn[a] = 1; }
    for (int d = 0; d < dim; ++d) {
        if (ne[d] < 1) throw std::runtime_error("vf_q2_create: at least one element per axis");
        const int a = 3 - dim + d;
        s->g.ne[a] = (int)ne[d]; s->g.nn[a] = 2 * (int)ne[d] + 1;
        s->dmin[d] = dmin[d]; s->dmax[d] = dmax[d]; s->h[d] = (dmax[d] - dmin[d]) / (double)ne[d];
    }
    s->g.numNodes = (long long)s->g.nn[0] * s->g.nn[1] * s->g.nn[2];
    s->g.numElems = (long long)s->g.ne[0] * s->g.ne[1] * s->g.ne[2];
    VF_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->ctx.stream = s->stream;
    const size_t nd = s->ndof();
    s->rho.alloc(s->g.numElems); s->E.alloc(s->g.numElems); s->energy.alloc(s->g.numElems);
    for (Buf<double> *b : {&s->tmpA, &s->tmpB, &s->diag, &s->r, &s->z, &s->p, &s->Ap}) b->alloc(nd);
    s->fixed.alloc(nd); s->scalar.alloc(4); s->scratch.alloc(reduce_scratch_doubles());
    launch_fill(s->ctx, s->g.numElems, 1.0, s->rho.p);
    // ETensor(1, 0) as the reference's default (TensorProductSimulator.hh:2114)
    std::memset(s->D, 0, sizeof(s->D));
    for (int i = 0; i < dim; ++i) s->D[i][i] = 1.0;
    for (int i = dim; i < (dim == 3 ? 6 : 3); ++i) s->D[i][i] = 0.5;
    s->updateK0(); s->updateModuli();
    VF_CUDA(cudaStreamSynchronize(s->stream));
    *out = s.release();
    Q2_CATCH
}
int vf_q2_destroy(vf_q2 *s) { Q2_TRY if (s) { cudaStreamSynchronize(s->stream); delete s; } Q2_CATCH }
int64_t vf_q2_num_nodes(const vf_q2 *s) { return s->g.numNodes; }
int64_t vf_q2_num_elements(const vf_q2 *s) { return s->g.numElems; }
int vf_q2_set_isotropic(vf_q2 *s, double Ey, double nu) {
    Q2_TRY
    const int N = s->N;
    double lambda = (nu * Ey) / ((1.0 + nu) * (1.0 - 2.0 * nu));
    const double mu = Ey / (2.0 + 2.0 * nu);
    if (N == 2) lambda = (nu * Ey) / (1.0 - nu * nu);
    std::memset(s->D, 0, sizeof(s->D));
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) s->D[i][j] = lambda;
    for (int i = 0; i < N; ++i) s->D[i][i] = lambda + 2 * mu;
    for (int i = N; i < (N == 3 ? 6 : 3); ++i) s->D[i][i] = mu;
    s->updateK0();
    Q2_CATCH
}
int vf_q2_set_elasticity_tensor(vf_q2 *s, const double *D) {
    Q2_TRY const int fl = s->N == 3 ? 6 : 3; std::memset(s->D, 0, sizeof(s->D));
    for (int i = 0; i < fl; ++i) for (int j = 0; j < fl; ++j) s->D[i][j] = D[i * fl + j];
    s->updateK0(); Q2_CATCH
}
int vf_q2_get_K0(const vf_q2 *s, double *out) { Q2_TRY std::copy(s->K0.begin(), s->K0.end(), out); Q2_CATCH }
int vf_q2_set_interpolation(vf_q2 *s, int law, double E_0, double E_min, double gamma, double q) {
    Q2_TRY s->law = law; s->E0 = E_0; s->Emin = E_min; s->gamma = gamma; s->q = q; s->updateModuli(); VF_CUDA(cudaStreamSynchronize(s->stream)); Q2_CATCH
}
int vf_q2_set_densities(vf_q2 *s, const double *rho) {
    Q2_TRY VF_CUDA(cudaMemcpyAsync(s->rho.p, rho, s->g.numElems * sizeof(double), cudaMemcpyHostToDevice, s->stream)); s->updateModuli(); VF_CUDA(cudaStreamSynchronize(s->stream)); Q2_CATCH
}
int vf_q2_get_young_moduli(const vf_q2 *s, double *E) {
    Q2_TRY VF_CUDA(cudaMemcpyAsync(E, s->E.p, s->g.numElems * sizeof(double), cudaMemcpyDeviceToHost, s->stream)); VF_CUDA(cudaStreamSynchronize(s->stream)); Q2_CATCH
}
// u, out: host arrays, component-major (c * numNodes + node) like every nodal field of this ABI
int vf_q2_apply_K(vf_q2 *s, const double *u, double *out, int zero_init, int negate) {
    Q2_TRY
    const size_t nd = s->ndof();
    VF_CUDA(cudaMemcpyAsync(s->tmpA.p, u, nd * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (!zero_init) VF_CUDA(cudaMemcpyAsync(s->tmpB.p, out, nd * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    s->apply(s->tmpA.p, s->tmpB.p, zero_init != 0, negate != 0);
    VF_CUDA(cudaMemcpyAsync(out, s->tmpB.p, nd * sizeof(double), cudaMemcpyDeviceToHost, s->stream)); VF_CUDA(cudaStreamSynchronize(s->stream));
    Q2_CATCH
}
// u_e^T K0 u_e per element (elementEnergyDensity up to the modulus; complianceGradient = -1/2 dE/drho * this)
int vf_q2_element_energies(vf_q2 *s, const double *u, double *energy) {
    Q2_TRY
    VF_CUDA(cudaMemcpyAsync(s->tmpA.p, u, s->ndof() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    count_launch();
    if (s->N == 3) k_q2_energy<3><<<(unsigned)s->g.numElems, 96, 0, s->stream>>>(s->g, s->K0dev.p, s->tmpA.p, s->energy.p);
    else           k_q2_energy<2><<<(unsigned)s->g.numElems, 32, 0, s->stream>>>(s->g, s->K0dev.p, s->tmpA.p, s->energy.p);
    VF_KERNEL_CHECK();
    VF_CUDA(cudaMemcpyAsync(energy, s->energy.p, s->g.numElems * sizeof(double), cudaMemcpyDeviceToHost, s->stream)); VF_CUDA(cudaStreamSynchronize(s->stream));
    Q2_CATCH
}
// Jacobi-preconditioned CG of K x = b with the components flagged in `fixed` (numNodes * N bytes, component-major) clamped to zero.
// x: initial guess in, solution out.  Stops when ||r|| <= tol ||b|| (masked norms) or after max_iter iterations.
int vf_q2_pcg(vf_q2 *s, double *x, const double *b, const uint8_t *fixed, int max_iter, double tol, int *iters, double *rel_residual) {
    Q2_TRY
    const size_t nd = s->ndof(); const long long n = (long long)nd; cudaStream_t st = s->stream;
    double *X = s->tmpA.p, *R = s->r.p, *Z = s->z.p, *P = s->p.p, *AP = s->Ap.p;
    VF_CUDA(cudaMemcpyAsync(X, x, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    VF_CUDA(cudaMemcpyAsync(R, b, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    VF_CUDA(cudaMemcpyAsync(s->fixed.p, fixed, nd, cudaMemcpyHostToDevice, st));
    // diag(K)_i = sum_e E_e K0_ii: the element scatter with K0 replaced by its diagonal, applied to the all-ones field
    {
        std::vector<double> Kd(s->K0.size(), 0.0); const int ke = s->ke();
        for (int i = 0; i < ke; ++i) Kd[(size_t)i * ke + i] = s->K0[(size_t)i * ke + i];
        Buf<double> full; full.alloc(s->K0.size());
        VF_CUDA(cudaMemcpyAsync(full.p, s->K0dev.p, s->K0.size() * sizeof(double), cudaMemcpyDeviceToDevice, st));
        VF_CUDA(cudaMemcpyAsync(s->K0dev.p, Kd.data(), Kd.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_fill(s->ctx, n, 1.0, Z);
        s->apply(Z, s->diag.p, true, false);                      // diag(K0) applied to the all-ones field = diag(K)
        VF_CUDA(cudaMemcpyAsync(s->K0dev.p, full.p, s->K0.size() * sizeof(double), cudaMemcpyDeviceToDevice, st));
        VF_CUDA(cudaStreamSynchronize(st));
    }
    k_q2_mask<<<q2_blocks(n), 256, 0, st>>>(n, s->fixed.p, X);
    k_q2_mask<<<q2_blocks(n), 256, 0, st>>>(n, s->fixed.p, R);
    const double bnorm2 = s->dot(R, R);
    s->apply(X, R, false, true);                                   // r = b - K x
    k_q2_mask<<<q2_blocks(n), 256, 0, st>>>(n, s->fixed.p, R);
    int it = 0; double rr = s->dot(R, R), rz = 0, rzOld = 0;
    while (it < max_iter && rr > tol * tol * bnorm2) {
        k_q2_jacobi<<<q2_blocks(n), 256, 0, st>>>(n, s->fixed.p, s->diag.p, R, Z);
        rzOld = rz; rz = s->dot(R, Z);
        if (it == 0) VF_CUDA(cudaMemcpyAsync(P, Z, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
        else k_q2_xpby<<<q2_blocks(n), 256, 0, st>>>(n, Z, rz / rzOld, P);
        s->apply(P, AP, true, false);
        k_q2_mask<<<q2_blocks(n), 256, 0, st>>>(n, s->fixed.p, AP);
        const double pAp = s->dot(P, AP);
        if (!(pAp > 0.0) || std::isnan(pAp)) throw std::logic_error("vf_q2_pcg: breakdown (p.Ap = " + std::to_string(pAp) + ")");
        const double alpha = rz / pAp;
        launch_axpy(s->ctx, n, alpha, P, X);
        launch_axpy(s->ctx, n, -alpha, AP, R);
        rr = s->dot(R, R);
        ++it;
    }
    VF_KERNEL_CHECK();
    VF_CUDA(cudaMemcpyAsync(x, X, nd * sizeof(double), cudaMemcpyDeviceToHost, st)); VF_CUDA(cudaStreamSynchronize(st));
    if (iters) *iters = it;
    if (rel_residual) *rel_residual = bnorm2 > 0 ? std::sqrt(rr / bnorm2) : 0.0;
    Q2_CATCH
}

} // extern "C"
