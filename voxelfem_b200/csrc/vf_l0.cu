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
#include <cstring>
#include <cmath>
#include <algorithm>

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
    pdl_prologue();
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
    pdl_prologue();
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

// ---------------------------------------------------------------------------
// 3D apply in the symmetry-adapted basis ("Walsh" kernel).
//
// The dense formulation above spends 576 DFMA per node and is FP64-pipe bound at 4x the HBM floor.  A voxel has three
// mirror planes, and for a material whose symmetry planes are those of the grid (isotropic, orthotropic) K0 commutes
// with the mirror group Z2^3 acting on node positions and displacement components.  In the basis of the 2x2x2 Walsh
// functions w_s(m) = (-1)^(s.m), K0 therefore splits into 8 blocks of 3x3 -- one per irreducible representation p,
// coupling the modes (component c, s = p ^ e_c) -- of which the 3 rigid translations vanish: 45..57 multiplies per
// ELEMENT instead of 576.  finalize_k0_param() verifies the block structure numerically; other materials keep the
// dense kernel.
//
// Mapping: one thread per element column (y, z), lanes along z, marching along x.  Per step (one element):
//   * load the 4 nodes of the next node plane, 2D Walsh transform (shared by the elements on either side of the plane)
//   * x butterfly -> 24 mode amplitudes, 8 blocks -> 24 mode forces
//   * inverse x butterfly fused with the modulus scaling and with the carry from the previous element of the column
//   * inverse 2D transform -> contributions to the 4 nodes of the finished node plane
//   * the thread owns node (y, z): contributions of the columns (y, z-1), (y-1, z), (y-1, z-1) arrive by warp shuffle
//     (z) and through shared memory (y); lane 0 and warp 0 of a block are halo columns.
// About 175 FP64 instructions per element including the node accumulation.  The loads of step ex + 1 are issued before
// the arithmetic of step ex (software pipeline) and two blocks are resident per SM (128 registers): measured at 256^3
// 0.45 ms (SET) / 0.58 ms (RESIDUAL) against 1.10 / 1.15 ms for the dense kernel; without the prefetch 0.56 ms, with one
// resident block 0.68 ms, with the carry in shared memory 0.50 ms (profiles/r01o_time_ab.log).
// ---------------------------------------------------------------------------
constexpr int kWalshBY = 8;      // warps (element rows) per block, one of them halo
constexpr int kWalshXC = 32;     // node planes per block along x (one extra element step to prime the carry)

__device__ __forceinline__ void wht2(const double (&r)[2][2], double (&P)[2][2]) {
    const double t0 = r[0][0] + r[0][1], t1 = r[0][0] - r[0][1], t2 = r[1][0] + r[1][1], t3 = r[1][0] - r[1][1];
    P[0][0] = t0 + t2; P[0][1] = t1 + t3; P[1][0] = t0 - t2; P[1][1] = t1 - t3;
}

template<int MODE, bool DOT, bool S7>
__global__ void __launch_bounds__(32 * kWalshBY, 2)
k_apply3w_l0(const __grid_constant__ GridDesc g, const __grid_constant__ KhatParam Kh, const double *__restrict__ u,
             const double *__restrict__ E, const double *__restrict__ b, const uint8_t *__restrict__ dmask,
             double *__restrict__ out, double *dotOut, double *scratch) {
    pdl_prologue();
    __shared__ double s_up[2][kWalshBY][3][32];
    const int lane = threadIdx.x, wy = threadIdx.y;
    const int cz = blockIdx.x * 31 + lane - 1;                 // element column / owned node coordinates (halo: lane 0, warp 0)
    const int cy = blockIdx.y * (kWalshBY - 1) + wy - 1;
    const int xs = blockIdx.z * kWalshXC, xe = min(xs + kWalshXC, g.nn[0]);
    const int nx = g.nn[0], ny = g.nn[1], nz = g.nn[2];
    const long long NN = g.numNodes;
    const bool colValid = cz >= 0 && cz < g.ne[2] && cy >= 0 && cy < g.ne[1];
    const bool owner = lane >= 1 && wy >= 1 && cz < nz && cy < ny;
    // clamped in-plane node offsets of the column's 4 nodes
    long long no[2][2];
    #pragma unroll
    for (int m1 = 0; m1 < 2; ++m1) {
        #pragma unroll
        for (int m2 = 0; m2 < 2; ++m2)
            no[m1][m2] = (long long)min(max(cy + m1, 0), ny - 1) * g.ns[1] + min(max(cz + m2, 0), nz - 1);
    }
    const long long eo = (long long)min(max(cy, 0), g.ne[1] - 1) * g.es[1] + min(max(cz, 0), g.ne[2] - 1);
    double Pp[3][2][2], carry[3][2][2], uown[3];
    {
        const long long xo = (long long)min(max(xs - 1, 0), nx - 1) * g.ns[0];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            double r[2][2];
            #pragma unroll
            for (int m1 = 0; m1 < 2; ++m1) {
                #pragma unroll
                for (int m2 = 0; m2 < 2; ++m2) r[m1][m2] = u[c * NN + xo + no[m1][m2]];
            }
            wht2(r, Pp[c]);
            uown[c] = r[0][0];
            #pragma unroll
            for (int i = 0; i < 4; ++i) carry[c][i >> 1][i & 1] = 0.0;
        }
    }
    // software pipeline: the node plane and modulus of step ex + 1 are requested before the arithmetic of step ex
    double rawN[3][2][2], EeN;
    {
        const long long xo = (long long)min(xs, nx - 1) * g.ns[0];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            #pragma unroll
            for (int m1 = 0; m1 < 2; ++m1) {
                #pragma unroll
                for (int m2 = 0; m2 < 2; ++m2) rawN[c][m1][m2] = u[c * NN + xo + no[m1][m2]];
            }
        }
        const bool ev = colValid && xs - 1 >= 0 && xs - 1 < g.ne[0];
        EeN = ev ? __ldg(E + ((long long)max(xs - 1, 0) * g.es[0] + eo)) : 0.0;
    }
    double dotv = 0.0;
    for (int ex = xs - 1; ex < xe; ++ex) {
        double raw[3][2][2];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            #pragma unroll
            for (int i = 0; i < 4; ++i) raw[c][i >> 1][i & 1] = rawN[c][i >> 1][i & 1];
        }
        const double Ee = EeN;
        if (ex + 1 < xe) {   // next node plane (clamped: beyond the grid the element is void and its modulus zero)
            const long long xo = (long long)min(ex + 2, nx - 1) * g.ns[0];
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                #pragma unroll
                for (int m1 = 0; m1 < 2; ++m1) {
                    #pragma unroll
                    for (int m2 = 0; m2 < 2; ++m2) rawN[c][m1][m2] = u[c * NN + xo + no[m1][m2]];
                }
            }
            const bool evn = colValid && ex + 1 < g.ne[0];
            EeN = evn ? __ldg(E + ((long long)(ex + 1) * g.es[0] + eo)) : 0.0;
        }
        // mode amplitudes uh[c][s], s = (s0 << 2) | (s1 << 1) | s2
        double uh[3][8];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            double Pn[2][2];
            wht2(raw[c], Pn);
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a = Pp[c][i >> 1][i & 1], bb = Pn[i >> 1][i & 1];
                uh[c][i] = a + bb; uh[c][4 + i] = a - bb;
                Pp[c][i >> 1][i & 1] = bb;
            }
        }
        // mode forces wh[c][s]: 8 blocks of 3x3, rigid translations (s = 0) vanish
        double wh[3][8];
        #pragma unroll
        for (int p = 0; p < 8; ++p) {
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int sc = p ^ (4 >> c);
                double acc = 0.0;
                if (sc != 0) {
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int sd = p ^ (4 >> d);
                        if (sd == 0) continue;
                        if (S7 && c != d && (sc == 7 || sd == 7)) continue;
                        acc = fma(Kh.v[p][c][d], uh[d][sd], acc);
                    }
                }
                wh[c][sc] = acc;
            }
        }
        // inverse x butterfly + modulus + carry: plane ex is complete, the m0 = 1 half is carried to plane ex + 1
        double gq[3][2][2];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            double Q[2][2];
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a = wh[c][i] + wh[c][4 + i], bb = wh[c][i] - wh[c][4 + i];
                Q[i >> 1][i & 1] = fma(Ee, a, carry[c][i >> 1][i & 1]);
                carry[c][i >> 1][i & 1] = Ee * bb;
            }
            wht2(Q, gq[c]);   // gq[c][m1][m2]: contribution of this column to node (ex, cy + m1, cz + m2)
        }
        if (ex >= xs) {
            const int buf = ex & 1;
            double acc[3];
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                acc[c] = gq[c][0][0] + __shfl_up_sync(0xffffffffu, gq[c][0][1], 1);
                s_up[buf][wy][c][lane] = gq[c][1][0] + __shfl_up_sync(0xffffffffu, gq[c][1][1], 1);
            }
            __syncthreads();
            if (owner) {
                const long long n = (long long)ex * g.ns[0] + (long long)cy * g.ns[1] + cz;
                if (cy >= g.nActive) { // detached layer: applyK<ZeroInit = true> zero-fills it (TPSStencils.hh:717-727)
                    if (MODE == APPLY_SET) {
                        #pragma unroll
                        for (int c = 0; c < 3; ++c) out[c * NN + n] = 0.0;
                    }
                } else {
                    const unsigned dm = dmask ? dmask[n] : 0u;
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double a = acc[c] + s_up[buf][wy - 1][c][lane];
                        double res;
                        if (MODE == APPLY_SET) res = a;
                        else if (MODE == APPLY_ADD) res = out[c * NN + n] + a;
                        else if (MODE == APPLY_SUB) res = out[c * NN + n] - a;
                        else res = b[c * NN + n] - a;
                        if ((dm >> c) & 1u) res = 0.0;
                        out[c * NN + n] = res;
                        if (DOT && ex >= g.ownLo && ex < g.ownHi) dotv = fma(uown[c], res, dotv);
                    }
                }
            }
        }
        #pragma unroll
        for (int c = 0; c < 3; ++c) uown[c] = raw[c][0][0];
    }
    if (DOT) grid_sum(dotv, scratch, dotOut);
}

// K0 in the Walsh basis: Khat = T K0 T^T / 64 with T = H (x) I3, H[s][m] = (-1)^popcount(s & m) over the local node index
// m = (m0 << 2) | (m1 << 1) | m2 (TensorProductSimulator.hh:1532-1651 node ordering, axis 0 slowest).
void finalize_k0_param(K0Param &K, int N) {
    K.walsh = 0; K.sparse7 = 0;
    std::memset(K.kh, 0, sizeof(K.kh));
    if (N != 3) return;
    static double Kt[24][24], Kh[24][24];
    auto sgn = [](int s, int m) { return (__builtin_popcount(s & m) & 1) ? -1.0 : 1.0; };
    for (int s = 0; s < 8; ++s) for (int c = 0; c < 3; ++c) for (int j = 0; j < 24; ++j) {
        double a = 0; for (int m = 0; m < 8; ++m) a += sgn(s, m) * K.v[(3 * m + c) * 24 + j];
        Kt[3 * s + c][j] = a;
    }
    double kmax = 0;
    for (int i = 0; i < 24; ++i) for (int s = 0; s < 8; ++s) for (int c = 0; c < 3; ++c) {
        double a = 0; for (int m = 0; m < 8; ++m) a += sgn(s, m) * Kt[i][3 * m + c];
        Kh[i][3 * s + c] = a / 64.0; kmax = std::max(kmax, std::fabs(a / 64.0));
    }
    bool blockDiag = true, s7 = true;
    const double tol = 1e-13 * kmax;
    for (int s1 = 0; s1 < 8; ++s1) for (int c1 = 0; c1 < 3; ++c1) for (int s2 = 0; s2 < 8; ++s2) for (int c2 = 0; c2 < 3; ++c2) {
        const double v = Kh[3 * s1 + c1][3 * s2 + c2];
        const bool same = (s1 ^ (4 >> c1)) == (s2 ^ (4 >> c2));
        if (!same || s1 == 0 || s2 == 0) { if (std::fabs(v) > tol) blockDiag = false; continue; }
        K.kh[s1 ^ (4 >> c1)][c1][c2] = v;
        if (c1 != c2 && (s1 == 7 || s2 == 7) && std::fabs(v) > tol) s7 = false;
    }
    // nodal form of the same symmetry (used by the row-unit smoother, vf_gs0.cu); valid whenever the Walsh form is block diagonal
    std::memset(K.vt, 0, sizeof(K.vt));
    for (int D = 0; D < 8; ++D) for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) K.vt[D][3 * c + d] = K.v[c * 24 + 3 * D + d];
    for (int e = 0; e < 8 && blockDiag; ++e) for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) for (int D = 0; D < 8; ++D) {
        const int ec = (e >> (2 - c)) & 1, ed = (e >> (2 - d)) & 1;
        const double sg = (c == d || ((ec + ed) & 1) == 0) ? 1.0 : -1.0;
        if (std::fabs(K.v[(3 * e + c) * 24 + 3 * (e ^ D) + d] - sg * K.vt[D][3 * c + d]) > 64 * tol) blockDiag = false;
    }
    const char *env = std::getenv("VF_L0_DENSE");
    K.walsh = blockDiag && !(env && env[0] == '1');
    K.sparse7 = s7;
}

template<bool DOT>
static void apply3w_launch(const LaunchCtx &ctx, dim3 grid, dim3 block, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                           const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    KhatParam Kh; std::memcpy(Kh.v, K.kh, sizeof(Kh.v));
#define VF_W_CASE(M) \
    if (mode == M) { \
        if (K.sparse7) VF_LAUNCH((k_apply3w_l0<M, DOT, true>), grid, block, 0, ctx.stream, g, Kh, u, E, b, dmask, out, dotOut, scratch); \
        else           VF_LAUNCH((k_apply3w_l0<M, DOT, false>), grid, block, 0, ctx.stream, g, Kh, u, E, b, dmask, out, dotOut, scratch); \
    }
    VF_W_CASE(APPLY_SET) VF_W_CASE(APPLY_ADD) VF_W_CASE(APPLY_SUB) VF_W_CASE(APPLY_RESIDUAL)
#undef VF_W_CASE
}

static void apply3w_l0_dispatch(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                                const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    dim3 block(32, kWalshBY, 1);
    dim3 grid((g.nn[2] + 30) / 31, (g.nn[1] + kWalshBY - 2) / (kWalshBY - 1), (g.nn[0] + kWalshXC - 1) / kWalshXC);
    const bool fused = dotOut && (size_t)grid.x * grid.y * grid.z <= (size_t)kReduceMaxBlocks;
    if (fused) apply3w_launch<true>(ctx, grid, block, g, K, u, E, b, dmask, out, mode, dotOut, scratch);
    else       apply3w_launch<false>(ctx, grid, block, g, K, u, E, b, dmask, out, mode, nullptr, nullptr);
    VF_KERNEL_CHECK();
    if (dotOut && !fused) launch_masked_dot(ctx, g, u, out, dotOut, scratch);
}

static void apply3_l0_dispatch(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                               const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    dim3 block(32, 4, 1);
    dim3 grid((g.nn[2] + 31) / 32, ((g.nn[1] + 1) / 2 + 3) / 4, g.nn[0]);
    // the fused u . out reduction keeps one partial per block; beyond the scratch capacity fall back to a separate dot
    const bool fused = dotOut && (size_t)grid.x * grid.y * grid.z <= (size_t)kReduceMaxBlocks;
#define VF_APPLY_CASE(M) \
    if (mode == M) { \
        if (fused) VF_LAUNCH((k_apply3_l0<M, true>), grid, block, 0, ctx.stream, g, K, u, E, b, dmask, out, dotOut, scratch); \
        else       VF_LAUNCH((k_apply3_l0<M, false>), grid, block, 0, ctx.stream, g, K, u, E, b, dmask, out, nullptr, nullptr); \
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
        if (fused) VF_LAUNCH((k_apply_l0<N, M, true>), grid, block, 0, ctx.stream, g, K, u, E, b, dmask, out, dotOut, scratch); \
        else       VF_LAUNCH((k_apply_l0<N, M, false>), grid, block, 0, ctx.stream, g, K, u, E, b, dmask, out, nullptr, nullptr); \
    }
    VF_APPLY_CASE(APPLY_SET) VF_APPLY_CASE(APPLY_ADD) VF_APPLY_CASE(APPLY_SUB) VF_APPLY_CASE(APPLY_RESIDUAL)
#undef VF_APPLY_CASE
    VF_KERNEL_CHECK();
    if (dotOut && !fused) launch_masked_dot(ctx, g, u, out, dotOut, scratch);
}

void launch_apply_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, const double *u, const double *E,
                     const double *b, const uint8_t *dmask, double *out, int mode, double *dotOut, double *scratch) {
    ProfScope ps(ctx, mode == APPLY_RESIDUAL ? PC_RESIDUAL_L0 : PC_APPLY_L0, (double)g.numNodes);
    if (g.N == 3 && K.walsh) apply3w_l0_dispatch(ctx, g, K, u, E, b, dmask, out, mode, dotOut, scratch);
    else if (g.N == 3) apply3_l0_dispatch(ctx, g, K, u, E, b, dmask, out, mode, dotOut, scratch);
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
    pdl_prologue();
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
    pdl_prologue();
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
    if (forward) VF_LAUNCH((k_gs3_color<true>), grid, block, 0, ctx.stream, g, K, col, u, b, E, dmask);
    else         VF_LAUNCH((k_gs3_color<false>), grid, block, 0, ctx.stream, g, K, col, u, b, E, dmask);
    VF_KERNEL_CHECK();
}

void launch_gs_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                  const uint8_t *dmask, int color, bool forward) {
    ColorDesc col;
    if (!make_color(g, color, col)) return;
    ProfScope ps(ctx, PC_GS_L0, (double)col.cnt[0] * col.cnt[1] * col.cnt[2]);
    dim3 block = (g.N == 3) ? dim3(32, 4, 2) : dim3(32, 8, 1);
    dim3 grid((col.cnt[2] + block.x - 1) / block.x, (col.cnt[1] + block.y - 1) / block.y, (col.cnt[0] + block.z - 1) / block.z);
    if (g.N == 3) VF_LAUNCH((k_gs_l0<3>), grid, block, 0, ctx.stream, g, K, col, u, b, E, dmask, forward ? 1 : 0);
    else          VF_LAUNCH((k_gs_l0<2>), grid, block, 0, ctx.stream, g, K, col, u, b, E, dmask, forward ? 1 : 0);
    VF_KERNEL_CHECK();
}

} // namespace vf
