// vf_gs0.cu -- level-0 multicoloured block Gauss-Seidel on ROW UNITS (sm_100a, 3D, mirror-symmetric K0).
//
// Replaces smoothingMulticoloredGS + NodeSmoothStencilFinest + m_smoothNode at level 0
// (MultigridSolver.hh:277-292, 347-378, 408-458).  The visiting order of the reference is kept: colour index
// 4 p0 + 2 p1 + p2 ascending (descending for backward sweeps), p_a = parity of the node coordinate along axis a.  Nodes of one
// colour do not couple, so any schedule in which a node of colour c sees its neighbours of colours < c updated and those of
// colours > c not yet updated produces the reference's result.
//
// Schedule: the colours (p0, p1, 0) and (p0, p1, 1) of one z-row (fixed x, y) only interact inside that row and with rows of
// other (x, y) parity classes, so a sweep is 4 launches -- one per (p0, p1) class -- in which one thread block owns one z-row and
// runs its two z-colours back to back out of shared memory.  Per row unit the block stages with cp.async (no register holds a
// pending load): the 3 x 3 neighbouring u rows (3 components each), the row's b, and the 2 x 2 adjacent rows of element moduli,
// every row split by z-parity so that the lanes of a colour read consecutive shared-memory words.  HBM traffic per sweep is 4
// passes over u/E instead of 8.
//
// Arithmetic: the 576 multiply-adds per node of the reference formulation, but with K0 expressed through its mirror symmetry
// (K0Param::vt): K0[(e,c),(e^D,d)] = +-vt[D][3c+d].  For a fixed neighbour-offset pattern D the 9 constants serve all incident
// elements and both mirror-image neighbours, so a node needs 72 constant-bank loads instead of 576 and one node per thread is
// enough (24 accumulators instead of 48 -> no register pressure, no spills).  The diagonal block is V[0] times signed sums of the
// 8 moduli.
#include "vf_internal.cuh"
#include <cstdlib>
#include <cmath>
#include <algorithm>

namespace vf {

constexpr int kRowThreads = 256;
constexpr int kRowArrays = 27 + 3 + 4;   // u rows (plane, row, component), b (component), moduli (layer, row)

struct RowPass { int px, py, cntX, cntY; };
// vt[][] of K0Param, plus -- for isotropic material on cubic voxels -- its 10 distinct values: vt[D][3c+d] = +-mag[cls(D,c,d)]
// (gs_vclass below), which then live in registers for the whole kernel instead of being fetched from the constant bank per use.
// v[h]: the table as seen by half h of a node's thread pair (below): h = 1 works on the mirror image along axis 0, which flips the sign
// of every entry coupling the x component with another one.
struct VtabParam { double v[2][8][10]; double mag[10]; };

__device__ __forceinline__ void cp_async8(unsigned dstSmem, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dstSmem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(unsigned dstSmem, const double *src, bool valid) {
    const int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dstSmem), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// x with its sign flipped if flip == 0x80000000 (integer pipe, not the FP64 pipe)
__device__ __forceinline__ double flip_sign(double x, int flip) { return __hiloint2double(__double2hiint(x) ^ flip, __double2loint(x)); }

// class / sign of vt[D][3c+d] for an isotropic material on cubic voxels; bit (2 - a) of D refers to axis a
__host__ __device__ constexpr int gs_dbit(int D, int a) { return (D >> (2 - a)) & 1; }
__host__ __device__ constexpr int gs_vclass(int D, int c, int d) {
    return c == d ? (gs_dbit(D, c) ? 3 : 0) + gs_dbit(D, (c + 1) % 3) + gs_dbit(D, (c + 2) % 3)
                  : 6 + ((gs_dbit(D, c) ^ gs_dbit(D, d)) ? 2 : 0) + gs_dbit(D, 3 - c - d);
}
__host__ __device__ constexpr bool gs_vneg(int D, int c, int d) { return c != d && gs_dbit(D, c) != 0; }

// Shared-memory layout: kRowArrays rows of 2 * HP doubles.  Row element z lives in half (z & 1) at index (z >> 1) + 1; index 0 of
// the odd half stands for z = -1 and the index after the last element for z = nz (both finite; the moduli there are zero).
template<int HP> __host__ __device__ constexpr int row_u(int p, int r, int c) { return ((p * 3 + r) * 3 + c) * 2 * HP; }
template<int HP> __host__ __device__ constexpr int row_b(int c) { return (27 + c) * 2 * HP; }
template<int HP> __host__ __device__ constexpr int row_e(int lx, int ly) { return (30 + lx * 2 + ly) * 2 * HP; }

// Partial sums of one side of a node: H = 0 takes the 4 incident elements on the +x side (node planes x, x + 1), H = 1 those on
// the -x side (planes x, x - 1).  Both are the SAME computation, written for the +x side: by the mirror symmetry of K0 along axis 0
// the -x side is that computation on the mirrored neighbourhood with the sign of every entry that couples the x component with
// another one flipped (folded into the constants at compile time).  part[c] = sum_e E_e (K0 u_e)[c] over the side's 4 elements.
template<int H, int HP, bool ISO>
__device__ __forceinline__ void gs_row_half(const VtabParam &V, const double *own, const double *oth, int zo, double (&part)[3], double (&uself)[3]) {
    constexpr int sideOff = (H ? -1 : 1) * 9 * 2 * HP;         // from the plane of the node to this side's other plane
    constexpr int eRow = row_e<HP>(1 - H, 0) - row_u<HP>(1, 0, 0); // modulus rows of this side's element layer (relative to own / oth)
    double t[4][3];
    #pragma unroll
    for (int e = 0; e < 4; ++e) { t[e][0] = 0.0; t[e][1] = 0.0; t[e][2] = 0.0; }
    // bit (2 - a) of D / sg / e refers to axis a (the reference's local node numbering, TensorProductSimulator.hh:1532-1651)
    #pragma unroll
    for (int D = 0; D < 8; ++D) {
        #pragma unroll
        for (int sg = 0; sg < 4; ++sg) {
            if (sg & ~D) continue;                         // sg: axes (within D) along which the neighbour lies at -1; never axis 0 here
            const int d0 = (D & 4) ? 1 : 0;
            const int d1 = (D & 2) ? ((sg & 2) ? -1 : 1) : 0;
            const int d2 = (D & 1) ? ((sg & 1) ? -1 : 1) : 0;
            double un[3];
            #pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int ro = ((d1 + 1) * 3 + d) * 2 * HP + (d0 ? sideOff : 0);
                un[d] = (d2 == 0) ? own[ro] : ((d2 < 0) ? oth[ro] : oth[ro + 1]);
            }
            if (D == 0) { uself[0] = un[0]; uself[1] = un[1]; uself[2] = un[2]; }
            #pragma unroll
            for (int e = 0; e < 4; ++e) {
                if ((e & D) != sg) continue;               // element e holds the node at local coordinate e_a: the neighbour at -1 needs e_a = 1
                #pragma unroll
                for (int c = 0; c < 3; ++c) {
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        bool neg = (c != d) && ((((e >> (2 - c)) ^ (e >> (2 - d))) & 1) != 0);
                        double k;
                        if (ISO) { k = V.mag[gs_vclass(D, c, d)]; neg = (neg != gs_vneg(D, c, d)) != (H == 1 && c != d && (c == 0 || d == 0)); }
                        else k = V.v[H][D][3 * c + d + zo];
                        t[e][c] = fma(neg ? -k : k, un[d], t[e][c]);
                    }
                }
            }
        }
    }
    // moduli of this side's 4 elements: element e lies at offset -e_a along axes 1, 2
    double Ee[4];
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ro = eRow + (1 - ((e >> 1) & 1)) * 2 * HP;
        Ee[e] = (e & 1) ? oth[ro] : own[ro];
    }
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
        double acc = 0.0;
        #pragma unroll
        for (int e = 0; e < 4; ++e) acc = fma(Ee[e], t[e][c], acc);
        part[c] = acc;
    }
}
// signed sums of the 4 moduli of one side for the diagonal block M = sum_e E_e K0[(e,c),(e,c')] = vt[0][3c+c'] * sum_e (+-)E_e
template<int H, int HP>
__device__ __forceinline__ void gs_row_modsums(const double *own, const double *oth, double (&sums)[4]) {
    constexpr int eRow = row_e<HP>(1 - H, 0) - row_u<HP>(1, 0, 0);
    double Ee[4];
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ro = eRow + (1 - ((e >> 1) & 1)) * 2 * HP;
        Ee[e] = (e & 1) ? oth[ro] : own[ro];
    }
    const double a = (Ee[0] + Ee[1]) - (Ee[2] + Ee[3]), bb = (Ee[0] - Ee[1]) + (Ee[2] - Ee[3]);
    sums[0] += (Ee[0] + Ee[1]) + (Ee[2] + Ee[3]);              // sum_e E_e
    sums[1] += H ? -a : a;                                     // sum_e (-1)^(e_0 + e_1) E_e
    sums[2] += H ? -bb : bb;                                   // sum_e (-1)^(e_0 + e_2) E_e
    sums[3] += (Ee[0] - Ee[1]) - (Ee[2] - Ee[3]);              // sum_e (-1)^(e_1 + e_2) E_e
}

constexpr int kRowPairs = kRowThreads / 64;                    // warp pairs per block; a pair covers 32 nodes of a colour per trip

// Two WARPS per 32 nodes: warp `pair` takes the +x side, warp `pair + kRowPairs` the -x side (gs_row_half<0 / 1>, so that the
// constants are warp-uniform operands).  The -x warp subtracts its partial sums from the staged b row in place and meets the +x
// warp at a 64-thread named barrier; the +x warp then forms the residual, solves the 3x3 block and writes the node.
template<bool FWD, int HP, bool ISO>
__global__ void __launch_bounds__(kRowThreads, 3)
k_gs3_rows(const __grid_constant__ GridDesc g, const __grid_constant__ VtabParam V, const __grid_constant__ RowPass rp,
           double *u, const double *__restrict__ b, const double *__restrict__ E, const uint8_t *__restrict__ dmask) {
    extern __shared__ __align__(16) double S[];
    pdl_prologue();
    const int tid = threadIdx.x;
    const int x = rp.px + 2 * (int)blockIdx.y, y = rp.py + 2 * (int)blockIdx.x;
    if (x < g.cmpLo || x >= g.cmpHi) return;              // ghost planes of a slab window are received, not computed
    const int nx = g.nn[0], ny = g.nn[1], nz = g.nn[2];
    const long long NN = g.numNodes;

    // Source table: one entry per shared-memory row -- u rows are clamped into the grid (an element outside the grid has modulus
    // zero, so whatever finite value a clamped row holds is multiplied by zero); modulus rows outside the grid are zero-filled.
    __shared__ const double *s_src[kRowArrays];
    __shared__ int s_len[kRowArrays];
    if (tid < kRowArrays) {
        const double *src; int len = nz;
        if (tid < 27) {
            const int p = tid / 9, r = (tid / 3) % 3, c = tid % 3;
            src = u + c * NN + (long long)min(max(x + p - 1, 0), nx - 1) * g.ns[0] + (long long)min(max(y + r - 1, 0), ny - 1) * g.ns[1];
        } else if (tid < 30) {
            src = b + (tid - 27) * NN + (long long)x * g.ns[0] + (long long)y * g.ns[1];
        } else {
            const int ex = x - 1 + ((tid - 30) >> 1), ey = y - 1 + ((tid - 30) & 1);
            const bool rowOk = ex >= 0 && ex < g.ne[0] && ey >= 0 && ey < g.ne[1];
            src = E + (rowOk ? (long long)ex * g.es[0] + (long long)ey * g.es[1] : 0);
            len = rowOk ? g.ne[2] : 0;
        }
        s_src[tid] = src; s_len[tid] = len;
        // pads: z = -1 (odd half, index 0) and z = nz
        double *row = S + tid * 2 * HP;
        row[HP] = 0.0;
        row[(nz & 1) * HP + (nz >> 1) + 1] = 0.0;
    }
    __syncthreads();
    {
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(S);
        const unsigned zh = sbase + 8u * ((tid & 1) * HP + (tid >> 1) + 1);   // position of z = tid; z += 256 moves it by 128 doubles
        const int nzMain = nz & ~(kRowThreads - 1);          // 0 or 256: rows hold at most 2 * HP - 4 <= 260 elements
        if (tid < nzMain) {
            #pragma unroll
            for (int j = 0; j < 30; ++j) cp_async8(zh + (unsigned)(j * 2 * HP * 8), s_src[j] + tid);   // u and b rows: every z valid
            #pragma unroll
            for (int j = 30; j < kRowArrays; ++j) {          // modulus rows: zero beyond the last element / outside the grid
                const bool ok = tid < s_len[j];
                cp_async8_zfill(zh + (unsigned)(j * 2 * HP * 8), ok ? s_src[j] + tid : s_src[j], ok);
            }
        }
        const int rem = nz - nzMain;                       // ragged end of the rows
        if (rem >= 32) {                                   // one partial trip per row
            if (tid < rem) {
                #pragma unroll 1
                for (int j = 0; j < kRowArrays; ++j) {
                    const int z = nzMain + tid, len = s_len[j];
                    cp_async8_zfill(zh + 8u * (unsigned)(j * 2 * HP + nzMain / 2), z < len ? s_src[j] + z : s_src[j], z < len);
                }
            }
        } else
        #pragma unroll 1
        for (int it = tid; it < kRowArrays * rem; it += kRowThreads) {   // a few elements per row: (row, z) pairs spread over the block
            const int j = it / rem, z = nzMain + it % rem;
            const double *src = s_src[j]; const int len = s_len[j];
            cp_async8_zfill(sbase + 8u * (unsigned)(j * 2 * HP + (z & 1) * HP + (z >> 1) + 1), z < len ? src + z : src, z < len);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int pair = warp % kRowPairs, h = warp / kRowPairs;   // h = 0: elements at x (planes x, x+1); 1: elements at x-1 (planes x, x-1)
    const int rot = ((int)blockIdx.x + (int)blockIdx.y) % kRowPairs;   // the pair that takes the ragged end of a colour (spread over the SM sub-partitions)
    const long long nrow = (long long)x * g.ns[0] + (long long)y * g.ns[1];
    #pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
        const int pz = FWD ? ph : 1 - ph;
        const int cnt = (nz + 1 - pz) >> 1;
        #pragma unroll 1
        for (int base = 0; base < cnt; base += 32 * kRowPairs) {
            // a full trip covers 32 nodes per pair; what is left over is taken 32 nodes at a time by pairs rot, rot + 1, ...
            int i;
            if (base + 32 * kRowPairs <= cnt) i = base + pair * 32 + lane;
            else i = base + ((pair - rot + kRowPairs) % kRowPairs) * 32 + lane;
            // hasFullDirichlet nodes are skipped (MultigridSolver.hh:350).  The skip is taken per WARP PAIR (both warps see the same
            // nodes, so they agree); idle lanes compute on node 0 of the row and write nothing.
            const bool inRange = i < cnt;
            bool active = inRange;
            if (!inRange) i = 0;
            const int z = 2 * i + pz;
            const unsigned dm = dmask[nrow + z];
            active = active && dm != 7u;
            if (__ballot_sync(0xffffffffu, active) == 0u) continue;
            const double *own = S + row_u<HP>(1, 0, 0) + pz * HP + i + 1;        // plane x: own[row]: element z of a row
            const double *oth = S + row_u<HP>(1, 0, 0) + (1 - pz) * HP + i + pz; // oth[row]: element z - 1, oth[row + 1]: element z + 1
            // An always-zero, loop-variant offset into the constant table: without it ptxas hoists all 72 constants out of the node
            // loop, runs out of uniform registers and spills them to local memory.
            const int zo = ISO ? 0 : (int)((unsigned)i >> 30);
            // named barrier of this pair and trip: the -x warp only arrives, so consecutive trips of a colour alternate between two
            // barriers (a colour has at most two trips: rows hold <= 2 * HP - 4 nodes); colours are separated by __syncthreads
            const int barId = 1 + pair + kRowPairs * (base ? 1 : 0);
            double part[3], uself[3];
            double *brow = S + row_b<HP>(0) + pz * HP + i + 1;
            if (h) {
                gs_row_half<1, HP, ISO>(V, own, oth, zo, part, uself);
                #pragma unroll
                for (int c = 0; c < 3; ++c) if (inRange) brow[c * 2 * HP] -= part[c];
                __threadfence_block();
                asm volatile("bar.arrive %0, 64;" ::"r"(barId) : "memory");   // producer side: does not wait for the +x warp
                continue;
            }
            gs_row_half<0, HP, ISO>(V, own, oth, zo, part, uself);
            double sums[4] = {0.0, 0.0, 0.0, 0.0};
            gs_row_modsums<0, HP>(own, oth, sums);
            gs_row_modsums<1, HP>(own, oth, sums);
            asm volatile("bar.sync %0, 64;" ::"r"(barId) : "memory");
            if (!active) continue;
            double rhs[3], M[3][3];
            #pragma unroll
            for (int c = 0; c < 3; ++c) rhs[c] = brow[c * 2 * HP] - part[c];
            // M = sum_e E_e K0[(e,c),(e,c')] = vt[0][3c+c'] * sum_e (+-)E_e
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                #pragma unroll
                for (int c2 = c; c2 < 3; ++c2) {
                    const double sE = (c == c2) ? sums[0] : sums[c + c2];
                    double k;
                    if (ISO) { k = V.mag[gs_vclass(0, c, c2)]; if (gs_vneg(0, c, c2)) k = -k; } else k = V.v[0][0][3 * c + c2 + zo];
                    M[c][c2] = k * sE;
                    M[c2][c] = M[c][c2];
                }
            }
            double du[3];
            gs_node_update<3>(M, rhs, dm, FWD, du);
            double *mine = S + row_u<HP>(1, 1, 0) + pz * HP + i + 1;
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double v = uself[c] + du[c];
                u[c * NN + nrow + z] = v;
                mine[c * 2 * HP] = v;
            }
        }
        if (ph == 0) __syncthreads();                      // the second colour reads the first colour's new values of this row
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// NEIGHBOUR FORM of the row-unit smoother (isotropic material on cubic voxels): 243 multiply-adds per node instead of 576.
//
// By the mirror symmetry of K0 the coefficient of u_d(n + delta) in (K u)_c(n) is
//     vt[D][3c+d] * S_cd(delta),   S_cd(delta) = sum over the elements a shared by n and n + delta of (+-) E_a,
// with D the pattern of non-zero offsets, a in {0,1}^3 the position of an incident element (a_k = 1: on the - side of the node along
// axis k) and the sign (-1)^(a_c + a_d) for c != d, + for c == d.  Only 4 sign patterns occur (none, axes 01, 02, 12), so per neighbour
// there are 4 signed sums of at most 8 moduli, and all 27 x 4 of them come out of one three-stage butterfly over the 8 moduli
// (52 additions: stage k either picks the - side, the + side, the sum or the difference along axis k).  For an isotropic material
// vt[D][3c+d] = +-mag[class(D, c, d)] with 10 classes (gs_vclass), so the products S * u are accumulated per (component, class) and
// the 10 magnitudes are applied once at the end: 243 + 52 + 30 FP64 operations per node plus the 3 x 3 solve, against 600+ for the
// element form above (gs_row_half), and one thread per node (no pairing, no named barriers).  Same staging, same visiting order.
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int kNbThreads = 160;   // 5 warps: a colour of a 257-node row has 129 nodes

// variant of the stage along one axis: 0 = the + side element (offset +1), 1 = the - side element (offset -1), 2 = sum, 3 = difference
__host__ __device__ constexpr int nb_variant(int delta, int s) { return delta > 0 ? 0 : (delta < 0 ? 1 : 2 + s); }
// bits (s0, s1, s2) of sign pattern index 0..3: none, axes 01, axes 02, axes 12
__host__ __device__ constexpr int nb_sbit(int sidx, int axis) { return sidx == 0 ? 0 : (sidx == 1 ? (axis != 2) : (sidx == 2 ? (axis != 1) : (axis != 0))); }

template<int HP>
__device__ __forceinline__ void gs_nb_node(const VtabParam &V, const double *own, const double *oth, double (&Ku)[3], double (&uself)[3], double (&M)[3][3]) {
    constexpr int PL = 9 * 2 * HP;                                  // from a node plane to the next one
    // stage 1 (axis 2): T1[v2][a0][a1]
    double T1[4][2][2];
    #pragma unroll
    for (int a0 = 0; a0 < 2; ++a0) {
        #pragma unroll
        for (int a1 = 0; a1 < 2; ++a1) {
            const int ro = row_e<HP>(1 - a0, 1 - a1) - row_u<HP>(1, 0, 0);
            const double ep = own[ro], em = oth[ro];                // element layers z and z - 1
            T1[0][a0][a1] = ep; T1[1][a0][a1] = em; T1[2][a0][a1] = ep + em; T1[3][a0][a1] = ep - em;
        }
    }
    // stage 2 (axis 1): T2[v1][v2][a0]
    double T2[4][4][2];
    #pragma unroll
    for (int v2 = 0; v2 < 4; ++v2) {
        #pragma unroll
        for (int a0 = 0; a0 < 2; ++a0) {
            const double ep = T1[v2][a0][0], em = T1[v2][a0][1];
            T2[0][v2][a0] = ep; T2[1][v2][a0] = em; T2[2][v2][a0] = ep + em; T2[3][v2][a0] = ep - em;
        }
    }
    double A[3][10];
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
        #pragma unroll
        for (int k = 0; k < 10; ++k) A[c][k] = 0.0;
    }
    #pragma unroll
    for (int d0 = -1; d0 <= 1; ++d0) {
        #pragma unroll
        for (int d1 = -1; d1 <= 1; ++d1) {
            #pragma unroll
            for (int d2 = -1; d2 <= 1; ++d2) {
                const int D = ((d0 != 0) << 2) | ((d1 != 0) << 1) | (d2 != 0);
                double un[3];
                #pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int ro = d0 * PL + ((d1 + 1) * 3 + d) * 2 * HP;
                    un[d] = (d2 == 0) ? own[ro] : ((d2 < 0) ? oth[ro] : oth[ro + 1]);
                }
                // stage 3 (axis 0): the 4 signed sums of this neighbour
                double Sv[4]; bool sneg[4];
                #pragma unroll
                for (int si = 0; si < 4; ++si) {
                    const int s0 = nb_sbit(si, 0), s1 = nb_sbit(si, 1), s2 = nb_sbit(si, 2);
                    const int v0 = nb_variant(d0, s0), v1 = nb_variant(d1, s1), v2 = nb_variant(d2, s2);
                    Sv[si] = v0 < 2 ? T2[v1][v2][v0] : (v0 == 2 ? T2[v1][v2][0] + T2[v1][v2][1] : T2[v1][v2][0] - T2[v1][v2][1]);
                    sneg[si] = ((d0 < 0 && s0) != (d1 < 0 && s1)) != (d2 < 0 && s2);
                }
                #pragma unroll
                for (int c = 0; c < 3; ++c) {
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int si = c == d ? 0 : c + d;
                        const bool neg = sneg[si] != gs_vneg(D, c, d);
                        const int k = gs_vclass(D, c, d);
                        A[c][k] = fma(neg ? -Sv[si] : Sv[si], un[d], A[c][k]);
                    }
                }
                if (D == 0) {
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        uself[c] = un[c];
                        #pragma unroll
                        for (int c2 = c; c2 < 3; ++c2) {
                            const double k = V.mag[gs_vclass(0, c, c2)];
                            M[c][c2] = (gs_vneg(0, c, c2) ? -k : k) * Sv[c == c2 ? 0 : c + c2];
                            M[c2][c] = M[c][c2];
                        }
                    }
                }
            }
        }
    }
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
        double acc = 0.0;
        #pragma unroll
        for (int k = 0; k < 10; ++k) acc = fma(V.mag[k], A[c][k], acc);
        Ku[c] = acc;
    }
}

template<bool FWD, int HP>
__global__ void __launch_bounds__(kNbThreads, 3)
k_gs3_nb(const __grid_constant__ GridDesc g, const __grid_constant__ VtabParam V, const __grid_constant__ RowPass rp,
         double *u, const double *__restrict__ b, const double *__restrict__ E, const uint8_t *__restrict__ dmask) {
    extern __shared__ __align__(16) double S[];
    pdl_prologue();
    const int tid = threadIdx.x;
    const int x = rp.px + 2 * (int)blockIdx.y, y = rp.py + 2 * (int)blockIdx.x;
    if (x < g.cmpLo || x >= g.cmpHi) return;              // ghost planes of a slab window are received, not computed
    const int nx = g.nn[0], ny = g.nn[1], nz = g.nn[2];
    const long long NN = g.numNodes;
    const long long nrow = (long long)x * g.ns[0] + (long long)y * g.ns[1];
    if (tid < kRowArrays) {                               // pads: z = -1 (odd half, index 0) and z = nz
        double *row = S + tid * 2 * HP;
        row[HP] = 0.0;
        row[(nz & 1) * HP + (nz >> 1) + 1] = 0.0;
    }
    {
        // u rows are clamped into the grid (an element outside the grid has modulus zero, so whatever finite value a clamped row
        // holds is multiplied by zero); modulus rows outside the grid are zero-filled
        long long uoff[3][3];
        #pragma unroll
        for (int p = 0; p < 3; ++p) {
            #pragma unroll
            for (int r = 0; r < 3; ++r) uoff[p][r] = (long long)min(max(x + p - 1, 0), nx - 1) * g.ns[0] + (long long)min(max(y + r - 1, 0), ny - 1) * g.ns[1];
        }
        long long eoff[2][2]; int elen[2][2];
        #pragma unroll
        for (int lx = 0; lx < 2; ++lx) {
            #pragma unroll
            for (int ly = 0; ly < 2; ++ly) {
                const int ex = x - 1 + lx, ey = y - 1 + ly;
                const bool rowOk = ex >= 0 && ex < g.ne[0] && ey >= 0 && ey < g.ne[1];
                eoff[lx][ly] = rowOk ? (long long)ex * g.es[0] + (long long)ey * g.es[1] : 0;
                elen[lx][ly] = rowOk ? g.ne[2] : 0;
            }
        }
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(S);
        #pragma unroll 1
        for (int z = tid; z < nz; z += kNbThreads) {
            const unsigned zh = sbase + 8u * (unsigned)((z & 1) * HP + (z >> 1) + 1);
            #pragma unroll
            for (int p = 0; p < 3; ++p) {
                #pragma unroll
                for (int r = 0; r < 3; ++r) {
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) cp_async8(zh + (unsigned)(row_u<HP>(p, r, c) * 8), u + c * NN + uoff[p][r] + z);
                }
            }
            #pragma unroll
            for (int c = 0; c < 3; ++c) cp_async8(zh + (unsigned)(row_b<HP>(c) * 8), b + c * NN + nrow + z);
            #pragma unroll
            for (int lx = 0; lx < 2; ++lx) {
                #pragma unroll
                for (int ly = 0; ly < 2; ++ly) {
                    const bool ok = z < elen[lx][ly];
                    cp_async8_zfill(zh + (unsigned)(row_e<HP>(lx, ly) * 8), E + eoff[lx][ly] + (ok ? z : 0), ok);
                }
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    #pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
        const int pz = FWD ? ph : 1 - ph;
        const int cnt = (nz + 1 - pz) >> 1;
        #pragma unroll 1
        for (int i = tid; i < cnt; i += kNbThreads) {
            const int z = 2 * i + pz;
            const unsigned dm = dmask[nrow + z];
            if (dm == 7u) continue;                       // hasFullDirichlet nodes are skipped (MultigridSolver.hh:350)
            const double *own = S + row_u<HP>(1, 0, 0) + pz * HP + i + 1;        // plane x: own[row]: element z of a row
            const double *oth = S + row_u<HP>(1, 0, 0) + (1 - pz) * HP + i + pz; // oth[row]: element z - 1, oth[row + 1]: element z + 1
            double Ku[3], uself[3], M[3][3], rhs[3], du[3];
            gs_nb_node<HP>(V, own, oth, Ku, uself, M);
            const double *brow = S + row_b<HP>(0) + pz * HP + i + 1;
            #pragma unroll
            for (int c = 0; c < 3; ++c) rhs[c] = brow[c * 2 * HP] - Ku[c];
            gs_node_update<3>(M, rhs, dm, FWD, du);
            double *mine = S + row_u<HP>(1, 1, 0) + pz * HP + i + 1;
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double v = uself[c] + du[c];
                u[c * NN + nrow + z] = v;
                mine[c * 2 * HP] = v;
            }
        }
        if (ph == 0) __syncthreads();                      // the second colour reads the first colour's new values of this row
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// TMA-staged neighbour form.  k_gs3_nb above is bound by the shared-memory instruction path (ncu: 27 % of the warp samples issue the
// 34 x 257 eight-byte cp.async of a row unit, mio_throttle is the second stall reason): the parity-split rows need one LDGSTS per
// element.  Here every row arrives RAW (z order) by ONE bulk copy (cp.async.bulk + mbarrier, no LSU instruction at all), and the
// lanes -- which still own every other node of the row -- fetch what they need as 16-byte aligned pairs: a pair holds two of the
// three z-neighbours (z - 1, z, z + 1) of a (row, component), the third is a second load.  Which pair is aligned depends on the
// parity of the row's first element in global memory and on the colour; with odd node counts along axes 1 and 2 (any grid that can
// be coarsened) that parity is (px + py + plane + row + CN * component) mod 2, CN = numNodes mod 2 (0 for a slab window with an even
// number of planes) -- compile-time constants per launch class (templates Q, CN).
// A row starts at element P + o of its buffer (o: parity of its global offset, P = 2): copies start and end on 16-byte boundaries,
// and the slots of z = -1 / z = nz that no copy covers are zeroed by hand (a covered slot holds a neighbouring row's value: finite,
// and always multiplied by a zero modulus).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int kRawP = 2;
// threads / resident blocks per SM by row-buffer length: a colour of a row has (nz + 1) / 2 nodes
__host__ __device__ constexpr int nbt_threads(int RS) { return RS > 140 ? 160 : (RS > 76 ? 96 : 64); }
__host__ __device__ constexpr int nbt_blocks(int RS) { return RS > 140 ? 3 : (RS > 76 ? 5 : 8); }
template<int RS> __host__ __device__ constexpr int raw_u(int p, int r, int c) { return ((p * 3 + r) * 3 + c) * RS; }
template<int RS> __host__ __device__ constexpr int raw_b(int c) { return (27 + c) * RS; }
template<int RS> __host__ __device__ constexpr int raw_e(int lx, int ly) { return (30 + lx * 2 + ly) * RS; }

// values at z - 1, z, z + 1 of a raw row whose element z' sits at row[kRawP + O + z']; z = 2 i + PZ
template<int O, int PZ>
__device__ __forceinline__ void raw_load3(const double *row, int i, double (&v)[3]) {
    if (((PZ + O) & 1) == 0) {
        const double2 a = *reinterpret_cast<const double2 *>(row + kRawP + O + PZ + 2 * i);        // (z, z + 1)
        v[1] = a.x; v[2] = a.y; v[0] = row[kRawP + O + PZ + 2 * i - 1];
    } else {
        const double2 a = *reinterpret_cast<const double2 *>(row + kRawP + O + PZ + 2 * i - 1);    // (z - 1, z)
        v[0] = a.x; v[1] = a.y; v[2] = row[kRawP + O + PZ + 2 * i + 1];
    }
}

template<int RS, int Q, int CN, int PZ>
__device__ __forceinline__ void gs_nbt_node(const VtabParam &V, const double *S, int i, int z, int nz, double (&Ku)[3], double (&uself)[3], double (&M)[3][3]) {
    // stage 1 (axis 2): T1[v2][a0][a1]; a_k = 1: the element on the - side of the node along axis k
    double T1[4][2][2];
    #pragma unroll
    for (int a0 = 0; a0 < 2; ++a0) {
        #pragma unroll
        for (int a1 = 0; a1 < 2; ++a1) {
            const double *row = S + raw_e<RS>(1 - a0, 1 - a1) + kRawP;   // element layers have even row offsets: o = 0
            double ep, em;
            if (PZ == 1) { const double2 a = *reinterpret_cast<const double2 *>(row + 2 * i); em = a.x; ep = a.y; }   // elements z - 1 = 2 i, z
            else { em = row[2 * i - 1]; ep = row[2 * i]; }
            T1[0][a0][a1] = ep; T1[1][a0][a1] = em; T1[2][a0][a1] = ep + em; T1[3][a0][a1] = ep - em;
        }
    }
    double T2[4][4][2];
    #pragma unroll
    for (int v2 = 0; v2 < 4; ++v2) {
        #pragma unroll
        for (int a0 = 0; a0 < 2; ++a0) {
            const double ep = T1[v2][a0][0], em = T1[v2][a0][1];
            T2[0][v2][a0] = ep; T2[1][v2][a0] = em; T2[2][v2][a0] = ep + em; T2[3][v2][a0] = ep - em;
        }
    }
    double A[3][10];
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
        #pragma unroll
        for (int k = 0; k < 10; ++k) A[c][k] = 0.0;
    }
    #pragma unroll
    for (int d0 = -1; d0 <= 1; ++d0) {
        #pragma unroll
        for (int d1 = -1; d1 <= 1; ++d1) {
            double un[3][3];                                     // [component][d2 + 1]
            // parity of the row's global offset: (Q + plane + row + CN * component) mod 2
            if ((Q + d0 + d1) & 1) { raw_load3<1, PZ>(S + raw_u<RS>(d0 + 1, d1 + 1, 0), i, un[0]); raw_load3<1 ^ CN, PZ>(S + raw_u<RS>(d0 + 1, d1 + 1, 1), i, un[1]); raw_load3<1, PZ>(S + raw_u<RS>(d0 + 1, d1 + 1, 2), i, un[2]); }
            else                   { raw_load3<0, PZ>(S + raw_u<RS>(d0 + 1, d1 + 1, 0), i, un[0]); raw_load3<0 ^ CN, PZ>(S + raw_u<RS>(d0 + 1, d1 + 1, 1), i, un[1]); raw_load3<0, PZ>(S + raw_u<RS>(d0 + 1, d1 + 1, 2), i, un[2]); }
            #pragma unroll
            for (int d2 = -1; d2 <= 1; ++d2) {
                const int D = ((d0 != 0) << 2) | ((d1 != 0) << 1) | (d2 != 0);
                double Sv[4]; bool sneg[4];
                #pragma unroll
                for (int si = 0; si < 4; ++si) {
                    const int s0 = nb_sbit(si, 0), s1 = nb_sbit(si, 1), s2 = nb_sbit(si, 2);
                    const int v0 = nb_variant(d0, s0), v1 = nb_variant(d1, s1), v2 = nb_variant(d2, s2);
                    Sv[si] = v0 < 2 ? T2[v1][v2][v0] : (v0 == 2 ? T2[v1][v2][0] + T2[v1][v2][1] : T2[v1][v2][0] - T2[v1][v2][1]);
                    sneg[si] = ((d0 < 0 && s0) != (d1 < 0 && s1)) != (d2 < 0 && s2);
                }
                #pragma unroll
                for (int c = 0; c < 3; ++c) {
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int si = c == d ? 0 : c + d;
                        const bool neg = sneg[si] != gs_vneg(D, c, d);
                        const int k = gs_vclass(D, c, d);
                        A[c][k] = fma(neg ? -Sv[si] : Sv[si], un[d][d2 + 1], A[c][k]);
                    }
                }
                if (D == 0) {
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        uself[c] = un[c][1];
                        #pragma unroll
                        for (int c2 = c; c2 < 3; ++c2) {
                            const double k = V.mag[gs_vclass(0, c, c2)];
                            M[c][c2] = (gs_vneg(0, c, c2) ? -k : k) * Sv[c == c2 ? 0 : c + c2];
                            M[c2][c] = M[c][c2];
                        }
                    }
                }
            }
        }
    }
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
        double acc = 0.0;
        #pragma unroll
        for (int k = 0; k < 10; ++k) acc = fma(V.mag[k], A[c][k], acc);
        Ku[c] = acc;
    }
}

template<bool FWD, int RS, int Q, int CN, int PZ>
__device__ __forceinline__ void gs_nbt_phase(const GridDesc &g, const VtabParam &V, double *S, double *u, const uint8_t *__restrict__ dmask, long long nrow, int tid, unsigned dmFirst) {
    const int nz = g.nn[2];
    const long long NN = g.numNodes;
    const int cnt = (nz + 1 - PZ) >> 1;
    #pragma unroll 1
    for (int i = tid; i < cnt; i += nbt_threads(RS)) {
        const int z = 2 * i + PZ;
        const unsigned dm = i == tid ? dmFirst : (unsigned)dmask[nrow + z];   // first trip: fetched while the rows were in flight
        if (dm == 7u) continue;                           // hasFullDirichlet nodes are skipped (MultigridSolver.hh:350)
        double Ku[3], uself[3], M[3][3], rhs[3], du[3];
        gs_nbt_node<RS, Q, CN, PZ>(V, S, i, z, nz, Ku, uself, M);
        #pragma unroll
        for (int c = 0; c < 3; ++c) rhs[c] = S[raw_b<RS>(c) + kRawP + ((Q + CN * c) & 1) + z] - Ku[c];
        gs_node_update<3>(M, rhs, dm, FWD, du);
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = uself[c] + du[c];
            u[c * NN + nrow + z] = v;
            S[raw_u<RS>(1, 1, c) + kRawP + ((Q + CN * c) & 1) + z] = v;
        }
    }
}

template<bool FWD, int RS, int Q, int CN>
__global__ void __launch_bounds__(nbt_threads(RS), nbt_blocks(RS))
k_gs3_nbt(const __grid_constant__ GridDesc g, const __grid_constant__ VtabParam V, const __grid_constant__ RowPass rp,
          double *u, const double *__restrict__ b, const double *__restrict__ E, const uint8_t *__restrict__ dmask) {
    extern __shared__ __align__(16) double S[];
    __shared__ alignas(8) unsigned long long mbar;
    pdl_prologue();
    const int tid = threadIdx.x;
    const int x = rp.px + 2 * (int)blockIdx.y, y = rp.py + 2 * (int)blockIdx.x;
    if (x < g.cmpLo || x >= g.cmpHi) return;              // ghost planes of a slab window are received, not computed
    const int nx = g.nn[0], ny = g.nn[1], nz = g.nn[2];
    const long long NN = g.numNodes;
    const long long nrow = (long long)x * g.ns[0] + (long long)y * g.ns[1];
    if (tid == 0) mbar_init(&mbar, kRowArrays);
    __syncthreads();
    if (tid < kRowArrays) {
        // one bulk copy per row.  Rows of u outside the grid are mirrored back into it (keeps the parity of the row offset; their
        // values only ever meet zero moduli); element rows outside the grid are zero-filled below.
        const double *base; long long e0, total; int len; bool valid = true;
        if (tid < 27) {
            const int p = tid / 9, r = (tid / 3) % 3, c = tid % 3;
            int xx = x + p - 1, yy = y + r - 1;
            if (xx < 0 || xx >= nx) xx = x - (p - 1);
            if (yy < 0 || yy >= ny) yy = y - (r - 1);
            base = u; e0 = c * NN + (long long)xx * g.ns[0] + (long long)yy * g.ns[1]; total = 3 * NN; len = nz;
        } else if (tid < 30) {
            base = b; e0 = (tid - 27) * NN + nrow; total = 3 * NN; len = nz;
        } else {
            const int ex = x - 1 + ((tid - 30) >> 1), ey = y - 1 + ((tid - 30) & 1);
            valid = ex >= 0 && ex < g.ne[0] && ey >= 0 && ey < g.ne[1];
            base = E; e0 = valid ? (long long)ex * g.es[0] + (long long)ey * g.es[1] : 0; total = g.numElems; len = g.ne[2];
        }
        double *row = S + tid * RS;
        row[kRawP - 1] = 0.0;                             // z = -1 of a row with o = 0 (never covered by a copy)
        if (valid) {
            const int o = (int)(e0 & 1);
            const long long start = e0 - o;
            int cnt = (o + len + 1) & ~1;
            row[kRawP + cnt] = 0.0;                       // z = len of a row with o + len even
            if (start + cnt > total) {                    // the very last row of the array: the rounded-up copy would leave the allocation
                cnt -= 2;
                row[kRawP + cnt] = base[start + cnt];
                row[kRawP + cnt + 1] = 0.0;
            }
            tma_load_1d(row + kRawP, base + start, (unsigned)cnt * 8u, &mbar);
        } else mbar_arrive(&mbar);
    }
    #pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ex = x - 1 + (j >> 1), ey = y - 1 + (j & 1);
        if (!(ex >= 0 && ex < g.ne[0] && ey >= 0 && ey < g.ne[1]))
            for (int k = tid; k < RS; k += nbt_threads(RS)) S[(30 + j) * RS + k] = 0.0;
    }
    // Dirichlet masks of this thread's first node of either colour: fetched while the rows are in flight
    const int zA = 2 * tid + (FWD ? 0 : 1), zB = 2 * tid + (FWD ? 1 : 0);
    const unsigned dmA = zA < nz ? dmask[nrow + zA] : 7u, dmB = zB < nz ? dmask[nrow + zB] : 7u;
    mbar_wait(&mbar, 0);
    __syncthreads();                                      // hand-written pad / tail slots
    if (FWD) {
        gs_nbt_phase<FWD, RS, Q, CN, 0>(g, V, S, u, dmask, nrow, tid, dmA);
        __syncthreads();                                  // the second colour reads the first colour's new values of this row
        gs_nbt_phase<FWD, RS, Q, CN, 1>(g, V, S, u, dmask, nrow, tid, dmB);
    } else {
        gs_nbt_phase<FWD, RS, Q, CN, 1>(g, V, S, u, dmask, nrow, tid, dmA);
        __syncthreads();
        gs_nbt_phase<FWD, RS, Q, CN, 0>(g, V, S, u, dmask, nrow, tid, dmB);
    }
}

static int gs_rows_hp(const GridDesc &g) {
    const int need = g.nn[2] / 2 + 2;
    if (need <= 36) return 36;
    if (need <= 68) return 68;
    if (need <= 132) return 132;
    return 0;
}

static bool gs_rows_iso(const K0Param &K, VtabParam &V);
bool gs_rows_supported(const GridDesc &g, const K0Param &K) {
    static const int mode = [] { const char *e = std::getenv("VF_GS_ROWS"); return e ? std::atoi(e) : 1; }();
    if (mode == 0 || g.N != 3 || !K.walsh || g.bd != 1) return false;
    if (gs_rows_hp(g) == 0) return false;
    if (mode == 2 || g.nn[2] >= 200) return true;
    // Shorter rows: the element form leaves its 256-thread block half idle (128^3: 0.35 vs 0.26 ms per sweep for the per-colour kernel),
    // the neighbour form with its row-length dependent block size does not (128^3: 0.145 ms, 64^3: 0.042 vs 0.085 ms).
    static const bool nbOff = [] { const char *e = std::getenv("VF_GS_NB"); return e && e[0] == '0'; }();
    VtabParam V;
    return !nbOff && g.nn[2] >= 9 && gs_rows_iso(K, V);
}

template<int HP>
static void gs_nb_launch(const LaunchCtx &ctx, const GridDesc &g, const VtabParam &V, const RowPass &rp, double *u, const double *b,
                         const double *E, const uint8_t *dmask, bool forward) {
    const size_t smem = (size_t)kRowArrays * 2 * HP * sizeof(double);
    static PerDeviceFlags attr;
    if (first_use_on_device(attr)) {
        VF_CUDA(cudaFuncSetAttribute(k_gs3_nb<true, HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VF_CUDA(cudaFuncSetAttribute(k_gs3_nb<false, HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(rp.cntY, rp.cntX), block(kNbThreads);
    if (forward) VF_LAUNCH((k_gs3_nb<true, HP>), grid, block, smem, ctx.stream, g, V, rp, u, b, E, dmask);
    else         VF_LAUNCH((k_gs3_nb<false, HP>), grid, block, smem, ctx.stream, g, V, rp, u, b, E, dmask);
    VF_KERNEL_CHECK();
}

// raw-row buffer length (doubles) for rows of nz nodes: elements -1 .. nz at offsets kRawP + o + z, one spare pair for the unused half of a load
static int gs_nbt_rs(const GridDesc &g) {
    const int need = g.nn[2] + 5;
    if (need <= 76) return 76;
    if (need <= 140) return 140;
    if (need <= 262) return 262;
    return 0;
}
// the TMA-staged kernel needs 16-byte aligned arrays and odd node counts per axis (compile-time pair alignment, see above)
static bool gs_nbt_usable(const GridDesc &g, const double *u, const double *b, const double *E) {
    static const bool off = [] { const char *e = std::getenv("VF_GS_TMA"); return e && e[0] == '0'; }();
    if (off || gs_nbt_rs(g) == 0) return false;
    if (!(g.nn[1] & 1) || !(g.nn[2] & 1) || g.ns[2] != 1 || g.ns[1] != g.nn[2] || g.ns[0] != (long long)g.nn[1] * g.nn[2]) return false;
    if ((g.es[1] & 1) || (g.es[0] & 1)) return false;
    return (((uintptr_t)u | (uintptr_t)b | (uintptr_t)E) & 15) == 0;
}
template<int RS, int Q, int CN>
static void gs_nbt_launch(const LaunchCtx &ctx, const GridDesc &g, const VtabParam &V, const RowPass &rp, double *u, const double *b,
                          const double *E, const uint8_t *dmask, bool forward) {
    const size_t smem = (size_t)kRowArrays * RS * sizeof(double);
    static PerDeviceFlags attr;
    if (first_use_on_device(attr)) {
        VF_CUDA(cudaFuncSetAttribute(k_gs3_nbt<true, RS, Q, CN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VF_CUDA(cudaFuncSetAttribute(k_gs3_nbt<false, RS, Q, CN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(rp.cntY, rp.cntX), block(nbt_threads(RS));
    if (forward) VF_LAUNCH((k_gs3_nbt<true, RS, Q, CN>), grid, block, smem, ctx.stream, g, V, rp, u, b, E, dmask);
    else         VF_LAUNCH((k_gs3_nbt<false, RS, Q, CN>), grid, block, smem, ctx.stream, g, V, rp, u, b, E, dmask);
    VF_KERNEL_CHECK();
}

template<int HP, bool ISO>
static void gs_rows_launch(const LaunchCtx &ctx, const GridDesc &g, const VtabParam &V, const RowPass &rp, double *u, const double *b,
                           const double *E, const uint8_t *dmask, bool forward) {
    const size_t smem = (size_t)kRowArrays * 2 * HP * sizeof(double);
    static PerDeviceFlags attr;
    if (first_use_on_device(attr)) {
        VF_CUDA(cudaFuncSetAttribute(k_gs3_rows<true, HP, ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VF_CUDA(cudaFuncSetAttribute(k_gs3_rows<false, HP, ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(rp.cntY, rp.cntX), block(kRowThreads);
    if (forward) VF_LAUNCH((k_gs3_rows<true, HP, ISO>), grid, block, smem, ctx.stream, g, V, rp, u, b, E, dmask);
    else         VF_LAUNCH((k_gs3_rows<false, HP, ISO>), grid, block, smem, ctx.stream, g, V, rp, u, b, E, dmask);
    VF_KERNEL_CHECK();
}

// vt as +-mag[class] (isotropic material, cubic voxels)?  Fills V.mag and returns true if every entry matches to rounding.
static bool gs_rows_iso(const K0Param &K, VtabParam &V) {
    bool have[10] = {}; double scale = 0.0;
    for (int D = 0; D < 8; ++D) for (int k = 0; k < 9; ++k) scale = std::max(scale, std::fabs(K.vt[D][k]));
    for (int D = 0; D < 8; ++D) for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) {
        const int cls = gs_vclass(D, c, d);
        const double v = gs_vneg(D, c, d) ? -K.vt[D][3 * c + d] : K.vt[D][3 * c + d];
        if (!have[cls]) { V.mag[cls] = v; have[cls] = true; }
        else if (std::fabs(V.mag[cls] - v) > 1e-14 * scale) return false;
    }
    static const bool off = [] { const char *e = std::getenv("VF_GS_ISO"); return e && e[0] == '0'; }();
    return !off;
}

void launch_gs_rows_l0(const LaunchCtx &ctx, const GridDesc &g, const K0Param &K, double *u, const double *b, const double *E,
                       const uint8_t *dmask, int cls, bool forward) {
    RowPass rp;
    rp.px = (cls >> 1) & 1; rp.py = cls & 1;
    const int limY = std::min(g.nActive, g.nn[1]);
    if (g.nn[0] - 1 - rp.px < 0 || limY - 1 - rp.py < 0) return;
    rp.cntX = (g.nn[0] - 1 - rp.px) / 2 + 1;
    rp.cntY = (limY - 1 - rp.py) / 2 + 1;
    ProfScope ps(ctx, PC_GS_L0, (double)rp.cntX * rp.cntY * g.nn[2]);
    VtabParam V;
    for (int h = 0; h < 2; ++h) for (int D = 0; D < 8; ++D) for (int k = 0; k < 10; ++k) {
        const int c = k / 3, d = k % 3;
        const bool mirrored = h == 1 && k < 9 && c != d && (c == 0 || d == 0);
        V.v[h][D][k] = mirrored ? -K.vt[D][k] : K.vt[D][k];
    }
    for (double &m : V.mag) m = 0.0;
    const bool iso = gs_rows_iso(K, V);
    static const bool nbForm = [] { const char *e = std::getenv("VF_GS_NB"); return !(e && e[0] == '0'); }();   // 0: element form (gs_row_half)
    if (iso && nbForm && gs_nbt_usable(g, u, b, E)) {
        const int q = (rp.px + rp.py) & 1;
        const int cn = (int)(g.numNodes & 1);                 // does the component offset change the parity of a row's global offset?
#define VF_NBT_CASE(RS_) case RS_: if (q && cn) gs_nbt_launch<RS_, 1, 1>(ctx, g, V, rp, u, b, E, dmask, forward); else if (cn) gs_nbt_launch<RS_, 0, 1>(ctx, g, V, rp, u, b, E, dmask, forward); \
                                   else if (q) gs_nbt_launch<RS_, 1, 0>(ctx, g, V, rp, u, b, E, dmask, forward); else gs_nbt_launch<RS_, 0, 0>(ctx, g, V, rp, u, b, E, dmask, forward); return;
        switch (gs_nbt_rs(g)) { VF_NBT_CASE(76) VF_NBT_CASE(140) VF_NBT_CASE(262) default: break; }
#undef VF_NBT_CASE
    }
#define VF_ROWS_CASE(HP_) case HP_: if (iso && nbForm) gs_nb_launch<HP_>(ctx, g, V, rp, u, b, E, dmask, forward); \
                                    else if (iso) gs_rows_launch<HP_, true>(ctx, g, V, rp, u, b, E, dmask, forward); \
                                    else gs_rows_launch<HP_, false>(ctx, g, V, rp, u, b, E, dmask, forward); break;
    switch (gs_rows_hp(g)) {
        VF_ROWS_CASE(36) VF_ROWS_CASE(68) VF_ROWS_CASE(132)
        default: throw std::runtime_error("launch_gs_rows_l0: row too long for the shared-memory tile");
    }
#undef VF_ROWS_CASE
}

// ---------------------------------------------------------------------------
// FP64 FMA throughput of the device (roofline denominator for the FP64-pipe view): 16 independent register-resident DFMA
// chains per thread, 8 blocks of 256 threads per SM.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double m, double c) {
    double a[16];
    #pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    #pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int r = 0; r < 4; ++r) {
            #pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
        }
    }
    double s = 0.0;
    #pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}

double measure_dfma_peak(cudaStream_t stream) {
    int dev = 0, sms = 0; VF_CUDA(cudaGetDevice(&dev));
    VF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out = nullptr; VF_CUDA(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1; VF_CUDA(cudaEventCreate(&e0)); VF_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        VF_CUDA(cudaEventRecord(e0, stream));
        k_dfma_peak<<<blocks, 256, 0, stream>>>(out, iters, 0.999999, 1e-7);
        VF_CUDA(cudaEventRecord(e1, stream));
        VF_CUDA(cudaEventSynchronize(e1));
        float ms = 0; VF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)blocks * 256 * 64.0 * iters / (ms * 1e-3) / 1e12;
        if (rep > 0) best = std::max(best, rate);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return best;
}

} // namespace vf
