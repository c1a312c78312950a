// vf_reduce.cuh -- deterministic block/grid reductions shared by the kernels.
#pragma once
#include <cuda_runtime.h>

namespace vf {

constexpr int kReduceMaxBlocks = 1 << 18; // partials capacity of a reduction scratch buffer (doubles) + counter

__device__ __forceinline__ double warp_sum(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum `v` over the thread block (any 1D/2D/3D block whose size is a multiple of 32, <= 1024).
// Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s_part[32];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5;
    v = warp_sum(v);
    __syncthreads(); // protect s_part reuse across consecutive calls
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = nthreads >> 5;
        v = (lane < nw) ? s_part[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// Grid-wide deterministic sum: every block contributes `v` (per thread); the last block to finish adds
// the per-block partials in index order and writes result[0].  scratch: kReduceMaxBlocks doubles followed
// by one unsigned counter (zero-initialised once; atomicInc wraps it back to zero).
__device__ __forceinline__ void grid_sum(double v, double *scratch, double *result) {
    __shared__ bool s_last;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    double bs = block_sum(v);
    if (tid == 0) {
        scratch[bid] = bs;
        __threadfence();
        unsigned *counter = reinterpret_cast<unsigned *>(scratch + kReduceMaxBlocks);
        unsigned prev = atomicInc(counter, nblocks - 1);
        s_last = (prev == nblocks - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double s = 0.0;
        for (unsigned i = tid; i < nblocks; i += nthreads) s += __ldcg(scratch + i);
        s = block_sum(s);
        if (tid == 0) result[0] = s;
    }
}

} // namespace vf
