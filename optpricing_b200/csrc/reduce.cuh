// reduce.cuh -- warp-shuffle + shared-memory block reductions.  Only per-block
// partials ever reach HBM; a fixed-order finishing pass keeps every sum
// bit-reproducible for a given launch geometry.
#pragma once
#include <cuda_runtime.h>

namespace optmc {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum N values per thread over the block.  Result valid in thread 0.
// scratch: N * 32 doubles of shared memory.  All threads must call.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) scratch[k * 32 + warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double s = lane < n_warps ? scratch[k * 32 + lane] : 0.0;
            v[k] = warp_sum(s);
        }
    }
    __syncthreads();
}

}  // namespace optmc
