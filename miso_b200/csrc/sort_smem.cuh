// sort_smem.cuh — block-wide bitonic sort of unique 64-bit keys held in shared memory.
#pragma once
#include "common.cuh"

namespace mb {

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ascending; n must be a power of two; all threads of the block must call
__device__ inline void bitonic_sort_smem(unsigned long long* a, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = tid; i < (n >> 1); i += nt) {
                // i-th compare-exchange pair of this stage
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));   // j is a power of two
                const int hi = lo + j;
                const bool up = ((lo & k) == 0);
                const unsigned long long x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

}  // namespace mb
