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

// Number of keys < key in the ascending run a[0..n) (keys are unique).
__device__ __forceinline__ int lower_bound_smem(const unsigned long long* a, int n, unsigned long long key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Rank of element j of run s in the merge of S ascending runs stored back to back in `keys`
// (run r occupies [off[r], off[r+1])): its own index plus, for every other run, the number of
// keys below it. Replaces a full sort when the inputs are already sorted per run.
__device__ __forceinline__ int merged_rank(const unsigned long long* keys, const int* off, int S, int s, int j) {
    const unsigned long long key = keys[off[s] + j];
    int rank = j;
    for (int r = 0; r < S; ++r)
        if (r != s) rank += lower_bound_smem(keys + off[r], off[r + 1] - off[r], key);
    return rank;
}

}  // namespace mb
