// sort_smem.cuh — block-wide bitonic sort of unique 64-bit keys held in shared memory.
#pragma once
#include "common.cuh"

namespace mb {

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ascending; n must be a power of two; all threads of the block must call.
// A thread owns four consecutive elements: compare-exchange steps at distance 1 and 2 stay inside the thread, distances
// 4 .. 64 are exchanges with another lane of the warp (shuffles), only distances >= 128 go through shared memory with
// a barrier each. At n = 1024 that is 6 barrier steps and 4 register phases instead of 55 barrier steps, run by a
// quarter of the threads with four independent exchanges in flight each.
__device__ __forceinline__ void cmpx(unsigned long long& lo, unsigned long long& hi, bool up) {
    const unsigned long long a = lo, b = hi;
    const bool sw = (a > b) == up;
    lo = sw ? b : a; hi = sw ? a : b;
}

__device__ inline void bitonic_sort_smem(unsigned long long* a, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (n < 128 || (nt & 31) != 0) {         // tiny inputs / odd block shapes: the plain network
        for (int k = 2; k <= n; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (int i = tid; i < (n >> 1); i += nt) {
                    const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));   // j is a power of two
                    const int hi = lo + j;
                    const bool up = ((lo & k) == 0);
                    const unsigned long long x = a[lo], y = a[hi];
                    if ((x > y) == up) { a[lo] = y; a[hi] = x; }
                }
            }
        }
        __syncthreads();
        return;
    }
    // register phase: all steps at distance <= 64 of the levels kfrom .. kto (virtual thread v owns elements 4v .. 4v+3;
    // n / 4 and nt are multiples of 32, so whole warps iterate together and the shuffles see all their lanes)
    auto reg_steps = [&](int kfrom, int kto) {
        __syncthreads();
        for (int v = tid; v < (n >> 2); v += nt) {
            ulonglong2* p = reinterpret_cast<ulonglong2*>(a + 4 * v);
            const ulonglong2 p0 = p[0], p1 = p[1];
            unsigned long long x0 = p0.x, x1 = p0.y, x2 = p1.x, x3 = p1.y;
            for (int k = kfrom; k <= kto; k <<= 1) {
                const bool up = ((4 * v) & k) == 0;          // k >= 4: the same for the thread's four elements
                for (int j = min(k >> 1, 64); j >= 4; j >>= 1) {
                    const int d = j >> 2;                    // lane distance
                    const bool take_min = (((v & d) == 0) == up);
                    const unsigned long long y0 = __shfl_xor_sync(0xffffffffu, x0, d), y1 = __shfl_xor_sync(0xffffffffu, x1, d);
                    const unsigned long long y2 = __shfl_xor_sync(0xffffffffu, x2, d), y3 = __shfl_xor_sync(0xffffffffu, x3, d);
                    x0 = take_min ? (x0 < y0 ? x0 : y0) : (x0 > y0 ? x0 : y0);
                    x1 = take_min ? (x1 < y1 ? x1 : y1) : (x1 > y1 ? x1 : y1);
                    x2 = take_min ? (x2 < y2 ? x2 : y2) : (x2 > y2 ? x2 : y2);
                    x3 = take_min ? (x3 < y3 ? x3 : y3) : (x3 > y3 ? x3 : y3);
                }
                if (k >= 4) {
                    cmpx(x0, x2, up); cmpx(x1, x3, up);      // distance 2
                    cmpx(x0, x1, up); cmpx(x2, x3, up);      // distance 1
                } else {                                     // level k = 2: pairs (0,1) ascending, (2,3) descending
                    cmpx(x0, x1, true); cmpx(x2, x3, false);
                }
            }
            p[0] = make_ulonglong2(x0, x1); p[1] = make_ulonglong2(x2, x3);
        }
    };
    reg_steps(2, 128);
    for (int k = 256; k <= n; k <<= 1) {
        for (int j = k >> 1; j >= 128; j >>= 1) {
            __syncthreads();
            for (int i = tid; i < (n >> 1); i += nt) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo + j;
                const bool up = ((lo & k) == 0);
                const unsigned long long x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
        }
        reg_steps(k, k);
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

// merged_rank when only ranks below `cap` matter (top-cap of the merge): returns a value >= cap as soon as the rank is
// known to reach it. Element j of its own run has rank >= j; in every other run at most cap - rank keys can still
// precede it without pushing it out, so one probe at that position either ends the search or bounds it.
__device__ __forceinline__ int merged_rank_capped(const unsigned long long* keys, const int* off, int S, int s, int j, int cap) {
    if (j >= cap) return cap;
    const unsigned long long key = keys[off[s] + j];
    int rank = j;
    for (int r = 0; r < S; ++r) {
        if (r == s) continue;
        const unsigned long long* run = keys + off[r];
        const int len = off[r + 1] - off[r];
        const int lim = min(len, cap - rank);                  // more than this many smaller keys -> rank >= cap
        if (lim <= 0) return cap;
        if (run[lim - 1] < key) {                              // at least lim keys of this run precede
            if (lim == cap - rank) return cap;
            rank += lim;                                       // lim == len: the whole run precedes
            continue;
        }
        rank += lower_bound_smem(run, lim, key);
        if (rank >= cap) return cap;
    }
    return rank;
}

}  // namespace mb
