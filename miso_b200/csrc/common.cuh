// common.cuh — shared device helpers for libmisob200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/misob200.h"

#define MB_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

#define MB_CUDA(call)                                       \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

namespace mb {

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

// Ascending 32-bit key whose order equals "score descending, NaN first, -0 == +0".
// torch.sort(descending=True, stable=True) and torch.topk treat NaN as the largest value.
__device__ __forceinline__ uint32_t desc_score_key(float s) {
    if (s != s) return 0u;
    uint32_t u = __float_as_uint(s);
    if (u == 0x80000000u) u = 0u;
    uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;  // >= 0x007FFFFF for +inf, so 0 is reserved for NaN
}

// Warp-aggregated slot allocation: one atomic per warp instead of one per lane. Every lane of the
// warp must call it (pred says whether this lane wants a slot); returns the lane's slot.
template <typename CounterT>
__device__ __forceinline__ int warp_alloc_slot(CounterT* counter, bool pred) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// Bump allocator over a caller-provided workspace.
struct Carver {
    char* base;
    size_t off;
    size_t cap;
    __host__ Carver(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    template <typename T>
    __host__ T* take(size_t n) {
        off = align_up(off, 256);
        T* r = (T*)(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    __host__ bool ok() const { return base != nullptr && off <= cap; }
};

// smallest fp32 f with (double)f > thr  ==>  ((double)iou > thr)  <=>  (iou >= f) for fp32 iou
__host__ inline float strict_gt_threshold(double thr) {
    float f = (float)thr;
    if (!((double)f > thr)) f = nextafterf(f, INFINITY);
    return f;
}

}  // namespace mb
