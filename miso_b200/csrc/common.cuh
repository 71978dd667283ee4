// common.cuh — shared device helpers for libmisob200 (sm_100a only).
#pragma once
#include <cstdlib>
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

// Opt a kernel in to `bytes` of dynamic shared memory: the attribute is per device and sticky, so it is only set when a
// call site needs more than it has already asked for on that device (not on every launch).
#define MB_DYN_SMEM(kernel, bytes)                                                                        \
    do {                                                                                                  \
        static int have__[64] = {};                                                                       \
        int dev__ = 0;                                                                                    \
        MB_CUDA(cudaGetDevice(&dev__));                                                                   \
        const int want__ = (int)(bytes);                                                                  \
        if (want__ > have__[dev__ & 63]) {                                                                \
            MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want__));   \
            have__[dev__ & 63] = want__;                                                                  \
        }                                                                                                 \
    } while (0)

namespace mb {

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// The per-batch path is a chain of ~16 short dependent kernels; launched back to back each boundary costs the full
// drain + launch latency. Kernels of the chain begin with pdl_wait() (griddepcontrol.wait: returns once the preceding
// grid has completed and its writes are visible) followed by pdl_trigger() (griddepcontrol.launch_dependents), and
// are launched through launch_pdl() with programmatic stream serialization: the next grid's CTAs are scheduled while
// the current one drains and sit at their pdl_wait(). Only kernels that start with pdl_wait() may be launched this
// way; MB_PDL=0 turns the attribute off (A/B switch), the waits are then no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

inline bool pdl_enabled() {
    static const bool v = [] { const char* e = getenv("MB_PDL"); return !(e != nullptr && e[0] == '0'); }();
    return v;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// SM count of the current device (148 on a B200), asked once per device: grids are sized in multiples of it
inline int num_sms() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev];
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;                 // racing first calls store the same value
    }
    return v;
}

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
