// roi_common.cuh — RoI geometry shared by the RoIAlign kernels (roi_align.cu, roi_align_tma.cu).
// Arithmetic of tv-csrc:ops/cpu/roi_align_kernel.cpp:393: one fp32 rounding per operation.
#pragma once
#include <math.h>

#include "common.cuh"

namespace mb {

struct Tap {       // one sample along one axis
    int lo, hi;    // pixel indices
    float l, h;    // weights of hi / lo pixel (ly, hy in the reference)
    int valid;
};

struct RoiGeom {
    int level, batch;   // batch < 0: dead row (zeros)
    int H, W;
    float start_h, start_w, bin_h, bin_w;
    int grid_h, grid_w;
    float count;
};

__device__ __forceinline__ int roi_level(const float* r, const mb_roi_align_params& p) {
    // LevelMapper (tv:ops/poolers.py:73-84) as monotone thresholds on the fp32 box area
    const float area = __fmul_rn(__fsub_rn(r[3], r[1]), __fsub_rn(r[4], r[2]));
    int lvl = 0;
    for (int i = 0; i + 1 < p.num_levels; ++i) lvl += (area >= p.level_thresholds[i]) ? 1 : 0;
    return lvl;
}

// row k of the RoI array as (batch, x1, y1, x2, y2) for either layout
__device__ __forceinline__ void load_roi(const float* rois, long long k, const mb_roi_align_params& p, float r[5]) {
    if (p.boxes_per_image > 0) {
        const float4 b = reinterpret_cast<const float4*>(rois)[k];
        const int n = (int)(k / p.boxes_per_image);
        const bool live = p.box_counts == nullptr || (int)(k - (long long)n * p.boxes_per_image) < p.box_counts[n];
        r[0] = live ? (float)n : -1.0f;
        r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
    } else {
        const float* q = rois + k * 5;
        r[0] = q[0]; r[1] = q[1]; r[2] = q[2]; r[3] = q[3]; r[4] = q[4];
    }
}

__device__ __forceinline__ void roi_geometry(const float* r, const mb_roi_align_params& p, RoiGeom& g) {
    g.level = roi_level(r, p);
    g.batch = (int)r[0];
    g.H = p.height[g.level];
    g.W = p.width[g.level];
    const float scale = p.spatial_scale[g.level];
    const float off = p.aligned ? 0.5f : 0.0f;
    const float sw = __fsub_rn(__fmul_rn(r[1], scale), off);
    const float sh = __fsub_rn(__fmul_rn(r[2], scale), off);
    const float ew = __fsub_rn(__fmul_rn(r[3], scale), off);
    const float eh = __fsub_rn(__fmul_rn(r[4], scale), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    g.start_h = sh; g.start_w = sw;
    g.bin_h = __fdiv_rn(rh, (float)p.pooled_h);
    g.bin_w = __fdiv_rn(rw, (float)p.pooled_w);
    g.grid_h = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)p.pooled_h));
    g.grid_w = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)p.pooled_w));
    g.count = (float)max(g.grid_h * g.grid_w, 1);
}

// one axis of bilinear_interpolate / pre_calc_for_bilinear_interpolate
__device__ __forceinline__ Tap make_tap(float start, float bin, int p_idx, int i_idx, int grid, int size) {
    Tap t;
    float v = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p_idx, bin)),
                        __fdiv_rn(__fmul_rn(__fadd_rn((float)i_idx, 0.5f), bin), (float)grid));
    t.valid = !(v < -1.0f || v > (float)size);
    if (v <= 0.0f) v = 0.0f;
    int lo = (int)v, hi;
    if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
    t.lo = lo; t.hi = hi;
    t.l = __fsub_rn(v, (float)lo);
    t.h = __fsub_rn(1.0f, t.l);
    if (!t.valid) { t.lo = 0; t.hi = 0; t.l = 0.f; t.h = 0.f; }
    return t;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Rotate (a0..a3) so that element t of the result is a[(rot + t) & 3]: two select stages instead of a
// select chain per element. Lane groups of 8 use different rotations, which spreads the four lanes
// that share a bank (row pitch odd, 4 rows per lane) over four banks.
__device__ __forceinline__ void rotate4(float4& a, int rot) {
    if (rot & 1) { const float t = a.x; a.x = a.y; a.y = a.z; a.z = a.w; a.w = t; }
    if (rot & 2) { float t = a.x; a.x = a.z; a.z = t; t = a.y; a.y = a.w; a.w = t; }
}


// Packed fp32x2 multiply with explicit round-to-nearest (sm_100 FMUL2). The additions of the exact mode stay
// scalar on purpose: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad false (it
// does honour the rounding modifiers of scalar adds), and a contracted sum is not the reference's arithmetic.
// (Packed ops do not raise FP32 throughput here — an FMUL2 occupies the pipe like two FMULs — they only save
// issue slots; a fully packed exact variant via doubled weights and fma(p', 0.5, t) measured no faster.)
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
}


// distinct pixel indices of one bin along one axis, in the slot order the patterns expect
__device__ __forceinline__ int axis_pattern(const Tap& A, const Tap& B, int idx[4]) {
    idx[0] = A.lo; idx[1] = A.hi; idx[2] = B.lo; idx[3] = B.hi;
    if (B.lo == A.lo && B.hi == A.hi) return 0;                 // A: same cell
    if (B.lo == A.hi) { idx[2] = B.hi; return 1; }              // B: the samples share one pixel
    return 2;                                                   // C: four pixels
}


// Packed fp32x2 fused multiply-add with explicit round-to-nearest (sm_100 FFMA2). fma(p, 1.0f, t) with an OPAQUE 1.0f (a
// kernel parameter) is an exactly rounded add that ptxas cannot contract with the multiply that produced p.
__device__ __forceinline__ float2 fma2_rn(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return reinterpret_cast<float2&>(d);
}


}  // namespace mb
