// boxmath.cuh — per-box device math shared by boxops.cu, rpn.cu, detpost.cu, crop.cu.
// Every function keeps the reference's fp32 operation order (one rounding per op, no FMA
// contraction); only exp() differs from the CPU path (<= 2 ulp), hence the 1e-5 box tolerance.
#pragma once
#include "common.cuh"

namespace mb {

struct DecodeWeights { float wx, wy, ww, wh, clip; };
struct BaseAnchors { float4 v[MB_MAX_ANCHORS_PER_LOC]; };

// BoxCoder.decode_single (tv:models/detection/_utils.py:183-224); b = reference box, d = (dx,dy,dw,dh) raw codes
__device__ __forceinline__ float4 decode_box(const float4 b, const float4 d, const DecodeWeights w) {
    const float width = __fsub_rn(b.z, b.x), height = __fsub_rn(b.w, b.y);
    const float ctr_x = __fadd_rn(b.x, __fmul_rn(0.5f, width));
    const float ctr_y = __fadd_rn(b.y, __fmul_rn(0.5f, height));
    const float dx = __fdiv_rn(d.x, w.wx), dy = __fdiv_rn(d.y, w.wy);
    float dw = __fdiv_rn(d.z, w.ww), dh = __fdiv_rn(d.w, w.wh);
    dw = (dw != dw) ? dw : fminf(dw, w.clip);  // torch.clamp(max=) propagates NaN
    dh = (dh != dh) ? dh : fminf(dh, w.clip);
    const float pcx = __fadd_rn(__fmul_rn(dx, width), ctr_x);
    const float pcy = __fadd_rn(__fmul_rn(dy, height), ctr_y);
    const float pw = __fmul_rn(expf(dw), width);
    const float ph = __fmul_rn(expf(dh), height);
    const float cw = __fmul_rn(0.5f, pw), ch = __fmul_rn(0.5f, ph);
    return make_float4(__fsub_rn(pcx, cw), __fsub_rn(pcy, ch), __fadd_rn(pcx, cw), __fadd_rn(pcy, ch));
}

// clip_boxes_to_image (tv:ops/boxes.py:149-182): clamp(min=0, max=size)
__device__ __forceinline__ float clamp_keep_nan(float v, float lo, float hi) {
    return (v != v) ? v : fminf(fmaxf(v, lo), hi);
}
__device__ __forceinline__ float4 clip_box(const float4 b, float h, float w) {
    return make_float4(clamp_keep_nan(b.x, 0.f, w), clamp_keep_nan(b.y, 0.f, h), clamp_keep_nan(b.z, 0.f, w),
                       clamp_keep_nan(b.w, 0.f, h));
}

// remove_small_boxes predicate (tv:ops/boxes.py:143-145)
__device__ __forceinline__ bool box_not_small(const float4 b, float min_size) {
    return (__fsub_rn(b.z, b.x) >= min_size) && (__fsub_rn(b.w, b.y) >= min_size);
}

// resize_boxes (tv:models/detection/transform.py:306-319)
__device__ __forceinline__ float4 resize_box(const float4 b, float ratio_h, float ratio_w) {
    return make_float4(__fmul_rn(b.x, ratio_w), __fmul_rn(b.y, ratio_h), __fmul_rn(b.z, ratio_w), __fmul_rn(b.w, ratio_h));
}

// grid_anchors (tv:models/detection/anchor_utils.py:100-113): int32 shift + fp32 base, exact in fp32
__device__ __forceinline__ float4 grid_anchor(const float4 base, int h, int w, int stride_h, int stride_w) {
    const float sx = (float)(w * stride_w), sy = (float)(h * stride_h);
    return make_float4(__fadd_rn(sx, base.x), __fadd_rn(sy, base.y), __fadd_rn(sx, base.z), __fadd_rn(sy, base.w));
}

__device__ __forceinline__ float sigmoid_rn(float x) {  // torch.sigmoid: 1 / (1 + exp(-x))
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

}  // namespace mb
