// boxops.cu — element-wise box operators with the reference's fp32 operation order.
// float4-vectorised: one thread per box, 16-byte coalesced loads and stores.
//   mb_box_decode    tv:models/detection/_utils.py:183-224   (BoxCoder.decode_single)
//   mb_clip_boxes    tv:ops/boxes.py:149-182
//   mb_box_convert   tv:ops/boxes.py:185-270, tv:ops/_box_convert.py:5-81
//   mb_remove_small  tv:ops/boxes.py:123-146
//   mb_resize_boxes  tv:models/detection/transform.py:306-319
//   mb_grid_anchors  tv:models/detection/anchor_utils.py:84-113
#include "boxmath.cuh"

namespace mb {

__global__ void k_box_decode(const float* __restrict__ rel, const float4* __restrict__ boxes, long long M, int C,
                             DecodeWeights w, float4* __restrict__ out) {
    const long long total = M * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / C;
        const float4 d = reinterpret_cast<const float4*>(rel)[i];  // rel_codes [M, 4*C]: class-major groups of 4
        out[i] = decode_box(boxes[m], d, w);
    }
}

__global__ void k_clip_boxes(const float4* __restrict__ in, long long n, float h, float w, float4* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = clip_box(in[i], h, w);
}

__global__ void k_box_convert(const float4* __restrict__ in, long long n, int in_fmt, int out_fmt, float4* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 b = in[i];
        // to xyxy  (tv:ops/_box_convert.py:5-81)
        if (in_fmt == 1) b = make_float4(b.x, b.y, __fadd_rn(b.x, b.z), __fadd_rn(b.y, b.w));
        else if (in_fmt == 2) {
            const float hw = __fmul_rn(0.5f, b.z), hh = __fmul_rn(0.5f, b.w);
            b = make_float4(__fsub_rn(b.x, hw), __fsub_rn(b.y, hh), __fadd_rn(b.x, hw), __fadd_rn(b.y, hh));
        }
        if (out_fmt == 1) b = make_float4(b.x, b.y, __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        else if (out_fmt == 2)
            b = make_float4(__fdiv_rn(__fadd_rn(b.x, b.z), 2.0f), __fdiv_rn(__fadd_rn(b.y, b.w), 2.0f),
                            __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        out[i] = b;
    }
}

__global__ void k_resize_boxes(const float4* __restrict__ in, long long n, float rh, float rw, float4* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = resize_box(in[i], rh, rw);
}

// stable compaction in one CTA (auxiliary op; the fused paths filter in their own kernels)
__global__ void __launch_bounds__(1024) k_remove_small(const float4* __restrict__ boxes, long long n, float min_size,
                                                      long long* __restrict__ keep, long long* __restrict__ count) {
    __shared__ int warp_cnt[32];
    __shared__ long long base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < n; i0 += 1024) {
        const long long i = i0 + tid;
        const bool ok = (i < n) && box_not_small(boxes[i], min_size);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { const int c = warp_cnt[w]; pre += (w < wid) ? c : 0; tot += c; }
        if (ok) keep[base + pre + __popc(m & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (tid == 0) base += tot;
        __syncthreads();
    }
    if (tid == 0) *count = base;
}

__global__ void k_grid_anchors(BaseAnchors ba, int num_base, int gh, int gw, int sh, int sw, float4* __restrict__ out) {
    const long long total = (long long)gh * gw * num_base;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i % num_base);
        const long long loc = i / num_base;
        out[i] = grid_anchor(ba.v[a], (int)(loc / gw), (int)(loc % gw), sh, sw);
    }
}

static inline int grid_for(long long n) { return (int)min((long long)num_sms() * 16, max(1ll, ceil_div64(n, 256))); }

}  // namespace mb

using namespace mb;

extern "C" int mb_box_decode(const float* rel_codes, const float* boxes, int64_t num_boxes, int32_t num_classes,
                             float wx, float wy, float ww, float wh, float clip, float* out, mb_stream_t s) {
    if (num_boxes < 0 || num_classes < 1) return MB_ERR_INVALID_ARG;
    if (num_boxes == 0) return MB_OK;
    if (!rel_codes || !boxes || !out) return MB_ERR_INVALID_ARG;
    DecodeWeights w{wx, wy, ww, wh, clip};
    k_box_decode<<<grid_for(num_boxes * num_classes), 256, 0, (cudaStream_t)s>>>(rel_codes, (const float4*)boxes, num_boxes,
                                                                                  num_classes, w, (float4*)out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

extern "C" int mb_clip_boxes(const float* boxes, int64_t n, float height, float width, float* out, mb_stream_t s) {
    if (n < 0) return MB_ERR_INVALID_ARG;
    if (n == 0) return MB_OK;
    if (!boxes || !out) return MB_ERR_INVALID_ARG;
    k_clip_boxes<<<grid_for(n), 256, 0, (cudaStream_t)s>>>((const float4*)boxes, n, height, width, (float4*)out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

extern "C" int mb_box_convert(const float* boxes, int64_t n, int32_t in_fmt, int32_t out_fmt, float* out, mb_stream_t s) {
    if (n < 0 || in_fmt < 0 || in_fmt > 2 || out_fmt < 0 || out_fmt > 2) return MB_ERR_INVALID_ARG;
    if (n == 0) return MB_OK;
    if (!boxes || !out) return MB_ERR_INVALID_ARG;
    if (in_fmt == out_fmt) {
        MB_CUDA(cudaMemcpyAsync(out, boxes, sizeof(float) * 4 * n, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
        return MB_OK;
    }
    k_box_convert<<<grid_for(n), 256, 0, (cudaStream_t)s>>>((const float4*)boxes, n, in_fmt, out_fmt, (float4*)out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

extern "C" int mb_remove_small(const float* boxes, int64_t n, float min_size, int64_t* keep_out, int64_t* count_out,
                               mb_stream_t s) {
    if (n < 0 || !count_out) return MB_ERR_INVALID_ARG;
    if (n > 0 && (!boxes || !keep_out)) return MB_ERR_INVALID_ARG;
    k_remove_small<<<1, 1024, 0, (cudaStream_t)s>>>((const float4*)boxes, n, min_size, (long long*)keep_out,
                                                    (long long*)count_out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

extern "C" int mb_resize_boxes(const float* boxes, int64_t n, float ratio_h, float ratio_w, float* out, mb_stream_t s) {
    if (n < 0) return MB_ERR_INVALID_ARG;
    if (n == 0) return MB_OK;
    if (!boxes || !out) return MB_ERR_INVALID_ARG;
    k_resize_boxes<<<grid_for(n), 256, 0, (cudaStream_t)s>>>((const float4*)boxes, n, ratio_h, ratio_w, (float4*)out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

extern "C" int mb_grid_anchors(const float* base_anchors_host, int32_t num_base, int32_t gh, int32_t gw, int32_t sh,
                               int32_t sw, float* out, mb_stream_t s) {
    if (!base_anchors_host || num_base < 1 || num_base > MB_MAX_ANCHORS_PER_LOC || gh < 0 || gw < 0) return MB_ERR_INVALID_ARG;
    if ((long long)gh * gw == 0) return MB_OK;
    if (!out) return MB_ERR_INVALID_ARG;
    BaseAnchors ba;
    for (int a = 0; a < num_base; ++a)
        ba.v[a] = make_float4(base_anchors_host[4 * a], base_anchors_host[4 * a + 1], base_anchors_host[4 * a + 2],
                              base_anchors_host[4 * a + 3]);
    k_grid_anchors<<<grid_for((long long)gh * gw * num_base), 256, 0, (cudaStream_t)s>>>(ba, num_base, gh, gw, sh, sw, (float4*)out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}
