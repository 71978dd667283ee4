// mosaic.cu — the two element-wise ends of the tiled-mosaic exchange (no reference
// counterpart: the reference never tiles). One launch each, so that a rank's step costs the host
// three calls around the NCCL all-gather instead of a dozen tensor operations.
//
//   mb_mosaic_pack    this rank's per-tile detections -> one fixed-size block [rows, 6] =
//                     (x1, y1, x2, y2, score, label) in mosaic coordinates; rows that hold no detection
//                     or fail miso's `score > threshold` (ref:miso/object_detection/inference.py:53)
//                     get label -1. The filter commutes with greedy NMS, so applying it before the
//                     exchange keeps the result and shrinks the seam NMS.
//   mb_mosaic_unpack  the gathered blocks [rows, 6] -> boxes [rows, 4], scores [rows], int64 labels
//                     [rows] for mb_nms (mode 1 ignores negative labels).
#include "common.cuh"

namespace mb {

__global__ void __launch_bounds__(256) k_mosaic_pack(const float4* __restrict__ det_boxes, const float* __restrict__ det_scores,
                                                     const long long* __restrict__ det_labels, const int* __restrict__ det_counts,
                                                     const float2* __restrict__ origins_yx, int tiles, int dpi, float threshold,
                                                     int rows, float* __restrict__ block) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float o[6] = {0.f, 0.f, 0.f, 0.f, 0.f, -1.0f};
    if (i < tiles * dpi) {
        const int t = i / dpi, j = i - t * dpi;
        const float4 b = det_boxes[i];
        const float2 yx = origins_yx[t];
        const float s = det_scores[i];
        // one fp32 add per coordinate (tile origins are integers, exactly representable)
        o[0] = __fadd_rn(b.x, yx.y); o[1] = __fadd_rn(b.y, yx.x); o[2] = __fadd_rn(b.z, yx.y); o[3] = __fadd_rn(b.w, yx.x);
        o[4] = s;
        o[5] = (j < det_counts[t] && s > threshold) ? (float)det_labels[i] : -1.0f;
    }
    float2* dst = reinterpret_cast<float2*>(block + (size_t)i * 6);
    dst[0] = make_float2(o[0], o[1]); dst[1] = make_float2(o[2], o[3]); dst[2] = make_float2(o[4], o[5]);
}

__global__ void __launch_bounds__(256) k_mosaic_unpack(const float* __restrict__ block, int rows, float4* __restrict__ boxes,
                                                       float* __restrict__ scores, long long* __restrict__ labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float2* src = reinterpret_cast<const float2*>(block + (size_t)i * 6);
    const float2 a = src[0], b = src[1], c = src[2];
    boxes[i] = make_float4(a.x, a.y, b.x, b.y);
    scores[i] = c.x;
    labels[i] = (long long)c.y;
}

}  // namespace mb

extern "C" int mb_mosaic_pack(const float* det_boxes, const float* det_scores, const int64_t* det_labels,
                              const int32_t* det_counts, const float* origins_yx, int32_t tiles, int32_t dpi,
                              float threshold, int64_t rows, float* block_out, mb_stream_t stream) {
    if (tiles < 0 || dpi < 1 || rows < (int64_t)tiles * dpi || rows >= (1ll << 31)) return MB_ERR_INVALID_ARG;
    if (rows == 0) return MB_OK;
    if (!block_out || (tiles > 0 && (!det_boxes || !det_scores || !det_labels || !det_counts || !origins_yx))) return MB_ERR_INVALID_ARG;
    MB_CUDA(mb::launch_pdl(mb::k_mosaic_pack, (unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream,
                           reinterpret_cast<const float4*>(det_boxes), det_scores, reinterpret_cast<const long long*>(det_labels),
                           det_counts, reinterpret_cast<const float2*>(origins_yx), tiles, dpi, threshold, (int)rows, block_out));
    return MB_OK;
}

extern "C" int mb_mosaic_unpack(const float* block, int64_t rows, float* boxes_out, float* scores_out, int64_t* labels_out,
                                mb_stream_t stream) {
    if (rows < 0 || rows >= (1ll << 31)) return MB_ERR_INVALID_ARG;
    if (rows == 0) return MB_OK;
    if (!block || !boxes_out || !scores_out || !labels_out) return MB_ERR_INVALID_ARG;
    mb::k_mosaic_unpack<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        block, (int)rows, reinterpret_cast<float4*>(boxes_out), scores_out, reinterpret_cast<long long*>(labels_out));
    MB_LAUNCH_CHECK();
    return MB_OK;
}
