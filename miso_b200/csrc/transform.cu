// transform.cu — the input side of the model in one launch.
//
// Replaces, for uint8 HWC images already on the device: ToTensor (ref:miso/object_detection/inference.py:117,
// x / 255), GeneralizedRCNNTransform.normalize ((x - mean) / std, tv:models/detection/transform.py:165-173),
// .resize (bilinear F.interpolate with recompute_scale_factor, align_corners=False, :175-201 and
// _resize_image_and_masks :23-70) and .batch_images (zero-padded to a multiple of size_divisible, :231-255).
// The output sizes are computed on the host exactly as the reference does (fp32 scale factor, floor).
//
// Per channel a 256-entry table holds ((v / 255) - mean) / std for every uint8 value (the same three fp32
// operations the reference applies per pixel), so the bilinear blend reads normalised taps straight
// from shared memory. Blend arithmetic = ATen's CPU upsample_bilinear2d as compiled (see paste.cu):
// bit-identical to the reference on identical inputs. A thread produces 4 consecutive output pixels of
// all channels: 16-byte stores per plane; the padding is written as zeros by the same kernel.
#include "common.cuh"

namespace mb {

constexpr int kTfThreads = 256;

struct TfDev {
    const uint8_t* img[MB_MAX_IMAGES];
    int in_h[MB_MAX_IMAGES], in_w[MB_MAX_IMAGES], out_h[MB_MAX_IMAGES], out_w[MB_MAX_IMAGES];
    float mean[4], stdv[4];
    int channels, pad_h, pad_w;
};

__device__ __forceinline__ void tf_tap(float scale, int d, int size, int& i0, int& i1, float& w0, float& w1) {
    float src = fmaf(scale, __fadd_rn((float)d, 0.5f), -0.5f);       // area_pixel_compute_source_index (contracted build)
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, size - 1);
    w1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
    w0 = __fsub_rn(1.0f, w1);
    i1 = i0 + (i0 < size - 1 ? 1 : 0);
}

__global__ void __launch_bounds__(kTfThreads) k_image_transform(const TfDev p, float* __restrict__ out) {
    __shared__ float lut[4][256];
    const int n = blockIdx.z, y = blockIdx.y, tid = threadIdx.x;
    const int C = p.channels;
    for (int i = tid; i < C * 256; i += kTfThreads) {
        const int c = i >> 8, v = i & 255;
        lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), p.mean[c]), p.stdv[c]);
    }
    __syncthreads();
    const int x4 = (blockIdx.x * kTfThreads + tid) << 2;
    if (x4 >= p.pad_w) return;
    const int ih = p.in_h[n], iw = p.in_w[n], oh = p.out_h[n], ow = p.out_w[n];
    const size_t plane = (size_t)p.pad_h * p.pad_w;
    float* dst = out + (size_t)n * C * plane + (size_t)y * p.pad_w + x4;
    float v[4][4];                                                     // [channel][pixel]
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[c][j] = 0.f;
    if (y < oh && x4 < ow) {
        const float sy = __fdiv_rn((float)ih, (float)oh), sx = __fdiv_rn((float)iw, (float)ow);   // (float)input / output
        int y0, y1; float wy0, wy1;
        tf_tap(sy, y, ih, y0, y1, wy0, wy1);
        const uint8_t* r0 = p.img[n] + (size_t)y0 * iw * C;
        const uint8_t* r1 = p.img[n] + (size_t)y1 * iw * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x4 + j;
            if (x < ow) {
                int x0, x1; float wx0, wx1;
                tf_tap(sx, x, iw, x0, x1, wx0, wx1);
                for (int c = 0; c < C; ++c) {
                    const float v00 = lut[c][r0[x0 * C + c]], v01 = lut[c][r0[x1 * C + c]];
                    const float v10 = lut[c][r1[x0 * C + c]], v11 = lut[c][r1[x1 * C + c]];
                    const float t0 = fmaf(v00, wx0, __fmul_rn(v01, wx1));
                    const float t1 = fmaf(v10, wx0, __fmul_rn(v11, wx1));
                    v[c][j] = fmaf(t0, wy0, __fmul_rn(t1, wy1));
                }
            }
        }
    }
    const bool vec = (p.pad_w & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    for (int c = 0; c < C; ++c) {
        if (vec) {
            *reinterpret_cast<float4*>(dst + c * plane) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x4 + j < p.pad_w) dst[c * plane + j] = v[c][j];
        }
    }
}

}  // namespace mb

extern "C" int mb_image_transform(const mb_transform_params* pp, float* out, mb_stream_t stream) {
    if (!pp || !out) return MB_ERR_INVALID_ARG;
    const mb_transform_params& q = *pp;
    if (q.num_images < 1 || q.num_images > MB_MAX_IMAGES || q.channels < 1 || q.channels > 4 || q.pad_h < 1 || q.pad_w < 1)
        return MB_ERR_INVALID_ARG;
    mb::TfDev d;
    memset(&d, 0, sizeof(d));
    for (int i = 0; i < q.num_images; ++i) {
        if (!q.images[i] || q.in_h[i] < 1 || q.in_w[i] < 1 || q.out_h[i] < 1 || q.out_w[i] < 1 || q.out_h[i] > q.pad_h ||
            q.out_w[i] > q.pad_w)
            return MB_ERR_INVALID_ARG;
        d.img[i] = q.images[i];
        d.in_h[i] = q.in_h[i]; d.in_w[i] = q.in_w[i]; d.out_h[i] = q.out_h[i]; d.out_w[i] = q.out_w[i];
    }
    for (int c = 0; c < q.channels; ++c) { d.mean[c] = q.mean[c]; d.stdv[c] = q.std[c]; }
    d.channels = q.channels; d.pad_h = q.pad_h; d.pad_w = q.pad_w;
    if (q.pad_h > 65535 || q.num_images > 65535) return MB_ERR_UNSUPPORTED;
    dim3 grid((q.pad_w / 4 + mb::kTfThreads) / mb::kTfThreads, q.pad_h, q.num_images);
    mb::k_image_transform<<<grid, mb::kTfThreads, 0, (cudaStream_t)stream>>>(d, out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}
