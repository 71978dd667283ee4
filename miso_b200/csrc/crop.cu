// crop.cu — final score filter, annotation geometry and per-detection crop extraction.
//
// Replaces, on the device and without per-detection host round trips:
//   ref:miso/object_detection/inference.py:53-62   scores > threshold (strict), x, y, w=x2-x1, h=y2-y1 (fp32)
//   ref:miso/object_detection/dataset/annotation.py:120-127   coords = (x, y, x+w, y+h) in fp32;
//                                                   coords_int = int(np.round(c)), round half to even
//   ref:miso/object_detection/crop.py:28-30          crop = im[c[1]:c[3], c[0]:c[2], ...] (numpy slice rules)
//
//   k_crop_plan    one CTA: filter + rounding + slice resolution + stable compaction + byte offsets
//   k_crop_gather  persistent grid (SM-count multiple): crop slots strided over blockIdx.x, rows over
//                  blockIdx.y, lanes along the contiguous bytes of a source row (HWC uint8)
#include <math.h>

#include "common.cuh"

namespace mb {

struct CropDev {
    int N, cap, ch, xywh;
    float thr;
    int h[MB_MAX_IMAGES], w[MB_MAX_IMAGES];
    const unsigned char* img[MB_MAX_IMAGES];
};

// Python slice.indices(len) for step 1: returns begin and extent
__device__ __forceinline__ void resolve_slice(long long start, long long stop, int len, int& begin, int& extent) {
    if (start < 0) { start += len; if (start < 0) start = 0; } else if (start > len) start = len;
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; } else if (stop > len) stop = len;
    begin = (int)start;
    extent = stop > start ? (int)(stop - start) : 0;
}

__device__ __forceinline__ long long round_half_even_to_int(float v) {
    // np.round on float32 rounds half to even and stays float32; int() then truncates an integer value
    const float r = rintf(v);
    if (!(r == r)) return 0;
    if (r > 9.0e18f) return (long long)9.0e18;
    if (r < -9.0e18f) return -(long long)9.0e18;
    return (long long)r;
}

__global__ void __launch_bounds__(1024) k_crop_plan(const CropDev d, const float4* __restrict__ boxes,
                                                   const float* __restrict__ scores, const int* __restrict__ counts,
                                                   int4* __restrict__ rects, float4* __restrict__ xywh,
                                                   int* __restrict__ src, long long* __restrict__ offsets,
                                                   long long* __restrict__ totals) {
    __shared__ int wcnt[32];
    __shared__ long long wbytes[32];
    __shared__ int s_cnt;
    __shared__ long long s_bytes;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_cnt = 0; s_bytes = 0; }
    __syncthreads();
    const int total = d.N * d.cap;
    constexpr int kE = 4;                             // consecutive entries per thread: 4096 per round, one round for the usual N * cap
    for (int e0 = 0; e0 < total; e0 += 1024 * kE) {
        bool ok[kE];
        int4 rc[kE];
        float4 an[kE];
        long long nbytes[kE];
        int cnt = 0;
        long long bytes = 0;
#pragma unroll
        for (int k = 0; k < kE; ++k) {
            const int e = e0 + tid * kE + k;
            ok[k] = false; rc[k] = make_int4(0, 0, 0, 0); an[k] = make_float4(0, 0, 0, 0); nbytes[k] = 0;
            if (e < total) {
                const int n = e / d.cap, i = e - n * d.cap;
                if (i < counts[n] && scores[e] > d.thr) {
                    ok[k] = true;
                    const float4 b = boxes[e];
                    an[k] = d.xywh ? b : make_float4(b.x, b.y, __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
                    const long long c0 = round_half_even_to_int(an[k].x), c1 = round_half_even_to_int(an[k].y);
                    const long long c2 = round_half_even_to_int(__fadd_rn(an[k].x, an[k].z));
                    const long long c3 = round_half_even_to_int(__fadd_rn(an[k].y, an[k].w));
                    resolve_slice(c0, c2, d.w[n], rc[k].x, rc[k].z);
                    resolve_slice(c1, c3, d.h[n], rc[k].y, rc[k].w);
                    nbytes[k] = (long long)rc[k].z * rc[k].w * d.ch;
                    ++cnt; bytes += nbytes[k];
                }
            }
        }
        // stable compaction: exclusive scan of (count, bytes) over the threads of this round (entries of a thread
        // are consecutive, threads are in entry order)
        int cs = cnt;
        long long bs = bytes;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int yc = __shfl_up_sync(0xffffffffu, cs, o);
            const long long yb = __shfl_up_sync(0xffffffffu, bs, o);
            if (lane >= o) { cs += yc; bs += yb; }
        }
        if (lane == 31) { wcnt[wid] = cs; wbytes[wid] = bs; }
        __syncthreads();
        int pc = 0, tc = 0;
        long long pb = 0, tb = 0;
        for (int q = 0; q < 32; ++q) {
            pc += (q < wid) ? wcnt[q] : 0; tc += wcnt[q];
            pb += (q < wid) ? wbytes[q] : 0; tb += wbytes[q];
        }
        const int base_c = s_cnt;
        const long long base_b = s_bytes;
        int j = base_c + pc + cs - cnt;
        long long off = base_b + pb + bs - bytes;
#pragma unroll
        for (int k = 0; k < kE; ++k) {
            if (ok[k]) {
                rects[j] = rc[k];
                xywh[j] = an[k];
                src[j] = e0 + tid * kE + k;
                offsets[j] = off;
                ++j; off += nbytes[k];
            }
        }
        __syncthreads();
        if (tid == 0) { s_cnt = base_c + tc; s_bytes = base_b + tb; }
        __syncthreads();
    }
    if (tid == 0) {
        offsets[s_cnt] = s_bytes;
        totals[0] = s_cnt;
        totals[1] = s_bytes;
        totals[2] = 0;
    }
}

__global__ void __launch_bounds__(256) k_crop_gather(const CropDev d, const int4* __restrict__ rects,
                                                    const int* __restrict__ src, const long long* __restrict__ offsets,
                                                    long long* __restrict__ totals, unsigned char* __restrict__ out,
                                                    long long capacity) {
    const long long ncrops = totals[0];
    if (totals[1] > capacity) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) totals[2] = 1;  // caller re-runs with totals[1] bytes
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (long long j = blockIdx.x; j < ncrops; j += gridDim.x) {
        const int4 rc = rects[j];
        const int n = src[j] / d.cap;
        const int row_bytes = rc.z * d.ch;
        const unsigned char* im = d.img[n];
        unsigned char* dst = out + offsets[j];
        for (int y = blockIdx.y * nwarps + warp; y < rc.w; y += gridDim.y * nwarps) {
            const unsigned char* s = im + ((size_t)(rc.y + y) * d.w[n] + rc.x) * d.ch;
            unsigned char* o = dst + (size_t)y * row_bytes;
            // head bytes up to 4-byte alignment of the destination, funnel-shifted words, tail bytes
            const int head = min(row_bytes, (int)((4 - (reinterpret_cast<uintptr_t>(o) & 3)) & 3));
            if (lane < head) o[lane] = __ldg(s + lane);
            const int words = (row_bytes - head) >> 2;
            const unsigned char* sw = s + head;
            const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(sw) & 3) * 8;
            const unsigned int* sa = reinterpret_cast<const unsigned int*>(sw - (shift >> 3));
            unsigned int* oa = reinterpret_cast<unsigned int*>(o + head);
            // the last source word may straddle the end of the image buffer: leave it to the byte tail
            const int safe_words = shift ? max(words - 1, 0) : words;
            for (int i = lane; i < safe_words; i += 32) {
                const unsigned lo = __ldg(sa + i);
                const unsigned hi = shift ? __ldg(sa + i + 1) : 0u;
                oa[i] = shift ? __funnelshift_r(lo, hi, shift) : lo;
            }
            for (int i = head + safe_words * 4 + lane; i < row_bytes; i += 32) o[i] = __ldg(s + i);
        }
    }
}

static int make_crop(const mb_crop_params& p, CropDev& d, bool need_images) {
    if (p.num_images < 1 || p.num_images > MB_MAX_IMAGES || p.capacity < 1 || p.channels < 1) return MB_ERR_INVALID_ARG;
    d.N = p.num_images; d.cap = p.capacity; d.ch = p.channels; d.thr = p.threshold; d.xywh = p.boxes_are_xywh;
    for (int n = 0; n < d.N; ++n) {
        if (p.image_h[n] < 0 || p.image_w[n] < 0) return MB_ERR_INVALID_ARG;
        d.h[n] = p.image_h[n]; d.w[n] = p.image_w[n]; d.img[n] = p.images[n];
        if (need_images && !d.img[n]) return MB_ERR_INVALID_ARG;
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" int mb_crop_plan(const mb_crop_params* p, const float* det_boxes, const float* det_scores,
                            const int32_t* det_counts, int32_t* rects_out, float* xywh_out, int32_t* src_out,
                            int64_t* offsets_out, int64_t* totals_out, mb_stream_t stream) {
    if (!p || !det_boxes || !det_scores || !det_counts || !rects_out || !xywh_out || !src_out || !offsets_out || !totals_out)
        return MB_ERR_INVALID_ARG;
    CropDev d;
    int rc = make_crop(*p, d, false);
    if (rc != MB_OK) return rc;
    k_crop_plan<<<1, 1024, 0, (cudaStream_t)stream>>>(d, (const float4*)det_boxes, det_scores, det_counts, (int4*)rects_out,
                                                      (float4*)xywh_out, src_out, (long long*)offsets_out, (long long*)totals_out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

extern "C" int mb_crop_gather(const mb_crop_params* p, const int32_t* rects, const int32_t* src, const int64_t* offsets,
                              int64_t* totals, uint8_t* crops_out, int64_t crops_capacity_bytes, mb_stream_t stream) {
    if (!p || !rects || !src || !offsets || !totals || crops_capacity_bytes < 0) return MB_ERR_INVALID_ARG;
    if (crops_capacity_bytes > 0 && !crops_out) return MB_ERR_INVALID_ARG;
    CropDev d;
    int rc = make_crop(*p, d, true);
    if (rc != MB_OK) return rc;
    dim3 grid(kNumSMs * 2, 16);   // crop slots over x, row blocks over y: many short rows in flight
    k_crop_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(d, (const int4*)rects, src, (const long long*)offsets,
                                                         (long long*)totals, crops_out, crops_capacity_bytes);
    MB_LAUNCH_CHECK();
    return MB_OK;
}
