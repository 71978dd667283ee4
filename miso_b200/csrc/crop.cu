// crop.cu — final score filter, annotation geometry and per-detection crop extraction.
//
// Replaces, on the device and without per-detection host round trips:
//   ref:miso/object_detection/inference.py:53-62   scores > threshold (strict), x, y, w=x2-x1, h=y2-y1 (fp32)
//   ref:miso/object_detection/dataset/annotation.py:120-127   coords = (x, y, x+w, y+h) in fp32;
//                                                   coords_int = int(np.round(c)), round half to even
//   ref:miso/object_detection/crop.py:28-30          crop = im[c[1]:c[3], c[0]:c[2], ...] (numpy slice rules)
//
//   k_crop_plan    one CTA: filter + rounding + slice resolution + stable compaction + byte offsets
//   k_crop_plan_chunks  the same for many slots (mosaic rows): CTA per 4096 slots, chained by flagged aggregates
//   k_crop_gather  persistent grid (SM-count multiple): crop slots strided over blockIdx.x (the next slot's metadata
//                  prefetched), rows over the CTA's warps and blockIdx.y, 8 / 16 / 32 lanes per row by row length;
//                  a row moves as 16-byte destination vectors assembled from two aligned source vectors (word
//                  select + funnel shift), bytes only at its unaligned ends and at the image border (HWC uint8)
#include <math.h>

#include <algorithm>
#include <mutex>
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

// one detection slot: live and above the threshold? -> annotation bounds, integer crop rectangle, crop bytes
__device__ __forceinline__ bool plan_entry(const CropDev& d, const float4* __restrict__ boxes, const float* __restrict__ scores,
                                           const int* __restrict__ counts, int e, int total, int4& rc, float4& an, long long& nbytes) {
    rc = make_int4(0, 0, 0, 0); an = make_float4(0, 0, 0, 0); nbytes = 0;
    if (e >= total) return false;
    const int n = e / d.cap, i = e - n * d.cap;
    if (!(i < counts[n] && scores[e] > d.thr)) return false;
    const float4 b = boxes[e];
    an = d.xywh ? b : make_float4(b.x, b.y, __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const long long c0 = round_half_even_to_int(an.x), c1 = round_half_even_to_int(an.y);
    const long long c2 = round_half_even_to_int(__fadd_rn(an.x, an.z));
    const long long c3 = round_half_even_to_int(__fadd_rn(an.y, an.w));
    resolve_slice(c0, c2, d.w[n], rc.x, rc.z);
    resolve_slice(c1, c3, d.h[n], rc.y, rc.w);
    nbytes = (long long)rc.z * rc.w * d.ch;
    return true;
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
            ok[k] = plan_entry(d, boxes, scores, counts, e, total, rc[k], an[k], nbytes[k]);
            if (ok[k]) { ++cnt; bytes += nbytes[k]; }
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

// The same plan for many slots (a mosaic's gathered rows): one CTA per 4096 slots. Chunk ids are handed out in order of
// arrival, every chunk publishes its (count, bytes) aggregate as one flagged 64-bit word as soon as it has it, then
// sums the aggregates of all chunks before it (thread per predecessor, spinning on the flag: predecessors are resident
// by construction) — the stable compaction of the one-CTA kernel without its 27 serial rounds.
constexpr int kPlanChunk = 4096;
constexpr unsigned long long kPlanFlag = 1ull << 63;

__global__ void __launch_bounds__(1024) k_crop_plan_chunks(const CropDev d, const float4* __restrict__ boxes,
                                                          const float* __restrict__ scores, const int* __restrict__ counts,
                                                          int4* __restrict__ rects, float4* __restrict__ xywh,
                                                          int* __restrict__ src, long long* __restrict__ offsets,
                                                          long long* __restrict__ totals,
                                                          unsigned long long* __restrict__ agg, unsigned int* __restrict__ ticket) {
    __shared__ int wcnt[32];
    __shared__ long long wbytes[32];
    __shared__ int s_chunk;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_chunk = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int chunk = s_chunk;
    const int total = d.N * d.cap;
    constexpr int kE = kPlanChunk / 1024;
    bool ok[kE];
    int4 rc[kE];
    float4 an[kE];
    long long nbytes[kE];
    int cnt = 0;
    long long bytes = 0;
    const int e0 = chunk * kPlanChunk + tid * kE;
#pragma unroll
    for (int k = 0; k < kE; ++k) {
        ok[k] = plan_entry(d, boxes, scores, counts, e0 + k, total, rc[k], an[k], nbytes[k]);
        if (ok[k]) { ++cnt; bytes += nbytes[k]; }
    }
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
    if (tid == 0) {
        // count <= 4096 in bits 50..62, bytes (< 2^50) below, flag on top: one word, one store
        const unsigned long long word = kPlanFlag | ((unsigned long long)tc << 50) | (unsigned long long)tb;
        *reinterpret_cast<volatile unsigned long long*>(agg + chunk) = word;
    }
    // aggregates of the chunks before this one
    int bc = 0;
    long long bb = 0;
    for (int q = tid; q < chunk; q += 1024) {
        unsigned long long v;
        do { v = *reinterpret_cast<const volatile unsigned long long*>(agg + q); } while (!(v & kPlanFlag));
        bc += (int)((v >> 50) & 0x1fffull);
        bb += (long long)(v & ((1ull << 50) - 1ull));
    }
    __syncthreads();                                   // wcnt / wbytes are reused
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bc += __shfl_xor_sync(0xffffffffu, bc, o);
        bb += __shfl_xor_sync(0xffffffffu, bb, o);
    }
    if (lane == 0) { wcnt[wid] = bc; wbytes[wid] = bb; }
    __syncthreads();
    int base_c = 0;
    long long base_b = 0;
    for (int q = 0; q < 32; ++q) { base_c += wcnt[q]; base_b += wbytes[q]; }
    int j = base_c + pc + cs - cnt;
    long long off = base_b + pb + bs - bytes;
#pragma unroll
    for (int k = 0; k < kE; ++k) {
        if (ok[k]) {
            rects[j] = rc[k];
            xywh[j] = an[k];
            src[j] = e0 + k;
            offsets[j] = off;
            ++j; off += nbytes[k];
        }
    }
    if (tid == 0 && chunk == (int)gridDim.x - 1) {
        offsets[base_c + tc] = base_b + tb;
        totals[0] = base_c + tc;
        totals[1] = base_b + tb;
        totals[2] = 0;
    }
}

// 128-bit read-only load / store helpers for the row copies
__device__ __forceinline__ uint4 ldg128(const unsigned char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Rows of one crop, GS lanes per row (32 / GS rows per warp pass). A row is copied as: head bytes up to the 16-byte
// alignment of the destination; 16-byte destination vectors, each assembled from the two aligned source vectors that
// cover it (word select by the source misalignment's word part, funnel shift by its byte part); leftover bytes one by
// one. Every load stays inside the image row the crop row lies in: vectors whose aligned window would reach past either
// end of that row (crops touching the image's left / right edge) fall to the byte path, so nothing outside the image
// buffer — or, for a mosaic band addressed through a virtual base pointer, outside the band — is ever read.
template <int GS>
__device__ __forceinline__ void crop_rows(const unsigned char* __restrict__ im, unsigned char* __restrict__ dst, const int4 rc,
                                          int img_w, int ch, int row_bytes, int y0, int ystep, int lane) {
    constexpr int RPP = 32 / GS;
    const int gl = lane % GS, grp = lane / GS;
    // a vector's aligned 16- or 32-byte source window may reach past the crop's columns but must stay inside the
    // image row: only crops within 16 bytes of the left or 32 bytes of the right image edge need the per-row check
    const bool near_edge = rc.x * ch < 16 || (img_w - rc.x - rc.z) * ch < 32;
    const int y_first = y0 * RPP + grp;
    const size_t src_pitch = (size_t)img_w * ch;
    const unsigned char* s = im + ((size_t)(rc.y + y_first) * img_w + rc.x) * ch;
    unsigned char* o = dst + (size_t)y_first * row_bytes;
    const size_t s_step = (size_t)ystep * RPP * src_pitch, o_step = (size_t)ystep * RPP * row_bytes;
    for (int y = y_first; y < rc.w; y += ystep * RPP, s += s_step, o += o_step) {
        const int head = min(row_bytes, (int)((16u - ((unsigned)reinterpret_cast<uintptr_t>(o) & 15u)) & 15u));
        const unsigned char* sv = s + head;
        const int sh = (int)((unsigned)reinterpret_cast<uintptr_t>(sv) & 15u);
        const unsigned char* sa = sv - sh;
        const int nvec = (row_bytes - head) >> 4;
        int i0 = 0, nsafe = nvec;
        if (near_edge) {
            const unsigned char* row_lo = s - (size_t)rc.x * ch;
            const long long before = sa - row_lo;                          // < 0: vector 0's window starts left of the image row
            const long long after = (row_lo + src_pitch) - sa;             // bytes from sa to the end of the image row
            i0 = before < 0 ? 1 : 0;
            const long long fit = (after - (sh != 0 ? 32 : 16)) >> 4;      // last vector index whose window ends inside the row
            nsafe = fit < 0 ? 0 : (int)min((long long)nvec, fit + 1);
        }
        const int ws = sh >> 2;
        const unsigned bs = (unsigned)(sh & 3) * 8u;
        // bytes the vectors do not cover: [0, head + 16 * i0) and [head + 16 * max(nsafe, i0), row_bytes). Their loads
        // (two per lane cover the usual <= 30 bytes) are issued before the vector loop so that both kinds of load
        // share one memory latency; the rest, if any, follows in a loop.
        const int lo_end = min(row_bytes, head + 16 * i0);
        const int hi_begin = min(row_bytes, max(lo_end, head + 16 * max(nsafe, i0)));
        const int nbyte = lo_end + (row_bytes - hi_begin);
        const int q0 = gl, q1 = gl + GS;
        const int bi0 = q0 < lo_end ? q0 : hi_begin + (q0 - lo_end), bi1 = q1 < lo_end ? q1 : hi_begin + (q1 - lo_end);
        unsigned char byte0 = 0, byte1 = 0;
        if (q0 < nbyte) byte0 = __ldg(s + bi0);
        if (q1 < nbyte) byte1 = __ldg(s + bi1);
#pragma unroll 2
        for (int i = i0 + gl; i < nsafe; i += GS) {
            const uint4 a = ldg128(sa + 16 * i);
            uint4 b = make_uint4(0u, 0u, 0u, 0u);
            if (sh != 0) b = ldg128(sa + 16 * i + 16);
            unsigned w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w, w4 = b.x, w5 = b.y, w6 = b.z, w7 = b.w;
            if (ws & 1) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6; w6 = w7; }
            if (ws & 2) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; }
            uint4 r;
            r.x = __funnelshift_r(w0, w1, bs); r.y = __funnelshift_r(w1, w2, bs);
            r.z = __funnelshift_r(w2, w3, bs); r.w = __funnelshift_r(w3, w4, bs);
            *reinterpret_cast<uint4*>(o + head + 16 * i) = r;
        }
        if (q0 < nbyte) o[bi0] = byte0;
        if (q1 < nbyte) o[bi1] = byte1;
        for (int q = gl + 2 * GS; q < nbyte; q += GS) {
            const int bidx = q < lo_end ? q : hi_begin + (q - lo_end);
            o[bidx] = __ldg(s + bidx);
        }
    }
}

__global__ void __launch_bounds__(256, 5) k_crop_gather(const CropDev d, const int4* __restrict__ rects,
                                                    const int* __restrict__ src, const long long* __restrict__ offsets,
                                                    long long* __restrict__ totals, unsigned char* __restrict__ out,
                                                    long long capacity) {
    const long long ncrops = totals[0];
    if (totals[1] > capacity) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) totals[2] = 1;  // caller re-runs with totals[1] bytes
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int y0 = blockIdx.y * nwarps + warp, ystep = gridDim.y * nwarps;
    // the next crop's rectangle, source slot and offset are fetched while the current crop is copied: a crop then costs
    // one memory latency (its rows), not two
    int4 rc_next = make_int4(0, 0, 0, 0);
    int src_next = 0;
    long long off_next = 0;
    if ((long long)blockIdx.x < ncrops) { rc_next = rects[blockIdx.x]; src_next = src[blockIdx.x]; off_next = offsets[blockIdx.x]; }
    for (long long j = blockIdx.x; j < ncrops; j += gridDim.x) {
        const int4 rc = rc_next;
        const int n = src_next / d.cap;
        unsigned char* dst = out + off_next;
        const long long jn = j + gridDim.x;
        if (jn < ncrops) { rc_next = rects[jn]; src_next = src[jn]; off_next = offsets[jn]; }
        const int row_bytes = rc.z * d.ch;
        const unsigned char* im = d.img[n];
        // lanes per row by the row length (warp-uniform: every warp of the CTA works on crop j)
        // (two vectors per lane: more rows of a crop in flight per warp, and the loop unrolls to both loads first)
        if (row_bytes <= 16 * 16 + 15) crop_rows<8>(im, dst, rc, d.w[n], d.ch, row_bytes, y0, ystep, lane);
        else if (row_bytes <= 32 * 16 + 15) crop_rows<16>(im, dst, rc, d.w[n], d.ch, row_bytes, y0, ystep, lane);
        else crop_rows<32>(im, dst, rc, d.w[n], d.ch, row_bytes, y0, ystep, lane);
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

// Stream-ordered scratch for the chunked plan: one pool per device that keeps what it has allocated (release threshold
// = max), so a call costs no trip to the driver's allocator after the first; allocation and free are ordered on the
// caller's stream, so concurrent calls on different streams never share scratch.
static int scratch_pool(cudaMemPool_t* out) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    MB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MB_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(mu);
    if (pools[dev] == nullptr) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        MB_CUDA(cudaMemPoolCreate(&pools[dev], &props));
        unsigned long long keep = ~0ull;
        MB_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = pools[dev];
    return MB_OK;
}

extern "C" int mb_crop_plan(const mb_crop_params* p, const float* det_boxes, const float* det_scores,
                            const int32_t* det_counts, int32_t* rects_out, float* xywh_out, int32_t* src_out,
                            int64_t* offsets_out, int64_t* totals_out, mb_stream_t stream) {
    if (!p || !det_boxes || !det_scores || !det_counts || !rects_out || !xywh_out || !src_out || !offsets_out || !totals_out)
        return MB_ERR_INVALID_ARG;
    CropDev d;
    int rc = make_crop(*p, d, false);
    if (rc != MB_OK) return rc;
    const long long slots = (long long)d.N * d.cap;
    if (slots <= 2 * kPlanChunk) {
        k_crop_plan<<<1, 1024, 0, (cudaStream_t)stream>>>(d, (const float4*)det_boxes, det_scores, det_counts, (int4*)rects_out,
                                                          (float4*)xywh_out, src_out, (long long*)offsets_out, (long long*)totals_out);
        MB_LAUNCH_CHECK();
        return MB_OK;
    }
    // many slots: chunked plan; its per-chunk aggregates live in a stream-ordered scratch allocation
    if (slots > (1ll << 31) - kPlanChunk) return MB_ERR_UNSUPPORTED;
    const int chunks = (int)((slots + kPlanChunk - 1) / kPlanChunk);
    void* scratch = nullptr;
    const size_t sbytes = (size_t)chunks * sizeof(unsigned long long) + 16;
    cudaMemPool_t pool = nullptr;
    rc = scratch_pool(&pool);
    if (rc != MB_OK) return rc;
    MB_CUDA(cudaMallocFromPoolAsync(&scratch, sbytes, pool, (cudaStream_t)stream));
    MB_CUDA(cudaMemsetAsync(scratch, 0, sbytes, (cudaStream_t)stream));
    k_crop_plan_chunks<<<chunks, 1024, 0, (cudaStream_t)stream>>>(
        d, (const float4*)det_boxes, det_scores, det_counts, (int4*)rects_out, (float4*)xywh_out, src_out,
        (long long*)offsets_out, (long long*)totals_out, (unsigned long long*)scratch,
        (unsigned int*)((unsigned long long*)scratch + chunks));
    MB_LAUNCH_CHECK();
    MB_CUDA(cudaFreeAsync(scratch, (cudaStream_t)stream));
    return MB_OK;
}

extern "C" int mb_crop_gather(const mb_crop_params* p, const int32_t* rects, const int32_t* src, const int64_t* offsets,
                              int64_t* totals, uint8_t* crops_out, int64_t crops_capacity_bytes, mb_stream_t stream) {
    if (!p || !rects || !src || !offsets || !totals || crops_capacity_bytes < 0) return MB_ERR_INVALID_ARG;
    if (crops_capacity_bytes > 0 && !crops_out) return MB_ERR_INVALID_ARG;
    CropDev d;
    int rc = make_crop(*p, d, true);
    if (rc != MB_OK) return rc;
    // crops over x; rows of a crop over the CTA's warps and, when there are few crop slots, over y as well (a CTA that
    // finds no rows of a crop left still pays for reading its rectangle, so y only grows while x cannot fill the GPU)
    const long long slots = (long long)d.N * d.cap;
    const int gx = (int)std::min<long long>(std::max<long long>(slots, 1), (long long)num_sms() * 8);
    const int gy = (int)std::min<long long>(16, std::max<long long>(1, ((long long)num_sms() * 256) / std::max<long long>(slots, 1)));
    dim3 grid(gx, gy);
    k_crop_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(d, (const int4*)rects, src, (const long long*)offsets,
                                                         (long long*)totals, crops_out, crops_capacity_bytes);
    MB_LAUNCH_CHECK();
    return MB_OK;
}
