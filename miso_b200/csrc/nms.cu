// nms.cu — torchvision-signature NMS entry point (plain / per-group "vanilla" / coordinate trick).
// Replaces tv:ops/boxes.py:20-120 (torchvision::nms + both batched_nms strategies); semantics of
// the CPU kernel tv-csrc:ops/cpu/nms_kernel.cpp:116.
#include <math.h>

#include "nms_core.cuh"

namespace mb {

constexpr int kScoreBuckets = 4096;  // second-level sort of the kept set: top 12 key bits

struct NmsScratch {
    SegArrays seg;        // segments = groups
    SegArrays seg2;       // segments = score buckets of the kept set (vanilla finalize)
    int* seg_fill;        // [G]
    int* seg2_fill;       // [kScoreBuckets]
    unsigned long long* bkey;  // [K] bucketed
    float4* bbox;              // [K]
    int* bseg;                 // [K]
    unsigned long long* skey;  // [K] sorted inside segment
    float4* sbox;              // [K]
    unsigned long long* kkey;  // [K] kept keys bucketed by score
    int* kseg;                 // [K]
    unsigned long long* keepbits;  // [K/64 + G + 1]
    unsigned int* scalars;     // [8]: 0 max-coordinate key, 1 bad-group flag
};

static size_t carve_nms(Carver& c, NmsScratch& w, long long K, int G) {
    w.seg = carve_seg_arrays(c, G);
    w.seg2 = carve_seg_arrays(c, kScoreBuckets);
    w.seg_fill = c.take<int>(G);
    w.seg2_fill = c.take<int>(kScoreBuckets);
    w.bkey = c.take<unsigned long long>(K);
    w.bbox = c.take<float4>(K);
    w.bseg = c.take<int>(K);
    w.skey = c.take<unsigned long long>(K);
    w.sbox = c.take<float4>(K);
    w.kkey = c.take<unsigned long long>(K);
    w.kseg = c.take<int>(K);
    w.keepbits = c.take<unsigned long long>(K / 64 + G + 2);
    w.scalars = c.take<unsigned int>(8);
    return c.off;
}

__device__ __forceinline__ unsigned int float_order_key(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void k_box_max(const float* __restrict__ boxes, long long n4, unsigned int* out_key) {
    unsigned int best = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        best = max(best, float_order_key(boxes[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_key, best);
}

__global__ void k_group_hist(const long long* __restrict__ groups, long long K, int G, int* seg_count,
                             unsigned int* bad_flag) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < K; i += (long long)gridDim.x * blockDim.x) {
        const long long g = groups[i];
        if (g < 0) continue;                       // negative group: row is not a box (padding), ignored
        if (g >= G) { *bad_flag = 1u; continue; }
        atomicAdd(&seg_count[(int)g], 1);
    }
}

__global__ void k_single_segment(int* seg_count, long long K) { seg_count[0] = (int)K; }

// mode: 0 plain, 1 vanilla (bucket by group), 2 trick (offset, single segment)
__global__ void k_scatter_boxes(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                const long long* __restrict__ groups, long long K, int G, int mode,
                                const int* __restrict__ seg_start, int* seg_fill,
                                const unsigned int* __restrict__ max_key,
                                unsigned long long* bkey, float4* bbox, int* bseg) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < K; i += (long long)gridDim.x * blockDim.x) {
        float4 b = boxes[i];
        int g = 0;
        long long p = i;  // single segment: keep the original slot, rank sorts anyway
        if (mode == MB_NMS_VANILLA) {
            const long long gg = groups[i];
            if (gg < 0 || gg >= G) continue;
            g = (int)gg;
            p = seg_start[g] + atomicAdd(&seg_fill[g], 1);
        } else if (mode == MB_NMS_TRICK) {
            // offsets = idxs.to(boxes) * (max_coordinate + 1); boxes + offsets[:, None]   (tv:ops/boxes.py:99-101)
            const float m1 = __fadd_rn(float_from_order_key(*max_key), 1.0f);
            const float off = __fmul_rn((float)groups[i], m1);
            b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off);
            b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
        }
        bkey[p] = ((unsigned long long)desc_score_key(scores[i]) << 32) | (unsigned long long)(unsigned int)i;
        bbox[p] = b;
        bseg[p] = g;
    }
}

// kept set of a single segment -> keep_out in sweep order (already the global score order)
__global__ void __launch_bounds__(1024) k_emit_single(SegArrays s, const unsigned long long* __restrict__ skey,
                                                     const unsigned long long* __restrict__ keepbits,
                                                     long long* keep_out, long long* status,
                                                     const unsigned int* __restrict__ bad_flag) {
    __shared__ long long sh[64];
    __shared__ long long carry;
    const int tid = threadIdx.x;
    if (*bad_flag != 0u) {                 // a group index >= num_groups (vanilla mode with one group)
        if (tid == 0) { status[0] = -2; status[1] = s.totals[1]; }
        return;
    }
    if (s.totals[2] != 0) {
        if (tid == 0) { status[0] = -1; status[1] = s.totals[1]; }
        return;
    }
    if (tid == 0) carry = 0;
    __syncthreads();
    const int T = s.seg_words[0];
    for (int base = 0; base < T; base += 1024) {
        const int w = base + tid;
        const unsigned long long bits = (w < T) ? keepbits[w] : 0ull;
        long long tot;
        long long pre = block_excl_scan_1024(__popcll(bits), sh, tot);
        const long long c = carry;
        __syncthreads();
        if (tid == 0) carry = c + tot;
        unsigned long long b = bits;
        long long o = c + pre;
        while (b) {
            const int t = __ffsll((long long)b) - 1; b &= b - 1;
            keep_out[o++] = (long long)(skey[(long long)w * 64 + t] & 0xffffffffull);
        }
        __syncthreads();
    }
    if (tid == 0) { status[0] = carry; status[1] = s.totals[1]; }
}

// vanilla finalize, step A/B: histogram / scatter of the kept keys into score buckets
__global__ void k_kept_bucket(SegArrays s, const int* __restrict__ bseg, const unsigned long long* __restrict__ skey,
                              const unsigned long long* __restrict__ keepbits, long long K, int pass,
                              int* seg2_count, const int* __restrict__ seg2_start, int* seg2_fill,
                              unsigned long long* kkey, int* kseg) {
    if (s.totals[2] != 0) return;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < K; p += (long long)gridDim.x * blockDim.x) {
        const int g = bseg[p];
        if (g < 0) continue;
        const int q = (int)p - s.seg_start[g];
        if (q >= s.seg_count[g]) continue;
        const unsigned long long word = keepbits[s.keep_off[g] + (q >> 6)];
        if (!((word >> (q & 63)) & 1ull)) continue;
        const unsigned long long key = skey[p];
        const int b = (int)(key >> 52);
        if (pass == 0) atomicAdd(&seg2_count[b], 1);
        else {
            const int pos = seg2_start[b] + atomicAdd(&seg2_fill[b], 1);
            kkey[pos] = key;
            kseg[pos] = b;
        }
    }
}

__global__ void k_emit_sorted(SegArrays s, SegArrays s2, const unsigned long long* __restrict__ sorted_keys,
                              long long* keep_out, long long* status, unsigned int* bad_flag) {
    if (s.totals[2] != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { status[0] = -1; status[1] = s.totals[1]; }
        return;
    }
    const long long total = (long long)s2.seg_start[kScoreBuckets - 1] + s2.seg_count[kScoreBuckets - 1];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        keep_out[i] = (long long)(sorted_keys[i] & 0xffffffffull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        status[0] = (*bad_flag) ? -2 : total;
        status[1] = s.totals[1];
    }
}

}  // namespace mb

using namespace mb;

extern "C" size_t mb_nms_workspace_bytes(int64_t num_boxes, int32_t num_groups) {
    Carver c(nullptr, 0);
    NmsScratch w;
    const int G = num_groups > 0 ? num_groups : 1;
    return carve_nms(c, w, num_boxes > 0 ? num_boxes : 1, G) + 1024;
}

extern "C" int mb_nms(const float* boxes, const float* scores, const int64_t* groups, int64_t num_boxes,
                      int32_t num_groups, int32_t mode, double iou_threshold, int64_t* keep_out,
                      int64_t* status_out, void* workspace, size_t workspace_bytes, void* mask_workspace,
                      size_t mask_workspace_bytes, mb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_boxes < 0 || num_boxes >= (1ll << 31) || !status_out) return MB_ERR_INVALID_ARG;
    if (mode != MB_NMS_PLAIN && mode != MB_NMS_VANILLA && mode != MB_NMS_TRICK) return MB_ERR_INVALID_ARG;
    if (mode != MB_NMS_PLAIN && groups == nullptr) return MB_ERR_INVALID_ARG;
    if (num_boxes == 0) {
        MB_CUDA(cudaMemsetAsync(status_out, 0, 4 * sizeof(int64_t), stream));
        return MB_OK;
    }
    if (!boxes || !scores || !keep_out) return MB_ERR_INVALID_ARG;
    const int G = (mode == MB_NMS_VANILLA) ? num_groups : 1;
    if (G < 1 || G > 65536) return MB_ERR_UNSUPPORTED;
    const long long K = num_boxes;
    Carver c(workspace, workspace_bytes);
    NmsScratch w;
    carve_nms(c, w, K, G);
    if (!c.ok()) return MB_ERR_WORKSPACE;
    const long long mask_cap_words = (long long)(mask_workspace_bytes / 8);
    unsigned long long* mask = (unsigned long long*)mask_workspace;

    MB_CUDA(cudaMemsetAsync(workspace, 0, c.off, stream));
    MB_CUDA(cudaMemsetAsync(w.bseg, 0xff, sizeof(int) * K, stream));
    MB_CUDA(cudaMemsetAsync(w.kseg, 0xff, sizeof(int) * K, stream));
    const int grid = (int)min((long long)num_sms() * 8, ceil_div64(K, 256));
    if (mode == MB_NMS_TRICK) {
        k_box_max<<<grid, 256, 0, stream>>>(boxes, K * 4, w.scalars);
        MB_LAUNCH_CHECK();
    }
    if (mode == MB_NMS_VANILLA) k_group_hist<<<grid, 256, 0, stream>>>((const long long*)groups, K, G, w.seg.seg_count, w.scalars + 1);
    else k_single_segment<<<1, 1, 0, stream>>>(w.seg.seg_count, K);
    MB_LAUNCH_CHECK();
    MetaRule norule{};
    k_seg_meta<<<1, 1024, 0, stream>>>(w.seg, G, 1, mask_cap_words, norule);
    MB_LAUNCH_CHECK();
    k_scatter_boxes<<<grid, 256, 0, stream>>>((const float4*)boxes, scores, (const long long*)groups, K, G, mode,
                                              w.seg.seg_start, w.seg_fill, w.scalars, w.bkey, w.bbox, w.bseg);
    MB_LAUNCH_CHECK();
    k_rank_in_segment<<<(int)ceil_div64(K, kRankKeys), kRankThreads, 0, stream>>>(
        w.bkey, w.bbox, w.bseg, w.seg.seg_start, w.seg.seg_count, nullptr, (int)K, w.skey, w.sbox);
    MB_LAUNCH_CHECK();
    int rc = launch_mask_and_sweep(w.sbox, w.seg, G, (int)K, iou_threshold, mask, w.keepbits, 0, stream);
    if (rc != MB_OK) return rc;
    if (G == 1) {
        k_emit_single<<<1, 1024, 0, stream>>>(w.seg, w.skey, w.keepbits, (long long*)keep_out, (long long*)status_out, w.scalars + 1);
        MB_LAUNCH_CHECK();
        return MB_OK;
    }
    // vanilla: order the union of kept boxes by (score desc, index asc) — tv:ops/boxes.py:119-120
    k_kept_bucket<<<grid, 256, 0, stream>>>(w.seg, w.bseg, w.skey, w.keepbits, K, 0, w.seg2.seg_count,
                                            nullptr, nullptr, nullptr, nullptr);
    MB_LAUNCH_CHECK();
    k_seg_meta<<<1, 1024, 0, stream>>>(w.seg2, kScoreBuckets, 1, (1ll << 62), norule);
    MB_LAUNCH_CHECK();
    k_kept_bucket<<<grid, 256, 0, stream>>>(w.seg, w.bseg, w.skey, w.keepbits, K, 1, w.seg2.seg_count,
                                            w.seg2.seg_start, w.seg2_fill, w.kkey, w.kseg);
    MB_LAUNCH_CHECK();
    // reuse bkey as the destination of the second-level rank; bseg is still needed -> kseg marks holes
    MB_CUDA(cudaMemsetAsync(w.bkey, 0, sizeof(unsigned long long) * K, stream));
    k_rank_in_segment<<<(int)ceil_div64(K, kRankKeys), kRankThreads, 0, stream>>>(
        w.kkey, nullptr, w.kseg, w.seg2.seg_start, w.seg2.seg_count, nullptr, (int)K, w.bkey, nullptr);
    MB_LAUNCH_CHECK();
    k_emit_sorted<<<grid, 256, 0, stream>>>(w.seg, w.seg2, w.bkey, (long long*)keep_out, (long long*)status_out, w.scalars + 1);
    MB_LAUNCH_CHECK();
    return MB_OK;
}
