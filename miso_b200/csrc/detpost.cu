// detpost.cu — fused detection post-processing.
//
// Replaces RoIHeads.postprocess_detections (tv:models/detection/roi_heads.py:680-737) and the box
// part of GeneralizedRCNNTransform.postprocess (transform.py:257-277, resize_boxes :306-319).
//
//   k_det_candidates  one thread per (image, proposal): softmax over the class logits, then for
//                     every foreground class decode (weights 10,10,5,5) + clip + `score > thresh`
//                     + remove_small_boxes(1e-2); survivors are appended to the (image, class)
//                     segment with the unique key (score desc, flattened index r*(C-1)+(c-1) asc)
//   k_rank_in_segment / k_seg_meta / k_nms_mask / k_nms_sweep   (nms_core.cuh)
//   k_det_finalize    one CTA per image: order the kept boxes of all classes by (score desc,
//                     flattened index asc), keep detections_per_img, map to the original image size
#include <math.h>

#include "boxmath.cuh"
#include "nms_core.cuh"
#include "sort_smem.cuh"

namespace mb {

constexpr int kDetThreads = 256;
constexpr int kDetFinalThreads = 1024;
constexpr int kDetSortCap = 16384;
constexpr int kDetMergeMaxRuns = 8;

struct DetDev {
    int N, C, max_props, dpi;
    float score_thresh, min_size;
    DecodeWeights dw;
};
struct DetImages {
    int h[MB_MAX_IMAGES], w[MB_MAX_IMAGES];
    float ratio_h[MB_MAX_IMAGES], ratio_w[MB_MAX_IMAGES];  // original / network size, fp32 quotient
};

struct DetScratch {
    float* img_max;             // [N]
    unsigned long long* bkey;   // [P] bucketed candidates, P = N*(C-1)*max_props
    float4* bbox;               // [P]
    int* bseg;                  // [P]
    unsigned long long* skey;   // [P]
    float4* sbox;               // [P]
    float4* box_by_flat;        // [N*F], F = max_props*(C-1)
    float* score_by_flat;       // [N*F]
    float* seg_offset;          // [G]
    unsigned long long* keepbits;
    unsigned long long* mask;
    unsigned long long* gkeys;  // [N * gcap] sort scratch in global memory, only when the per-image bound exceeds shared memory
    int gcap;
    SegArrays seg;
    size_t zero_bytes;
    long long mask_words;
    long long P;
};

static void carve_det(Carver& c, DetScratch& w, const DetDev& d) {
    const int G = d.N * (d.C - 1);
    const long long P = (long long)G * d.max_props;
    w.P = P;
    w.img_max = c.take<float>(d.N);
    w.seg = carve_seg_arrays(c, G);       // seg_count must start at zero
    w.zero_bytes = align_up(c.off, 16);   // (the next item starts 256-byte aligned: the round-up only covers padding)
    w.bkey = c.take<unsigned long long>(P);
    w.bbox = c.take<float4>(P);
    w.bseg = c.take<int>(P);
    w.skey = c.take<unsigned long long>(P);
    w.sbox = c.take<float4>(P);
    w.box_by_flat = c.take<float4>(P);
    w.score_by_flat = c.take<float>(P);
    w.seg_offset = c.take<float>(G);
    w.keepbits = c.take<unsigned long long>((size_t)G * (d.max_props / 64 + 2));
    w.mask_words = (long long)G * d.max_props * ((d.max_props + 63) / 64);
    w.mask = c.take<unsigned long long>((size_t)w.mask_words);
    const long long per_seg = d.max_props < d.dpi ? d.max_props : d.dpi;
    const long long bound = (long long)(d.C - 1) * per_seg;
    w.gcap = bound > kDetSortCap ? next_pow2((int)bound) : 0;
    w.gkeys = w.gcap ? c.take<unsigned long long>((size_t)d.N * w.gcap) : nullptr;
}

__global__ void __launch_bounds__(kDetThreads) k_det_candidates(const DetDev d, const DetImages im,
                                                               const float* __restrict__ logits,
                                                               const float* __restrict__ reg,
                                                               const float4* __restrict__ proposals,
                                                               const int* __restrict__ prop_counts, int packed,
                                                               DetScratch w) {
    pdl_enter();
    const int t = blockIdx.x * kDetThreads + threadIdx.x;
    if (t >= d.N * d.max_props) return;
    const int n = t / d.max_props, r = t - n * d.max_props;
    if (r >= prop_counts[n]) return;
    long long row = (long long)n * d.max_props + r;
    if (packed) {
        row = r;
        for (int q = 0; q < n; ++q) row += prop_counts[q];
    }
    const int C = d.C;
    const float* lg = logits + row * C;
    // softmax over the last dim (max-subtracted, fp32 accumulation in class order)
    float mx = lg[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, lg[c]);
    float sum = 0.0f;
    for (int c = 0; c < C; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(lg[c], mx)));
    const float4 prop = proposals[(size_t)n * d.max_props + r];
    const float img_h = (float)im.h[n], img_w = (float)im.w[n];
    const int F = d.max_props * (C - 1);
    float vmax = 0.0f;
    for (int c = 1; c < C; ++c) {
        const float sc = __fdiv_rn(expf(__fsub_rn(lg[c], mx)), sum);
        if (!(sc > d.score_thresh)) continue;
        const float4 dl = *reinterpret_cast<const float4*>(reg + (row * C + c) * 4);
        const float4 box = clip_box(decode_box(prop, dl, d.dw), img_h, img_w);
        if (!box_not_small(box, d.min_size)) continue;
        const int g = n * (C - 1) + (c - 1);
        const int flat = r * (C - 1) + (c - 1);
        const int pos = g * d.max_props + atomicAdd(&w.seg.seg_count[g], 1);
        w.bkey[pos] = ((unsigned long long)desc_score_key(sc) << 32) | (unsigned int)flat;
        w.bbox[pos] = box;
        w.bseg[pos] = g;
        w.box_by_flat[(size_t)n * F + flat] = box;
        w.score_by_flat[(size_t)n * F + flat] = sc;
        vmax = fmaxf(vmax, fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)));
    }
    if (vmax > 0.0f) atomicMax((int*)&w.img_max[n], __float_as_int(vmax));
}

// one launch for the per-call initialisation: clear the counters region, mark every bucket slot as a hole (-1)
// and write the fixed segment starts (instead of two memsets and a kernel). seg_start lies INSIDE the cleared
// region: the thread that clears a 16-byte vector overlapping it writes the final values instead of zeros, so no
// word is written by two threads (no ordering between threads is assumed).
__global__ void __launch_bounds__(256) k_det_init(int G, int max_props, long long seg_start_vec, uint4* __restrict__ zero16,
                                                  long long zero_vecs, int* __restrict__ bseg, long long P) {
    pdl_enter();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const long long seg_vecs = ((long long)G + 3) / 4;
    for (long long j = i; j < zero_vecs; j += stride) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const long long s = j - seg_start_vec;
        if (s >= 0 && s < seg_vecs) {
            const long long g0 = 4 * s;
            v.x = g0 < G ? (unsigned)(g0 * max_props) : 0u;
            v.y = g0 + 1 < G ? (unsigned)((g0 + 1) * max_props) : 0u;
            v.z = g0 + 2 < G ? (unsigned)((g0 + 2) * max_props) : 0u;
            v.w = g0 + 3 < G ? (unsigned)((g0 + 3) * max_props) : 0u;
        }
        zero16[j] = v;
    }
    for (long long j = i; j < P; j += stride) bseg[j] = -1;
}

__global__ void __launch_bounds__(kDetFinalThreads) k_det_finalize(const DetDev d, const DetImages im, DetScratch w,
                                                                  float4* det_boxes, float4* det_boxes_net,
                                                                  float* det_scores, long long* det_labels,
                                                                  int* det_counts, int use_merge, int smem_keys) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned long long keys_smem[];
    unsigned long long* keys = keys_smem;
    __shared__ int s_cnt;
    __shared__ int run_off[kDetMergeMaxRuns + 1];
    __shared__ int wpre[kDetMergeMaxRuns][kSweepSmallMaxWords + 1];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int S = d.C - 1, F = d.max_props * S;
    int total;
    if (use_merge) {
        // few classes: every class's kept list is already sorted, so merge by rank (no sort);
        // the merged order lands in keys[total .. 2*total)
        for (int i = tid; i < S * (kSweepSmallMaxWords + 1); i += kDetFinalThreads) {   // independent loads, one latency
            const int c = i / (kSweepSmallMaxWords + 1), q = i - c * (kSweepSmallMaxWords + 1);
            const int g = n * S + c;
            wpre[c][q] = (q < w.seg.seg_words[g]) ? __popcll(w.keepbits[w.seg.keep_off[g] + q]) : 0;
        }
        __syncthreads();
        if (tid < S) {
            const int T = w.seg.seg_words[n * S + tid];
            int acc = 0;
            for (int q = 0; q < T; ++q) { const int cq = wpre[tid][q]; wpre[tid][q] = acc; acc += cq; }
            wpre[tid][T] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int c = 0; c < S; ++c) { run_off[c] = acc; acc += wpre[c][w.seg.seg_words[n * S + c]]; }
            run_off[S] = acc;
        }
        __syncthreads();
        total = run_off[S];
        for (int c = 0; c < S; ++c) {
            const int g = n * S + c;
            const int cnt = w.seg.seg_count[g], st = w.seg.seg_start[g];
            const unsigned long long* kb = w.keepbits + w.seg.keep_off[g];
            for (int q = tid; q < cnt; q += kDetFinalThreads) {
                const unsigned long long word = kb[q >> 6];
                if ((word >> (q & 63)) & 1ull)
                    keys[total + run_off[c] + wpre[c][q >> 6] + __popcll(word & ((1ull << (q & 63)) - 1ull))] = w.skey[st + q];
            }
        }
        __syncthreads();
        for (int e = tid; e < total; e += kDetFinalThreads) {
            int c = 0;
            while (e >= run_off[c + 1]) ++c;
            const int rank = merged_rank(keys + total, run_off, S, c, e - run_off[c]);
            keys[rank] = keys[total + e];
        }
        __syncthreads();
    } else {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        if (w.gcap) {
            // many classes (e.g. the 91-class COCO head at 300 detections): the worst case does not fit shared memory.
            // Count what was actually kept; only if THAT exceeds the shared buffer is the sort done in global memory.
            int local = 0;
            for (int c = 0; c < S; ++c) {
                const int g = n * S + c;
                const unsigned long long* kb = w.keepbits + w.seg.keep_off[g];
                for (int q = tid; q < w.seg.seg_words[g]; q += kDetFinalThreads) local += __popcll(kb[q]);
            }
            if (local) atomicAdd(&s_cnt, local);
            __syncthreads();
            if (next_pow2(max(s_cnt, 2)) > smem_keys) keys = w.gkeys + (size_t)n * w.gcap;
            __syncthreads();
            if (tid == 0) s_cnt = 0;
            __syncthreads();
        }
        for (int c = 0; c < S; ++c) {
            const int g = n * S + c;
            const int cnt = w.seg.seg_count[g], st = w.seg.seg_start[g];
            const unsigned long long* kb = w.keepbits + w.seg.keep_off[g];
            for (int q0 = 0; q0 < cnt; q0 += kDetFinalThreads) {
                const int q = q0 + tid;
                const bool kept = q < cnt && ((kb[q >> 6] >> (q & 63)) & 1ull);
                const int slot = warp_alloc_slot(&s_cnt, kept);
                if (kept) keys[slot] = w.skey[st + q];
            }
        }
        __syncthreads();
        total = s_cnt;
        const int np2 = next_pow2(max(total, 2));
        for (int i = total + tid; i < np2; i += kDetFinalThreads) keys[i] = ~0ull;
        bitonic_sort_smem(keys, np2);
    }
    const int nout = min(total, d.dpi);
    for (int i = tid; i < d.dpi; i += kDetFinalThreads) {
        float4 b = make_float4(0, 0, 0, 0);
        float s = 0.0f;
        long long lab = 0;
        if (i < nout) {
            const int flat = (int)(keys[i] & 0xffffffffull);
            b = w.box_by_flat[(size_t)n * F + flat];
            s = w.score_by_flat[(size_t)n * F + flat];
            lab = flat % S + 1;
        }
        const size_t o = (size_t)n * d.dpi + i;
        if (det_boxes_net != nullptr) det_boxes_net[o] = b;
        det_boxes[o] = (i < nout) ? resize_box(b, im.ratio_h[n], im.ratio_w[n]) : b;
        det_scores[o] = s;
        det_labels[o] = lab;
    }
    if (tid == 0) det_counts[n] = nout;
}

static int make_det(const mb_det_params& p, DetDev& d, DetImages& im) {
    if (p.num_images < 1 || p.num_images > MB_MAX_IMAGES || p.num_classes < 2 || p.max_props_per_image < 1 ||
        p.detections_per_img < 1)
        return MB_ERR_INVALID_ARG;
    d.N = p.num_images; d.C = p.num_classes; d.max_props = p.max_props_per_image; d.dpi = p.detections_per_img;
    d.score_thresh = p.score_thresh; d.min_size = p.min_size;
    d.dw = DecodeWeights{p.wx, p.wy, p.ww, p.wh, p.bbox_xform_clip};
    const long long per_seg = d.max_props < d.dpi ? d.max_props : d.dpi;
    if ((long long)(d.C - 1) * per_seg >= (1ll << 24)) return MB_ERR_UNSUPPORTED;
    if ((long long)d.N * (d.C - 1) * d.max_props >= (1ll << 30)) return MB_ERR_UNSUPPORTED;
    for (int n = 0; n < d.N; ++n) {
        im.h[n] = p.image_h[n]; im.w[n] = p.image_w[n];
        const bool rs = p.orig_h[n] > 0 && p.orig_w[n] > 0;
        // resize_boxes: torch.tensor(new, fp32) / torch.tensor(orig, fp32) per axis
        im.ratio_h[n] = rs ? (float)p.orig_h[n] / (float)p.image_h[n] : 1.0f;
        im.ratio_w[n] = rs ? (float)p.orig_w[n] / (float)p.image_w[n] : 1.0f;
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" size_t mb_det_workspace_bytes(const mb_det_params* p) {
    if (!p) return 0;
    DetDev d; DetImages im;
    if (make_det(*p, d, im) != MB_OK) return 0;
    Carver c(nullptr, 0);
    DetScratch w;
    carve_det(c, w, d);
    return c.off + 1024;
}

extern "C" int mb_det_postprocess(const mb_det_params* p, const float* class_logits, const float* box_regression,
                                  const float* proposals, const int32_t* prop_counts, int32_t packed, float* det_boxes,
                                  float* det_boxes_net, float* det_scores, int64_t* det_labels, int32_t* det_counts,
                                  void* workspace, size_t workspace_bytes, mb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !class_logits || !box_regression || !proposals || !prop_counts || !det_boxes || !det_scores ||
        !det_labels || !det_counts)
        return MB_ERR_INVALID_ARG;
    DetDev d; DetImages im;
    int rc = make_det(*p, d, im);
    if (rc != MB_OK) return rc;
    Carver c(workspace, workspace_bytes);
    DetScratch w;
    carve_det(c, w, d);
    if (!c.ok()) return MB_ERR_WORKSPACE;
    const int G = d.N * (d.C - 1);
    if ((w.zero_bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return MB_ERR_INVALID_ARG;
    const size_t seg_start_off = (size_t)((const char*)w.seg.seg_start - (const char*)workspace);
    if ((seg_start_off & 15) != 0 || seg_start_off + 4 * (size_t)((G + 3) / 4 * 4) > w.zero_bytes) return MB_ERR_INVALID_ARG;
    MB_CUDA(launch_pdl(k_det_init, max(ceil_div(G, 256), 32), 256, 0, stream, G, d.max_props, (long long)(seg_start_off / 16),
                       (uint4*)workspace, (long long)(w.zero_bytes / 16), w.bseg, (long long)w.P));
    MB_CUDA(launch_pdl(k_det_candidates, ceil_div(d.N * d.max_props, kDetThreads), kDetThreads, 0, stream,
                       d, im, class_logits, box_regression, (const float4*)proposals, prop_counts, packed, w));
    MetaRule rule{d.C - 1, 1, p->trick_numel, w.img_max, w.seg_offset};
    MB_CUDA(launch_pdl(k_seg_meta, 1, 1024, 0, stream, w.seg, G, 0, w.mask_words, rule));
    MB_CUDA(launch_pdl(k_rank_in_segment, (int)ceil_div64(w.P, kRankKeys), kRankThreads, 0, stream,
                       w.bkey, w.bbox, w.bseg, w.seg.seg_start, w.seg.seg_count, nullptr, (int)w.P, w.skey, w.sbox));
    rc = launch_mask_and_sweep(w.sbox, w.seg, G, d.max_props, p->nms_thresh, w.mask, w.keepbits, d.dpi, stream, w.seg_offset);
    if (rc != MB_OK) return rc;
    const long long per_seg = d.max_props < d.dpi ? d.max_props : d.dpi;
    const int bound = (int)((d.C - 1) * per_seg);
    const int use_merge = (d.C - 1) <= kDetMergeMaxRuns && d.max_props <= 64 * kSweepSmallMaxWords && 2 * bound <= kDetSortCap;
    int cap = use_merge ? 2 * bound : next_pow2(bound > 2 ? bound : 2);   // merge keeps source + merged order
    if (cap > kDetSortCap) cap = kDetSortCap;                             // larger actual counts sort in w.gkeys
    const int smem = (cap > 2 ? cap : 2) * (int)sizeof(unsigned long long);
    MB_DYN_SMEM(k_det_finalize, smem);
    MB_CUDA(launch_pdl(k_det_finalize, d.N, kDetFinalThreads, smem, stream, d, im, w, (float4*)det_boxes, (float4*)det_boxes_net,
                       det_scores, (long long*)det_labels, det_counts, use_merge, cap));
    return MB_OK;
}
