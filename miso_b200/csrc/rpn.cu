// rpn.cu — fused RPN post-head stage, straight from the head's NCHW outputs.
//
// Replaces AnchorGenerator.forward (tv:models/detection/anchor_utils.py:115-133),
// concat_box_prediction_layers (rpn.py:81-110), BoxCoder.decode (_utils.py:162-224) and
// RegionProposalNetwork.filter_proposals (rpn.py:231-297).
//
//   k_rpn_hist     per (image, level): 4096-bin histogram of the order-preserving logit key
//                  (warp-aggregated shared-memory atomics); the last CTA of each (image, level)
//                  finds the bin that holds the k-th largest logit          [radix-select pass 1]
//   k_rpn_select   second coalesced pass over the logits: keys above the threshold bin are
//                  taken, keys inside it become candidates (one global atomic per CTA and list);
//                  the last CTA of each (image, level) stages the candidates in shared memory,
//                  resolves them (remaining key bits + index, 8 bits per pass, stopping as soon as
//                  one candidate is left under the prefix) and bitonic-sorts the k winners in
//                  shared memory -> (logit desc, anchor index asc)            [pass 2 + sort]
//   k_rpn_decode   one CTA per (image, level): regenerate the anchor analytically, gather the 4
//                  deltas from NCHW, decode, clip, small-box and score filters, stable compaction
//   k_seg_meta / k_nms_mask / k_nms_fixpoint  (nms_core.cuh), segments = (image, level)
//   k_rpn_finalize one CTA per image: merge the kept boxes of all levels by (score desc, candidate
//                  order asc), keep post_nms_top_n (ranks computed only as far as that cut needs)
// The kernels of the chain start with pdl_enter() and are launched with programmatic stream
// serialization (common.cuh): the next kernel's launch overlaps the drain of the current one.
// Only the selected anchors are decoded: 0.64 MB of logits + 4507*16 B of deltas per image are
// read instead of the reference's 36 B x 159 882 anchors.
#include <math.h>

#include "boxmath.cuh"
#include "nms_core.cuh"
#include "sort_smem.cuh"

namespace mb {

constexpr int kHistBins = 4096;
constexpr int kRpnThreads = 1024;
constexpr int kRpnChunk = 4096;          // logits per CTA in the two streaming passes
constexpr int kSelectSortCap = 8192;     // keys the resolving CTA can sort in shared memory

struct RpnDev {
    int N, L;
    int H[MB_MAX_LEVELS], W[MB_MAX_LEVELS], A[MB_MAX_LEVELS];
    int stride_h[MB_MAX_LEVELS], stride_w[MB_MAX_LEVELS];
    int AL[MB_MAX_LEVELS];        // anchors per level = A*H*W
    int aoff[MB_MAX_LEVELS + 1];  // prefix of AL (flattened anchor index base of a level)
    int k[MB_MAX_LEVELS];         // min(pre_nms_top_n, AL)
    int koff[MB_MAX_LEVELS + 1];  // prefix of k
    int boff[MB_MAX_LEVELS + 1];  // prefix of streaming CTAs per level
    const float* obj[MB_MAX_LEVELS];
    const float* dlt[MB_MAX_LEVELS];
    float4 base[MB_MAX_LEVELS][MB_MAX_ANCHORS_PER_LOC];
    int post_nms_top_n;
    float score_thresh, min_size;
    DecodeWeights dw;
};

struct RpnImages { int h[MB_MAX_IMAGES], w[MB_MAX_IMAGES]; };

struct RpnScratch {
    int* hist;          // [N*L][4096]
    int* ticket;        // [2][N*L]
    int* thr_bin;       // [N*L]
    int* n_above;       // [N*L]
    int* n_sel;         // [N*L]  running count of definite picks
    int* n_cand;        // [N*L]
    unsigned long long* sel;   // [N][Ktot]   definite picks (unsorted), then the sorted winners
    unsigned long long* cand;  // [N][sumA]   keys falling into the threshold bin
    float4* rbox;       // [N*Ktot] clipped boxes, compacted per segment
    float* score;       // [N*Ktot]
    float* img_max;     // [N]
    float* seg_offset;  // [N*L]
    unsigned long long* keepbits;
    unsigned long long* mask;
    SegArrays seg;
    size_t zero_bytes;  // leading region that must be cleared per call
    long long mask_words;
};

static void carve_rpn(Carver& c, RpnScratch& w, const RpnDev& d) {
    const int G = d.N * d.L;
    const int Ktot = d.koff[d.L], sumA = d.aoff[d.L];
    w.hist = c.take<int>((size_t)G * kHistBins);
    w.ticket = c.take<int>(2 * G);
    w.n_sel = c.take<int>(G);
    w.n_cand = c.take<int>(G);
    w.img_max = c.take<float>(d.N);
    w.zero_bytes = c.off;
    w.thr_bin = c.take<int>(G);
    w.n_above = c.take<int>(G);
    w.sel = c.take<unsigned long long>((size_t)d.N * Ktot);
    w.cand = c.take<unsigned long long>((size_t)d.N * sumA);
    w.rbox = c.take<float4>((size_t)d.N * Ktot);
    w.score = c.take<float>((size_t)d.N * Ktot);
    w.seg_offset = c.take<float>(G);
    w.seg = carve_seg_arrays(c, G);
    w.keepbits = c.take<unsigned long long>((size_t)d.N * (Ktot / 64 + d.L + 1));
    long long words = 0;
    for (int l = 0; l < d.L; ++l) words += (long long)d.k[l] * ((d.k[l] + 63) / 64);
    w.mask_words = words * d.N;
    w.mask = c.take<unsigned long long>((size_t)w.mask_words);
}

// flattened anchor index inside a level (reference order (h*W + w)*A + a) from the memory
// order of the NCHW logits (a, h, w)
__device__ __forceinline__ int ref_index_from_mem(int m, int A, int HW) {
    const int a = m / HW, hw = m - a * HW;
    return hw * A + a;
}

__device__ __forceinline__ void block_to_level(const RpnDev& d, int b, int& n, int& l, int& chunk) {
    const int per_img = d.boff[d.L];
    n = b / per_img;
    const int r = b - n * per_img;
    l = 0;
    while (l + 1 < d.L && r >= d.boff[l + 1]) ++l;
    chunk = r - d.boff[l];
}

__global__ void __launch_bounds__(kRpnThreads) k_rpn_hist(const RpnDev d, RpnScratch w) {
    pdl_enter();
    __shared__ int sh[kHistBins];
    __shared__ int s_last;
    int n, l, chunk;
    block_to_level(d, blockIdx.x, n, l, chunk);
    const int g = n * d.L + l, tid = threadIdx.x;
    for (int i = tid; i < kHistBins; i += kRpnThreads) sh[i] = 0;
    __syncthreads();
    const float* obj = d.obj[l] + (size_t)n * d.AL[l];
    const int end = min(d.AL[l], (chunk + 1) * kRpnChunk);
    for (int m = chunk * kRpnChunk + tid; m < end; m += kRpnThreads)
        atomicAdd(&sh[desc_score_key(__ldg(obj + m)) >> 20], 1);
    __syncthreads();
    int* gh = w.hist + (size_t)g * kHistBins;
    for (int i = tid; i < kHistBins; i += kRpnThreads)
        if (sh[i]) atomicAdd(&gh[i], sh[i]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&w.ticket[g], 1) == d.boff[l + 1] - d.boff[l] - 1);
    __syncthreads();
    if (!s_last) return;
    // last CTA of this (image, level): locate the bin of the k-th largest logit
    __threadfence();
    __shared__ long long scan_sh[64];
    const int k = d.k[l];
    static_assert(kHistBins == 4 * kRpnThreads, "one thread per four consecutive bins");
    int v[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = __ldcg(&gh[4 * tid + q]); sum += v[q]; }
    long long tot;
    int acc = (int)block_excl_scan_1024(sum, scan_sh, tot);       // one block scan over the per-thread sums
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (acc < k && acc + v[q] >= k) { w.thr_bin[g] = 4 * tid + q; w.n_above[g] = acc; }
        acc += v[q];
    }
}

__global__ void __launch_bounds__(kRpnThreads) k_rpn_select(const RpnDev d, RpnScratch w, long long* topk_idx_out) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned long long keys[];  // [kSelectSortCap]
    __shared__ int s_last, s_need, s_prefix_ok, s_single;
    __shared__ unsigned long long s_thr;
    __shared__ int hist256[256];
    int n, l, chunk;
    block_to_level(d, blockIdx.x, n, l, chunk);
    const int g = n * d.L + l, tid = threadIdx.x;
    const int Ktot = d.koff[d.L], sumA = d.aoff[d.L];
    const int HW = d.H[l] * d.W[l];
    const int tb = w.thr_bin[g];
    const float* obj = d.obj[l] + (size_t)n * d.AL[l];
    unsigned long long* sel = w.sel + (size_t)n * Ktot + d.koff[l];
    unsigned long long* cand = w.cand + (size_t)n * sumA + d.aoff[l];
    const int end = min(d.AL[l], (chunk + 1) * kRpnChunk);
    // this CTA's logits: all loads first (one latency), the definite winners (bin < threshold bin) and the threshold-bin
    // candidates counted per thread, one block scan, ONE global atomic per list for the whole CTA, then the writes
    // (the order inside the lists is irrelevant: keys are unique and sorted / selected below)
    constexpr int kPer = kRpnChunk / kRpnThreads;
    unsigned int key[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int m = chunk * kRpnChunk + q * kRpnThreads + tid;
        key[q] = m < end ? desc_score_key(__ldg(obj + m)) : 0xffffffffu;
    }
    int nsel = 0, ncnd = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int m = chunk * kRpnChunk + q * kRpnThreads + tid;
        const int bin = m < end ? (int)(key[q] >> 20) : kHistBins;
        nsel += bin < tb; ncnd += bin == tb;
    }
    {
        __shared__ int s_wsel[32], s_wcnd[32], s_bsel, s_bcnd;
        const int lane = tid & 31, wid = tid >> 5;
        int xs = nsel, xc = ncnd;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ys = __shfl_up_sync(0xffffffffu, xs, o), yc = __shfl_up_sync(0xffffffffu, xc, o);
            if (lane >= o) { xs += ys; xc += yc; }
        }
        if (lane == 31) { s_wsel[wid] = xs; s_wcnd[wid] = xc; }
        __syncthreads();
        int ps = 0, pc = 0, ts = 0, tc = 0;
        for (int q = 0; q < kRpnThreads / 32; ++q) {
            ps += q < wid ? s_wsel[q] : 0; ts += s_wsel[q];
            pc += q < wid ? s_wcnd[q] : 0; tc += s_wcnd[q];
        }
        if (tid == 0) s_bsel = ts ? atomicAdd(&w.n_sel[g], ts) : 0;
        if (tid == 32) s_bcnd = tc ? atomicAdd(&w.n_cand[g], tc) : 0;
        __syncthreads();
        int js = s_bsel + ps + xs - nsel, jc = s_bcnd + pc + xc - ncnd;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int m = chunk * kRpnChunk + q * kRpnThreads + tid;
            if (m >= end) continue;
            const int bin = (int)(key[q] >> 20);
            const unsigned long long ck = ((unsigned long long)key[q] << 32) | (unsigned int)ref_index_from_mem(m, d.A[l], HW);
            if (bin < tb) sel[js++] = ck;
            else if (bin == tb) cand[jc++] = ck;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&w.ticket[d.N * d.L + g], 1) == d.boff[l + 1] - d.boff[l] - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last CTA: resolve the threshold bin, then sort exactly k winners ----
    const int k = d.k[l];
    const int na = w.n_above[g];
    const int nc = __ldcg(&w.n_cand[g]);
    int need = k - na;  // winners still to take from the candidates (1 <= need <= nc)
    unsigned long long prefix = 0x000fffffffffffffull;   // take every candidate unless narrowed below
    // (sorting all candidates together with the definite picks instead of narrowing first measured slower: the
    // bitonic sort doubles in size, the seven radix passes over a few hundred keys are cheap)
    // the candidates are read many times below: once from global memory into the upper half of the key buffer
    unsigned long long* scand = keys + kSelectSortCap / 2;
    const bool staged = nc <= kSelectSortCap / 2;
    if (staged) {
        for (int i = tid; i < nc; i += kRpnThreads) scand[i] = __ldcg(&cand[i]);
        __syncthreads();
    }
    if (nc > need) {
        // radix-select the need-th smallest candidate key over the remaining 52 bits, 8 bits per pass
        unsigned long long pmask = 0;
        prefix = 0;
        for (int shift = 44; shift >= -4; shift -= 8) {
            const int sft = shift < 0 ? 0 : shift;
            const int bits = shift < 0 ? 4 : 8;
            for (int i = tid; i < 256; i += kRpnThreads) hist256[i] = 0;
            __syncthreads();
            for (int i = tid; i < nc; i += kRpnThreads) {
                const unsigned long long ck = (staged ? scand[i] : __ldcg(&cand[i])) & 0x000fffffffffffffull;
                if ((ck & pmask) == prefix) atomicAdd(&hist256[(int)((ck >> sft) & ((1u << bits) - 1))], 1);
            }
            __syncthreads();
            if (tid < 32) {   // warp scan over 256 bins, 8 per lane
                int loc[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { loc[q] = hist256[tid * 8 + q]; sum += loc[q]; }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += y; }
                int cum = inc - sum;   // candidates in bins below this lane's first bin
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (cum < need && cum + loc[q] >= need) {
                        s_need = need - cum;
                        s_thr = prefix | ((unsigned long long)(tid * 8 + q) << sft);
                        s_single = loc[q] == 1;
                    }
                    cum += loc[q];
                }
            }
            __syncthreads();
            need = s_need;
            prefix = s_thr;
            pmask |= ((unsigned long long)((1u << bits) - 1)) << sft;
            const bool single = s_single != 0;
            __syncthreads();
            if (single && sft > 0) {
                // one candidate left under the prefix: it is the last winner; take its remaining bits directly
                // instead of resolving them eight at a time (typical: the score bits alone separate the candidates)
                for (int i = tid; i < nc; i += kRpnThreads) {
                    const unsigned long long ck = (staged ? scand[i] : __ldcg(&cand[i])) & 0x000fffffffffffffull;
                    if ((ck & pmask) == prefix) s_thr = ck;
                }
                __syncthreads();
                prefix = s_thr;
                break;
            }
        }
        // prefix is now the exact low-52-bit value of the last winner (keys are unique)
    }
    if (tid == 0) s_prefix_ok = 0;
    __syncthreads();
    for (int i = tid; i < na; i += kRpnThreads) keys[i] = __ldcg(&sel[i]);
    for (int i0 = 0; i0 < nc; i0 += kRpnThreads) {
        const int i = i0 + tid;
        unsigned long long ck = 0;
        bool win = false;
        if (i < nc) { ck = staged ? scand[i] : __ldcg(&cand[i]); win = (ck & 0x000fffffffffffffull) <= prefix; }
        const int slot = warp_alloc_slot(&s_prefix_ok, win);
        if (win) keys[na + slot] = ck;
    }
    __syncthreads();
    const int total = na + s_prefix_ok;  // == k
    const int np2 = next_pow2(max(total, 2));
    for (int i = total + tid; i < np2; i += kRpnThreads) keys[i] = ~0ull;
    bitonic_sort_smem(keys, np2);
    for (int i = tid; i < k; i += kRpnThreads) {
        sel[i] = keys[i];
        if (topk_idx_out != nullptr) topk_idx_out[(size_t)n * Ktot + d.koff[l] + i] = (long long)(keys[i] & 0xffffffffull) + d.aoff[l];
    }
}

// one CTA per (image, level): decode the winners in order, filter, compact
__global__ void __launch_bounds__(kRpnThreads) k_rpn_decode(const RpnDev d, const RpnImages im, RpnScratch w) {
    pdl_enter();
    __shared__ int warp_cnt[32];
    __shared__ int s_base;
    const int g = blockIdx.x, n = g / d.L, l = g - n * d.L;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int Ktot = d.koff[d.L];
    const int k = d.k[l], A = d.A[l], H = d.H[l], W = d.W[l], HW = H * W;
    const unsigned long long* sel = w.sel + (size_t)n * Ktot + d.koff[l];
    const float* obj = d.obj[l] + (size_t)n * d.AL[l];
    const float* dlt = d.dlt[l] + (size_t)n * 4 * d.AL[l];
    const int seg_start = n * Ktot + d.koff[l];
    const float img_h = (float)im.h[n], img_w = (float)im.w[n];
    if (tid == 0) s_base = 0;
    __syncthreads();
    float vmax = 0.0f;
    for (int i0 = 0; i0 < k; i0 += kRpnThreads) {
        const int i = i0 + tid;
        bool ok = false;
        float4 box = make_float4(0, 0, 0, 0);
        float sc = 0.0f;
        if (i < k) {
            const int idx = (int)(sel[i] & 0xffffffffull);
            const int a = idx % A, hw = idx / A;
            const int h = hw / W, x = hw - h * W;
            const float logit = __ldg(obj + (size_t)a * HW + hw);
            const float4 dl = make_float4(__ldg(dlt + (size_t)(4 * a + 0) * HW + hw), __ldg(dlt + (size_t)(4 * a + 1) * HW + hw),
                                          __ldg(dlt + (size_t)(4 * a + 2) * HW + hw), __ldg(dlt + (size_t)(4 * a + 3) * HW + hw));
            const float4 anchor = grid_anchor(d.base[l][a], h, x, d.stride_h[l], d.stride_w[l]);
            box = clip_box(decode_box(anchor, dl, d.dw), img_h, img_w);
            sc = sigmoid_rn(logit);
            ok = box_not_small(box, d.min_size) && (sc >= d.score_thresh);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int q = 0; q < 32; ++q) { const int c = warp_cnt[q]; pre += (q < wid) ? c : 0; tot += c; }
        const int base = s_base;
        if (ok) {
            const int pos = seg_start + base + pre + __popc(m & ((1u << lane) - 1));
            w.rbox[pos] = box;
            w.score[pos] = sc;
            vmax = fmaxf(vmax, fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)));
        }
        __syncthreads();
        if (tid == 0) s_base = base + tot;
        __syncthreads();
    }
    // boxes are clipped to [0, size], so the int view of the float preserves the order
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) atomicMax((int*)&w.img_max[n], __float_as_int(vmax));
    if (tid == 0) { w.seg.seg_start[g] = seg_start; w.seg.seg_count[g] = s_base; }
}

// one CTA per image: merge the kept boxes of all levels by (score desc, candidate order asc).
// Every level's kept list is already in that order (sweep order), so the global rank of a box is
// its index in its own list plus, per other level, a binary search — no sort.
__global__ void __launch_bounds__(kRpnThreads) k_rpn_finalize(const RpnDev d, RpnScratch w, float4* proposals_out,
                                                             float* scores_out, int* counts_out, int sort_cap) {
    pdl_enter();
    extern __shared__ unsigned long long keys[];
    __shared__ int run_off[MB_MAX_LEVELS + 1];
    __shared__ int wpre[MB_MAX_LEVELS][kSweepSmallMaxWords + 1];
    __shared__ unsigned long long kword[MB_MAX_LEVELS][kSweepSmallMaxWords];      // the keep words themselves
    __shared__ int s_cnt[MB_MAX_LEVELS], s_start[MB_MAX_LEVELS], s_T[MB_MAX_LEVELS];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int Ktot = d.koff[d.L];
    // per level: segment metadata, the keep words and their popcounts, staged in shared memory by one thread per
    // (level, word) — one round of independent loads; everything after this reads global memory only for the
    // scores and boxes of kept proposals
    if (tid < d.L) {
        const int g = n * d.L + tid;
        s_cnt[tid] = w.seg.seg_count[g]; s_start[tid] = w.seg.seg_start[g]; s_T[tid] = w.seg.seg_words[g];
    }
    {
        const int l = tid / (kSweepSmallMaxWords + 1), q = tid - l * (kSweepSmallMaxWords + 1);
        if (l < d.L && q < kSweepSmallMaxWords) {
            const int g = n * d.L + l;
            const int T = w.seg.seg_words[g];
            const unsigned long long word = (q < T) ? w.keepbits[w.seg.keep_off[g] + q] : 0ull;
            kword[l][q] = word;
            wpre[l][q] = __popcll(word);
        }
    }
    __syncthreads();
    {
        // exclusive scan of a level's word counts by one warp (two words per lane, kSweepSmallMaxWords == 64)
        const int l = tid >> 5, lane = tid & 31;
        if (l < d.L) {
            const int T = s_T[l];
            const int c0 = lane < T ? wpre[l][lane] : 0, c1 = lane + 32 < T ? wpre[l][lane + 32] : 0;
            int x0 = c0, x1 = c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
                if (lane >= o) { x0 += y0; x1 += y1; }
            }
            const int tot0 = __shfl_sync(0xffffffffu, x0, 31), tot1 = __shfl_sync(0xffffffffu, x1, 31);
            if (lane < T) wpre[l][lane] = x0 - c0;
            if (lane + 32 < T) wpre[l][lane + 32] = tot0 + x1 - c1;
            if (lane == 0) wpre[l][T] = tot0 + tot1;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < d.L; ++l) { run_off[l] = acc; acc += wpre[l][s_T[l]]; }
        run_off[d.L] = acc;
    }
    __syncthreads();
    // one flat loop over the candidate slots of all levels (slot e of the image belongs to level l with
    // koff[l] <= e < koff[l+1]); the kept test runs on shared memory, only kept slots load their score, and the
    // loads of a thread's slots are independent of each other
    for (int e0 = 0; e0 < Ktot; e0 += 4 * kRpnThreads) {
        int pos[4], pp[4];
        float sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * kRpnThreads + tid;
            pos[u] = -1; pp[u] = 0;
            if (e < Ktot) {
                int l = 0;
                while (l + 1 < d.L && e >= d.koff[l + 1]) ++l;
                const int q = e - d.koff[l];
                if (q < s_cnt[l]) {
                    const unsigned long long word = kword[l][q >> 6];
                    if ((word >> (q & 63)) & 1ull) {
                        pos[u] = run_off[l] + wpre[l][q >> 6] + __popcll(word & ((1ull << (q & 63)) - 1ull));
                        pp[u] = s_start[l] + q;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) sc[u] = pos[u] >= 0 ? __ldg(w.score + pp[u]) : 0.0f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (pos[u] >= 0) keys[pos[u]] = ((unsigned long long)desc_score_key(sc[u]) << 32) | (unsigned int)(pp[u] - n * Ktot);
    }
    __syncthreads();
    const int total = run_off[d.L];
    const int nout = min(total, d.post_nms_top_n);
    for (int e = tid; e < total; e += kRpnThreads) {
        int l = 0;
        while (e >= run_off[l + 1]) ++l;
        const int rank = merged_rank_capped(keys, run_off, d.L, l, e - run_off[l], nout);
        if (rank < nout) {
            const int p = n * Ktot + (int)(keys[e] & 0xffffffffull);
            proposals_out[(size_t)n * d.post_nms_top_n + rank] = w.rbox[p];
            scores_out[(size_t)n * d.post_nms_top_n + rank] = w.score[p];
        }
    }
    for (int i = nout + tid; i < d.post_nms_top_n; i += kRpnThreads) {
        proposals_out[(size_t)n * d.post_nms_top_n + i] = make_float4(0, 0, 0, 0);
        scores_out[(size_t)n * d.post_nms_top_n + i] = 0.0f;
    }
    if (tid == 0) counts_out[n] = nout;
    (void)sort_cap;
}

static int make_dev(const mb_rpn_params& p, RpnDev& d, RpnImages& im, bool need_ptrs) {
    if (p.num_images < 1 || p.num_images > MB_MAX_IMAGES || p.num_levels < 1 || p.num_levels > MB_MAX_LEVELS)
        return MB_ERR_INVALID_ARG;
    if (p.pre_nms_top_n < 1 || p.post_nms_top_n < 1) return MB_ERR_INVALID_ARG;
    d.N = p.num_images; d.L = p.num_levels;
    d.aoff[0] = d.koff[0] = d.boff[0] = 0;
    for (int l = 0; l < d.L; ++l) {
        if (p.feat_h[l] < 1 || p.feat_w[l] < 1 || p.anchors_per_loc[l] < 1 || p.anchors_per_loc[l] > MB_MAX_ANCHORS_PER_LOC)
            return MB_ERR_INVALID_ARG;
        d.H[l] = p.feat_h[l]; d.W[l] = p.feat_w[l]; d.A[l] = p.anchors_per_loc[l];
        d.stride_h[l] = p.stride_h[l]; d.stride_w[l] = p.stride_w[l];
        const long long al = (long long)d.A[l] * d.H[l] * d.W[l];
        if (al >= (1ll << 30)) return MB_ERR_UNSUPPORTED;
        d.AL[l] = (int)al;
        d.k[l] = (int)((long long)p.pre_nms_top_n < al ? p.pre_nms_top_n : al);
        if (d.k[l] > kSelectSortCap / 2) return MB_ERR_UNSUPPORTED;
        d.aoff[l + 1] = d.aoff[l] + d.AL[l];
        d.koff[l + 1] = d.koff[l] + d.k[l];
        d.boff[l + 1] = d.boff[l] + ceil_div(d.AL[l], kRpnChunk);
        d.obj[l] = p.objectness[l]; d.dlt[l] = p.deltas[l];
        if (need_ptrs && (!d.obj[l] || !d.dlt[l])) return MB_ERR_INVALID_ARG;
        for (int a = 0; a < d.A[l]; ++a)
            d.base[l][a] = make_float4(p.base_anchors[l][a][0], p.base_anchors[l][a][1], p.base_anchors[l][a][2], p.base_anchors[l][a][3]);
    }
    if (d.koff[d.L] > 16384) return MB_ERR_UNSUPPORTED;
    d.post_nms_top_n = p.post_nms_top_n;
    d.score_thresh = p.score_thresh; d.min_size = p.min_size;
    d.dw = DecodeWeights{p.wx, p.wy, p.ww, p.wh, p.bbox_xform_clip};
    for (int n = 0; n < d.N; ++n) { im.h[n] = p.image_h[n]; im.w[n] = p.image_w[n]; }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" size_t mb_rpn_workspace_bytes(const mb_rpn_params* p) {
    if (!p) return 0;
    RpnDev d; RpnImages im;
    if (make_dev(*p, d, im, false) != MB_OK) return 0;
    Carver c(nullptr, 0);
    RpnScratch w;
    carve_rpn(c, w, d);
    return c.off + 1024;
}

extern "C" int mb_rpn_proposals(const mb_rpn_params* p, float* proposals_out, float* scores_out, int32_t* counts_out,
                                int64_t* topk_idx_out, void* workspace, size_t workspace_bytes, mb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !proposals_out || !scores_out || !counts_out) return MB_ERR_INVALID_ARG;
    RpnDev d; RpnImages im;
    int rc = make_dev(*p, d, im, true);
    if (rc != MB_OK) return rc;
    Carver c(workspace, workspace_bytes);
    RpnScratch w;
    carve_rpn(c, w, d);
    if (!c.ok()) return MB_ERR_WORKSPACE;
    const int G = d.N * d.L;
    MB_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, stream));
    const int stream_blocks = d.N * d.boff[d.L];
    MB_CUDA(launch_pdl(k_rpn_hist, stream_blocks, kRpnThreads, 0, stream, d, w));
    const int sel_smem = kSelectSortCap * (int)sizeof(unsigned long long);
    MB_DYN_SMEM(k_rpn_select, sel_smem);
    MB_CUDA(launch_pdl(k_rpn_select, stream_blocks, kRpnThreads, sel_smem, stream, d, w, (long long*)topk_idx_out));
    MB_CUDA(launch_pdl(k_rpn_decode, G, kRpnThreads, 0, stream, d, im, w));
    MetaRule rule{d.L, 0, p->trick_numel, w.img_max, w.seg_offset};
    MB_CUDA(launch_pdl(k_seg_meta, 1, 1024, 0, stream, w.seg, G, 0, w.mask_words, rule));
    int max_k = 0;
    for (int l = 0; l < d.L; ++l) max_k = max_k > d.k[l] ? max_k : d.k[l];
    rc = launch_mask_and_sweep(w.rbox, w.seg, G, max_k, p->nms_thresh, w.mask, w.keepbits, d.post_nms_top_n, stream,
                               w.seg_offset);
    if (rc != MB_OK) return rc;
    const int cap = next_pow2(d.koff[d.L] > 2 ? d.koff[d.L] : 2);
    const int fin_smem = cap * (int)sizeof(unsigned long long);
    MB_DYN_SMEM(k_rpn_finalize, fin_smem);
    MB_CUDA(launch_pdl(k_rpn_finalize, d.N, kRpnThreads, fin_smem, stream, d, w, (float4*)proposals_out, scores_out, counts_out, cap));
    return MB_OK;
}
