// seam.cu — cross-tile ("seam") NMS of the tiled-mosaic path (BASELINE config 5) as a SPARSE problem.
//
// Semantics (the reference never tiles, so it is defined as a composition of reference
// operations): torchvision.ops.boxes._batched_nms_vanilla (tv:ops/boxes.py:106-120) over ALL gathered rows on
// raw mosaic coordinates, i.e. per label a greedy NMS with the CPU kernel's arithmetic
// (tv-csrc:ops/cpu/nms_kernel.cpp:116, SURVEY.md Appendix B.1: fp32 IoU with one rounding per operation,
// `(double)iou > thr`, candidates in stable descending-score order).
//
// The dense formulation (mb_nms mode 1: an n x n/64 bit matrix per label) costs 13.6 ms for the 108 300 rows of
// a 16384^2 mosaic although almost every pair is disjoint: rows come in groups of `rows_per_tile` consecutive
// rows (one tile's detections) and two rows can only interact if their boxes intersect. Greedy NMS is the
// unique fixed point of   kept(j) <=> no kept i precedes j with same label and IoU(i, j) > thr,   so it can be
// solved on the sparse "i may suppress j" graph:
//   k_seam_prep     per tile: bounding box of its live rows (label >= 0), live count
//   k_seam_pairs    CTA = tile t. The neighbour tiles (bounding box intersects t's) are found by all threads in
//                   parallel; their live rows that intersect t's bounding box are compacted into shared memory;
//                   thread = own row j tests them 32 at a time: cheap tests (label, precedence, boxes not disjoint)
//                   without branching, then the exact IoU predicate on the lane's own marked candidates. The first
//                   four suppressors of a row stay in registers, one edge segment per tile is reserved with one
//                   atomic; a second pass over the candidates runs only for rows with more: no capacity guess per row.
//   k_seam_resolve  CTA = tile. Rounds of: every edge (i -> j) with j undecided looks at state[i]
//                   (kept -> j removed, undecided -> j blocked); undecided rows that are neither become kept.
//                   Decisions are final and order-independent, so tiles run asynchronously on global state;
//                   a few launches (stream order = the barrier between rounds) settle all chains of real data,
//   k_seam_finish   one CTA loops until nothing is undecided (usually nothing to do).
//   k_seam_emit     optional ascending list of kept rows + count; k_seam_select feeds the crop stage.
// No geometric assumption about tiles is made (the bounding boxes come from the data), so the result equals
// the dense per-label NMS for ANY input with iou_threshold >= 0; negative thresholds (where disjoint boxes
// suppress each other) are rejected — callers use mb_nms for them.
#include "common.cuh"
#include "nms_core.cuh"

namespace mb {

constexpr int kSeamChunk = 1536;        // candidate rows staged per chunk (36 KB of shared memory)

struct SeamWs {
    float4* tile_bbox;      // [tiles]
    int* tile_live;         // [tiles]
    int* seg_start;         // [tiles]
    int* seg_count;         // [tiles]
    int* counters;          // [16]: 0 edges reserved, 1 overflow flag, 2.. undecided rows after round r
    int2* edges;            // [edge_cap]
};

__device__ __forceinline__ bool boxes_touch(const float4 a, const float4 b) {
    return !(a.z < b.x || b.z < a.x || a.w < b.y || b.w < a.y);      // closed intersection; NaN -> true (conservative)
}

__global__ void __launch_bounds__(128) k_seam_prep(const float* __restrict__ block, int rows, int rpt, SeamWs w,
                                                  int* __restrict__ state) {
    const int t = blockIdx.x, tid = threadIdx.x;
    float x1 = INFINITY, y1 = INFINITY, x2 = -INFINITY, y2 = -INFINITY;
    int live = 0;
    bool nan = false;
    for (int r = tid; r < rpt; r += blockDim.x) {
        const int row = t * rpt + r;
        if (row >= rows) break;
        const float2* q = reinterpret_cast<const float2*>(block + (size_t)row * 6);
        const float2 a = q[0], b = q[1], c = q[2];
        const bool ok = c.y >= 0.0f;
        state[row] = ok ? 0 : 3;
        if (ok) {
            ++live;
            nan = nan || !(a.x == a.x && a.y == a.y && b.x == b.x && b.y == b.y);
            x1 = fminf(x1, fminf(a.x, b.x)); x2 = fmaxf(x2, fmaxf(a.x, b.x));      // also correct for flipped boxes
            y1 = fminf(y1, fminf(a.y, b.y)); y2 = fmaxf(y2, fmaxf(a.y, b.y));
        }
    }
    __shared__ float sx1[4], sy1[4], sx2[4], sy2[4];
    __shared__ int sl[4], sn[4];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        x1 = fminf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fminf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        x2 = fmaxf(x2, __shfl_xor_sync(0xffffffffu, x2, o)); y2 = fmaxf(y2, __shfl_xor_sync(0xffffffffu, y2, o));
        live += __shfl_xor_sync(0xffffffffu, live, o);
    }
    int anynan = __any_sync(0xffffffffu, nan);
    if ((tid & 31) == 0) { sx1[tid >> 5] = x1; sy1[tid >> 5] = y1; sx2[tid >> 5] = x2; sy2[tid >> 5] = y2; sl[tid >> 5] = live; sn[tid >> 5] = anynan; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < 4; ++q) {
            x1 = fminf(x1, sx1[q]); y1 = fminf(y1, sy1[q]); x2 = fmaxf(x2, sx2[q]); y2 = fmaxf(y2, sy2[q]);
            live += sl[q]; anynan |= sn[q];
        }
        // a NaN coordinate makes the box "touch" everything: give the tile an unbounded box
        if (anynan) { x1 = y1 = -INFINITY; x2 = y2 = INFINITY; }
        w.tile_bbox[t] = make_float4(x1, y1, x2, y2);
        w.tile_live[t] = live;
        if (t == 0) {
            for (int q = 0; q < 16; ++q) w.counters[q] = 0;
        }
    }
}

// candidate row staged in shared memory
struct SeamCand {
    float4 box;
    float area;
    unsigned key;     // desc_score_key: ascending key = descending score
    int label;
    int row;
};

constexpr int kSeamNbCap = 96;          // neighbour tiles listed per round of the pairs kernel
constexpr int kSeamRegEdges = 4;        // suppressors a thread keeps in registers during the first pass

__global__ void __launch_bounds__(1024) k_seam_pairs(const float* __restrict__ block, int rows, int rpt, int tiles,
                                                    float thr_up, SeamWs w, int edge_cap) {
    extern __shared__ __align__(16) unsigned char seam_smem[];
    SeamCand* cand = reinterpret_cast<SeamCand*>(seam_smem);
    __shared__ int s_n, s_base, s_nnb, s_more;
    __shared__ int s_nb[kSeamNbCap];
    __shared__ int s_wcnt[32];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (w.tile_live[t] == 0) {
        if (tid == 0) { w.seg_start[t] = 0; w.seg_count[t] = 0; }
        return;
    }
    const float4 bb = w.tile_bbox[t];
    // own row
    const int my_row = t * rpt + tid;
    bool mine = tid < rpt && my_row < rows;
    float4 mb_ = make_float4(0, 0, 0, 0);
    float marea = 0.f;
    unsigned mkey = 0;
    int mlabel = -1;
    if (mine) {
        const float2* q = reinterpret_cast<const float2*>(block + (size_t)my_row * 6);
        const float2 a = q[0], b = q[1], c = q[2];
        mb_ = make_float4(a.x, a.y, b.x, b.y);
        marea = box_area_rn(mb_);
        mkey = desc_score_key(c.x);
        mlabel = (int)c.y;
        mine = c.y >= 0.0f;
    }
    // Pass 0 finds every suppressor once, keeps the first kSeamRegEdges of a row in registers and counts them; the
    // tile then reserves one contiguous edge segment. Pass 1 runs only if some row had more (rare) and writes the rest.
    int my_count = 0, my_off = 0;
    int reg_e[kSeamRegEdges];
#pragma unroll
    for (int e = 0; e < kSeamRegEdges; ++e) reg_e[e] = -1;
    for (int pass = 0; pass < 2; ++pass) {
        int seen = 0;
        auto process = [&]() {
            const int n = s_n;
            if (mine) {
                // 32 candidates at a time. Pass A marks the ones that pass the cheap tests (same label, precedes me,
                // boxes not disjoint — disjoint boxes have IoU 0, or NaN for two zero-area boxes: never > thr >= 0) without
                // branching; pass B runs the exact IoU on the few marked ones, each lane walking its own bits. In one
                // loop a warp paid for the division whenever any of its 32 rows passed the cheap tests.
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int m1 = min(32, n - i0);
                    unsigned int hit = 0;
#pragma unroll 4
                    for (int k = 0; k < m1; ++k) {
                        const SeamCand c = cand[i0 + k];
                        const bool pre = c.key < mkey || (c.key == mkey && c.row < my_row);      // c must precede me
                        const bool apart = c.box.z < mb_.x || mb_.z < c.box.x || c.box.w < mb_.y || mb_.w < c.box.y;
                        if (c.label == mlabel && pre && !apart) hit |= 1u << k;
                    }
                    while (hit) {
                        const int k = __ffs((int)hit) - 1;
                        hit &= hit - 1u;
                        const SeamCand c = cand[i0 + k];
                        if (!iou_suppresses(c.box, c.area, mb_, marea, thr_up)) continue;
                        if (pass == 0) {
#pragma unroll
                            for (int e = 0; e < kSeamRegEdges; ++e)
                                if (my_count == e) reg_e[e] = c.row;
                            ++my_count;
                        } else {
                            // the neighbour list's order may differ between the passes, so a row that overflowed its
                            // registers rewrites its whole segment
                            const int slot = my_off + seen;
                            if (my_count > kSeamRegEdges && slot < edge_cap) w.edges[slot] = make_int2(c.row, my_row);
                            ++seen;
                        }
                    }
                }
            }
        };
        // stage the live rows of tile uu that touch my tile's box, flushing the chunk through process() when full
        auto stage_tile = [&](int uu) {
            for (int r0 = 0; r0 < rpt; r0 += blockDim.x) {
                if (s_n + (int)blockDim.x > kSeamChunk) {
                    process();
                    __syncthreads();
                    if (tid == 0) s_n = 0;
                    __syncthreads();
                }
                const int r = r0 + tid, row = uu * rpt + r;
                bool ok = r < rpt && row < rows;
                SeamCand c;
                if (ok) {
                    const float2* q = reinterpret_cast<const float2*>(block + (size_t)row * 6);
                    const float2 a = q[0], b = q[1], s = q[2];
                    c.box = make_float4(a.x, a.y, b.x, b.y);
                    ok = s.y >= 0.0f && boxes_touch(c.box, bb);
                    c.area = box_area_rn(c.box);
                    c.key = desc_score_key(s.x);
                    c.label = (int)s.y;
                    c.row = row;
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) s_wcnt[wid] = __popc(m);
                __syncthreads();
                int base = s_n, tot = 0;
                for (int q = 0; q < nw; ++q) {
                    base += q < wid ? s_wcnt[q] : 0;
                    tot += s_wcnt[q];
                }
                if (ok) cand[base + __popc(m & ((1u << lane) - 1u))] = c;
                __syncthreads();
                if (tid == 0) s_n += tot;
                __syncthreads();
            }
        };
        if (tid == 0) s_n = 0;
        for (int u0 = 0; u0 < tiles; u0 += blockDim.x) {
            // neighbour tiles among u0 .. u0 + blockDim.x - 1: every thread tests one tile, the hits go to a short list
            // (in any order: the edges are a set)
            __syncthreads();
            if (tid == 0) s_nnb = 0;
            __syncthreads();
            const int u = u0 + tid;
            if (u < tiles && w.tile_live[u] != 0 && boxes_touch(bb, w.tile_bbox[u])) {
                const int slot = atomicAdd(&s_nnb, 1);
                if (slot < kSeamNbCap) s_nb[slot] = u;
            }
            __syncthreads();
            const int nnb = s_nnb;
            if (nnb <= kSeamNbCap) {
                for (int k = 0; k < nnb; ++k) stage_tile(s_nb[k]);
            } else {
                // more neighbours than the list holds (boxes far larger than a tile): walk the range in order
                const int u1 = min(u0 + (int)blockDim.x, tiles);
                for (int uu = u0; uu < u1; ++uu)
                    if (w.tile_live[uu] != 0 && boxes_touch(bb, w.tile_bbox[uu])) stage_tile(uu);
            }
        }
        __syncthreads();
        process();
        __syncthreads();
        if (pass == 0) {
            // reserve one contiguous edge segment for the tile; per-row offsets by a block scan
            int x = my_count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wcnt[wid] = x;
            if (tid == 0) s_more = 0;
            __syncthreads();
            int pre = 0, tot = 0;
            for (int q = 0; q < nw; ++q) { pre += (q < wid) ? s_wcnt[q] : 0; tot += s_wcnt[q]; }
            if (tid == 0) {
                s_base = atomicAdd(&w.counters[0], tot);
                w.seg_start[t] = s_base;
                w.seg_count[t] = tot;
                if (s_base + tot > edge_cap) w.counters[1] = 1;
            }
            if (my_count > kSeamRegEdges) s_more = 1;
            __syncthreads();
            my_off = s_base + pre + x - my_count;
#pragma unroll
            for (int e = 0; e < kSeamRegEdges; ++e)
                if (e < my_count && my_off + e < edge_cap) w.edges[my_off + e] = make_int2(reg_e[e], my_row);
            if (tot == 0 || !s_more) return;
        }
    }
}

// one resolve sweep of a tile's edges; returns the number of its rows still undecided (valid on thread 0)
__device__ __forceinline__ int seam_resolve_tile(int t, int rpt, int rows, const int2* __restrict__ edges, int e0, int ne,
                                                 volatile int* state, int* s_flag /*[rpt]*/, int max_iters) {
    __shared__ int s_und, s_changed;
    const int tid = threadIdx.x;
    const int row0 = t * rpt;
    int und = 0;
    for (int iter = 0; iter < max_iters; ++iter) {
        for (int r = tid; r < rpt; r += blockDim.x) s_flag[r] = 0;
        if (tid == 0) { s_und = 0; s_changed = 0; }
        __syncthreads();
        for (int e = tid; e < ne; e += blockDim.x) {
            const int2 ed = edges[e0 + e];
            if (state[ed.y] != 0) continue;
            const int s = state[ed.x];
            if (s == 1) atomicOr(&s_flag[ed.y - row0], 2);        // a kept predecessor suppresses me
            else if (s == 0) atomicOr(&s_flag[ed.y - row0], 1);   // undecided predecessor: wait
        }
        __syncthreads();
        int my_und = 0, my_changed = 0;
        for (int r = tid; r < rpt; r += blockDim.x) {
            const int row = row0 + r;
            if (row >= rows || state[row] != 0) continue;
            const int f = s_flag[r];
            if (f & 2) { state[row] = 2; my_changed = 1; }
            else if (f == 0) { state[row] = 1; my_changed = 1; }
            else ++my_und;
        }
        if (my_und) atomicAdd(&s_und, my_und);
        if (my_changed) s_changed = 1;
        __syncthreads();
        und = s_und;
        const int changed = s_changed;
        __syncthreads();
        if (und == 0 || !changed) break;
    }
    return und;
}

__global__ void __launch_bounds__(256) k_seam_resolve(int rows, int rpt, SeamWs w, int* state, int round, int edge_cap) {
    extern __shared__ __align__(16) unsigned char seam_smem[];
    int* s_flag = reinterpret_cast<int*>(seam_smem);
    const int t = blockIdx.x;
    if (w.tile_live[t] == 0 || w.counters[1]) return;
    const int und = seam_resolve_tile(t, rpt, rows, w.edges, w.seg_start[t], w.seg_count[t], state, s_flag, 4);
    if (threadIdx.x == 0 && und) atomicAdd(&w.counters[2 + round], und);
}

__global__ void __launch_bounds__(1024) k_seam_finish(int rows, int rpt, int tiles, SeamWs w, int* state, int last_round,
                                                     long long* status) {
    extern __shared__ __align__(16) unsigned char seam_smem[];
    int* s_flag = reinterpret_cast<int*>(seam_smem);
    __shared__ int s_total;
    if (w.counters[1]) {                      // edge workspace too small: report how many edges are needed
        if (threadIdx.x == 0) { status[0] = -1; status[1] = w.counters[0]; }
        return;
    }
    if (w.counters[2 + last_round] != 0) {    // rare: dependency chains longer than the fixed rounds resolve
        // every sweep decides at least the highest-precedence undecided row, so `rows` sweeps always suffice
        for (int guard = 0;; ++guard) {
            if (guard > rows) {
                if (threadIdx.x == 0) { status[0] = -4; status[1] = w.counters[0]; }
                return;
            }
            if (threadIdx.x == 0) s_total = 0;
            __syncthreads();
            for (int t = 0; t < tiles; ++t) {
                if (w.tile_live[t] == 0) continue;
                const int und = seam_resolve_tile(t, rpt, rows, w.edges, w.seg_start[t], w.seg_count[t], state, s_flag, 64);
                if (threadIdx.x == 0) s_total += und;
            }
            __syncthreads();
            const int tot = s_total;
            __syncthreads();
            if (tot == 0) break;
        }
    }
    if (threadIdx.x == 0) { status[0] = 0; status[1] = w.counters[0]; }
}

// ascending list of kept rows (int64) + count in status[0]; one CTA, 8 rows per thread and round
__global__ void __launch_bounds__(1024) k_seam_emit(int rows, const int* __restrict__ state, long long* __restrict__ keep,
                                                   long long* status) {
    __shared__ int wsum[32];
    __shared__ int s_base;
    if (status[0] < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    constexpr int kE = 8;
    for (int r0 = 0; r0 < rows; r0 += 1024 * kE) {
        int flags = 0, cnt = 0;
#pragma unroll
        for (int k = 0; k < kE; ++k) {
            const int row = r0 + tid * kE + k;
            if (row < rows && state[row] == 1) { flags |= 1 << k; ++cnt; }
        }
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        int pre = 0, tot = 0;
        for (int q = 0; q < 32; ++q) { pre += (q < wid) ? wsum[q] : 0; tot += wsum[q]; }
        int j = s_base + pre + x - cnt;
#pragma unroll
        for (int k = 0; k < kE; ++k)
            if (flags & (1 << k)) keep[j++] = r0 + tid * kE + k;
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) status[0] = s_base;
}

// rows [row_lo, row_lo + n) of the block -> boxes [n, 4] and scores [n] for mb_crop_plan; rows that were not kept get
// score -inf, so the crop plan's `score > threshold` filter drops them and its `src` output is the row offset
__global__ void __launch_bounds__(256) k_seam_select(const float* __restrict__ block, const int* __restrict__ state,
                                                    int row_lo, int n, float4* __restrict__ boxes, float* __restrict__ scores) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2* q = reinterpret_cast<const float2*>(block + (size_t)(row_lo + i) * 6);
    const float2 a = q[0], b = q[1], c = q[2];
    boxes[i] = make_float4(a.x, a.y, b.x, b.y);
    scores[i] = state[row_lo + i] == 1 ? c.x : -INFINITY;
}

static inline int seam_tiles(int64_t rows, int rpt) { return (int)((rows + rpt - 1) / rpt); }

static SeamWs carve_seam(Carver& c, int tiles, int64_t edge_cap) {
    SeamWs w;
    w.tile_bbox = c.take<float4>(tiles);
    w.tile_live = c.take<int>(tiles);
    w.seg_start = c.take<int>(tiles);
    w.seg_count = c.take<int>(tiles);
    w.counters = c.take<int>(16);
    w.edges = c.take<int2>((size_t)edge_cap);
    return w;
}

}  // namespace mb

using namespace mb;

constexpr int kSeamRounds = 3;

extern "C" size_t mb_seam_nms_workspace_bytes(int64_t rows, int32_t rows_per_tile, int64_t edge_capacity) {
    if (rows < 0 || rows_per_tile < 1 || edge_capacity < 0) return 0;
    Carver c(nullptr, 0);
    carve_seam(c, seam_tiles(rows, rows_per_tile) + 1, edge_capacity);
    return c.off + 256;
}

extern "C" int mb_seam_nms(const float* block, int64_t rows, int32_t rows_per_tile, double iou_threshold,
                           int64_t edge_capacity, int32_t* state_out, int64_t* keep_out, int64_t* status_out,
                           void* workspace, size_t workspace_bytes, mb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || rows >= (1ll << 31) || rows_per_tile < 1 || edge_capacity < 1 || edge_capacity >= (1ll << 31))
        return MB_ERR_INVALID_ARG;
    if (!(iou_threshold >= 0.0)) return MB_ERR_UNSUPPORTED;     // disjoint boxes would suppress each other: dense mb_nms
    if (rows_per_tile > 1024) return MB_ERR_UNSUPPORTED;
    if (!status_out) return MB_ERR_INVALID_ARG;
    if (rows == 0) {
        MB_CUDA(cudaMemsetAsync(status_out, 0, 4 * sizeof(int64_t), stream));
        return MB_OK;
    }
    if (!block || !state_out || !workspace) return MB_ERR_INVALID_ARG;
    const int tiles = seam_tiles(rows, rows_per_tile);
    if (workspace_bytes < mb_seam_nms_workspace_bytes(rows, rows_per_tile, edge_capacity)) return MB_ERR_WORKSPACE;
    Carver c(workspace, workspace_bytes);
    SeamWs w = carve_seam(c, tiles + 1, edge_capacity);
    if (!c.ok()) return MB_ERR_WORKSPACE;
    const float thr_up = strict_gt_threshold(iou_threshold);
    k_seam_prep<<<tiles, 128, 0, stream>>>(block, (int)rows, rows_per_tile, w, state_out);
    MB_LAUNCH_CHECK();
    const int threads = ((rows_per_tile + 31) / 32) * 32;
    const int pair_smem = kSeamChunk * (int)sizeof(SeamCand);
    MB_DYN_SMEM(k_seam_pairs, pair_smem);
    k_seam_pairs<<<tiles, threads, pair_smem, stream>>>(block, (int)rows, rows_per_tile, tiles, thr_up, w, (int)edge_capacity);
    MB_LAUNCH_CHECK();
    const int flag_smem = rows_per_tile * (int)sizeof(int);
    for (int r = 0; r < kSeamRounds; ++r) {
        k_seam_resolve<<<tiles, 256, flag_smem, stream>>>((int)rows, rows_per_tile, w, state_out, r, (int)edge_capacity);
        MB_LAUNCH_CHECK();
    }
    k_seam_finish<<<1, 1024, flag_smem, stream>>>((int)rows, rows_per_tile, tiles, w, state_out, kSeamRounds - 1,
                                                 (long long*)status_out);
    MB_LAUNCH_CHECK();
    if (keep_out) {
        k_seam_emit<<<1, 1024, 0, stream>>>((int)rows, state_out, (long long*)keep_out, (long long*)status_out);
        MB_LAUNCH_CHECK();
    }
    return MB_OK;
}

extern "C" int mb_seam_select(const float* block, const int32_t* state, int64_t row_lo, int64_t num_rows,
                              float* boxes_out, float* scores_out, mb_stream_t stream) {
    if (row_lo < 0 || num_rows < 0 || row_lo + num_rows >= (1ll << 31)) return MB_ERR_INVALID_ARG;
    if (num_rows == 0) return MB_OK;
    if (!block || !state || !boxes_out || !scores_out) return MB_ERR_INVALID_ARG;
    k_seam_select<<<(unsigned)((num_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        block, state, (int)row_lo, (int)num_rows, (float4*)boxes_out, scores_out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}
