/*
 * oracle_kernels.c — CPU restatement of the two native kernels on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under miso_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / CPU baseline.
 *
 * The arithmetic restated here lives in a third-party dependency of the reference
 * (torchvision, unpinned by ref:setup.py:16-24; de-facto pin = the installed
 * torchvision 0.26.0+cu128). Its native sources are not on disk, so this file follows
 * the published algorithm as pinned in SURVEY.md Appendix B:
 *   - oracle_nms            <- torchvision::nms CPU  (tv-csrc:ops/cpu/nms_kernel.cpp:116),
 *                              called from tv:ops/boxes.py:48
 *   - oracle_roi_align      <- torchvision::roi_align CPU forward
 *                              (tv-csrc:ops/cpu/roi_align_kernel.cpp:393),
 *                              called from tv:ops/roi_align.py:258-260
 * Parity pin: tests/test_oracle_pin.py checks both against the real torchvision CPU
 * ops (importable in the build container and on the GPU box) and against the golden
 * vectors in tests/golden/ that tests/gen_golden.py produced from those ops.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/build.py).
 * -ffp-contract=off matters: every product and sum is rounded to fp32 separately,
 * exactly as the reference's x86-64 build does (no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- stable descending argsort, NaN first (torch.sort(stable=True, descending=True)) ---- */
static int score_before(float a, float b) {
    /* strict "a sorts before b" for a descending order in which NaN is the largest */
    int an = isnan(a), bn = isnan(b);
    if (an || bn) return an && !bn;
    return a > b;
}

static void merge_sort_idx(const float* s, int64_t* idx, int64_t* tmp, int64_t lo, int64_t hi) {
    if (hi - lo < 2) return;
    int64_t mid = lo + (hi - lo) / 2;
    merge_sort_idx(s, idx, tmp, lo, mid);
    merge_sort_idx(s, idx, tmp, mid, hi);
    int64_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi) {
        /* take from the right run only if it sorts strictly before the left head */
        if (score_before(s[idx[j]], s[idx[i]])) tmp[k++] = idx[j++];
        else tmp[k++] = idx[i++];
    }
    while (i < mid) tmp[k++] = idx[i++];
    while (j < hi) tmp[k++] = idx[j++];
    memcpy(idx + lo, tmp + lo, (size_t)(hi - lo) * sizeof(int64_t));
}

void oracle_argsort_desc_stable(const float* scores, int64_t n, int64_t* order) {
    int64_t* tmp = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    merge_sort_idx(scores, order, tmp, 0, n);
    free(tmp);
}

/*
 * Greedy NMS. boxes [n,4] xyxy fp32, scores [n], threshold compared in double with a
 * strict '>'. Writes kept indices (descending score order) to keep[n]; returns count.
 */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold,
                   int64_t* keep) {
    if (n <= 0) return 0;
    int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    float* area = (float*)malloc((size_t)n * sizeof(float));
    uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
    oracle_argsort_desc_stable(scores, n, order);
    for (int64_t k = 0; k < n; ++k) {
        volatile float w = boxes[4 * k + 2] - boxes[4 * k + 0];
        volatile float h = boxes[4 * k + 3] - boxes[4 * k + 1];
        area[k] = w * h;
    }
    int64_t nk = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = order[_i];
        if (suppressed[i]) continue;
        keep[nk++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2],
                    iy2 = boxes[4 * i + 3], iarea = area[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j];
            if (suppressed[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float uni = iarea + area[j];
            uni = uni - inter;
            float ovr = inter / uni;
            if ((double)ovr > iou_threshold) suppressed[j] = 1;
        }
    }
    free(order); free(area); free(suppressed);
    return nk;
}

/*
 * RoIAlign forward, NCHW fp32. rois [K,5] = (batch, x1, y1, x2, y2). out [K,C,PH,PW].
 * Operation order follows SURVEY.md Appendix B.2 exactly (each op rounded to fp32).
 */
void oracle_roi_align(const float* in, int64_t N, int64_t C, int64_t H, int64_t W,
                      const float* rois, int64_t K, float spatial_scale, int PH, int PW,
                      int sampling_ratio, int aligned, float* out) {
    (void)N;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t k = 0; k < K; ++k) {
        const float* r = rois + 5 * k;
        const int64_t b = (int64_t)r[0];
        const float off = aligned ? 0.5f : 0.0f;
        const float sw = r[1] * spatial_scale - off;
        const float sh = r[2] * spatial_scale - off;
        const float ew = r[3] * spatial_scale - off;
        const float eh = r[4] * spatial_scale - off;
        float rw = ew - sw, rh = eh - sh;
        if (!aligned) { rw = rw > 1.0f ? rw : 1.0f; rh = rh > 1.0f ? rh : 1.0f; }
        const float bh = rh / (float)PH, bw = rw / (float)PW;
        const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
        const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
        const float count = (float)((gh * gw) > 1 ? (gh * gw) : 1);
        for (int64_t c = 0; c < C; ++c) {
            const float* plane = in + (b * C + c) * H * W;
            for (int ph = 0; ph < PH; ++ph) {
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.0f;
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = (sh + (float)ph * bh) + (((float)iy + 0.5f) * bh) / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = (sw + (float)pw * bw) + (((float)ix + 0.5f) * bw) / (float)gw;
                            float yy = y;
                            if (yy < -1.0f || yy > (float)H || x < -1.0f || x > (float)W) continue;
                            if (yy <= 0.0f) yy = 0.0f;
                            if (x <= 0.0f) x = 0.0f;
                            int yl = (int)yy, xl = (int)x, yh, xh;
                            if (yl >= H - 1) { yh = yl = (int)H - 1; yy = (float)yl; } else yh = yl + 1;
                            if (xl >= W - 1) { xh = xl = (int)W - 1; x = (float)xl; } else xh = xl + 1;
                            float ly = yy - (float)yl, lx = x - (float)xl;
                            float hy = 1.0f - ly, hx = 1.0f - lx;
                            float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                            float t = w1 * plane[yl * W + xl];
                            t = t + w2 * plane[yl * W + xh];
                            t = t + w3 * plane[yh * W + xl];
                            t = t + w4 * plane[yh * W + xh];
                            acc = acc + t;
                        }
                    }
                    out[((k * C + c) * PH + ph) * PW + pw] = acc / count;
                }
            }
        }
    }
}

int oracle_abi_version(void) { return 1; }
