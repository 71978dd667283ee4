"""bench.py's roofline numerator: roi_align_algorithmic_bytes (vectorised torch rasteriser of the pixels a launch
touches) against a plain per-RoI, per-sample loop over the reference's sampling formula (SURVEY.md Appendix B.2 /
tv-csrc roi_align: aligned=False, sampling_ratio 2) on a small pyramid. Runs on the CPU."""
import math

import numpy as np
import torch

import bench

F = np.float32


def touched_by_loop(props, counts, grids, thresholds, scales, pooled, sr):
    maps = [np.zeros((len(props), gh, gw), bool) for gh, gw in grids]
    live = 0
    for n in range(len(props)):
        for r in range(int(counts[n])):
            x1, y1, x2, y2 = (F(v) for v in props[n, r])
            live += 1
            area = F(F(x2 - x1) * F(y2 - y1))
            lvl = sum(1 for t in thresholds if area >= F(t))
            gh, gw = grids[lvl]
            sc = F(scales[lvl])

            def taps(lo_c, hi_c, size):
                s0, e0 = F(lo_c * sc), F(hi_c * sc)
                ext = max(F(e0 - s0), F(1.0))
                b = F(ext / F(pooled))
                out = []
                for p in range(pooled):
                    for i in range(sr):
                        v = F(F(s0 + F(F(p) * b)) + F(F(F(F(i) + F(0.5)) * b) / F(sr)))
                        if v < -1.0 or v > size:
                            out.append(None)
                            continue
                        v = max(v, F(0.0))
                        lo = min(int(v), size - 1)
                        out.append((lo, min(lo + 1, size - 1)))
                return out

            ys, xs = taps(y1, y2, gh), taps(x1, x2, gw)
            for ty in ys:
                for tx in xs:
                    if ty is None or tx is None:
                        continue
                    for yy in ty:
                        for xx in tx:
                            maps[lvl][n, yy, xx] = True
    return live, sum(int(m.sum()) for m in maps)


def test_algorithmic_bytes_equal_the_per_sample_loop():
    rng = np.random.default_rng(3)
    n, R, C, P, sr = 2, 40, 16, 7, 2
    hw = (160, 224)
    grids = [(hw[0] // s, hw[1] // s) for s in (4, 8, 16, 32)]
    scales = [1.0 / s for s in (4, 8, 16, 32)]
    # LevelMapper thresholds on the area: k = floor(4 + log2(sqrt(area) / 224)) clamped to 2..5 -> area >= (224 * 2^(k-4))^2
    thresholds = [float((224.0 * 2.0 ** (k - 4)) ** 2) for k in (3, 4, 5)]
    c = rng.uniform(-10, [hw[1] + 10, hw[0] + 10], (n, R, 2))
    s = np.exp(rng.uniform(np.log(3), np.log(260), (n, R, 2)))
    props = np.concatenate([c - s / 2, c + s / 2], 2).astype(F)
    props[0, 0] = [-50, -50, -20, -20]                       # entirely outside: every sample invalid
    props[0, 1] = [0, 0, hw[1], hw[0]]                       # the whole image on the coarsest level
    counts = np.array([R, R - 7], np.int32)
    got, k_live, touched = bench.roi_align_algorithmic_bytes(torch.from_numpy(props), torch.from_numpy(counts), grids,
                                                             thresholds, scales, C, P, sr)
    live, want = touched_by_loop(props, counts, grids, thresholds, scales, P, sr)
    assert k_live == live == int(counts.sum())
    assert touched == want and want > 0
    assert got == 20 * live + 4 * C * P * P * live + 4 * C * want
    assert not math.isnan(got)
