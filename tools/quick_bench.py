"""Quick device timings of individual kernels (development aid; bench.py is the contract)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import ops  # noqa: E402
from tests import cases  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return {"median_ms": t[len(t) // 2], "min_ms": t[0]}


def main():
    res = {"gpu": torch.cuda.get_device_name(0)}
    rng = np.random.default_rng(0)
    n, c = 4, 256
    feats = [torch.randn(n, c, 800 // s, 800 // s, device=DEV) for s in (4, 8, 16, 32)]
    x = {str(i): f for i, f in enumerate(feats)}
    shapes = [(800, 800)] * n
    for P, per in ((7, 1000), (14, 100)):
        boxes = [torch.from_numpy(cases.stress_rois(rng, per, (800, 800))).to(DEV) for _ in range(n)]
        for exact in (True, False):
            pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2, exact=exact)
            pool(x, boxes, shapes)
            rois = ops._f32c(ops.convert_boxes_to_roi_format(boxes))
            fn = lambda: ops._roi_align_launch(feats, rois, pool.scales, pool.thresholds, pool.output_size, 2, False, exact)
            r = timeit(fn)
            out_bytes = rois.shape[0] * c * P * P * 4
            in_bytes = sum(f.numel() * 4 for f in feats)
            r["GBps_out_plus_maps"] = (out_bytes + in_bytes) / r["median_ms"] / 1e6
            res[f"roi_align_P{P}_exact{int(exact)}"] = r
        # torchvision's own CUDA path on the same inputs (kernel to beat)
        import torchvision
        tvp = torchvision.ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)
        res[f"tv_cuda_roi_align_P{P}"] = timeit(lambda: tvp(x, boxes, shapes))
    for k in (1000, 4507, 20000):
        b = torch.from_numpy(cases.random_boxes(rng, k, extent=1024.0)).to(DEV)
        s = torch.from_numpy(cases.distinct_scores(rng, k)).to(DEV)
        res[f"nms_{k}"] = timeit(lambda: ops.nms(b, s, 0.7))
        import torchvision
        res[f"tv_cuda_nms_{k}"] = timeit(lambda: torchvision.ops.nms(b, s, 0.7))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
