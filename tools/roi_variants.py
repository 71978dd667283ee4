"""Time RoIAlign kernel variants (env-selected) on config-2 stress RoIs."""
import json
import os
import sys

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
    return round(t[len(t) // 2], 4)


rng = np.random.default_rng(0)
n, c = 4, 256
feats = [torch.randn(n, c, 800 // s, 800 // s, device=DEV) for s in (4, 8, 16, 32)]
res = {}
VARIANTS = [("legacy_vec4", {"MB_ROI_KERNEL": "legacy", "MB_ROI_VARIANT": "0"}),
            ("legacy_scalar8", {"MB_ROI_KERNEL": "legacy", "MB_ROI_VARIANT": "1"}),
            ("legacy_scalar16", {"MB_ROI_KERNEL": "legacy", "MB_ROI_VARIANT": "3"})]
for P, per in ((7, 1000), (14, 100)):
    boxes = [torch.from_numpy(cases.stress_rois(rng, per, (800, 800))).to(DEV) for _ in range(n)]
    rois = ops._f32c(ops.convert_boxes_to_roi_format(boxes))
    pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)
    pool._setup(feats, [(800, 800)] * n)
    ref = None
    for name, env in VARIANTS:
        os.environ.update(env)
        fn = lambda: ops._roi_align_launch(feats, rois, pool.scales, pool.thresholds, pool.output_size, 2, False, True)
        out = fn()
        if ref is None:
            ref = out.clone()
        res[f"P{P}_{name}"] = {"ms": timeit(fn), "same": bool(torch.equal(out, ref))}
    feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    for nv in ("scalar", "vec"):
      os.environ["MB_ROI_NHWC"] = nv
      for exact in (True, False):
        fn = lambda: ops._roi_align_launch(feats_cl, rois, pool.scales, pool.thresholds, pool.output_size, 2, False, exact)
        out = fn()
        res[f"P{P}_nhwc_{nv}_exact{int(exact)}"] = {"ms": timeit(fn), "same": bool(torch.equal(out, ref)) if exact else bool(torch.allclose(out, ref, rtol=1e-5, atol=5e-5))}
print(json.dumps(res, indent=1))
with open(os.path.join(ROOT, "gpurun_out", "roi_variants.json"), "w") as fh:
    json.dump(res, fh, indent=1)
