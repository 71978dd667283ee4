"""A/B timing of the two channels-last RoIAlign kernels on the bench input (config 2: 4000 proposals of the
RPN stage, 256-channel pyramid): k_roi_align_tma vs k_roi_align_nhwc4d, exact and FMA modes. CUDA events
around each launch; the 218 MB pyramid + 201 MB output exceed L2, so every launch runs cold."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import pipeline, workload  # noqa: E402

dev = torch.device("cuda:0")
w = workload.faster_rcnn_batch(num_images=4, seed=0, features_layout="channels_last")
hp = pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold, crop_capacity_bytes=64 << 20, device=dev)
d = workload.to_device(w, dev)
hp.bind(d["objectness"], d["deltas"], d["features"], d["class_logits"][0], d["box_regression"][0], d["images"])
hp.rpn()
torch.cuda.synchronize()
res = {}
ref = None
for force in (1, 2):
    for exact in (1, 0):
        hp.roi_params.force_gather, hp.roi_params.exact = force, exact
        for _ in range(3):
            hp.roi_align()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            a.record(); hp.roi_align(); b.record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in evs)
        name = {1: "gather", 2: "tma"}[force] + ("_exact" if exact else "_fma")
        res[name] = {"median_ms": t[len(t) // 2], "min_ms": t[0]}
        if exact:
            if ref is None:
                ref = hp.box_features.clone()
            else:
                res[name + "_equals_first_exact"] = bool(torch.equal(ref, hp.box_features))
print(json.dumps(res))
