"""One HotPath step of BASELINE config 2 (the program behind the ncu launch list)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import pipeline, workload  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda:0"
w = workload.faster_rcnn_batch(num_images=4, seed=0, pin=False, features_layout="channels_last")
hp = pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold, crop_capacity_bytes=512 << 20, device=dev)
d = workload.to_device(w, dev)
hp.bind(d["objectness"], d["deltas"], d["features"], d["class_logits"][0], d["box_regression"][0], d["images"])
for _ in range(steps):
    hp.step()
torch.cuda.synchronize()
print("ok", hp.crop_totals.tolist(), hp.prop_counts.tolist(), hp.det_counts.tolist())
