"""Experiment: one HotPath.step() captured in a CUDA graph vs plain stream launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import pipeline, workload  # noqa: E402

w = workload.faster_rcnn_batch(num_images=4, seed=0, features_layout="channels_last", pin=False)
hp = pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold)
d = workload.to_device(w, "cuda:0")
hp.bind(d["objectness"], d["deltas"], d["features"], d["class_logits"][0], d["box_regression"][0], d["images"])
for _ in range(3):
    hp.step()
torch.cuda.synchronize()


def timed(fn, k=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


print("stream launches: %.4f ms/step" % timed(hp.step))
ref = [t.clone() for t in (hp.det_boxes, hp.det_counts, hp.crop_totals)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    hp.step()
    with torch.cuda.graph(g, stream=s):
        hp.step()
torch.cuda.current_stream().wait_stream(s)
print("graph replay:    %.4f ms/step" % timed(g.replay))
torch.cuda.synchronize()
print("same results:", all(torch.equal(a, b) for a, b in zip(ref, (hp.det_boxes, hp.det_counts, hp.crop_totals))))
