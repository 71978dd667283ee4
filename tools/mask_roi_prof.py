"""Mask-head RoIAlign of BASELINE config 3 (batch 8, 100 detections/img, 256 ch, 14x14): ours vs torchvision CUDA."""
import os
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import ops  # noqa: E402
from tests import cases  # noqa: E402

DEV = "cuda:0"
rng = np.random.default_rng(0)
n, per = 8, 100
feats = [torch.randn(n, 256, 800 // s, 800 // s, device=DEV) for s in (4, 8, 16, 32)]
boxes = [torch.from_numpy(cases.stress_rois(rng, per, (800, 800), side=(24.0, 400.0))).to(DEV) for _ in range(n)]
shapes = [(800, 800)] * n


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for layout in ("nchw", "channels_last"):
    x = {str(i): (f if layout == "nchw" else f.contiguous(memory_format=torch.channels_last)) for i, f in enumerate(feats)}
    for P in (14, 7):
        ours = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)
        tv = torchvision.ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)
        t1, o1 = timed(lambda: ours(x, boxes, shapes))
        t2, o2 = timed(lambda: tv(x, boxes, shapes))
        print(f"{layout:14s} P={P:2d} K={n * per}: ours {t1:.4f} ms, torchvision CUDA {t2:.4f} ms, out {o1.numel() * 4 / 1e6:.0f} MB, max |diff| {float((o1 - o2).abs().max()):.2g}")
