"""Experiment: does the ORDER of the RoIs (which RoIs run on the same SM at the same time) change the
RoIAlign kernel time? Orders: as given (random), spatially sorted, sorted + interleaved so that the
blocks initially resident on one SM are neighbours in the sorted order."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import ops  # noqa: E402
from tests import cases  # noqa: E402

DEV = "cuda:0"
rng = np.random.default_rng(0)
n, sms = 4, 148
exact = len(sys.argv) < 2 or sys.argv[1] != "fast"
feats = [torch.randn(n, 256, 800 // s, 800 // s, device=DEV).contiguous(memory_format=torch.channels_last) for s in (4, 8, 16, 32)]
boxes = np.concatenate([np.concatenate([np.full((1000, 1), i, np.float32), cases.stress_rois(rng, 1000, (800, 800))], 1) for i in range(n)])
scales = [0.25, 0.125, 0.0625, 0.03125]
thr = ops.level_thresholds(2, 5)
area = (boxes[:, 3] - boxes[:, 1]) * (boxes[:, 4] - boxes[:, 2])
lvl = sum((area >= t).astype(np.int64) for t in thr)
cy, cx = (boxes[:, 2] + boxes[:, 4]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
cell = 64.0 * (2.0 ** lvl)                                     # 16 feature pixels at the RoI's level
key = np.lexsort((cx // cell, cy // cell, boxes[:, 0], lvl))   # level, image, y cell, x cell
K = len(boxes)
per = -(-K // sms)
inter = np.array([min((b % sms) * per + b // sms, K - 1) for b in range(K)])   # block b -> sorted position
orders = {"random": np.arange(K), "sorted": key, "sorted_interleaved": key[np.argsort(np.argsort(inter), kind="stable")] if False else key[inter]}


def run(order):
    r = torch.from_numpy(boxes[order]).to(DEV)
    for _ in range(3):
        out = ops._roi_align_launch(feats, r, scales, thr, (7, 7), 2, False, exact)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        ev[0].record(); out = ops._roi_align_launch(feats, r, scales, thr, (7, 7), 2, False, exact); ev[1].record()
        torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    return sorted(ts)[len(ts) // 2], out


base, ref = run(orders["random"])
for name, order in orders.items():
    ms, out = run(order)
    same = torch.equal(out, ref[torch.from_numpy(np.argsort(np.arange(K))[order]).to(DEV)]) if name != "random" else True
    print(f"{name:20s} {ms:.4f} ms  same={same}")
