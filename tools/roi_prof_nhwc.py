"""RoIAlign (channels-last features), config-2 stress shapes — profiled with ncu."""
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
n = 4
feats = [torch.randn(n, 256, 800 // s, 800 // s, device=DEV).contiguous(memory_format=torch.channels_last) for s in (4, 8, 16, 32)]
boxes = [torch.from_numpy(cases.stress_rois(rng, 1000, (800, 800))).to(DEV) for _ in range(n)]
pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, exact=(len(sys.argv) < 2 or sys.argv[1] != "fast"))
x = {str(i): f for i, f in enumerate(feats)}
for _ in range(3):
    out = pool(x, boxes, [(800, 800)] * n)
torch.cuda.synchronize()
print("ok", out.shape, float(out.abs().mean()))
