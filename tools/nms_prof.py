"""NMS timings (BASELINE config 4 and smaller): ours vs torchvision's CUDA kernels on the same inputs."""
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


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


rng = np.random.default_rng(0)
for n, classes, extent in ((200_000, 80, 4096), (30_000, 80, 2048), (20_000, 1, 2048), (4_000, 1, 1024), (1_000, 1, 800)):
    c = rng.uniform(0, extent, (n, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(256), (n, 2)))
    b = torch.from_numpy(np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)).to(DEV)
    s = torch.from_numpy(cases.distinct_scores(rng, n)).to(DEV)
    idx = torch.from_numpy(rng.integers(0, classes, n).astype(np.int64)).to(DEV)
    if classes > 1:
        ours, k1 = timed(lambda: ops.batched_nms(b, s, idx, 0.5, strategy="vanilla"))
        ours_t, _ = timed(lambda: ops.batched_nms(b, s, idx, 0.5, strategy="trick"))
        ref, k2 = timed(lambda: torchvision.ops.batched_nms(b, s, idx, 0.5))
        print(f"batched_nms n={n} classes={classes}: ours vanilla {ours:.3f} ms, ours trick {ours_t:.3f} ms, torchvision CUDA {ref:.3f} ms; kept {k1.numel()} / {k2.numel()}")
    else:
        ours, k1 = timed(lambda: ops.nms(b, s, 0.5))
        ref, k2 = timed(lambda: torchvision.ops.nms(b, s, 0.5))
        print(f"nms n={n}: ours {ours:.3f} ms, torchvision CUDA {ref:.3f} ms; kept {k1.numel()} / {k2.numel()}")
