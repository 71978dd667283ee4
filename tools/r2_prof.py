"""Round-2 ncu target: one warm launch set of (a) RoIAlign on the default gather kernel, (b) RoIAlign on the opt-in TMA
route, (c) the sparse seam NMS over a 19 x 19 synthetic block. Config-2 stress shapes, channels-last maps.
  ncu --set full --clock-control none --import-source on -k regex:'k_roi_align|k_roi_geom|k_seam_pairs' python tools/r2_prof.py
profiles the last of the three passes with --launch-skip (see profiles/README.md)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import mosaic, ops  # noqa: E402
from tests import cases  # noqa: E402
from tests.test_gpu_seam import seam_block  # noqa: E402

DEV = "cuda:0"
rng = np.random.default_rng(0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4          # images (tiles) per launch
feats = [torch.randn(n, 256, 800 // s, 800 // s, device=DEV).contiguous(memory_format=torch.channels_last) for s in (4, 8, 16, 32)]
boxes = [torch.from_numpy(cases.stress_rois(rng, 1000, (800, 800))).to(DEV) for _ in range(n)]
x = {str(i): f for i, f in enumerate(feats)}
pools = [ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, exact=True, force_gather=r) for r in ("gather", "tma")]
block, dpi = seam_block(rng, nty=19, ntx=19, dpi=300, objects=110 * 361, thr=0.3)
g = torch.from_numpy(block).to(DEV)
seam = mosaic.SparseSeamNms(block.shape[0], dpi, DEV)
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(passes):
    outs = [p(x, boxes, [(800, 800)] * n) for p in pools]
    seam.launch(g, 0.5)
torch.cuda.synchronize()
print("ok", bool(torch.equal(outs[0], outs[1])), seam.check())
