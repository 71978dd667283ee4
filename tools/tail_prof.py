"""The mosaic tail on a synthetic 19 x 19 tile block: sparse seam NMS, survivor selection, crop plan, crop gather from
a 16384 x 16384 RGB band. ncu target for the per-kernel split of bench.py's `select_crop` stage."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import mosaic  # noqa: E402
from tests.test_gpu_seam import seam_block  # noqa: E402

DEV = "cuda:0"
rng = np.random.default_rng(0)
block, dpi = seam_block(rng, nty=19, ntx=19, dpi=300, objects=200 * 361, thr=0.3)
g = torch.from_numpy(block).to(DEV)
S = 896 * 18 + 1024
band = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, device=DEV)
seam = mosaic.SparseSeamNms(block.shape[0], dpi, DEV)
crops = mosaic.MosaicCrops(block.shape[0], (S, S), 3, 0.5, 6 << 30, DEV)
crops.bind_band(band, 0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    seam.launch(g, 0.5)
    crops.launch(g, seam.state, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); crops.launch(g, seam.state, 0); e1.record(); torch.cuda.synchronize()
r = crops.results()
print("ok", r["count"], r["bytes"], "select+plan+gather ms", e0.elapsed_time(e1))
