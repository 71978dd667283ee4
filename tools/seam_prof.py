"""Time the seam NMS (sparse mb_seam_nms vs dense unpack + mb_nms mode 1) on a synthetic 19 x 19 tile mosaic's
gathered block (108 300 rows, every tile full, objects repeated in every tile they fall into)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import mosaic  # noqa: E402
from tests.test_gpu_seam import seam_block  # noqa: E402

DEV = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
side = int(sys.argv[2]) if len(sys.argv) > 2 else 19
rng = np.random.default_rng(0)
block, dpi = seam_block(rng, nty=side, ntx=side, dpi=300, objects=110 * side * side, thr=0.3)
g = torch.from_numpy(block).to(DEV)
live = int((block[:, 5] >= 0).sum())
sparse = mosaic.SparseSeamNms(block.shape[0], dpi, DEV, want_keep=True)
dense = mosaic.SeamNms(block.shape[0], 3, DEV)
for name, fn in (("sparse", lambda: sparse.launch(g, 0.5)), ("dense", lambda: dense.launch(g, 0.5))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: rows {block.shape[0]} live {live}  {e0.elapsed_time(e1) / reps:.4f} ms", flush=True)
n, edges = sparse.check()
rows = dense.finish()[3]
print("kept sparse", n, "dense", int(rows.numel()), "edges", edges, "equal", bool(torch.equal(torch.sort(rows).values, sparse.keep[:n])))
