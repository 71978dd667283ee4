"""Time mosaic.SeamNms (unpack + mb_nms mode 1) for the gathered block sizes of 2/4/8 ranks."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import mosaic  # noqa: E402
from tests.test_mosaic_cpu import synth_tiles  # noqa: E402

DEV = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cases_ = [(w_, 4 * w_) for w_ in (1, 2, 4, 8)] + ([(0, 361)] if len(sys.argv) > 2 else [])
for world, tiles in cases_:
    dpi = 300
    b, s, l, c, o = synth_tiles(num_tiles=tiles, dpi=dpi, seed=1)
    c[:] = dpi
    block = mosaic.pack_block(*(torch.from_numpy(x).to(DEV) for x in (b, s, l)), torch.from_numpy(c).to(DEV).to(torch.int32),
                              torch.from_numpy(o).to(DEV), 0.0, tiles * dpi)
    seam = mosaic.SeamNms(block.shape[0], 3, DEV)
    for _ in range(3):
        seam.launch(block, 0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        seam.launch(block, 0.5)
    e1.record(); torch.cuda.synchronize()
    kept = int(seam.nms.status[0])
    print(f"world {world} tiles {tiles}: rows {block.shape[0]} kept {kept}  {e0.elapsed_time(e1) / reps:.4f} ms per seam NMS", flush=True)
