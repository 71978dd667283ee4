"""paste_masks_in_image: 100 detections of one 1024^2 image (419 MB of fp32 masks), ours vs torchvision's loop."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import ops  # noqa: E402
from tests import cases  # noqa: E402

DEV = "cuda:0"
rng = np.random.default_rng(0)
R, H, W = 100, 1024, 1024
mk = torch.rand((R, 1, 28, 28), device=DEV)
bx = torch.from_numpy(cases.stress_rois(rng, R, (H, W), side=(16.0, 400.0))).to(DEV)
for _ in range(3):
    out = ops.paste_masks_in_image(mk, bx, (H, W))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    out = ops.paste_masks_in_image(mk, bx, (H, W))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = out.numel() * 4
print(f"mb_paste_masks: {ms:.4f} ms per call (incl. torch.empty), {nbytes / 1e6:.0f} MB written, {nbytes / ms / 1e6:.0f} GB/s")
if len(sys.argv) > 1 and sys.argv[1] == "ref":
    from torchvision.models.detection.roi_heads import paste_masks_in_image as tv_paste
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ref = tv_paste(mk, bx, (H, W)); torch.cuda.synchronize()
    print(f"torchvision loop on the GPU: {1e3 * (time.perf_counter() - t0):.1f} ms")
    t0 = time.perf_counter(); refc = tv_paste(mk.cpu(), bx.cpu(), (H, W))
    print(f"torchvision loop on the CPU: {1e3 * (time.perf_counter() - t0):.1f} ms; max |ours - cpu| = {float((out.cpu() - refc).abs().max()):.3g}")
