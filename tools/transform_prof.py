"""Input transform of BASELINE config 2 (4 x 1024^2 uint8 -> [4,3,800,800] fp32): ours (one launch) vs torchvision ops on the GPU."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
u8 = [torch.randint(0, 256, (1024, 1024, 3), dtype=torch.uint8, generator=g).to(DEV) for _ in range(4)]
mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
for _ in range(3):
    out, sizes = ops.transform_images(u8, 800, 1333, mean, std)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    out, sizes = ops.transform_images(u8, 800, 1333, mean, std)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
nbytes = out.numel() * 4 + sum(a.numel() for a in u8)
print(f"mb_image_transform: {ms:.4f} ms per batch of 4 ({nbytes / 1e6:.1f} MB moved, {nbytes / ms / 1e6:.0f} GB/s), out {tuple(out.shape)}")
from torchvision.models.detection.transform import GeneralizedRCNNTransform
tr = GeneralizedRCNNTransform(800, 1333, mean, std).eval().to(DEV)
fl = lambda: tr([a.permute(2, 0, 1).to(torch.float32) / 255 for a in u8])[0]
for _ in range(3):
    ref = fl()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    ref = fl()
torch.cuda.synchronize()
print(f"torchvision ToTensor + transform on the GPU: {1e3 * (time.perf_counter() - t0) / 20:.3f} ms per batch; max |diff| {float((ref.tensors - out).abs().max()):.3g}")
