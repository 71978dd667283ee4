"""Context numbers: whole Faster R-CNN R50-FPN forward, batch 4 x 1024^2, random-init weights, on one B200:
stock torchvision CUDA path vs the patched model (everything after the heads in libmisob200)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso.object_detection.models import get_object_detection_model  # noqa: E402
from miso_b200.patch import forward_uint8, patch_model, unpatch_model  # noqa: E402

DEV = "cuda:0"
torch.manual_seed(0)
model = get_object_detection_model(3).eval().to(DEV)
with torch.no_grad():
    model.roi_heads.box_predictor.cls_score.weight.mul_(8.0)
g = torch.Generator().manual_seed(0)
u8 = [torch.randint(0, 256, (1024, 1024, 3), dtype=torch.uint8, generator=g).to(DEV) for _ in range(4)]
fl = [a.permute(2, 0, 1).to(torch.float32) / 255 for a in u8]


def timed(fn, reps=10):
    with torch.inference_mode():
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, out


t_stock, o = timed(lambda: model(fl))
print(f"stock torchvision (CUDA ops):            {t_stock:8.2f} ms / batch of 4   dets {[len(x['boxes']) for x in o]}")
patch_model(model)
t_p, o = timed(lambda: model(fl))
print(f"patched (post-head path in libmisob200): {t_p:8.2f} ms / batch of 4   dets {[len(x['boxes']) for x in o]}")
t_u, o = timed(lambda: forward_uint8(model, u8))
print(f"patched + fused uint8 input transform:   {t_u:8.2f} ms / batch of 4")
model = model.to(memory_format=torch.channels_last)
t_c, o = timed(lambda: forward_uint8(model, u8))
print(f"same, channels_last backbone:            {t_c:8.2f} ms / batch of 4")
with torch.inference_mode():
    def backbone_only():
        from miso_b200 import ops
        tr = model.transform
        b, _ = ops.transform_images(u8, tr.min_size[-1], tr.max_size, tr.image_mean, tr.image_std, tr.size_divisible)
        return model.backbone(b)
t_b, _ = timed(backbone_only)
print(f"backbone + FPN alone (cuDNN, fp32):      {t_b:8.2f} ms / batch of 4")
