"""BASELINE.json configurations 3 and 4 at FULL size (the oracle still finishes in seconds for
these), plus size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import detection as D
from oracle import native
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_config4_200k_boxes_80_classes_bit_exact():
    """NMS / top-k stress: 200 000 boxes, 80 classes, distinct scores."""
    from miso_b200 import ops
    rng = np.random.default_rng(0)
    n = 200_000
    c = rng.uniform(0, 4096, (n, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(256), (n, 2)))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.permutation(np.linspace(0, 1, n)).astype(np.float32)
    idx = rng.integers(0, 80, n).astype(np.int64)
    tb, ts, ti = cu(b), cu(s), cu(idx)
    for thr in (0.3, 0.5, 0.7):
        keep = ops.batched_nms(tb, ts, ti, thr)                 # auto: > 100k elements -> per-class strategy
        ref = D.batched_nms_vanilla(b, s, idx, thr)
        assert np.array_equal(keep.cpu().numpy(), ref), thr
        # properties: sorted by score, idempotent
        ks = ts[keep]
        assert torch.all(ks[:-1] >= ks[1:])
        again = ops.batched_nms(tb[keep], ts[keep], ti[keep], thr, strategy="vanilla")
        assert again.numel() == keep.numel()


def test_config4_tie_heavy_single_class():
    from miso_b200 import ops
    rng = np.random.default_rng(1)
    n = 50_000
    c = rng.uniform(0, 4096, (n, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(256), (n, 2)))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = (np.floor(rng.uniform(0, 1, n) * 256) / 256).astype(np.float32)      # ~195 boxes per score value
    keep = ops.nms(cu(b), cu(s), 0.5).cpu().numpy()
    assert np.array_equal(keep, native.nms(b, s, 0.5))


def test_config3_mask_roi_align_batch8_full_size():
    """Mask R-CNN mask head pooling: batch 8, 100 detections/img, 14x14, 256-channel 800^2 pyramid."""
    from miso_b200 import ops
    rng = np.random.default_rng(3)
    n, c = 8, 256
    feats = [torch.randn(n, c, 800 // st, 800 // st, device=DEV) for st in (4, 8, 16, 32)]
    boxes = [cu(cases.stress_rois(rng, 100, (800, 800))) for _ in range(n)]
    shapes = [(800, 800)] * n
    pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 14, 2)
    x = {str(i): f for i, f in enumerate(feats)}
    out, levels = pool(x, boxes, shapes, return_levels=True)
    assert out.shape == (800, 256, 14, 14)
    assert set(levels.cpu().tolist()) == {0, 1, 2, 3}
    assert torch.equal(pool({k: -v for k, v in x.items()}, boxes, shapes), -out)      # odd symmetry is exact in fp32
    fast = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 14, 2, exact=False)(x, boxes, shapes)
    assert torch.allclose(fast, out, rtol=1e-5, atol=5e-5)
    pick = np.sort(rng.choice(800, 48, replace=False))
    sub = [boxes[i].cpu().numpy()[pick[pick // 100 == i] % 100] for i in range(n)]
    ref = D.multiscale_roi_align([f.cpu().numpy() for f in feats], sub, shapes, 14, 2)
    assert np.array_equal(out[torch.from_numpy(pick).to(DEV)].cpu().numpy(), ref)
