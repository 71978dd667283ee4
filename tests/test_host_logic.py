"""Host-side logic that runs without a GPU: LevelMapper thresholds, scale inference, base anchors."""
import numpy as np
import pytest
import torch

from miso_b200 import ops


def test_level_thresholds_match_survey_constants():
    thr = ops.level_thresholds(2, 5)
    bits = [int(np.float32(t).view(np.uint32)) for t in thr]
    # exact equivalents of the fp32 sqrt/log2/floor mapper for k_min=2, k_max=5
    assert bits == [0x4643FFEC, 0x4743FFEE, 0x4843FFEB]


@pytest.mark.parametrize("kmin,kmax", [(2, 5), (2, 4), (3, 6), (0, 3)])
def test_level_thresholds_reproduce_torchvision_mapper(kmin, kmax):
    from torchvision.ops.poolers import LevelMapper
    rng = np.random.default_rng(3)
    bx = rng.uniform(0, 600, (200000, 2)); wh = np.exp(rng.uniform(-1, 8, (200000, 2)))
    boxes = torch.from_numpy(np.concatenate([bx, bx + wh], 1).astype(np.float32))
    thr = ops.level_thresholds(kmin, kmax)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    mine = sum((area >= t).to(torch.int64) for t in thr)
    assert torch.equal(mine, LevelMapper(kmin, kmax)([boxes]))
    # +-2000 ulp around every threshold
    for t in thr:
        b = np.float32(t).view(np.uint32).astype(np.int64) + np.arange(-2000, 2001)
        a = b.astype(np.uint32).view(np.float32)
        bb = torch.from_numpy(np.stack([np.zeros_like(a), np.zeros_like(a), a, np.ones_like(a)], 1))
        area = (bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])
        mine = sum((area >= tt).to(torch.int64) for tt in thr)
        assert torch.equal(mine, LevelMapper(kmin, kmax)([bb]))


def test_infer_scale_and_base_anchors():
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    assert ops.infer_scale((1, 256, 200, 200), (800, 800)) == 0.25
    assert ops.infer_scale((1, 256, 25, 25), (800, 800)) == 0.03125
    assert ops.infer_scale((1, 256, 13, 13), (800, 800)) == 0.015625
    ag = AnchorGenerator(((32,), (64,)), ((0.5, 1.0, 2.0), (0.5, 1.0, 2.0)))
    for i, s in enumerate((32, 64)):
        assert torch.equal(ops.base_anchors((s,), (0.5, 1.0, 2.0)), ag.cell_anchors[i])


def test_resized_image_size_matches_torchvision_transform():
    """ops.resized_image_size (host arithmetic of the fused input transform) against the sizes torchvision's
    GeneralizedRCNNTransform produces on the CPU, for shapes on both sides of the min/max rule."""
    import torch
    from torchvision.models.detection.transform import GeneralizedRCNNTransform
    from miso_b200 import ops
    rng = np.random.default_rng(0)
    for mn, mx in ((800, 1333), (96, 150), (224, 224)):
        tr = GeneralizedRCNNTransform(mn, mx, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]).eval()
        for _ in range(12):
            h, w = int(rng.integers(17, 400)), int(rng.integers(17, 400))
            il, _ = tr([torch.zeros(3, h, w)])
            assert tuple(il.image_sizes[0]) == ops.resized_image_size(h, w, mn, mx), (h, w, mn, mx)


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first) needs no GPU and no CUDA library: one JSON
    line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "detections_per_sec_posthead_path" and line["value"] > 0
    for key in ("unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0
