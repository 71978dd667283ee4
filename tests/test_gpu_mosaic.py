"""GPU parity tests of the tiled-mosaic plan (config 5 at a size the oracle finishes in seconds): per-tile
detections against the oracle pipeline, the seam NMS and the crops of the plan against the CPU composition of
SURVEY.md §8(e) evaluated on the same gathered rows, world-size independence (2 and 3 emulated ranks on one GPU,
and 2 real NCCL ranks when two GPUs are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import pipeline as ref_pipeline
from tests import mosaic_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_mosaic():
    from miso_b200 import workload
    return workload.mosaic(2816, 3712, resized=256, channels=16, post_nms_top_n=200, detections_per_img=60)


def make_plan(w, rank, world, batch=4):
    from miso_b200 import mosaic, pipeline
    grid = mosaic.tile_grid(w.height, w.width, w.tile, w.overlap)
    mine = list(mosaic.rank_tiles(len(grid), world, rank))
    sizes = [min(batch, len(mine) - i) for i in range(0, len(mine), batch)]

    def make_hp(n):
        return pipeline.HotPath(w.shapes(n), w.base.rpn, w.base.det, threshold=w.base.threshold, crop_capacity_bytes=1 << 20, device=DEV)

    plan = mosaic.MosaicPlan(grid, w.tile, (w.height, w.width), make_hp, sizes, rank=rank, world=world,
                             threshold=w.base.threshold, iou_threshold=w.base.det.nms_thresh, crop_capacity_bytes=256 << 20, device=DEV)
    batches, i0 = [], 0
    for n in sizes:
        batches.append(w.batch_inputs(mine[i0:i0 + n], DEV))
        i0 += n
    y0, y1 = mosaic.rank_band(grid, w.tile, w.height, world, rank)
    plan.bind_band(w.band(y0, y1, DEV), y0)
    return plan, batches, grid, mine


def crops_list(res):
    rects, offs, pix = res["rects"].cpu().numpy(), res["offsets"].cpu().numpy(), res["pixels"].cpu().numpy()
    return [pix[offs[j]:offs[j + 1]].reshape(rects[j][3], rects[j][2], 3) for j in range(res["count"])]


def test_mosaic_plan_world1_matches_the_cpu_composition():
    w = small_mosaic()
    plan, batches, grid, mine = make_plan(w, 0, 1)
    assert len(grid) == 12 and plan.batch_sizes == [4, 4, 4]
    plan.run(batches)
    res = plan.results()
    block = plan.gathered.cpu().numpy()
    dpi = plan.dpi
    # (1) per-tile detections against the oracle pipeline on the same head outputs
    s = w.base.shapes
    for t in (0, 5, 11):
        h = {k: [x.cpu().numpy() for x in v] for k, v in w.tile_inputs(t, DEV).items()}
        tile_px = np.zeros((w.tile, w.tile, 3), np.uint8)
        ref = ref_pipeline.run(h["objectness"], h["deltas"], h["features"], h["class_logits"][0], h["box_regression"][0], [tile_px],
                               padded_image_size=s.padded_image_size, image_sizes=s.image_sizes, original_image_sizes=s.original_image_sizes,
                               sizes=w.base.rpn.sizes, aspect_ratios=w.base.rpn.aspect_ratios, pre_nms_top_n=w.base.rpn.pre_nms_top_n,
                               post_nms_top_n=w.base.rpn.post_nms_top_n, detections_per_img=dpi, threshold=w.base.threshold)[0]
        rows = block[t * dpi:(t + 1) * dpi]
        nd = len(ref["boxes"])
        live = ref["scores"] > np.float32(w.base.threshold)
        assert np.array_equal(rows[:nd, 5] >= 0, live) and not (rows[nd:, 5] >= 0).any()
        assert np.array_equal(rows[:nd][live, 5].astype(np.int64), ref["labels"][live])
        off = np.array([grid[t][1], grid[t][0], grid[t][1], grid[t][0]], np.float32)
        want = (ref["boxes"] + off).astype(np.float32)
        assert np.max(np.abs(rows[:nd, :4] - want)) <= 1e-5 * 1024 + 1e-3        # fp32 ulp at mosaic coordinates ~ 2e-4
        assert np.max(np.abs(rows[:nd, 4] - ref["scores"])) <= 1e-6
    # (2) seam NMS and (3) crops: the CPU composition on the very same gathered rows, bit for bit
    keep = R.seam_keep_rows(block, w.base.det.nms_thresh)
    state = plan.seam.state.cpu().numpy()
    assert np.array_equal(np.nonzero(state == 1)[0], keep)
    assert 0 < len(keep) <= int((block[:, 5] >= 0).sum())      # independent random tiles: (almost) nothing to suppress; duplicates are tests/test_gpu_seam.py's job
    mosaic_px = w.band(0, w.height, DEV).cpu().numpy()
    xywh, ci, crops = R.crops_of_rows(mosaic_px, block, keep)
    assert res["count"] == len(keep) and np.array_equal(res["src"].cpu().numpy(), keep)
    assert np.array_equal(res["xywh"].cpu().numpy(), xywh)
    got = crops_list(res)
    assert sum(c.size for c in crops) == res["bytes"] > 0
    for a, b in zip(got, crops):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("world", [2, 3])
def test_mosaic_plan_emulated_ranks_equal_world1(world):
    """Ranks emulated on one GPU: every rank's plan runs its own tiles, the rank blocks are concatenated in rank
    order (what the all-gather produces), every rank runs the seam NMS + its crops. The concatenation of the
    per-rank outputs must equal the world-size-1 output."""
    from miso_b200 import mosaic
    w = small_mosaic()
    p1, b1, grid, _ = make_plan(w, 0, 1)
    p1.run(b1)
    r1 = p1.results()
    g1 = p1.gathered.cpu().numpy()
    dpi = p1.dpi
    plans = [make_plan(w, r, world) for r in range(world)]
    for p, b, _, _ in plans:
        p.run_tiles(b)
    gathered = torch.cat([p.block for p, _, _, _ in plans], dim=0).contiguous()
    to_w = np.array([mosaic.gathered_row(r // dpi, r % dpi, len(grid), world, dpi) for r in range(len(grid) * dpi)])
    assert np.array_equal(gathered.cpu().numpy()[to_w], g1)
    outs = []
    for p, _, _, _ in plans:
        p.run_tail(gathered)
        outs.append(p.results())
    keep1 = np.nonzero(p1.seam.state.cpu().numpy() == 1)[0]
    for p, _, _, _ in plans:
        assert np.array_equal(np.nonzero(p.seam.state.cpu().numpy() == 1)[0], np.sort(to_w[keep1]))
    assert sum(o["count"] for o in outs) == r1["count"]
    assert torch.equal(torch.cat([o["rects"] for o in outs]), r1["rects"])
    assert torch.equal(torch.cat([o["xywh"] for o in outs]), r1["xywh"])
    assert torch.equal(torch.cat([o["pixels"] for o in outs]), r1["pixels"])
    src = np.concatenate([o["src"].cpu().numpy() + r * plans[0][0].block_rows for r, o in enumerate(outs)])
    assert np.array_equal(src, to_w[r1["src"].cpu().numpy()])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_mosaic_two_nccl_ranks_equal_world1():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "mosaic_nccl_worker.py")],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "world 2 == world 1: ok" in out.stdout
