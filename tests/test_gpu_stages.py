"""GPU parity tests of the fused stages (mb_rpn_proposals, mb_det_postprocess, mb_crop_*) against
the golden vectors (torchvision CPU / miso reference) and the CPU oracle, on identical inputs.
Top-k indices, NMS keep decisions, labels and crop bytes: bit-exact. Boxes: <= 1e-5 of the box
scale. Scores (sigmoid / softmax use CUDA expf): <= 1e-6 absolute."""
import os

import numpy as np
import pytest
import torch

from oracle import detection as D
from oracle import miso_path as M
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.fixture(scope="module")
def det():
    from miso_b200 import detection
    detection._lib.load()
    return detection


def _rpn_inputs():
    obj, dlt, grids, image_sizes, padded = cases.rpn_case()
    objectness, deltas = D.concat_rpn_head_outputs(obj, dlt)
    napl = [o.shape[1] * o.shape[2] * o.shape[3] for o in obj]
    return obj, dlt, objectness, napl, image_sizes, padded


def test_rpn_matches_torchvision_golden(det, golden_dir):
    g = load(golden_dir, "rpn")
    obj, dlt, objectness, napl, image_sizes, padded = _rpn_inputs()
    cfg = det.RpnConfig(cases.RPN_SIZES, cases.RPN_RATIOS, pre_nms_top_n=300, post_nms_top_n=200, nms_thresh=0.7)
    out = det.rpn_proposals([cu(o) for o in obj], [cu(d) for d in dlt], image_sizes, padded, cfg, return_topk_idx=True)
    assert np.array_equal(out.topk_idx.cpu().numpy(), g["top_idx"])      # tie-free logits: exact indices
    boxes, scores = out.as_lists()
    for i in range(len(boxes)):
        b, s = boxes[i].cpu().numpy(), scores[i].cpu().numpy()
        assert b.shape == g[f"boxes{i}"].shape, (b.shape, g[f"boxes{i}"].shape)
        assert cases.box_rel_err(b, g[f"boxes{i}"]) < 1e-5
        assert np.max(np.abs(s - g[f"scores{i}"])) < 1e-6
    # padding rows are zero, counts are int32 on the device
    cnt = out.counts.cpu().numpy()
    assert out.counts.dtype == torch.int32
    for i, c in enumerate(cnt):
        assert torch.count_nonzero(out.proposals[i, c:]) == 0


@pytest.mark.parametrize("pre,post,rule", [(1000, 1000, 4000), (1000, 300, -1), (120, 50, 100000), (4000, 2000, 4000)])
def test_rpn_matches_oracle_both_strategies(det, pre, post, rule):
    """Same stage against the oracle composition with the strategy rule forced on both sides."""
    obj, dlt, objectness, napl, image_sizes, padded = _rpn_inputs()
    cfg = det.RpnConfig(cases.RPN_SIZES, cases.RPN_RATIOS, pre_nms_top_n=pre, post_nms_top_n=post, nms_thresh=0.7,
                        trick_numel=rule)
    out = det.rpn_proposals([cu(o) for o in obj], [cu(d) for d in dlt], image_sizes, padded, cfg, return_topk_idx=True)
    n = objectness.shape[0]
    for i in range(n):
        assert np.array_equal(out.topk_idx[i].cpu().numpy(), D.rpn_top_n_idx(objectness[i], napl, pre))
    grids = [o.shape[-2:] for o in obj]
    anchors = D.grid_anchors(padded, grids, cases.RPN_SIZES, cases.RPN_RATIOS)
    _, deltas = D.concat_rpn_head_outputs(obj, dlt)
    decoded = np.stack([D.decode_boxes(deltas[i], anchors)[:, 0] for i in range(n)])
    device_rule = {4000: "cpu", 100000: "cuda"}.get(rule)
    if device_rule is None:   # forced vanilla
        orig = D.batched_nms
        D.batched_nms = lambda b, s, idx, thr, device_rule="cpu": D.batched_nms_vanilla(b, s, idx, thr)
    try:
        rb, rs = D.filter_proposals(decoded, objectness, image_sizes, napl, pre_nms_top_n=pre, post_nms_top_n=post,
                                    nms_thresh=0.7, device_rule=device_rule or "cpu")
    finally:
        if device_rule is None:
            D.batched_nms = orig
    boxes, scores = out.as_lists()
    for i in range(n):
        assert boxes[i].shape == rb[i].shape
        assert cases.box_rel_err(boxes[i].cpu().numpy(), rb[i]) < 1e-5
        assert np.max(np.abs(scores[i].cpu().numpy() - rs[i])) < 1e-6


def test_rpn_ties_and_degenerate_logits(det):
    """All-equal logits (every key lands in one histogram bin -> radix-select fallback) and NaNs:
    the winners must be (value desc, index asc) with NaN treated as the largest value."""
    obj, dlt, objectness, napl, image_sizes, padded = _rpn_inputs()
    obj = [np.zeros_like(o) for o in obj]
    obj[0][0, 1, 3, 5] = np.nan
    obj[0][1, 2, 7, 2] = 1.0
    obj[1][0].reshape(-1)[::3] = -0.0
    objectness, _ = D.concat_rpn_head_outputs(obj, dlt)
    cfg = det.RpnConfig(cases.RPN_SIZES, cases.RPN_RATIOS, pre_nms_top_n=700, post_nms_top_n=100)
    out = det.rpn_proposals([cu(o) for o in obj], [cu(d) for d in dlt], image_sizes, padded, cfg, return_topk_idx=True)
    for i in range(objectness.shape[0]):
        assert np.array_equal(out.topk_idx[i].cpu().numpy(), D.rpn_top_n_idx(objectness[i], napl, 700))


def test_det_postprocess_matches_golden(det, golden_dir):
    g = load(golden_dir, "det")
    logits, reg, proposals, shapes = cases.det_case()
    n = len(proposals)
    r = max(len(p) for p in proposals)
    padded = np.zeros((n, r, 4), np.float32)
    for i, p in enumerate(proposals):
        padded[i, :len(p)] = p
    counts = torch.tensor([len(p) for p in proposals], dtype=torch.int32, device=DEV)
    cfg = det.DetConfig(detections_per_img=100)
    out = det.postprocess_detections(cu(logits), cu(reg), cu(padded), counts, shapes, cfg,
                                     original_image_sizes=[(287, 369)] * n, packed=True)
    boxes, scores, labels = out.as_lists()
    for i in range(n):
        assert np.array_equal(labels[i].cpu().numpy(), g[f"labels{i}"])
        assert labels[i].dtype == torch.int64
        assert cases.box_rel_err(out.boxes_net[i, :len(labels[i])].cpu().numpy(), g[f"boxes{i}"]) < 1e-5
        assert cases.box_rel_err(boxes[i].cpu().numpy(), g[f"resized{i}"]) < 1e-5
        assert np.max(np.abs(scores[i].cpu().numpy() - g[f"scores{i}"])) < 1e-6


# (91, 1000, 300): the stock COCO head at miso's box_detections_per_img=300 (ref:miso/object_detection/models.py:9) — 90 x 300
# candidate slots per image exceed the shared-memory sort; the kernel counts what was kept and sorts in global memory if needed
@pytest.mark.parametrize("ncls,props,dpi,rule", [(3, 1000, 300, 4000), (3, 1000, 300, -1), (11, 400, 100, 4000), (2, 50, 100, 100000),
                                                  (91, 1000, 300, 4000)])
def test_det_postprocess_matches_oracle(det, ncls, props, dpi, rule):
    logits, reg, proposals, shapes = cases.det_case(n_img=3, props=props, num_classes=ncls, seed=5 + ncls)
    n = len(proposals)
    r = props
    padded = np.zeros((n, r, 4), np.float32)
    rows_l = np.zeros((n * r, ncls), np.float32); rows_r = np.zeros((n * r, 4 * ncls), np.float32)
    off = 0
    for i, p in enumerate(proposals):
        padded[i, :len(p)] = p
        rows_l[i * r:i * r + len(p)] = logits[off:off + len(p)]
        rows_r[i * r:i * r + len(p)] = reg[off:off + len(p)]
        off += len(p)
    counts = torch.tensor([len(p) for p in proposals], dtype=torch.int32, device=DEV)
    cfg = det.DetConfig(detections_per_img=dpi, trick_numel=rule)
    out = det.postprocess_detections(cu(rows_l), cu(rows_r), cu(padded), counts, shapes, cfg, packed=False)  # strided rows
    device_rule = {4000: "cpu", 100000: "cuda"}.get(rule)
    if device_rule is None:
        orig = D.batched_nms
        D.batched_nms = lambda b, s, idx, thr, device_rule="cpu": D.batched_nms_vanilla(b, s, idx, thr)
    try:
        ref = D.postprocess_detections(logits, reg, proposals, shapes, detections_per_img=dpi, device_rule=device_rule or "cpu")
    finally:
        if device_rule is None:
            D.batched_nms = orig
    boxes, scores, labels = out.as_lists()
    for i in range(n):
        assert np.array_equal(labels[i].cpu().numpy(), ref[i][2])
        assert cases.box_rel_err(boxes[i].cpu().numpy(), ref[i][0]) < 1e-5      # no original size: boxes == boxes_net
        assert np.max(np.abs(scores[i].cpu().numpy() - ref[i][1])) < 1e-6


def test_det_postprocess_many_classes_global_sort_path(det):
    """91 classes, score threshold 0: every (proposal, class) pair is a candidate, each class keeps up to 300 boxes, so
    far more than 16384 keys per image reach the final ordering — the global-memory sort path. Vanilla strategy on both sides."""
    ncls, props, dpi = 91, 600, 300
    logits, reg, proposals, shapes = cases.det_case(n_img=2, props=props, num_classes=ncls, image_hw=(512, 640), seed=77)
    n, r = len(proposals), props
    padded = np.zeros((n, r, 4), np.float32)
    for i, p in enumerate(proposals):
        padded[i, :len(p)] = p
    counts = torch.tensor([len(p) for p in proposals], dtype=torch.int32, device=DEV)
    cfg = det.DetConfig(detections_per_img=dpi, trick_numel=-1, score_thresh=0.0)
    out = det.postprocess_detections(cu(logits), cu(reg), cu(padded), counts, shapes, cfg, packed=True)
    orig = D.batched_nms
    D.batched_nms = lambda b, s, idx, thr, device_rule="cpu": D.batched_nms_vanilla(b, s, idx, thr)
    try:
        ref = D.postprocess_detections(logits, reg, proposals, shapes, score_thresh=0.0, detections_per_img=dpi)
    finally:
        D.batched_nms = orig
    boxes, scores, labels = out.as_lists()
    for i in range(n):
        assert len(ref[i][2]) == dpi
        assert np.array_equal(labels[i].cpu().numpy(), ref[i][2])
        assert cases.box_rel_err(boxes[i].cpu().numpy(), ref[i][0]) < 1e-5
        assert np.max(np.abs(scores[i].cpu().numpy() - ref[i][1])) < 1e-6


@pytest.mark.parametrize("tag,ch", [("rgb", 3), ("gray", 1)])
def test_filter_and_crop_bit_exact(det, golden_dir, tag, ch):
    g = load(golden_dir, "crop_" + tag)
    img, boxes, scores, labels = cases.crop_case(channels=ch)
    cap = 64
    db = np.zeros((2, cap, 4), np.float32); ds = np.zeros((2, cap), np.float32)
    db[1, :len(boxes)] = boxes; ds[1, :len(boxes)] = scores
    db[0, :3] = boxes[:3]; ds[0, :3] = 0.99          # image 0: three detections, but count says 2
    counts = torch.tensor([2, len(boxes)], dtype=torch.int32, device=DEV)
    img0 = np.ascontiguousarray(img[:150, :200])
    out = det.filter_and_crop([cu(img0), cu(img)], cu(db), cu(ds), counts, 0.5)
    crops = out.to_host(ch)
    first = [c for c in crops if c[0] == 0]
    second = [c for c in crops if c[0] == 1]
    assert len(first) == 2
    rb, rs, rl, rxywh, rci, rcrops = M.filter_and_crop(img0, boxes[:2], np.array([0.99, 0.99], np.float32), labels[:2], 0.5)
    for (n, i, xywh, arr), ref in zip(first, rcrops):
        assert np.array_equal(arr, ref)
    assert len(second) == len(g["coords"])
    assert np.array_equal(np.stack([c[2] for c in second]), g["bounds"])
    assert np.array_equal(np.array([c[3].shape[:2] for c in second]), g["sizes"])
    assert np.array_equal(np.concatenate([c[3].reshape(-1) for c in second]), g["pixels"])
    # fixed-capacity, sync-free variant: too small a buffer must flag, not corrupt
    small = det.filter_and_crop([cu(img0), cu(img)], cu(db), cu(ds), counts, 0.5, capacity_bytes=16)
    assert small.totals.tolist()[2] == 1
    big = det.filter_and_crop([cu(img0), cu(img)], cu(db), cu(ds), counts, 0.5, capacity_bytes=1 << 22)
    assert np.array_equal(np.concatenate([c[3].reshape(-1) for c in big.to_host(ch) if c[0] == 1]), g["pixels"])


@pytest.mark.parametrize("hw,ch", [((1024, 1024), 3), ((1000, 1021), 1), ((777, 1003), 4), ((640, 999), 3)])
def test_crop_full_size_roundtrip(det, hw, ch):
    """Full-size images (odd widths: every source / destination alignment occurs; 1, 3 and 4 channels: rows from a
    few bytes to 1.6 KB, all three lanes-per-row variants of the copy), 300 detections incl. boxes clipped at the
    image edges: every crop equals the oracle's slice."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (hw[0], hw[1], ch), dtype=np.uint8)
    boxes = cases.stress_rois(rng, 300, hw, side=(2.0, 400.0))
    boxes[:8, 0] = 0.0; boxes[8:16, 2] = hw[1]; boxes[16:24, 3] = hw[0]; boxes[24:28] = [0, 0, hw[1], 3]
    scores = rng.uniform(0.3, 1.0, 300).astype(np.float32)
    counts = torch.tensor([300], dtype=torch.int32, device=DEV)
    out = det.filter_and_crop([cu(img)], cu(boxes[None]), cu(scores[None]), counts, 0.5)
    crops = out.to_host(ch)
    _, _, _, _, ci, ref = M.filter_and_crop(img if ch > 1 else img[:, :, 0], boxes, scores, np.ones(300, np.int64), 0.5)
    assert len(crops) == len(ref)
    for (n, i, xywh, arr), r in zip(crops, ref):
        assert arr.shape == r.shape and np.array_equal(arr, r)


def test_dense_seam_nms_and_pack_match_the_cpu_restatement(det):
    """mb_mosaic_pack == tests/mosaic_ref.pack_block bit for bit; mosaic.SeamNms (dense per-label mb_nms, padding rows
    ignored on the device, no host sync before finish) == the oracle's per-class NMS over the live rows."""
    from miso_b200 import mosaic
    from tests import mosaic_ref as R
    from tests.test_mosaic_cpu import synth_tiles
    b, s, l, c, o = synth_tiles(num_tiles=6, dpi=60, seed=3)
    block = mosaic.pack_block(cu(b), cu(s), cu(l), cu(c).to(torch.int32), cu(o), 0.5, 6 * 60 + 17)
    host_block = R.pack_block(b, s, l, c, o, 0.5, 6 * 60 + 17)
    assert np.array_equal(block.cpu().numpy(), host_block)
    seam = mosaic.SeamNms(block.shape[0], 3, DEV)
    seam.launch(block, 0.5)
    gb, gs, gl, rows = seam.finish()
    live = np.nonzero(host_block[:, 5] >= 0)[0]
    keep = D.batched_nms_vanilla(host_block[live, :4], host_block[live, 4], host_block[live, 5].astype(np.int64), 0.5)
    assert np.array_equal(rows.cpu().numpy(), live[keep])            # batched_nms order: score desc, row asc
    assert np.array_equal(gb.cpu().numpy(), host_block[live[keep], :4])
    assert np.array_equal(np.sort(rows.cpu().numpy()), R.seam_keep_rows(host_block, 0.5))
    seam.launch(block, 0.5)                                          # idempotent and reusable
    assert torch.equal(seam.finish()[3], rows)


@pytest.mark.parametrize("n,cap", [(24, 200), (60, 520)])
def test_filter_and_crop_many_slots_multi_round(det, n, cap):
    """24 images x 200 slots = 4800 detection slots: more than one 4096-entry round of the one-CTA plan kernel;
    60 x 520 = 31 200 slots: the chunked plan kernel (8 CTAs chained by their aggregates). Every image's crops must
    equal the oracle's, in order, with ragged counts and scores on both sides of the threshold."""
    rng = np.random.default_rng(17)
    hw = (96, 128)
    imgs = [rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8) for _ in range(n)]
    db = np.zeros((n, cap, 4), np.float32); ds = np.zeros((n, cap), np.float32)
    cnt = rng.integers(0, cap + 1, n).astype(np.int32)
    for i in range(n):
        db[i] = cases.stress_rois(rng, cap, hw, side=(2.0, 90.0))
        ds[i] = rng.uniform(0.0, 1.0, cap).astype(np.float32)
    out = det.filter_and_crop([cu(a) for a in imgs], cu(db), cu(ds), cu(cnt), 0.5)
    crops = out.to_host(3)
    k = 0
    for i in range(n):
        _, _, _, rxywh, _, rcrops = M.filter_and_crop(imgs[i], db[i, :cnt[i]], ds[i, :cnt[i]], np.ones(cnt[i], np.int64), 0.5)
        for xy, ref in zip(rxywh, rcrops):
            img_idx, _, gxy, arr = crops[k]
            assert img_idx == i and np.array_equal(gxy, xy) and arr.shape == ref.shape and np.array_equal(arr, ref)
            k += 1
    assert k == len(crops) and k > 1000
