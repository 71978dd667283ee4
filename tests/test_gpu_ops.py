"""GPU parity tests of the torchvision-signature operators (through the C ABI) against the CPU
oracle and the committed golden vectors. Bit-exact for indices; RoIAlign exact mode bit-exact,
fast mode within 1e-5; decoded boxes within 1e-5 of each box's scale."""
import os

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


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.fixture(scope="module")
def ops():
    from miso_b200 import ops as o
    o._lib.load()
    return o


def test_nms_cases_bit_exact(ops, golden_dir):
    g = load(golden_dir, "nms")
    for name, b, s, thr in cases.nms_cases():
        keep = ops.nms(cu(b), cu(s), thr).cpu().numpy()
        assert keep.dtype == np.int64
        assert np.array_equal(keep, g[name + "/keep"]), name
        assert np.array_equal(keep, native.nms(b, s, thr)), name


def test_nms_empty_and_shapes(ops):
    e = ops.nms(torch.zeros((0, 4), device=DEV), torch.zeros((0,), device=DEV), 0.5)
    assert e.shape == (0,) and e.dtype == torch.int64 and e.device.type == "cuda"
    with pytest.raises(Exception):
        ops.nms(torch.zeros((3, 5), device=DEV), torch.zeros((3,), device=DEV), 0.5)


def test_batched_nms_both_strategies_bit_exact(ops, golden_dir):
    g = load(golden_dir, "batched_nms")
    for name, b, s, idx, thr in cases.batched_nms_cases():
        v = ops.batched_nms(cu(b), cu(s), cu(idx), thr, strategy="vanilla").cpu().numpy()
        t = ops.batched_nms(cu(b), cu(s), cu(idx), thr, strategy="trick").cpu().numpy()
        assert np.array_equal(v, g[name + "/vanilla"]), name
        assert np.array_equal(t, g[name + "/trick"]), name
    # auto follows torchvision's CUDA rule (trick below 100k elements)
    name, b, s, idx, thr = cases.batched_nms_cases()[1]
    assert np.array_equal(ops.batched_nms(cu(b), cu(s), cu(idx), thr).cpu().numpy(), g[name + "/trick"])
    e = ops.batched_nms(torch.zeros((0, 4), device=DEV), torch.zeros((0,), device=DEV),
                        torch.zeros((0,), dtype=torch.int64, device=DEV), 0.5)
    assert e.numel() == 0


def test_batched_nms_arbitrary_category_values(ops):
    rng = np.random.default_rng(8)
    b, s = cases.random_boxes(rng, 700, extent=150.0), cases.distinct_scores(rng, 700)
    idx = rng.choice(np.array([-7, 3, 10**9, 123456789012], dtype=np.int64), 700)
    v = ops.batched_nms(cu(b), cu(s), cu(idx), 0.5, strategy="vanilla").cpu().numpy()
    assert np.array_equal(v, D.batched_nms_vanilla(b, s, idx, 0.5))


def test_nms_medium_stress(ops):
    """20k boxes, single class, tie-heavy scores; 30k boxes over 80 classes (scaled-down cfg 4)."""
    rng = np.random.default_rng(2024)
    n = 20000
    c = rng.uniform(0, 2048, (n, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(256), (n, 2)))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = (np.floor(rng.uniform(0, 1, n) * 4096) / 4096).astype(np.float32)
    assert np.array_equal(ops.nms(cu(b), cu(s), 0.5).cpu().numpy(), native.nms(b, s, 0.5))
    # 4097..12k boxes in one segment: the wide tiled sweep with two cp.async tile buffers (20k above: one buffer)
    for n2, extent in ((9001, 1024), (4100, 256)):
        c = rng.uniform(0, extent, (n2, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(128), (n2, 2)))
        b2 = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        s2 = cases.distinct_scores(rng, n2)
        for thr in (0.3, 0.6):
            assert np.array_equal(ops.nms(cu(b2), cu(s2), thr).cpu().numpy(), native.nms(b2, s2, thr)), (n2, thr)
    n = 30000
    c = rng.uniform(0, 2048, (n, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(256), (n, 2)))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = cases.distinct_scores(rng, n)
    idx = rng.integers(0, 80, n).astype(np.int64)
    for thr in (0.3, 0.7):
        v = ops.batched_nms(cu(b), cu(s), cu(idx), thr, strategy="vanilla").cpu().numpy()
        assert np.array_equal(v, D.batched_nms_vanilla(b, s, idx, thr))


def test_roi_align_exact_matches_golden_and_oracle(ops, golden_dir):
    g = load(golden_dir, "roi_align")
    for name, x, rois, scale, P, sr, aligned in cases.roi_align_cases():
        out = ops.roi_align(cu(x), cu(rois), P, scale, sr, aligned, exact=True).cpu().numpy()
        assert out.shape == (rois.shape[0], x.shape[1], P, P)
        assert np.array_equal(out[:, cases.ROI_GOLDEN_CHANNELS], g[name]), name
        assert np.array_equal(out, native.roi_align(x, rois, scale, P, P, sr, aligned)), name


def test_roi_align_fast_mode_within_tolerance(ops):
    for name, x, rois, scale, P, sr, aligned in cases.roi_align_cases():
        if sr != 2:
            continue
        out = ops.roi_align(cu(x), cu(rois), P, scale, sr, aligned, exact=False).cpu().numpy()
        ref = native.roi_align(x, rois, scale, P, P, sr, aligned)
        tol = 1e-5 * np.abs(ref) + 1e-5 * np.abs(x).max()   # rtol 1e-5, atol 1e-5 of the feature scale
        assert np.all(np.abs(out - ref) <= tol), name


def test_roi_align_list_input_and_errors(ops):
    x = torch.randn(2, 8, 20, 20, device=DEV)
    b0 = torch.tensor([[1.0, 2.0, 10.0, 12.0]], device=DEV)
    b1 = torch.tensor([[0.0, 0.0, 5.0, 5.0], [3.0, 3.0, 19.0, 18.0]], device=DEV)
    out = ops.roi_align(x, [b0, b1], (7, 5), 0.5, 2)
    rois = np.array([[0, 1, 2, 10, 12], [1, 0, 0, 5, 5], [1, 3, 3, 19, 18]], np.float32)
    assert np.array_equal(out.cpu().numpy(), native.roi_align(x.cpu().numpy(), rois, 0.5, 7, 5, 2, False))
    with pytest.raises(Exception):
        ops.roi_align(x, torch.zeros((3, 4), device=DEV), 7)
    assert ops.roi_align(x, torch.zeros((0, 5), device=DEV), 7).shape == (0, 8, 7, 7)


@pytest.mark.parametrize("tag,P", [("box7", 7), ("mask14", 14)])
def test_multiscale_roi_align_single_launch(ops, golden_dir, tag, P):
    g = load(golden_dir, "multiscale_" + tag)
    feats, boxes, shapes = cases.multiscale_case()
    pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)
    x = {str(i): cu(f) for i, f in enumerate(feats)}
    x["pool"] = torch.zeros(1, device=DEV)   # ignored like in the reference (not in featmap_names)
    out, levels = pool(x, [cu(b) for b in boxes], shapes, return_levels=True)
    assert np.array_equal(levels.cpu().numpy().astype(np.int64), g["levels"])
    assert pool.scales == list(g["scales"])
    out = out.cpu().numpy()
    assert np.array_equal(out[:, ::4], g["out"])
    assert np.array_equal(out, D.multiscale_roi_align(feats, boxes, shapes, P, 2))


def test_roi_align_full_size_properties(ops):
    """BASELINE config 2 shapes (4 x 256 ch pyramid of an 800^2 image, 1000 RoIs/img): properties
    that need no oracle run — linearity in the features and agreement of exact/fast modes — plus
    a bit-exact oracle check on a random subset of RoIs."""
    rng = np.random.default_rng(0)
    n, c = 4, 256
    feats = [torch.randn(n, c, 800 // s, 800 // s, device=DEV) for s in (4, 8, 16, 32)]
    boxes = [cu(cases.stress_rois(rng, 1000, (800, 800))) for _ in range(n)]
    shapes = [(800, 800)] * n
    pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    x = {str(i): f for i, f in enumerate(feats)}
    out, levels = pool(x, boxes, shapes, return_levels=True)
    assert out.shape == (4000, 256, 7, 7)
    assert set(levels.cpu().tolist()) == {0, 1, 2, 3}
    x2 = {k: 2.0 * v for k, v in x.items()}
    assert torch.equal(pool(x2, boxes, shapes), 2.0 * out)          # exact in fp32: scaling by 2
    fast = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, exact=False)(x, boxes, shapes)
    assert torch.allclose(fast, out, rtol=1e-5, atol=5e-5)
    pick = np.sort(rng.choice(4000, 64, replace=False))
    sub_boxes = [boxes[i].cpu().numpy()[pick[pick // 1000 == i] % 1000] for i in range(n)]
    ref = D.multiscale_roi_align([f.cpu().numpy() for f in feats], sub_boxes, shapes, 7, 2)
    assert np.array_equal(out[torch.from_numpy(pick).to(DEV)].cpu().numpy(), ref)


def test_box_ops(ops, golden_dir):
    g = load(golden_dir, "boxes")
    b, rel, wild = g["boxes"], g["rel"], g["wild"]
    d1 = ops.decode_boxes(cu(rel[:, :4].copy()), [cu(b)], (1.0, 1.0, 1.0, 1.0)).cpu().numpy()
    d2 = ops.decode_boxes(cu(rel), [cu(b[:1200]), cu(b[1200:])], (10.0, 10.0, 5.0, 5.0)).cpu().numpy()
    assert d1.shape == (2000, 1, 4) and d2.shape == (2000, 3, 4)
    assert cases.box_rel_err(d1, g["decode_rpn"]) < 1e-5
    assert cases.box_rel_err(d2, g["decode_roi"]) < 1e-5
    clip = ops.clip_boxes_to_image(cu(wild), (480, 640))
    assert np.array_equal(clip.cpu().numpy(), g["clip"])
    assert np.array_equal(ops.remove_small_boxes(clip, 1e-3).cpu().numpy(), g["small_1e-3"])
    assert np.array_equal(ops.remove_small_boxes(clip, 20.0).cpu().numpy(), g["small_20"])
    for a in ("xyxy", "xywh", "cxcywh"):
        for c in ("xyxy", "xywh", "cxcywh"):
            assert np.array_equal(ops.box_convert(cu(b), a, c).cpu().numpy(), g[f"convert_{a}_{c}"]), (a, c)
    with pytest.raises(ValueError):
        ops.box_convert(cu(b), "xyxy", "bogus")
    assert np.array_equal(ops.resize_boxes(cu(b), [800, 800], [1024, 1024]).cpu().numpy(), g["resize"])
    assert np.array_equal(ops.resize_boxes(cu(b), [800, 1216], [683, 1024]).cpu().numpy(), g["resize2"])


def test_grid_anchors(ops, golden_dir):
    g = load(golden_dir, "anchors")
    outs = []
    for lvl, (gh, gw) in enumerate(g["grids"]):
        base = ops.base_anchors(cases.RPN_SIZES[lvl], cases.RPN_RATIOS[lvl])
        outs.append(ops.grid_anchors(base, (gh, gw), (224 // gh, 288 // gw), DEV).cpu().numpy())
    assert np.array_equal(np.concatenate(outs), g["anchors"])


def test_roi_align_channels_last_features_bit_identical(ops, golden_dir):
    """Feature maps in torch.channels_last memory format are consumed in place by the NHWC kernel;
    the result must be bit-identical to the NCHW path (and hence to the CPU reference)."""
    g = load(golden_dir, "multiscale_box7")
    feats, boxes, shapes = cases.multiscale_case()
    pool = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    x_nchw = {str(i): cu(f) for i, f in enumerate(feats)}
    x_nhwc = {k: v.contiguous(memory_format=torch.channels_last) for k, v in x_nchw.items()}
    assert not x_nhwc["0"].is_contiguous()
    b = [cu(bb) for bb in boxes]
    a = pool(x_nchw, b, shapes)
    c = pool(x_nhwc, b, shapes)
    assert torch.equal(a, c)
    assert np.array_equal(c.cpu().numpy()[:, ::4], g["out"])
    assert c.is_contiguous()                               # output stays [K, C, P, P] contiguous for the box head
    for name, x, rois, scale, P, sr, aligned in cases.roi_align_cases():
        if sr != 2:
            continue
        xc = cu(x).contiguous(memory_format=torch.channels_last)
        out = ops.roi_align(xc, cu(rois), P, scale, sr, aligned).cpu().numpy()
        assert np.array_equal(out, native.roi_align(x, rois, scale, P, P, sr, aligned)), name


def test_roi_align_separable_fast_mode_channels_last(ops):
    """exact=False on channels-last maps takes the separable merged-tap kernel (k_roi_align_nhwc_sep):
    same mathematical sum, different association -> within 1e-5 (relative + feature scale), the
    tolerance north_star states for RoIAlign; covers border clamps, out-of-image samples, tiny RoIs
    (several samples on one pixel row), aligned/unaligned, 7x7 and 14x14, dead slots."""
    for name, x, rois, scale, P, sr, aligned in cases.roi_align_cases():
        if sr != 2 or x.shape[1] % 4:
            continue
        xc = cu(x).contiguous(memory_format=torch.channels_last)
        out = ops.roi_align(xc, cu(rois), P, scale, sr, aligned, exact=False).cpu().numpy()
        ref = native.roi_align(x, rois, scale, P, P, sr, aligned)
        tol = 1e-5 * np.abs(ref) + 1e-5 * np.abs(x).max()
        assert np.all(np.abs(out - ref) <= tol), name
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 64, 50, 60)).astype(np.float32)
    tiny = np.concatenate([rng.integers(0, 2, (300, 1)).astype(np.float32),
                           cases.stress_rois(rng, 300, (200, 240))], 1).astype(np.float32)
    tiny[:100, 3:] = tiny[:100, 1:3] + rng.uniform(0.0, 6.0, (100, 2)).astype(np.float32)   # sub-bin-sized boxes
    tiny[100:120, 1:] = np.array([-50, -50, -20, -20], np.float32)                       # fully outside
    for P in (7, 14):
        xc = cu(x).contiguous(memory_format=torch.channels_last)
        out = ops.roi_align(xc, cu(tiny), P, 0.25, 2, False, exact=False).cpu().numpy()
        ref = native.roi_align(x, tiny, 0.25, P, P, 2, False)
        assert np.all(np.abs(out - ref) <= 1e-5 * np.abs(ref) + 1e-5 * np.abs(x).max()), P
    feats, boxes, shapes = cases.multiscale_case()
    xm = {str(i): cu(f).contiguous(memory_format=torch.channels_last) for i, f in enumerate(feats)}
    for P in (7, 14):
        fast = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2, exact=False)(xm, [cu(b) for b in boxes], shapes).cpu().numpy()
        ref = D.multiscale_roi_align(feats, boxes, shapes, P, 2)
        fmax = max(float(np.abs(f).max()) for f in feats)
        assert np.all(np.abs(fast - ref) <= 1e-5 * np.abs(ref) + 1e-5 * fmax), P


def test_paste_masks_in_image(ops, golden_dir):
    """mb_paste_masks against the golden torchvision output and the oracle: same pixels written, values
    within 2 ulp (see tests/test_oracle_pin.py::test_paste_masks_golden); plus properties at full size."""
    g = load(golden_dir, "paste_masks")
    masks, boxes, hw = cases.paste_case()
    out = ops.paste_masks_in_image(cu(masks), cu(boxes), hw).cpu().numpy()
    assert out.shape == g["out"].shape and out.dtype == np.float32
    assert np.array_equal(out != 0, g["out"] != 0)
    assert np.max(np.abs(out - g["out"])) <= 2.4e-7
    ref = D.paste_masks_in_image(masks, boxes, hw)
    assert np.array_equal(out, ref)                       # same arithmetic as the oracle: bit-identical
    # odd width (scalar store path), other mask size / padding, empty input, box outside the image -> zeros
    m2 = np.random.default_rng(1).random((5, 1, 14, 14)).astype(np.float32)
    b2 = np.array([[3, 4, 60, 70], [0, 0, 10, 10], [50, 50, 101, 99], [-40, -40, -5, -5], [20.2, 30.7, 25.1, 90.3]], np.float32)
    o2 = ops.paste_masks_in_image(cu(m2), cu(b2), (99, 101), padding=2).cpu().numpy()
    keep = [0, 1, 2, 4]
    assert np.array_equal(o2[keep], D.paste_masks_in_image(m2[keep], b2[keep], (99, 101), padding=2))
    assert not o2[3].any()
    assert ops.paste_masks_in_image(torch.zeros((0, 1, 28, 28), device=DEV), torch.zeros((0, 4), device=DEV), (64, 64)).shape == (0, 1, 64, 64)
    # full size: 100 masks into 1024^2 (419 MB): values in [0, 1], nothing outside the expanded integer boxes
    rng = np.random.default_rng(2)
    mk = torch.rand((100, 1, 28, 28), device=DEV)
    bx = cu(cases.stress_rois(rng, 100, (1024, 1024), side=(16.0, 400.0)))
    big = ops.paste_masks_in_image(mk, bx, (1024, 1024))
    assert big.shape == (100, 1, 1024, 1024) and float(big.min()) >= 0.0 and float(big.max()) <= 1.0
    sub = big[:8].cpu().numpy()
    assert np.array_equal(sub, D.paste_masks_in_image(mk[:8].cpu().numpy(), bx[:8].cpu().numpy(), (1024, 1024)))


def test_transform_images(ops, golden_dir):
    """mb_image_transform (ToTensor + normalize + bilinear resize + zero-padded batch in one launch) against
    the torchvision golden fixture and the oracle — bit-identical — and at the BASELINE size (1024^2 -> 800^2)."""
    g = load(golden_dir, "transform")
    imgs, mn, mx, mean, std = cases.transform_case()
    batch, sizes = ops.transform_images([cu(a) for a in imgs], mn, mx, mean, std)
    assert np.array_equal(np.array(sizes), g["sizes"])
    assert np.array_equal(batch.cpu().numpy(), g["batch"])
    rng = np.random.default_rng(0)
    big = [rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8) for _ in range(2)] + [rng.integers(0, 256, (700, 1000, 3), dtype=np.uint8)]
    b2, s2 = ops.transform_images([cu(a) for a in big], 800, 1333, mean, std)
    rb, rs = D.transform_images(big, 800, 1333, mean, std)
    assert s2 == rs and tuple(b2.shape) == rb.shape == (3, 3, 800, 1152)
    assert np.array_equal(b2.cpu().numpy(), rb)


@pytest.mark.parametrize("out_size", [(10, 6), (16, 16), (3, 3), (1, 1), (5, 14)])
def test_roi_align_channels_last_other_output_sizes(ops, out_size):
    """The 16-byte-gather kernel splits large outputs into bands of pooled rows (16x16 -> 6 + 6 + 4 rows) and
    uses the plain copy-out when the staged chunk is not the output layout: still bit-exact."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 24, 40, 52)).astype(np.float32)
    rois = np.concatenate([rng.integers(0, 2, (60, 1)).astype(np.float32), cases.stress_rois(rng, 60, (160, 208), side=(4.0, 200.0))], 1)
    xc = cu(x).contiguous(memory_format=torch.channels_last)
    for aligned in (False, True):
        out = ops.roi_align(xc, cu(rois), out_size, 0.25, 2, aligned).cpu().numpy()
        assert np.array_equal(out, native.roi_align(x, rois, 0.25, out_size[0], out_size[1], 2, aligned))


def test_transform_images_gray(ops):
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (90, 70, 1), dtype=np.uint8), rng.integers(0, 256, (64, 120, 1), dtype=np.uint8)]
    b, s = ops.transform_images([cu(a) for a in imgs], 128, 200, [0.5], [0.25])
    rb, rs = D.transform_images(imgs, 128, 200, [0.5], [0.25])
    assert s == rs and np.array_equal(b.cpu().numpy(), rb)


def test_maskrcnn_inference_selected_channel_sigmoid(ops):
    """mb_mask_prob against torchvision's maskrcnn_inference (tv:models/detection/roi_heads.py:56-82) on the CPU."""
    from torchvision.models.detection.roi_heads import maskrcnn_inference as tv_inf
    tv_inf = getattr(__import__("torchvision.models.detection.roi_heads", fromlist=["x"]), "_miso_b200_orig_maskrcnn_inference", tv_inf)
    rng = np.random.default_rng(8)
    x = torch.from_numpy((rng.standard_normal((37, 3, 28, 28)) * 4).astype(np.float32))
    labels = [torch.from_numpy(rng.integers(1, 3, k)) for k in (20, 0, 17)]
    ref = tv_inf(x, labels)
    got = ops.maskrcnn_inference(x.to(DEV), [l.to(DEV) for l in labels])
    assert [tuple(g.shape) for g in got] == [tuple(r.shape) for r in ref]
    for g, r in zip(got, ref):
        assert float((g.cpu() - r).abs().max()) <= 1e-6 if r.numel() else True


def test_nms_reports_group_index_outside_range(ops):
    """mb_nms in per-group mode must flag a group index >= num_groups for any group count — also for a single group,
    where the kept set is emitted by the one-segment kernel."""
    from miso_b200 import MisoB200Error
    rng = np.random.default_rng(2)
    b, s = cases.random_boxes(rng, 300, extent=100.0), cases.distinct_scores(rng, 300)
    for groups, num_groups in ((np.r_[np.zeros(299, np.int64), 1], 1), (rng.integers(0, 4, 300), 3)):
        with pytest.raises(MisoB200Error):
            ops._nms_impl(cu(b), cu(s), cu(groups.astype(np.int64)), num_groups, 1, 0.5)
    ok = ops._nms_impl(cu(b), cu(s), cu(np.zeros(300, np.int64)), 1, 1, 0.5)
    assert np.array_equal(ok.cpu().numpy(), D.nms(b, s, 0.5))
