"""Module-level drop-in and the kept entry points: a torchvision Faster R-CNN /
Mask R-CNN with random-init weights is patched in place; every patched stage is compared with the
UNPATCHED torchvision CPU code on the same stage inputs (captured from the GPU forward), and the
`python -m miso.cli infer-object-detector-directory` command is run end to end."""
import os

import numpy as np
import pytest
import torch

from oracle import miso_path as M
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_model(kind="faster"):
    from miso.object_detection.models import get_instance_segmentation_model, get_object_detection_model
    torch.manual_seed(0)
    m = get_object_detection_model(3) if kind == "faster" else get_instance_segmentation_model(3)
    # sharpen the random box classifier so that scores spread over (0, 1) instead of sitting at 1/3
    # (moderately: saturated scores would tie, and the reference's final sort is not stable on ties)
    with torch.no_grad():
        m.roi_heads.box_predictor.cls_score.weight.mul_(8.0)
    return m.eval()


def images(n=2, size=(512, 640), seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(3, size[0] + 32 * i, size[1], generator=g) for i in range(n)]


def match_rate(a, b, la, lb, tol=1e-3):
    """fraction of boxes in a that have a same-label box in b within tol (relative to box scale)"""
    if len(a) == 0:
        return 1.0 if len(b) == 0 else 0.0
    hit = 0
    for i in range(len(a)):
        d = np.abs(b - a[i]).max(axis=1) / max(np.abs(a[i]).max(), 1.0)
        hit += bool(np.any((d < tol) & (lb == la[i])))
    return hit / len(a)


# ("faster", 1, (1024, 1024)) is BASELINE config 1 at full size: one 1024^2 image through the patched Faster R-CNN R50-FPN
@pytest.mark.parametrize("kind,n_img,size", [("faster", 2, (512, 640)), ("mask", 2, (512, 640)), ("faster", 1, (1024, 1024))])
def test_patched_model_matches_reference_stage_by_stage(kind, n_img, size):
    from miso_b200.patch import patch_model, unpatch_model
    model = make_model(kind).to(DEV)
    imgs = [im.to(DEV) for im in images(n_img, size)]
    cap = {}

    # capture the stage inputs of the GPU forward
    def rpn_head_hook(mod, inp, out):
        cap["objectness"], cap["deltas"] = [o.detach().cpu() for o in out[0]], [d.detach().cpu() for d in out[1]]
    h1 = model.rpn.head.register_forward_hook(rpn_head_hook)
    orig_pp = model.roi_heads.postprocess_detections

    patch_model(model)
    assert getattr(model, "_miso_b200_patched")
    patched_pp = model.roi_heads.postprocess_detections

    def spy_pp(class_logits, box_regression, proposals, image_shapes):
        # the patched forward never waits for the proposal counts: `proposals` is a LazyProposals and the pooler ran on
        # its padded [N, R, 4] layout, so logits / regression have N*R rows (rows beyond an image's count are ignored)
        from miso_b200.patch import LazyProposals
        assert isinstance(proposals, LazyProposals)
        n_img, cap_r = proposals.padded.shape[:2]
        assert class_logits.shape[0] == n_img * cap_r
        cnt = proposals.counts.tolist()
        live = torch.cat([torch.arange(c) + i * cap_r for i, c in enumerate(cnt)])
        cap["live_rows"] = live
        cap["logits"], cap["reg"] = class_logits.detach().cpu()[live], box_regression.detach().cpu()[live]
        cap["proposals"], cap["shapes"] = [p.detach().cpu() for p in proposals], list(image_shapes)
        res = patched_pp(class_logits, box_regression, proposals, image_shapes)
        cap["pp_out"] = [[t.detach().cpu() for t in r] for r in res]
        return res
    model.roi_heads.postprocess_detections = spy_pp
    box_pool = model.roi_heads.box_roi_pool

    def pool_hook(mod, inp, out):
        cap["pool_in"] = ({k: v.detach().cpu() for k, v in inp[0].items()}, [b.detach().cpu() for b in inp[1]], list(inp[2]))
        cap["pool_out"] = out.detach().cpu()
    h2 = box_pool.register_forward_hook(pool_hook)
    if kind == "mask":      # capture the per-image mask probabilities that reach transform.postprocess
        patched_tp = model.transform.postprocess

        def spy_tp(result, image_shapes, original_image_sizes):
            cap["mask_probs"] = [r["masks"].detach().cpu().clone() for r in result]
            return patched_tp(result, image_shapes, original_image_sizes)
        model.transform.postprocess = spy_tp
    with torch.inference_mode():
        out = model(imgs)
    h1.remove(); h2.remove()
    assert len(out) == len(imgs) and set(out[0]) >= {"boxes", "labels", "scores"}
    if kind == "mask":
        assert out[0]["masks"].shape[-2:] == imgs[0].shape[-2:]
        # mb_paste_masks inside the patched transform.postprocess against torchvision's CPU paste on the same
        # (captured) 28x28 mask probabilities and final boxes
        from torchvision.models.detection.roi_heads import paste_masks_in_image as tv_paste
        for i, o in enumerate(out):
            if len(o["boxes"]) == 0:
                continue
            ref_m = tv_paste(cap["mask_probs"][i], o["boxes"].cpu(), tuple(imgs[i].shape[-2:]))
            got_m = o["masks"].cpu()
            assert got_m.shape == ref_m.shape
            assert torch.equal(got_m != 0, ref_m != 0)
            assert float((got_m - ref_m).abs().max()) <= 2.4e-7

    # ---- reference CPU code on the captured inputs ----
    cpu = make_model(kind)          # same seed -> same hyper-parameters (weights are irrelevant here)
    from torchvision.models.detection.image_list import ImageList
    from torchvision.models.detection.rpn import concat_box_prediction_layers
    padded = (cap["objectness"][0].shape[-2] * 4, cap["objectness"][0].shape[-1] * 4)
    il = ImageList(torch.zeros(len(imgs), 3, *padded), cap["shapes"])
    with torch.inference_mode():
        anchors = cpu.rpn.anchor_generator(il, cap["objectness"])
        napl = [o.shape[1] * o.shape[2] * o.shape[3] for o in cap["objectness"]]
        objectness, deltas = concat_box_prediction_layers(list(cap["objectness"]), list(cap["deltas"]))
        decoded = cpu.rpn.box_coder.decode(deltas, anchors).view(len(imgs), -1, 4)
        ref_props, _ = cpu.rpn.filter_proposals(decoded, objectness, il.image_sizes, napl)
    # A random-init RPN emits logits within ~1e-3 of zero: many DISTINCT logits share one fp32 sigmoid
    # value, and the reference's final cross-level sort is not stable on such ties.
    # Compare the proposals as sets above the last (possibly tied) score, in a canonical order.
    from miso_b200 import detection
    cfg = detection.RpnConfig.from_model(cpu.rpn)
    gout = detection.rpn_proposals([o.to(DEV) for o in cap["objectness"]], [d.to(DEV) for d in cap["deltas"]],
                                   cap["shapes"], padded, cfg)
    gb, gs = gout.as_lists()
    with torch.inference_mode():
        ref_props, ref_scores = cpu.rpn.filter_proposals(decoded, objectness, il.image_sizes, napl)

    def canon(b, s, cut):
        b, s = b[s > cut], s[s > cut]
        order = np.lexsort((b[:, 3], b[:, 2], b[:, 1], b[:, 0], -s))
        return b[order], s[order]
    for i, (p_gpu, p_ref) in enumerate(zip(cap["proposals"], ref_props)):
        assert p_gpu.shape == p_ref.shape
        assert torch.equal(p_gpu, gb[i].cpu())                       # the patched model used the same stage
        cut = max(float(gs[i].min()), float(ref_scores[i].min())) + 1e-7
        ab, as_ = canon(gb[i].cpu().numpy(), gs[i].cpu().numpy(), cut)
        rb_, rs_ = canon(p_ref.numpy(), ref_scores[i].numpy(), cut)
        assert ab.shape == rb_.shape and len(ab) > 100
        assert np.max(np.abs(as_ - rs_)) < 1e-6
        assert cases.box_rel_err(ab, rb_) < 1e-4   # canonical order of near-equal boxes; exactness is tested per stage
    # RoIAlign: bit-exact against torchvision's CPU pooler on identical inputs
    with torch.inference_mode():
        ref_pool = cpu.roi_heads.box_roi_pool(cap["pool_in"][0], cap["pool_in"][1], cap["pool_in"][2])
    assert cap["pool_out"].shape[0] == len(imgs) * model.rpn.post_nms_top_n()     # padded layout, no RoI tensor was built
    assert torch.equal(cap["pool_out"][cap["live_rows"]], ref_pool)
    dead = torch.ones(cap["pool_out"].shape[0], dtype=torch.bool)
    dead[cap["live_rows"]] = False
    assert not cap["pool_out"][dead].any()
    # detection post-processing on identical inputs
    with torch.inference_mode():
        rb, rs, rl = cpu.roi_heads.postprocess_detections(cap["logits"], cap["reg"], cap["proposals"], cap["shapes"])
    for i in range(len(imgs)):
        gb, gs, gl = (cap["pp_out"][j][i].numpy() for j in range(3))
        assert len(gl) == len(rl[i])
        # same tie caveat as above (the reference's last sort is unstable): canonical order above the cut
        cut = max(gs.min(), float(rs[i].min())) + 1e-7 if len(gs) else 0.0

        def canon_det(b, s_, l):
            m = s_ > cut
            b, s_, l = b[m], s_[m], l[m]
            order = np.lexsort((b[:, 3], b[:, 2], b[:, 1], b[:, 0], l, -s_))
            return b[order], s_[order], l[order]
        ab, as_, al = canon_det(gb, gs, gl)
        rb_, rs_, rl_ = canon_det(rb[i].numpy(), rs[i].numpy(), rl[i].numpy())
        assert np.array_equal(al, rl_)
        assert np.max(np.abs(as_ - rs_)) < 1e-6 if len(as_) else True
        assert cases.box_rel_err(ab, rb_) < 1e-4

    # ---- whole model: patched vs unpatched torchvision CUDA path ----
    model.roi_heads.postprocess_detections = patched_pp
    with torch.inference_mode():
        a = model(imgs)
        unpatch_model(model)
        assert model.roi_heads.postprocess_detections == orig_pp
        # keep the backbone in the memory format the patched run used: cuDNN's NHWC and NCHW convolutions round
        # differently, and a random-init detector amplifies that into different boxes — not what this compares
        model.backbone.to(memory_format=torch.channels_last)
        b = model(imgs)
    for x, y in zip(a, b):
        r = match_rate(x["boxes"].cpu().numpy(), y["boxes"].cpu().numpy(), x["labels"].cpu().numpy(), y["labels"].cpu().numpy())
        assert r > 0.9, r


def test_forward_uint8_equals_float_forward():
    """patch.forward_uint8 (fused ToTensor + normalize + resize + batch) feeds the backbone the same tensor,
    bit for bit, as the reference transform on the float images — so the detections are identical."""
    from miso_b200.patch import forward_uint8, patch_model
    model = patch_model(make_model("faster").to(DEV))
    g = torch.Generator().manual_seed(5)
    u8 = [torch.randint(0, 256, (300 + 40 * i, 420 - 30 * i, 3), dtype=torch.uint8, generator=g).to(DEV) for i in range(2)]
    with torch.inference_mode():
        il, _ = model.transform([a.permute(2, 0, 1).to(torch.float32) / 255 for a in u8])
        from miso_b200 import ops
        tr = model.transform
        batch, sizes = ops.transform_images(u8, tr.min_size[-1], tr.max_size, tr.image_mean, tr.image_std, tr.size_divisible)
        assert sizes == [tuple(s) for s in il.image_sizes]
        # torchvision's CUDA interpolate rounds differently from its CPU kernel; the parity target is the CPU path
        # (tests/test_gpu_ops.py::test_transform_images is bit-exact against it), so here: close, same shape
        assert batch.shape == il.tensors.shape and float((batch - il.tensors).abs().max()) < 1e-4
        a = forward_uint8(model, u8)
        # the model's own forward on the float images, with its transform handing over the same batch: everything
        # around the transform (original sizes, backbone call, postprocess) must then agree exactly
        from torchvision.models.detection.image_list import ImageList
        model.transform.forward = lambda images, targets=None: (ImageList(batch, sizes), targets)
        b = model([x.permute(2, 0, 1).to(torch.float32) / 255 for x in u8])
        del model.transform.forward
    for x, y in zip(a, b):
        assert torch.equal(x["boxes"], y["boxes"]) and torch.equal(x["labels"], y["labels"]) and torch.equal(x["scores"], y["scores"])


def test_forward_uint8_fixed_size_and_skip_resize_follow_the_models_transform():
    """GeneralizedRCNNTransform's other two modes (fixed_size, _skip_resize) are not what the fused transform kernel
    implements: forward_uint8 must hand those to the model's own transform and return what model(float images)
    returns."""
    from miso_b200.patch import forward_uint8, patch_model
    model = patch_model(make_model("faster").to(DEV))
    g = torch.Generator().manual_seed(9)
    u8 = [torch.randint(0, 256, (320, 384, 3), dtype=torch.uint8, generator=g).to(DEV) for _ in range(2)]
    fl = [x.permute(2, 0, 1).to(torch.float32) / 255 for x in u8]
    for attr, value in (("fixed_size", (352, 416)), ("_skip_resize", True)):
        old = getattr(model.transform, attr, None)
        setattr(model.transform, attr, value)
        try:
            with torch.inference_mode():
                a, b = forward_uint8(model, u8), model(fl)
        finally:
            setattr(model.transform, attr, old)
        for x, y in zip(a, b):
            assert torch.equal(x["boxes"], y["boxes"]) and torch.equal(x["labels"], y["labels"]) and torch.equal(x["scores"], y["scores"])


def test_dispatcher_override_routes_torchvision_ops():
    import torchvision
    from miso_b200.patch import override_torchvision_ops
    from oracle import native
    rng = np.random.default_rng(1)
    b, s = cases.random_boxes(rng, 800), cases.distinct_scores(rng, 800)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        override_torchvision_ops()
    keep = torchvision.ops.nms(torch.from_numpy(b).to(DEV), torch.from_numpy(s).to(DEV), 0.5)
    assert np.array_equal(keep.cpu().numpy(), native.nms(b, s, 0.5))
    x = torch.randn(1, 8, 30, 30, device=DEV)
    rois = torch.tensor([[0, 2.5, 3.5, 20.0, 25.0]], device=DEV)
    out = torchvision.ops.roi_align(x, rois, 7, 1.0, 2)
    assert np.array_equal(out.cpu().numpy(), native.roi_align(x.cpu().numpy(), rois.cpu().numpy(), 1.0, 7, 7, 2, False))


def test_cli_infer_directory_writes_reference_crops(tmp_path):
    from click.testing import CliRunner
    from PIL import Image
    import miso.cli as cli
    model = make_model("faster")
    mdir = tmp_path / "models" / "m1"
    mdir.mkdir(parents=True)
    torch.save(model, mdir / "model.pt")                    # whole pickled module, like the reference
    (mdir / "labels.txt").write_text("0,Coccolith\n1,Coccosphere\n")
    idir = tmp_path / "in" / "sub"
    idir.mkdir(parents=True)
    rng = np.random.default_rng(0)
    arrs = {}
    for name in ("a.png", "b.png"):
        arr = rng.integers(0, 256, (384, 512, 3), dtype=np.uint8)
        Image.fromarray(arr).save(idir / name)
        arrs[name] = arr
    odir = tmp_path / "out"
    res = CliRunner().invoke(cli.cli, ["infer-object-detector-directory", "-i", str(tmp_path / "in"), "-o", str(odir),
                                       "--model-dir", str(tmp_path / "models"), "--model", "m1", "--threshold", "0.3",
                                       "--batch-size", "2"], catch_exceptions=False)
    assert res.exit_code == 0, res.output
    files = sorted(p for p in odir.rglob("*.png"))
    assert files, "no crops written"
    for f in files:
        assert f.parent.name in ("Coccolith", "Coccosphere") and f.parent.parent.name == "sub"
        stem, x, y, w, h = f.stem.rsplit("_", 4)
        crop = np.asarray(Image.open(f))
        src = arrs[stem + ".png"]
        assert crop.shape[0] <= src.shape[0] and crop.shape[1] <= src.shape[1] and crop.ndim == 3
        # the file name carries the rounded bounds; the pixels must be a slice of the source image there
        xs, ys = int(x), int(y)
        found = any(np.array_equal(crop, src[yy:yy + crop.shape[0], xx:xx + crop.shape[1]])
                    for yy in range(max(ys - 1, 0), ys + 2) for xx in range(max(xs - 1, 0), xs + 2))
        assert found


def test_infer_mosaic_equals_per_tile_composition():
    """Config 5 driver on a small mosaic (4 overlapping tiles): its result must equal the explicit composition:
    per-tile detections of the same model -> + tile origin -> score filter -> the oracle's per-class NMS -> miso's
    rounding + numpy slices of the mosaic (tests/mosaic_ref.py)."""
    from miso_b200 import mosaic
    from miso_b200.patch import forward_uint8, patch_model
    from tests import mosaic_ref as R
    model = patch_model(make_model("faster").to(DEV))
    g = torch.Generator().manual_seed(5)
    mos = torch.randint(0, 256, (1536, 1536, 3), dtype=torch.uint8, generator=g)
    mos_dev = mos.to(DEV)
    thr = 0.3
    res = mosaic.infer_mosaic(model, mos_dev, (1536, 1536), tile=1024, overlap=128, threshold=thr, batch_size=2)
    grid = mosaic.tile_grid(1536, 1536, 1024, 128)
    assert grid == [(0, 0), (0, 512), (512, 0), (512, 512)]
    dpi = int(model.roi_heads.detections_per_img)
    block = np.zeros((4 * dpi, 6), np.float32)
    block[:, 5] = -1
    with torch.inference_mode():
        for i0 in range(0, 4, 2):
            # the same forward infer_mosaic uses (fused uint8 input transform); the float path differs from it by
            # ~1e-6 at the backbone input (torchvision's CUDA interpolate), which a random-init model amplifies
            tiles = [mos_dev[y:y + 1024, x:x + 1024] for y, x in grid[i0:i0 + 2]]
            for t, ((y, x), r) in enumerate(zip(grid[i0:i0 + 2], forward_uint8(model, tiles))):
                off = torch.tensor([x, y, x, y], dtype=torch.float32, device=DEV)
                k = r["boxes"].shape[0]
                rows = block[(i0 + t) * dpi:(i0 + t) * dpi + k]
                rows[:, :4] = (r["boxes"] + off).cpu().numpy()
                rows[:, 4] = r["scores"].cpu().numpy()
                rows[:, 5] = np.where(r["scores"].cpu().numpy() > np.float32(thr), r["labels"].cpu().numpy().astype(np.float32), -1.0)
    assert np.array_equal(res["gathered"].cpu().numpy(), block)
    keep = R.seam_keep_rows(block, float(model.roi_heads.nms_thresh))
    assert len(keep) > 0
    assert np.array_equal(np.nonzero(res["state"].cpu().numpy() == 1)[0], keep)
    xywh, ci, ref = R.crops_of_rows(mos.numpy(), block, keep)
    c = res["crops"]
    assert c["count"] == len(keep) and np.array_equal(c["src"].cpu().numpy(), keep) and np.array_equal(c["xywh"].cpu().numpy(), xywh)
    rects, offs, pix = c["rects"].cpu().numpy(), c["offsets"].cpu().numpy(), c["pixels"].cpu().numpy()
    for j, r in enumerate(ref):
        assert np.array_equal(pix[offs[j]:offs[j + 1]].reshape(rects[j][3], rects[j][2], 3), r)


def test_patched_model_runs_channels_last_and_never_syncs_before_the_results():
    """patch_model moves backbone + FPN to torch.channels_last, so cuDNN hands the pooler NHWC maps (gathered in place
    by the RoIAlign kernels: no transpose launch), and the RPN hands over a LazyProposals that nobody
    materialises during the forward pass."""
    from miso_b200 import _lib
    from miso_b200.patch import LazyProposals, patch_model
    model = patch_model(make_model("faster").to(DEV))
    seen = {}
    pool = model.roi_heads.box_roi_pool

    def pre_hook(mod, inp):
        feats = [v for k, v in inp[0].items() if k in mod.featmap_names]
        seen["nhwc"] = all((not f.is_contiguous()) and f.is_contiguous(memory_format=torch.channels_last) for f in feats)
        seen["lazy"] = isinstance(inp[1], LazyProposals) and not inp[1].materialised
        seen["props"] = inp[1]
    h = pool.register_forward_pre_hook(pre_hook)
    with torch.inference_mode():
        out = model([im.to(DEV) for im in images()])
    h.remove()
    assert seen["nhwc"] and seen["lazy"]
    assert not seen["props"].materialised                      # the whole forward ran without reading the proposal counts
    assert len(out) == 2 and out[0]["boxes"].shape[1] == 4


def test_crop_objects_any_dtype_palette_and_float64_bounds(tmp_path):
    """crop_objects on what skimage.io.imread hands the reference: a 16-bit TIFF (cropped as bytes and viewed back),
    a palette PNG (expanded to RGB), an RGB PNG — with float64 annotation bounds whose x + w lands on .5 only in
    float64, a negative corner (numpy's wrap-around slice) and a box past the image edge. Every written crop equals
    the reference's `im[c[1]:c[3], c[0]:c[2], ...]` with c = int(np.round(.)) of the Python-float sums."""
    from PIL import Image
    from miso.object_detection.crop import crop_objects
    from miso.object_detection.dataset.annotation import RectangleAnnotation
    from miso.object_detection.dataset.image import ImageMetadata
    from miso.object_detection.dataset.project import Project
    rng = np.random.default_rng(5)
    idir = tmp_path / "in"
    idir.mkdir()
    im16 = rng.integers(0, 65536, (90, 130), dtype=np.uint16)
    Image.fromarray(im16).save(idir / "deep.tif")
    pal = Image.fromarray(rng.integers(0, 256, (70, 95), dtype=np.uint8), mode="P")
    pal.putpalette([int(v) for v in rng.integers(0, 256, 768)])
    pal.save(idir / "pal.png")
    rgb = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(idir / "rgb.png")
    want = {"deep.tif": im16, "pal.png": np.array(pal.convert("RGB")), "rgb.png": rgb}
    # 56.828.. + 17.671.. = 74.50000016 in float64 (-> 75) but 74.5 in float32 (-> 74, half to even); the second box
    # is 29.4999981 in float64 (-> 29) and 29.5 in float32 (-> 30): XML / CVAT annotations are Python floats
    boxes = [(56.82802170936761, 5.0, 17.671978449834047, 30.0), (12.996859387349325, 12.996859387349325, 16.50313874181763, 16.50313874181763),
             (0.3, 0.3, 2.2, 2.2), (-4.0, 3.0, 10.0, 12.0), (50.0, 40.0, 500.0, 500.0), (3.5, 2.5, 4.0, 5.0)]
    assert int(np.round(boxes[0][0] + boxes[0][2])) != int(np.round(np.float32(boxes[0][0]) + np.float32(boxes[0][2])))
    project = Project()
    for name in want:
        meta = ImageMetadata(name, str(idir))
        meta.boxes = [RectangleAnnotation(*b, label="obj") for b in boxes]
        project.add_image(meta)
    odir = tmp_path / "out"
    crop_objects(project, str(odir))
    seen = 0
    for name, im in want.items():
        stem, suffix = name.rsplit(".", 1)
        for x, y, w, h in boxes:
            c = tuple(int(np.round(v)) for v in (x, y, x + w, y + h))
            ref = im[c[1]:c[3], c[0]:c[2], ...]
            f = odir / "obj" / f"{stem}_{x:.0f}_{y:.0f}_{w:.0f}_{h:.0f}.{suffix}"
            if ref.size == 0:
                assert not f.exists()
                continue
            got = np.array(Image.open(f))
            assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref), (name, (x, y, w, h))
            seen += 1
    assert seen >= 15
