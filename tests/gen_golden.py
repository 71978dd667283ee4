"""Generate tests/golden/*.npz from the reference arithmetic itself.

Run in the build container (where /root/reference and torchvision 0.26.0 are importable):

    python tests/gen_golden.py

Sources of truth:
  * torchvision 0.26.0 CPU ops and detection modules (the third-party code the reference's
    pipelines execute: SURVEY.md §2.2) for nms / batched_nms / roi_align / MultiScaleRoIAlign /
    AnchorGenerator / BoxCoder / filter_proposals / postprocess_detections / resize_boxes;
  * the reference's own RectangleAnnotation (ref:miso/object_detection/dataset/annotation.py),
    imported from /root/reference with `lxml` stubbed out, for coords_int / bounds, and the
    slice expression of ref:miso/object_detection/crop.py:30 for the crop pixels.
The fixtures hold inputs and outputs, so neither /root/reference nor torchvision is needed to
check the oracle or the CUDA path against them.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torchvision
from torchvision.models.detection._utils import BoxCoder
from torchvision.models.detection.anchor_utils import AnchorGenerator
from torchvision.models.detection.image_list import ImageList
from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead, concat_box_prediction_layers
from torchvision.models.detection.roi_heads import RoIHeads
from torchvision.models.detection.transform import resize_boxes
from torchvision.ops import boxes as tvb
from torchvision.ops.poolers import MultiScaleRoIAlign

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tests import cases  # noqa: E402

OUT = os.path.join(HERE, "golden")
T = torch.from_numpy


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(f"{name}: {sum(a.nbytes for a in arrays.values()) / 1e6:.2f} MB raw")


def gen_nms():
    d = {}
    for name, b, s, thr in cases.nms_cases():
        d[name + "/keep"] = torch.ops.torchvision.nms(T(b), T(s), thr).numpy()
    save("nms", **d)
    d = {}
    for name, b, s, idx, thr in cases.batched_nms_cases():
        d[name + "/vanilla"] = tvb._batched_nms_vanilla(T(b), T(s), T(idx), thr).numpy()
        d[name + "/trick"] = tvb._batched_nms_coordinate_trick(T(b), T(s), T(idx), thr).numpy()
    save("batched_nms", **d)


def gen_roi_align():
    d = {}
    for name, x, rois, scale, P, sr, aligned in cases.roi_align_cases():
        # channels 0, 13 and 39 of 40 keep the fixture small; the planes are independent
        d[name] = torchvision.ops.roi_align(T(x), T(rois), P, scale, sr, aligned).numpy()[:, cases.ROI_GOLDEN_CHANNELS]
    save("roi_align", **d)
    feats, boxes, shapes = cases.multiscale_case()
    for P, tag in ((7, "box7"), (14, "mask14")):
        pool = MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)
        x = {str(i): T(f) for i, f in enumerate(feats)}
        out = pool(x, [T(b) for b in boxes], shapes).numpy()
        lv = pool.map_levels([T(b) for b in boxes]).numpy()
        save("multiscale_" + tag, out=out[:, ::4], levels=lv, scales=np.array(pool.scales, np.float64))


def gen_boxes():
    rng = np.random.default_rng(5)
    b = cases.random_boxes(rng, 2000, extent=700.0, wh=(1.0, 400.0))
    rel = (rng.standard_normal((2000, 12)) * 2).astype(np.float32)
    rel[::7] *= 10
    d = {"boxes": b, "rel": rel}
    d["decode_rpn"] = BoxCoder((1.0, 1.0, 1.0, 1.0)).decode(T(rel[:, :4].copy()), [T(b)]).numpy()
    d["decode_roi"] = BoxCoder((10.0, 10.0, 5.0, 5.0)).decode(T(rel), [T(b)]).numpy()
    wild = (b + rng.uniform(-300, 300, b.shape)).astype(np.float32)
    d["wild"] = wild
    d["clip"] = tvb.clip_boxes_to_image(T(wild), (480, 640)).numpy()
    d["small_1e-3"] = tvb.remove_small_boxes(T(d["clip"]), 1e-3).numpy()
    d["small_20"] = tvb.remove_small_boxes(T(d["clip"]), 20.0).numpy()
    for a in ("xyxy", "xywh", "cxcywh"):
        for c in ("xyxy", "xywh", "cxcywh"):
            d[f"convert_{a}_{c}"] = tvb.box_convert(T(b), a, c).numpy()
    d["resize"] = resize_boxes(T(b), [800, 800], [1024, 1024]).numpy()
    d["resize2"] = resize_boxes(T(b), [800, 1216], [683, 1024]).numpy()
    save("boxes", **d)
    # anchors
    ag = AnchorGenerator(cases.RPN_SIZES, cases.RPN_RATIOS)
    grids = [(56, 72), (28, 36), (14, 18), (7, 9), (4, 5)]
    il = ImageList(torch.zeros(1, 3, 224, 288), [(224, 288)])
    anchors = ag(il, [torch.zeros(1, 1, *g) for g in grids])[0].numpy()
    save("anchors", anchors=anchors, grids=np.array(grids), cell=np.stack([c.numpy() for c in ag.cell_anchors]))


def gen_rpn():
    obj, dlt, grids, image_sizes, padded = cases.rpn_case()
    n = obj[0].shape[0]
    ag = AnchorGenerator(cases.RPN_SIZES, cases.RPN_RATIOS)
    rpn = RegionProposalNetwork(ag, RPNHead(8, 3), 0.7, 0.3, 256, 0.5,
                                dict(training=2000, testing=300), dict(training=2000, testing=200), 0.7)
    rpn.eval()
    il = ImageList(torch.zeros(n, 3, *padded), image_sizes)
    anchors = ag(il, [T(o) for o in obj])
    num_anchors_per_level = [o.shape[1] * o.shape[2] * o.shape[3] for o in obj]
    objectness, deltas = concat_box_prediction_layers([T(o) for o in obj], [T(d) for d in dlt])
    proposals = rpn.box_coder.decode(deltas.detach(), anchors).view(n, -1, 4)
    top_idx = rpn._get_top_n_idx(objectness.reshape(n, -1), num_anchors_per_level)
    boxes, scores = rpn.filter_proposals(proposals, objectness, image_sizes, num_anchors_per_level)
    d = {"top_idx": top_idx.numpy(), "decoded": proposals.numpy()}
    for i, (b, s) in enumerate(zip(boxes, scores)):
        d[f"boxes{i}"] = b.numpy(); d[f"scores{i}"] = s.numpy()
    save("rpn", **d)


def gen_det():
    logits, reg, proposals, shapes = cases.det_case()
    rh = RoIHeads(None, None, None, 0.5, 0.5, 512, 0.25, (10.0, 10.0, 5.0, 5.0), 0.05, 0.5, 100)
    boxes, scores, labels = rh.postprocess_detections(T(logits), T(reg), [T(p) for p in proposals], shapes)
    d = {}
    for i, (b, s, l) in enumerate(zip(boxes, scores, labels)):
        d[f"boxes{i}"] = b.numpy(); d[f"scores{i}"] = s.numpy(); d[f"labels{i}"] = l.numpy()
        d[f"resized{i}"] = resize_boxes(b, list(shapes[i]), [287, 369]).numpy()
    save("det", **d)


def gen_crop():
    # the reference's own annotation class, with its lxml import stubbed (lxml is absent here)
    lxml = types.ModuleType("lxml")
    lxml.etree = types.ModuleType("lxml.etree")
    sys.modules.setdefault("lxml", lxml)
    sys.modules.setdefault("lxml.etree", lxml.etree)
    sys.path.insert(0, "/root/reference")
    from miso.object_detection.dataset.annotation import RectangleAnnotation

    for tag, ch in (("rgb", 3), ("gray", 1)):
        img, boxes, scores, labels = cases.crop_case(channels=ch)
        thr = 0.5
        res = {"boxes": T(boxes), "scores": T(scores), "labels": T(labels)}
        # ref:miso/object_detection/inference.py:53-62
        kb = res["boxes"][res["scores"] > thr].cpu().numpy()
        kl = res["labels"][res["scores"] > thr].cpu().numpy()
        coords, bounds, pix, sizes = [], [], [], []
        for box, _ in zip(kb, kl):
            ann = RectangleAnnotation(box[0], box[1], box[2] - box[0], box[3] - box[1], "x")
            c = ann.coords_int
            crop = img[c[1]:c[3], c[0]:c[2], ...]       # ref:miso/object_detection/crop.py:30
            coords.append(c); bounds.append(ann.bounds)
            sizes.append(crop.shape[:2]); pix.append(crop.reshape(-1))
        save("crop_" + tag, kept_boxes=kb, kept_labels=kl, coords=np.array(coords, np.int64),
             bounds=np.array(bounds, np.float32), sizes=np.array(sizes, np.int64),
             pixels=np.concatenate(pix) if pix else np.zeros(0, np.uint8))


def gen_paste():
    from torchvision.models.detection.roi_heads import paste_masks_in_image
    masks, boxes, hw = cases.paste_case()
    out = paste_masks_in_image(T(masks), T(boxes), hw, padding=1).numpy()
    save("paste_masks", out=out)


def gen_transform():
    from torchvision.models.detection.transform import GeneralizedRCNNTransform
    imgs, mn, mx, mean, std = cases.transform_case()
    tr = GeneralizedRCNNTransform(mn, mx, mean, std).eval()
    il, _ = tr([T(a).permute(2, 0, 1).to(torch.float32) / 255 for a in imgs])      # ToTensor + transform
    save("transform", batch=il.tensors.numpy(), sizes=np.array(il.image_sizes, np.int64))


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(4)
    gen_nms(); gen_roi_align(); gen_boxes(); gen_rpn(); gen_det(); gen_crop(); gen_paste(); gen_transform()
