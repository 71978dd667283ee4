"""numpy restatement of the tensor-level stages of the torchvision detection hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Every function cites the reference
code it follows. All arithmetic is carried in np.float32 with one rounding per operation,
mirroring the fp32 aten elementwise ops (no FMA on the CPU path).
"""
from __future__ import annotations

import math

import numpy as np

from . import native

F = np.float32


# --------------------------------------------------------------------------------------
# a1  AnchorGenerator  (tv:models/detection/anchor_utils.py:58-74, :84-113, :115-133)
# --------------------------------------------------------------------------------------
def base_anchors(scales, aspect_ratios) -> np.ndarray:
    """generate_anchors: round([-ws,-hs,ws,hs]/2), round-half-even (anchor_utils.py:58-74)."""
    scales = np.asarray(scales, dtype=F)
    ar = np.asarray(aspect_ratios, dtype=F)
    h_ratios = np.sqrt(ar).astype(F)
    w_ratios = (F(1) / h_ratios).astype(F)
    ws = (w_ratios[:, None] * scales[None, :]).astype(F).reshape(-1)
    hs = (h_ratios[:, None] * scales[None, :]).astype(F).reshape(-1)
    base = (np.stack([-ws, -hs, ws, hs], axis=1) / F(2)).astype(F)
    return np.round(base).astype(F)          # np.round is half-to-even like torch.round


def grid_anchors(padded_image_size, grid_sizes, sizes, aspect_ratios) -> np.ndarray:
    """grid_anchors + forward: anchors for ONE image, all levels concatenated.
    Index order inside a level is (h*W + w)*A + a (anchor_utils.py:100-113)."""
    out = []
    for (gh, gw), sc, ar in zip(grid_sizes, sizes, aspect_ratios):
        sh = padded_image_size[0] // gh
        sw = padded_image_size[1] // gw
        base = base_anchors(sc, ar)
        sx = (np.arange(gw, dtype=np.int32) * np.int32(sw))
        sy = (np.arange(gh, dtype=np.int32) * np.int32(sh))
        yy, xx = np.meshgrid(sy, sx, indexing="ij")
        shifts = np.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], axis=1)
        a = (shifts.reshape(-1, 1, 4).astype(F) + base.reshape(1, -1, 4)).astype(F)
        out.append(a.reshape(-1, 4))
    return np.concatenate(out, axis=0)


# --------------------------------------------------------------------------------------
# a2  concat_box_prediction_layers  (tv:models/detection/rpn.py:81-110)
# --------------------------------------------------------------------------------------
def concat_rpn_head_outputs(objectness_levels, delta_levels):
    """[N,A,H,W] / [N,4A,H,W] per level -> objectness [N, sumA], deltas [N, sumA, 4]
    in the reference's (h, w, a) anchor order."""
    obj, dlt = [], []
    for o, d in zip(objectness_levels, delta_levels):
        n, a, h, w = o.shape
        obj.append(np.transpose(o.reshape(n, a, 1, h, w), (0, 3, 4, 1, 2)).reshape(n, -1))
        dlt.append(np.transpose(d.reshape(n, a, 4, h, w), (0, 3, 4, 1, 2)).reshape(n, -1, 4))
    return np.concatenate(obj, axis=1).astype(F), np.concatenate(dlt, axis=1).astype(F)


# --------------------------------------------------------------------------------------
# a3  BoxCoder.decode  (tv:models/detection/_utils.py:162-224)
# --------------------------------------------------------------------------------------
BBOX_XFORM_CLIP = math.log(1000.0 / 16)


def decode_boxes(rel_codes, boxes, weights=(1.0, 1.0, 1.0, 1.0), clip=BBOX_XFORM_CLIP) -> np.ndarray:
    """rel_codes [M, 4*C], boxes [M, 4] -> [M, C, 4] (decode_single, _utils.py:183-224)."""
    rel = np.asarray(rel_codes, dtype=F).reshape(len(boxes), -1)
    b = np.asarray(boxes, dtype=F)
    widths = (b[:, 2] - b[:, 0]).astype(F)
    heights = (b[:, 3] - b[:, 1]).astype(F)
    ctr_x = (b[:, 0] + (F(0.5) * widths).astype(F)).astype(F)
    ctr_y = (b[:, 1] + (F(0.5) * heights).astype(F)).astype(F)
    wx, wy, ww, wh = (F(v) for v in weights)
    dx = (rel[:, 0::4] / wx).astype(F)
    dy = (rel[:, 1::4] / wy).astype(F)
    dw = np.minimum((rel[:, 2::4] / ww).astype(F), F(clip))
    dh = np.minimum((rel[:, 3::4] / wh).astype(F), F(clip))
    pcx = ((dx * widths[:, None]).astype(F) + ctr_x[:, None]).astype(F)
    pcy = ((dy * heights[:, None]).astype(F) + ctr_y[:, None]).astype(F)
    pw = (np.exp(dw).astype(F) * widths[:, None]).astype(F)
    ph = (np.exp(dh).astype(F) * heights[:, None]).astype(F)
    cw = (F(0.5) * pw).astype(F)
    ch = (F(0.5) * ph).astype(F)
    return np.stack([pcx - cw, pcy - ch, pcx + cw, pcy + ch], axis=2).astype(F)


# --------------------------------------------------------------------------------------
# a6  clip_boxes_to_image / remove_small_boxes  (tv:ops/boxes.py:149-182, :123-146)
# --------------------------------------------------------------------------------------
def clip_boxes_to_image(boxes, size) -> np.ndarray:
    b = np.array(boxes, dtype=F, copy=True)
    h, w = size
    b[..., 0::2] = np.clip(b[..., 0::2], F(0), F(w))
    b[..., 1::2] = np.clip(b[..., 1::2], F(0), F(h))
    return b


def remove_small_boxes(boxes, min_size: float) -> np.ndarray:
    b = np.asarray(boxes, dtype=F)
    ws = (b[..., 2] - b[..., 0]).astype(F)
    hs = (b[..., 3] - b[..., 1]).astype(F)
    # the reference compares an fp32 tensor with a Python float: the scalar is cast to fp32
    return np.nonzero((ws >= F(min_size)) & (hs >= F(min_size)))[0].astype(np.int64)


# --------------------------------------------------------------------------------------
# a9  box_convert  (tv:ops/boxes.py:185-270, tv:ops/_box_convert.py:5-81)
# --------------------------------------------------------------------------------------
def box_convert(boxes, in_fmt: str, out_fmt: str) -> np.ndarray:
    allowed = ("xyxy", "xywh", "cxcywh")
    if in_fmt not in allowed or out_fmt not in allowed:
        raise ValueError("Unsupported Bounding Box Conversions for given in_fmt and out_fmt")
    b = np.array(boxes, dtype=F, copy=True)
    if in_fmt == out_fmt:
        return b
    if in_fmt != "xyxy" and out_fmt != "xyxy":
        b = box_convert(b, in_fmt, "xyxy")
        in_fmt = "xyxy"
    a0, a1, a2, a3 = (b[..., i] for i in range(4))
    if in_fmt == "xyxy" and out_fmt == "xywh":
        r = [a0, a1, a2 - a0, a3 - a1]
    elif in_fmt == "xywh":
        r = [a0, a1, a0 + a2, a1 + a3]
    elif in_fmt == "xyxy" and out_fmt == "cxcywh":
        r = [(a0 + a2) / F(2), (a1 + a3) / F(2), a2 - a0, a3 - a1]
    else:  # cxcywh -> xyxy
        r = [a0 - F(0.5) * a2, a1 - F(0.5) * a3, a0 + F(0.5) * a2, a1 + F(0.5) * a3]
    return np.stack(r, axis=-1).astype(F)


# --------------------------------------------------------------------------------------
# a4  per-level top-k  (tv:models/detection/rpn.py:231-240)
# --------------------------------------------------------------------------------------
def topk_desc(values, k: int) -> np.ndarray:
    """Indices of the k largest values, ordered (value desc, index asc); NaN is largest.
    torch.topk's tie order on CPU is unspecified, so exact index equality
    with torch is only defined on tie-free inputs; this is the order the CUDA path defines."""
    return native.argsort_desc_stable(values)[:k]


def rpn_top_n_idx(objectness, num_anchors_per_level, pre_nms_top_n: int) -> np.ndarray:
    """_get_top_n_idx for one image: objectness [sumA] -> indices [sum min(k, A_l)]."""
    r, off = [], 0
    for a in num_anchors_per_level:
        k = min(pre_nms_top_n, a)
        r.append(topk_desc(objectness[off:off + a], k) + off)
        off += a
    return np.concatenate(r)


# --------------------------------------------------------------------------------------
# a7/a8  nms / batched_nms  (tv:ops/boxes.py:20-120)
# --------------------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold: float) -> np.ndarray:
    return native.nms(boxes, scores, iou_threshold)


def batched_nms_coordinate_trick(boxes, scores, idxs, iou_threshold: float) -> np.ndarray:
    """_batched_nms_coordinate_trick (boxes.py:86-103): offsets computed and added in fp32."""
    b = np.asarray(boxes, dtype=F).reshape(-1, 4)
    if b.size == 0:
        return np.empty((0,), dtype=np.int64)
    m = b.max()
    offsets = (np.asarray(idxs).astype(F) * F(m + F(1))).astype(F)
    return nms((b + offsets[:, None]).astype(F), scores, iou_threshold)


def batched_nms_vanilla(boxes, scores, idxs, iou_threshold: float) -> np.ndarray:
    """_batched_nms_vanilla (boxes.py:106-120). The reference's final sort is not stable;
    this restatement breaks score ties by ascending index (the order the CUDA path defines)."""
    b = np.asarray(boxes, dtype=F).reshape(-1, 4)
    s = np.asarray(scores, dtype=F).reshape(-1)
    idxs = np.asarray(idxs)
    keep_mask = np.zeros(s.shape[0], dtype=bool)
    for cid in np.unique(idxs):
        cur = np.nonzero(idxs == cid)[0]
        k = nms(b[cur], s[cur], iou_threshold)
        keep_mask[cur[k]] = True
    kept = np.nonzero(keep_mask)[0]
    return kept[native.argsort_desc_stable(s[kept])].astype(np.int64)


def batched_nms(boxes, scores, idxs, iou_threshold: float, device_rule: str = "cpu") -> np.ndarray:
    """batched_nms strategy switch (boxes.py:80): vanilla iff numel > 4000 (cpu) / 100000 (cuda)."""
    b = np.asarray(boxes, dtype=F).reshape(-1, 4)
    if b.size > (4000 if device_rule == "cpu" else 100_000):
        return batched_nms_vanilla(b, scores, idxs, iou_threshold)
    return batched_nms_coordinate_trick(b, scores, idxs, iou_threshold)


# --------------------------------------------------------------------------------------
# a5  RegionProposalNetwork.filter_proposals  (tv:models/detection/rpn.py:242-297)
# --------------------------------------------------------------------------------------
def sigmoid(x) -> np.ndarray:
    x = np.asarray(x, dtype=F)
    return (F(1) / (F(1) + np.exp(-x).astype(F))).astype(F)


def filter_proposals(proposals, objectness, image_shapes, num_anchors_per_level, *,
                     pre_nms_top_n=1000, post_nms_top_n=1000, nms_thresh=0.7,
                     score_thresh=0.0, min_size=1e-3, device_rule="cpu"):
    """proposals [N, sumA, 4], objectness [N, sumA] logits -> per image (boxes, scores)."""
    proposals = np.asarray(proposals, dtype=F)
    objectness = np.asarray(objectness, dtype=F).reshape(proposals.shape[0], -1)
    levels = np.concatenate([np.full(a, i, dtype=np.int64) for i, a in enumerate(num_anchors_per_level)])
    out_b, out_s = [], []
    for n in range(proposals.shape[0]):
        top = rpn_top_n_idx(objectness[n], num_anchors_per_level, pre_nms_top_n)
        boxes, scores, lvl = proposals[n][top], sigmoid(objectness[n][top]), levels[top]
        boxes = clip_boxes_to_image(boxes, image_shapes[n])
        keep = remove_small_boxes(boxes, min_size)
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = np.nonzero(scores >= F(score_thresh))[0]
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = batched_nms(boxes, scores, lvl, nms_thresh, device_rule)[:post_nms_top_n]
        out_b.append(boxes[keep]); out_s.append(scores[keep])
    return out_b, out_s


# --------------------------------------------------------------------------------------
# a10/a11  LevelMapper + MultiScaleRoIAlign  (tv:ops/poolers.py:47-84, :87-95, :147-227)
# --------------------------------------------------------------------------------------
def box_area(boxes) -> np.ndarray:
    b = np.asarray(boxes, dtype=F)
    return ((b[:, 2] - b[:, 0]).astype(F) * (b[:, 3] - b[:, 1]).astype(F)).astype(F)


def map_levels(boxes, k_min=2, k_max=5, canonical_scale=224, canonical_level=4, eps=1e-6) -> np.ndarray:
    """LevelMapper.__call__ (poolers.py:73-84). log2/sqrt here are numpy's; the pin test
    checks the result against torchvision's mapper, including ulp sweeps at the boundaries."""
    s = np.sqrt(box_area(boxes)).astype(F)
    t = np.floor((F(canonical_level) + np.log2((s / F(canonical_scale)).astype(F)).astype(F)).astype(F)
                 + F(eps)).astype(F)
    t = np.clip(t, F(k_min), F(k_max))
    return (t.astype(np.int64) - k_min).astype(np.int64)


def infer_scales(feature_shapes, image_shapes):
    """_setup_scales / _infer_scale (poolers.py:98-134): 2**round(log2(feat/img))."""
    max_h = max(s[0] for s in image_shapes)
    scales = []
    for fs in feature_shapes:
        approx = float(fs[0]) / float(max_h)
        scales.append(2 ** float(np.round(np.log2(F(approx)))))
    k_min = int(-math.log2(scales[0]))
    k_max = int(-math.log2(scales[-1]))
    return scales, k_min, k_max


def multiscale_roi_align(features, boxes_per_image, image_shapes, output_size, sampling_ratio=2,
                         canonical_scale=224, canonical_level=4) -> np.ndarray:
    """features: list of [N,C,H_l,W_l]; boxes_per_image: list of [R_i,4] -> [sumR, C, P, P]."""
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    rois = np.concatenate([np.concatenate([np.full((len(b), 1), i, dtype=F), np.asarray(b, dtype=F).reshape(-1, 4)],
                                          axis=1) for i, b in enumerate(boxes_per_image)], axis=0)
    scales, k_min, k_max = infer_scales([f.shape[-2:] for f in features], image_shapes)
    if len(features) == 1:
        return native.roi_align(features[0], rois, scales[0], ph, pw, sampling_ratio, False)
    levels = map_levels(rois[:, 1:], k_min, k_max, canonical_scale, canonical_level)
    out = np.zeros((len(rois), features[0].shape[1], ph, pw), dtype=F)
    for lvl, (f, sc) in enumerate(zip(features, scales)):
        idx = np.nonzero(levels == lvl)[0]
        out[idx] = native.roi_align(f, rois[idx], sc, ph, pw, sampling_ratio, False)
    return out


# --------------------------------------------------------------------------------------
# a12  RoIHeads.postprocess_detections  (tv:models/detection/roi_heads.py:680-737)
# --------------------------------------------------------------------------------------
def softmax_lastdim(x) -> np.ndarray:
    x = np.asarray(x, dtype=F)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp((x - m).astype(F)).astype(F)
    # aten's CPU softmax accumulates the row sum in fp32, left to right
    ssum = np.zeros(e.shape[:-1] + (1,), dtype=F)
    for c in range(e.shape[-1]):
        ssum = (ssum + e[..., c:c + 1]).astype(F)
    return (e / ssum).astype(F)


def postprocess_detections(class_logits, box_regression, proposals, image_shapes, *,
                           box_weights=(10.0, 10.0, 5.0, 5.0), score_thresh=0.05, nms_thresh=0.5,
                           detections_per_img=100, device_rule="cpu"):
    class_logits = np.asarray(class_logits, dtype=F)
    num_classes = class_logits.shape[-1]
    counts = [len(p) for p in proposals]
    pred_boxes = decode_boxes(box_regression, np.concatenate(proposals, axis=0), box_weights)
    pred_scores = softmax_lastdim(class_logits)
    out, off = [], 0
    for n, cnt in enumerate(counts):
        boxes = clip_boxes_to_image(pred_boxes[off:off + cnt], image_shapes[n])
        scores = pred_scores[off:off + cnt]
        off += cnt
        labels = np.broadcast_to(np.arange(num_classes, dtype=np.int64)[None, :], scores.shape)
        boxes, scores, labels = boxes[:, 1:].reshape(-1, 4), scores[:, 1:].reshape(-1), labels[:, 1:].reshape(-1)
        inds = np.nonzero(scores > F(score_thresh))[0]
        boxes, scores, labels = boxes[inds], scores[inds], labels[inds]
        keep = remove_small_boxes(boxes, 1e-2)
        boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
        keep = batched_nms(boxes, scores, labels, nms_thresh, device_rule)[:detections_per_img]
        out.append((boxes[keep], scores[keep], labels[keep]))
    return out


# --------------------------------------------------------------------------------------
# a14  resize_boxes  (tv:models/detection/transform.py:306-319)
# --------------------------------------------------------------------------------------
def resize_boxes(boxes, original_size, new_size) -> np.ndarray:
    b = np.asarray(boxes, dtype=F).reshape(-1, 4)
    rh = F(F(new_size[0]) / F(original_size[0]))
    rw = F(F(new_size[1]) / F(original_size[1]))
    return np.stack([b[:, 0] * rw, b[:, 1] * rh, b[:, 2] * rw, b[:, 3] * rh], axis=1).astype(F)


def _fma32(a, b, c) -> np.ndarray:
    """fp32 fused multiply-add (the product of two fp32 numbers is exact in fp64; one extra rounding to fp32)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _upsample_axis(in_size: int, out_size: int):
    """ATen area_pixel_compute_source_index + guard_index_and_lambda for align_corners=False
    (aten/src/ATen/native/UpSample.h), fp32, with the scale*(i+0.5)-0.5 contraction the CPU build applies."""
    f = np.float32
    scale = f(in_size) / f(out_size)
    i = np.arange(out_size, dtype=f)
    src = _fma32(scale, i + f(0.5), f(-0.5))
    src = np.where(src < 0, f(0), src).astype(f)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    lam = np.clip(src - i0.astype(f), f(0), f(1)).astype(f)
    i1 = i0 + (i0 < in_size - 1)
    return i0, i1, (f(1) - lam).astype(f), lam


def interpolate_bilinear(mask, out_h: int, out_w: int) -> np.ndarray:
    """F.interpolate(mask[None, None], size=(h, w), mode="bilinear", align_corners=False) on the CPU
    (aten/src/ATen/native/cpu/UpSampleKernel.cpp, generic 2-d linear kernel, FMA-contracted build)."""
    m = np.asarray(mask, np.float32)
    y0, y1, wy0, wy1 = _upsample_axis(m.shape[0], out_h)
    x0, x1, wx0, wx1 = _upsample_axis(m.shape[1], out_w)
    r0, r1 = m[y0], m[y1]
    t0 = _fma32(r0[:, x0], wx0[None, :], (r0[:, x1] * wx1[None, :]).astype(np.float32))
    t1 = _fma32(r1[:, x0], wx0[None, :], (r1[:, x1] * wx1[None, :]).astype(np.float32))
    return _fma32(t0, wy0[:, None], (t1 * wy1[:, None]).astype(np.float32))


def paste_masks_in_image(masks, boxes, img_shape, padding: int = 1) -> np.ndarray:
    """tv:models/detection/roi_heads.py:375-501: expand_masks, expand_boxes (fp32) -> int64, per detection
    bilinear resize to (h, w) = box size (+1, at least 1) and paste into a zero image."""
    f = np.float32
    masks = np.asarray(masks, f)
    boxes = np.asarray(boxes, f)
    r, m = masks.shape[0], masks.shape[-1]
    im_h, im_w = int(img_shape[0]), int(img_shape[1])
    scale = f(float(m + 2 * padding) / m)
    padded = np.pad(masks[:, 0], ((0, 0), (padding, padding), (padding, padding)))
    w_half = ((boxes[:, 2] - boxes[:, 0]) * f(0.5)).astype(f) * scale
    h_half = ((boxes[:, 3] - boxes[:, 1]) * f(0.5)).astype(f) * scale
    x_c = ((boxes[:, 2] + boxes[:, 0]) * f(0.5)).astype(f)
    y_c = ((boxes[:, 3] + boxes[:, 1]) * f(0.5)).astype(f)
    exp = np.stack([x_c - w_half, y_c - h_half, x_c + w_half, y_c + h_half], axis=1).astype(f)
    ib = np.trunc(exp).astype(np.int64)
    out = np.zeros((r, 1, im_h, im_w), f)
    for i in range(r):
        bx0, by0, bx1, by1 = (int(v) for v in ib[i])
        w, h = max(bx1 - bx0 + 1, 1), max(by1 - by0 + 1, 1)
        x_0, x_1 = max(bx0, 0), min(bx1 + 1, im_w)
        y_0, y_1 = max(by0, 0), min(by1 + 1, im_h)
        if x_1 <= x_0 or y_1 <= y_0:
            continue
        resized = interpolate_bilinear(padded[i], h, w)
        out[i, 0, y_0:y_1, x_0:x_1] = resized[y_0 - by0:y_1 - by0, x_0 - bx0:x_1 - bx0]
    return out


def transform_images(images_u8, min_size: int, max_size: int, image_mean, image_std, size_divisible: int = 32):
    """ToTensor (x / 255) + GeneralizedRCNNTransform.forward in eval mode (tv:models/detection/transform.py:
    normalize :165-173, resize :175-201 via _resize_image_and_masks :23-70, batch_images :231-255) for uint8
    HWC arrays. Returns (batch [N, C, H_pad, W_pad] fp32, resized sizes)."""
    f = np.float32
    mean, std = np.asarray(image_mean, f), np.asarray(image_std, f)
    outs, sizes = [], []
    for a in images_u8:
        h, w = a.shape[:2]
        # eager (non-scripted) torchvision 0.26 computes the factor with Python ints/floats (double precision)
        sf = min(float(min_size) / min(h, w), float(max_size) / max(h, w))
        oh, ow = int(math.floor(float(h) * sf)), int(math.floor(float(w) * sf))
        x = (a.astype(f) / f(255)).astype(f)
        x = ((x - mean) / std).astype(f)
        outs.append(np.stack([interpolate_bilinear(x[..., c], oh, ow) for c in range(x.shape[2])]))
        sizes.append((oh, ow))
    d = float(size_divisible)
    ph = int(math.ceil(max(s[0] for s in sizes) / d) * d)
    pw = int(math.ceil(max(s[1] for s in sizes) / d) * d)
    batch = np.zeros((len(outs), outs[0].shape[0], ph, pw), f)
    for i, o in enumerate(outs):
        batch[i, :, : o.shape[1], : o.shape[2]] = o
    return batch, sizes
