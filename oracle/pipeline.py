"""CPU composition of the whole post-head hot path out of the oracle's restated stages.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for tests/ and smoke(), and the
timed `cpu_baseline` ("port") leg of bench.py. Mirrors SURVEY.md §3.2 between the CNN heads and
the crops, taking the cuDNN/cuBLAS parts (head outputs, features, box-head outputs) as inputs.
"""
from __future__ import annotations

import time

import numpy as np

from . import detection as D
from . import miso_path as M


def run(obj, dlt, features, class_logits, box_regression, images, *, padded_image_size, image_sizes,
        original_image_sizes, sizes, aspect_ratios, pre_nms_top_n=1000, post_nms_top_n=1000, rpn_nms_thresh=0.7,
        rpn_score_thresh=0.0, rpn_min_size=1e-3, pooled=7, sampling_ratio=2, box_score_thresh=0.05,
        box_nms_thresh=0.5, detections_per_img=100, threshold=0.5, device_rule="cpu", timings=None):
    """obj/dlt: RPN head outputs per level (numpy NCHW); features: pooler levels; class_logits
    [N*R, C] and box_regression [N*R, 4C] are indexed by proposal SLOT (image n, slot r -> row
    n*R + r with R = post_nms_top_n). Returns per image a dict with proposals, box features,
    detections and the filtered crops."""
    t0 = time.perf_counter()
    n = obj[0].shape[0]
    grids = [o.shape[-2:] for o in obj]
    objectness, deltas = D.concat_rpn_head_outputs(obj, dlt)
    anchors = D.grid_anchors(padded_image_size, grids, sizes, aspect_ratios)
    decoded = np.stack([D.decode_boxes(deltas[i], anchors)[:, 0] for i in range(n)])
    napl = [o.shape[1] * o.shape[2] * o.shape[3] for o in obj]
    props, pscores = D.filter_proposals(decoded, objectness, image_sizes, napl, pre_nms_top_n=pre_nms_top_n,
                                        post_nms_top_n=post_nms_top_n, nms_thresh=rpn_nms_thresh,
                                        score_thresh=rpn_score_thresh, min_size=rpn_min_size, device_rule=device_rule)
    t1 = time.perf_counter()
    feats = D.multiscale_roi_align(features, props, image_sizes, pooled, sampling_ratio)
    t2 = time.perf_counter()
    R = post_nms_top_n
    lg = np.concatenate([class_logits[i * R:i * R + len(p)] for i, p in enumerate(props)])
    rg = np.concatenate([box_regression[i * R:i * R + len(p)] for i, p in enumerate(props)])
    dets = D.postprocess_detections(lg, rg, props, image_sizes, score_thresh=box_score_thresh,
                                    nms_thresh=box_nms_thresh, detections_per_img=detections_per_img,
                                    device_rule=device_rule)
    out, off = [], 0
    for i in range(n):
        b, s, l = dets[i]
        rb = D.resize_boxes(b, image_sizes[i], original_image_sizes[i])
        kb, ks, kl, xywh, ci, crops = M.filter_and_crop(images[i], rb, s, l, threshold)
        out.append({"proposals": props[i], "proposal_scores": pscores[i],
                    "box_features": feats[off:off + len(props[i])], "boxes": rb, "boxes_net": b, "scores": s,
                    "labels": l, "kept_labels": kl, "kept_scores": ks, "xywh": xywh, "coords": ci, "crops": crops})
        off += len(props[i])
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(rpn_s=t1 - t0, roi_align_s=t2 - t1, det_crop_s=t3 - t2, total_s=t3 - t0)
    return out
