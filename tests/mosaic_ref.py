"""CPU restatement of the tiled-mosaic composition for the tests: block packing, the seam NMS as
torchvision's per-label strategy over all live rows, miso's rounding and the crop slice from the mosaic array.
TEST INFRASTRUCTURE ONLY — the product (miso_b200/mosaic.py) has no CPU branch."""
from __future__ import annotations

import numpy as np

from oracle import detection as D
from oracle import miso_path as M

F = np.float32


def pack_block(det_boxes, det_scores, det_labels, det_counts, origins, threshold, rows):
    """[T, dpi, *] numpy detections -> [rows, 6] block; dead / filtered / padding rows get label -1."""
    t, dpi = det_scores.shape
    live = (np.arange(dpi)[None, :] < det_counts[:, None]) & (det_scores > F(threshold))
    off = np.stack([origins[:, 1], origins[:, 0], origins[:, 1], origins[:, 0]], axis=1).astype(F)[:, None, :]
    boxes = (det_boxes.astype(F) + off).astype(F)            # one fp32 add per coordinate
    lab = np.where(live, det_labels.astype(F), F(-1.0))
    block = np.concatenate([boxes, det_scores[..., None].astype(F), lab[..., None].astype(F)], axis=2).reshape(t * dpi, 6)
    if block.shape[0] < rows:
        pad = np.zeros((rows - block.shape[0], 6), F)
        pad[:, 5] = -1.0
        block = np.concatenate([block, pad], axis=0)
    return np.ascontiguousarray(block, dtype=F)


def seam_keep_rows(block, iou_threshold):
    """Rows kept by _batched_nms_vanilla over the live rows of a gathered block, ascending."""
    live = np.nonzero(block[:, 5] >= 0)[0]
    if live.size == 0:
        return live
    keep = D.batched_nms_vanilla(block[live, :4], block[live, 4], block[live, 5].astype(np.int64), iou_threshold)
    return np.sort(live[keep])


def crops_of_rows(mosaic, block, rows):
    """miso's annotation + coords_int + slice (ref:miso/object_detection/inference.py:56-60,
    dataset/annotation.py:120-127, crop.py:28-30) on mosaic coordinates for the given rows."""
    xywh = M.annotations_xywh(block[rows, :4])
    ci = M.coords_int(xywh)
    return xywh, ci, [M.crop(mosaic, c) for c in ci]
