"""numpy restatement of miso's own part of the hot path: the final score filter, the
xyxy -> (x, y, w, h) annotation, coords_int rounding and the per-detection crop slice.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). miso.object_detection.{inference,crop}
cannot be imported in this image (lxml / scikit-image are absent), so the
~15 lines they contribute are restated here line for line.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def score_filter(boxes, scores, labels, threshold: float):
    """ref:miso/object_detection/inference.py:53-55 — `scores > threshold`, strict.
    The comparison is an fp32 tensor against a Python float, i.e. against fp32(threshold)."""
    s = np.asarray(scores, dtype=F)
    m = s > F(threshold)
    return np.asarray(boxes, dtype=F)[m], s[m], np.asarray(labels)[m]


def annotations_xywh(boxes) -> np.ndarray:
    """ref:miso/object_detection/inference.py:56-60 — RectangleAnnotation(x, y, x2-x, y2-y),
    all np.float32 scalars taken from the `.cpu().numpy()` array."""
    b = np.asarray(boxes, dtype=F).reshape(-1, 4)
    return np.stack([b[:, 0], b[:, 1], (b[:, 2] - b[:, 0]).astype(F), (b[:, 3] - b[:, 1]).astype(F)],
                    axis=1).astype(F)


def coords_int(xywh) -> np.ndarray:
    """ref:miso/object_detection/dataset/annotation.py:120-127 — coords = (x, y, x+w, y+h) in
    fp32, then int(np.round(c)) (round half to even)."""
    a = np.asarray(xywh, dtype=F).reshape(-1, 4)
    c = np.stack([a[:, 0], a[:, 1], (a[:, 0] + a[:, 2]).astype(F), (a[:, 1] + a[:, 3]).astype(F)], axis=1)
    return np.round(c).astype(np.int64)


def crop(image: np.ndarray, c) -> np.ndarray:
    """ref:miso/object_detection/crop.py:28-30 — im[c[1]:c[3], c[0]:c[2], ...] with numpy
    slice semantics (ends clamp to the image, negative indices wrap, empty if end <= start)."""
    return image[int(c[1]):int(c[3]), int(c[0]):int(c[2]), ...]


def slice_bounds(start: int, stop: int, length: int):
    """What a numpy basic slice [start:stop] resolves to on an axis of `length`
    (Python slice.indices with step 1); returns (begin, extent)."""
    b, e, _ = slice(int(start), int(stop), 1).indices(int(length))
    return b, max(e - b, 0)


def filter_and_crop(image: np.ndarray, boxes, scores, labels, threshold: float):
    """Composition used as the checker for the fused score-filter + crop kernel:
    returns (kept boxes, kept scores, kept labels, xywh, int coords, list of crop arrays)."""
    b, s, l = score_filter(boxes, scores, labels, threshold)
    xywh = annotations_xywh(b)
    ci = coords_int(xywh)
    crops = [crop(image, c) for c in ci]
    return b, s, l, xywh, ci, crops
