"""Tiled mosaic path (BASELINE config 5): overlapping tiles sharded over the GPUs of one box,
one NCCL all-gather of fixed-size detection blocks, then a cross-tile "seam" NMS.

The reference has no counterpart (ref:miso/object_detection/inference.py:86-88 feeds whole
images, which the model's transform shrinks to <= 1333 px); the semantics are DEFINED as the
single-process composition of reference operations:

    for every tile, row-major: detections of the tile (reference path) -> add the tile origin in
    fp32 -> concatenate -> torchvision.ops.boxes._batched_nms_vanilla(boxes, scores, labels, iou)
    -> miso's `score > threshold` filter (applied before the exchange; greedy NMS only lets
    higher-scored boxes suppress lower-scored ones, so this does not change the result).

Sharding: tiles in row-major order, contiguous blocks per rank, so concatenating the rank blocks
in rank order reproduces the single-GPU tile order (deterministic tie-breaking for any world size).
The only collective is one all_gather_into_tensor of [tiles_per_rank_max * dpi, 6] fp32 rows.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor


def tile_starts(extent: int, tile: int, overlap: int) -> List[int]:
    """Starts 0, stride, 2*stride, ... plus a clamped last tile (16384/1024/128 -> 19 starts)."""
    if extent <= tile:
        return [0]
    stride = tile - overlap
    starts = list(range(0, extent - tile, stride))
    starts.append(extent - tile)
    return starts


def tile_grid(height: int, width: int, tile: int, overlap: int) -> List[Tuple[int, int]]:
    """Row-major (y, x) origins."""
    return [(y, x) for y in tile_starts(height, tile, overlap) for x in tile_starts(width, tile, overlap)]


def rank_tiles(num_tiles: int, world: int, rank: int) -> range:
    """Contiguous block of rank r: [floor(T*r/G), floor(T*(r+1)/G))."""
    return range(num_tiles * rank // world, num_tiles * (rank + 1) // world)


def tiles_per_rank_max(num_tiles: int, world: int) -> int:
    return max(len(rank_tiles(num_tiles, world, r)) for r in range(world))


def pack_block(det_boxes: Tensor, det_scores: Tensor, det_labels: Tensor, det_counts: Tensor, origins: Tensor,
               threshold: float, rows: int, out: Optional[Tensor] = None) -> Tensor:
    """[T, dpi, *] detections of this rank's tiles -> one fixed-size block [rows, 6] =
    (x1, y1, x2, y2, score, label) in mosaic coordinates. Rows that hold no detection, or one
    that fails `score > threshold`, get label -1. origins: [T, 2] = (y, x) as fp32 (exact).
    CUDA tensors: one mb_mosaic_pack launch (into `out` if given); CPU tensors (the gloo tests of
    the host logic): the same arithmetic with tensor operations."""
    t, dpi = det_scores.shape
    if det_boxes.is_cuda:
        from . import _lib
        from .ops import _ptr, _stream
        if out is None:
            out = torch.empty((rows, 6), dtype=torch.float32, device=det_boxes.device)
        for x in (det_boxes, det_scores, det_labels, det_counts, origins, out):
            torch._assert(x.is_contiguous(), "pack_block: contiguous tensors expected")
        torch._assert(det_labels.dtype == torch.int64 and det_counts.dtype == torch.int32 and out.shape == (rows, 6),
                      "pack_block: int64 labels, int32 counts, [rows, 6] output expected")
        _lib.check(_lib.load().mb_mosaic_pack(_ptr(det_boxes), _ptr(det_scores), _ptr(det_labels), _ptr(det_counts),
                                              _ptr(origins), int(t), int(dpi), float(threshold), int(rows), _ptr(out),
                                              _stream(det_boxes)), "mb_mosaic_pack")
        return out
    idx = torch.arange(dpi, device=det_scores.device)[None, :]
    live = (idx < det_counts[:, None]) & (det_scores > threshold)
    off = torch.stack([origins[:, 1], origins[:, 0], origins[:, 1], origins[:, 0]], dim=1)[:, None, :]
    boxes = det_boxes + off                      # fp32 add, rounded once (origins are exact integers)
    lab = torch.where(live, det_labels.to(torch.float32), torch.full_like(det_scores, -1.0))
    block = torch.cat([boxes, det_scores[..., None], lab[..., None]], dim=2).reshape(t * dpi, 6)
    if block.shape[0] < rows:
        pad = torch.zeros((rows - block.shape[0], 6), dtype=block.dtype, device=block.device)
        pad[:, 5] = -1.0
        block = torch.cat([block, pad], dim=0)
    return block.contiguous()


def exchange(block: Tensor, world: int, group=None, out: Optional[Tensor] = None) -> Tensor:
    """The path's one collective: all-gather of the per-rank blocks, rank order preserved."""
    if world == 1:
        return block
    import torch.distributed as dist
    if out is None:
        out = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block, group=group)
    return out


class SeamNms:
    """Sync-free seam NMS for CUDA tensors: one mb_mosaic_unpack launch plus one prepared mb_nms
    launch sequence over ALL gathered rows (padding rows carry label -1 and are ignored on the
    device), enqueued on the same stream right behind the all-gather; every buffer is allocated
    once. `finish()` reads the count (the step's only host sync)."""

    def __init__(self, rows: int, num_classes: int, device):
        from .ops import PreparedBatchedNms
        self.nms = PreparedBatchedNms(rows, num_classes, device)
        self.rows = int(rows)
        self.boxes = torch.empty((rows, 4), dtype=torch.float32, device=device)
        self.scores = torch.empty((rows,), dtype=torch.float32, device=device)
        self.labels = torch.empty((rows,), dtype=torch.int64, device=device)
        self.gathered = None

    def launch(self, gathered: Tensor, iou_threshold: float):
        from . import _lib
        from .ops import _ptr, _stream
        torch._assert(gathered.is_contiguous() and gathered.shape == (self.rows, 6), "SeamNms: [rows, 6] contiguous block expected")
        self.gathered = gathered
        _lib.check(_lib.load().mb_mosaic_unpack(_ptr(gathered), self.rows, _ptr(self.boxes), _ptr(self.scores),
                                                _ptr(self.labels), _stream(gathered)), "mb_mosaic_unpack")
        return self.nms(self.boxes, self.scores, self.labels, iou_threshold)

    def finish(self):
        keep, status = self.nms.keep, self.nms.status
        n = int(status[0])
        keep = keep[:n]
        return self.boxes[keep], self.scores[keep], self.labels[keep]


def seam_nms(gathered: Tensor, iou_threshold: float, nms_fn: Optional[Callable] = None):
    """Cross-tile NMS over every live row, always the per-class ("vanilla") strategy on raw
    mosaic coordinates. Returns (boxes, scores, labels) in (score desc, gathered order asc) order."""
    if nms_fn is None:
        from .ops import _batched_nms_vanilla as nms_fn
    live = gathered[:, 5] >= 0
    rows = gathered[live]                        # compaction: the step's one host sync
    boxes, scores, labels = rows[:, :4].contiguous(), rows[:, 4].contiguous(), rows[:, 5].to(torch.int64)
    keep = nms_fn(boxes, scores, labels, iou_threshold)
    return boxes[keep], scores[keep], labels[keep]


def infer_mosaic(model, mosaic_u8: Tensor, tile: int = 1024, overlap: int = 128, threshold: float = 0.5,
                 batch_size: int = 4, rank: int = 0, world: int = 1, crops: bool = True):
    """Detect on a large slide/mosaic image (config 5): cut it into overlapping tiles, run this
    rank's contiguous share of tiles through the (patched) model, exchange the per-rank detection
    blocks with one all-gather, run the seam NMS and cut the crops of the surviving detections.

    mosaic_u8: uint8 [H, W, C] on the device (every rank holds the pixels it needs; here the whole
    mosaic). Returns (boxes [M,4] in mosaic coordinates, scores [M], labels [M], crops) where crops
    is the CropOutput of this rank's share (detections rank::world) or None."""
    from . import detection
    dev = mosaic_u8.device
    h, w = int(mosaic_u8.shape[0]), int(mosaic_u8.shape[1])
    grid = tile_grid(h, w, tile, overlap)
    mine = list(rank_tiles(len(grid), world, rank))
    dpi = int(model.roi_heads.detections_per_img)
    tmax = tiles_per_rank_max(len(grid), world)
    boxes = torch.zeros((len(mine), dpi, 4), dtype=torch.float32, device=dev)
    scores = torch.zeros((len(mine), dpi), dtype=torch.float32, device=dev)
    labels = torch.zeros((len(mine), dpi), dtype=torch.int64, device=dev)
    counts = torch.zeros((len(mine),), dtype=torch.int32, device=dev)
    with torch.inference_mode():
        for i0 in range(0, len(mine), batch_size):
            idx = mine[i0:i0 + batch_size]
            tiles = [mosaic_u8[grid[t][0]:grid[t][0] + tile, grid[t][1]:grid[t][1] + tile] for t in idx]
            if getattr(model, "_miso_b200_patched", False) and mosaic_u8.dim() == 3 and mosaic_u8.shape[2] <= 4:
                from .patch import forward_uint8
                res = forward_uint8(model, tiles)        # ToTensor + normalize + resize + batch in one kernel
            else:
                res = model([t.permute(2, 0, 1).to(torch.float32) / 255 for t in tiles])
            for j, r in enumerate(res):
                k = int(r["boxes"].shape[0])
                boxes[i0 + j, :k], scores[i0 + j, :k], labels[i0 + j, :k] = r["boxes"], r["scores"], r["labels"]
                counts[i0 + j] = k
    origins = torch.tensor([[float(grid[t][0]), float(grid[t][1])] for t in mine], dtype=torch.float32, device=dev).reshape(-1, 2)
    block = pack_block(boxes, scores, labels, counts, origins, threshold, tmax * dpi)
    gathered = exchange(block, world)
    num_classes = int(model.roi_heads.box_predictor.cls_score.out_features)
    seam = SeamNms(gathered.shape[0], num_classes, dev)
    seam.launch(gathered, float(model.roi_heads.nms_thresh))
    fb, fs, fl = seam.finish()
    out_crops = None
    if crops:
        share = torch.arange(rank, fb.shape[0], world, device=dev)
        sb = fb[share]
        if sb.shape[0] > 0:
            out_crops = detection.filter_and_crop([mosaic_u8], sb[None].contiguous(), torch.ones((1, sb.shape[0]), device=dev),
                                                  torch.tensor([sb.shape[0]], dtype=torch.int32, device=dev), 0.5)
    return fb, fs, fl, out_crops
