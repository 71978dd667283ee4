"""Synthetic inputs of the BASELINE.json configurations (random-init shapes, no datasets).

Config 2 ("Faster R-CNN ResNet-50-FPN batch-4 inference, 1000 RPN proposals/image, post-proc +
RoIAlign"): 1024^2 images resized to 800^2 by the model's transform, FPN pyramid 200/100/50/25/13
x 256 channels, 3 anchors per location, 3 logits (Coccolith, Coccosphere + background), miso's
300 detections per image (ref:miso/object_detection/models.py:9). Everything upstream of the
hot path (backbone, RPN head, box head) is replaced by seeded random tensors of the right shape.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from .detection import DetConfig, RpnConfig
from .pipeline import HotPathShapes

RPN_SIZES = ((32,), (64,), (128,), (256,), (512,))
RPN_RATIOS = ((0.5, 1.0, 2.0),) * 5


@dataclass
class Workload:
    name: str
    shapes: HotPathShapes
    rpn: RpnConfig
    det: DetConfig
    threshold: float
    host: Dict[str, List[torch.Tensor]]   # pinned host tensors: objectness, deltas, features, logits, regression, images

    def input_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for ts in self.host.values() for t in ts)


def faster_rcnn_batch(num_images: int = 4, original: int = 1024, resized: int = 800, channels: int = 256,
                      num_classes: int = 3, post_nms_top_n: int = 1000, detections_per_img: int = 300,
                      threshold: float = 0.5, seed: int = 0, pin: bool = True,
                      features_layout: str = "nchw") -> Workload:
    g = torch.Generator().manual_seed(seed)
    padded = -(-resized // 32) * 32
    grids = [(padded // s, padded // s) for s in (4, 8, 16, 32)]
    pool = (-(-grids[-1][0] // 2), -(-grids[-1][1] // 2))          # LastLevelMaxPool: kernel 1, stride 2
    rpn_grids = grids + [pool]
    n = num_images

    def rn(*shape, scale=1.0, channels_last=False):
        t = torch.randn(*shape, generator=g) * scale
        if channels_last:   # same logical [N,C,H,W] tensor, NHWC strides (what a channels_last cuDNN backbone emits)
            t = t.contiguous(memory_format=torch.channels_last)
        return t.pin_memory() if pin else t

    host = {
        "objectness": [rn(n, 3, gh, gw, scale=2.0) for gh, gw in rpn_grids],
        "deltas": [rn(n, 12, gh, gw, scale=0.5) for gh, gw in rpn_grids],
        "features": [rn(n, channels, gh, gw, channels_last=(features_layout == "channels_last")) for gh, gw in grids],
        "class_logits": [rn(n * post_nms_top_n, num_classes, scale=3.0)],
        "box_regression": [rn(n * post_nms_top_n, 4 * num_classes, scale=0.5)],
        "images": [],
    }
    for _ in range(n):
        im = torch.randint(0, 256, (original, original, 3), dtype=torch.uint8, generator=g)
        host["images"].append(im.pin_memory() if pin else im)
    shapes = HotPathShapes(num_images=n, padded_image_size=(padded, padded), image_sizes=[(resized, resized)] * n,
                           original_image_sizes=[(original, original)] * n, rpn_grids=rpn_grids, feature_grids=grids,
                           channels=channels, num_classes=num_classes)
    rpn = RpnConfig(RPN_SIZES, RPN_RATIOS, pre_nms_top_n=1000, post_nms_top_n=post_nms_top_n)
    det = DetConfig(detections_per_img=detections_per_img)
    name = (f"Faster R-CNN R50-FPN post-head path, batch {n}, {original}^2 images -> {resized}^2, "
            f"{post_nms_top_n} RPN proposals/img, {channels} ch ({features_layout} FPN maps), RoIAlign 7x7, "
            f"{detections_per_img} dets/img")
    return Workload(name, shapes, rpn, det, threshold, host)


def to_device(w: Workload, device) -> Dict[str, List[torch.Tensor]]:
    # .to() keeps the memory format (channels_last stays channels_last)
    return {k: [t.to(device, non_blocking=True) for t in v] for k, v in w.host.items()}


# ------------------------------------------------------------------------------------------
# Config 5: the tiled mosaic. Every tile's head outputs are a function of the tile index alone, so any
# partition of the tiles over ranks sees the same per-tile inputs (world-size-independent results).
# ------------------------------------------------------------------------------------------
@dataclass
class MosaicWorkload:
    name: str
    height: int
    width: int
    tile: int
    overlap: int
    base: Workload                 # shapes / configs of one tile (num_images = 1)
    features_layout: str

    def shapes(self, n: int) -> HotPathShapes:
        s = self.base.shapes
        return HotPathShapes(num_images=n, padded_image_size=s.padded_image_size, image_sizes=[s.image_sizes[0]] * n,
                             original_image_sizes=[s.original_image_sizes[0]] * n, rpn_grids=s.rpn_grids,
                             feature_grids=s.feature_grids, channels=s.channels, num_classes=s.num_classes,
                             pooled=s.pooled, sampling_ratio=s.sampling_ratio, image_channels=s.image_channels)

    def tile_inputs(self, t: int, device) -> Dict[str, List[torch.Tensor]]:
        """Seeded (by tile index) head outputs of one tile, generated on `device`."""
        s = self.base.shapes
        g = torch.Generator(device=device).manual_seed(7919 * (t + 1))
        R = self.base.rpn.post_nms_top_n

        def rn(*shape, scale=1.0):
            return torch.randn(*shape, generator=g, device=device) * scale

        return {
            "objectness": [rn(1, 3, gh, gw, scale=2.0) for gh, gw in s.rpn_grids],
            "deltas": [rn(1, 12, gh, gw, scale=0.5) for gh, gw in s.rpn_grids],
            "features": [rn(1, s.channels, gh, gw) for gh, gw in s.feature_grids],
            "class_logits": [rn(R, s.num_classes, scale=3.0)],
            "box_regression": [rn(R, 4 * s.num_classes, scale=0.5)],
        }

    def batch_inputs(self, tiles: List[int], device) -> Dict[str, List[torch.Tensor]]:
        """Head outputs of a batch of tiles (concatenated per level; FPN maps in self.features_layout)."""
        per = [self.tile_inputs(t, device) for t in tiles]
        out = {k: [torch.cat([p[k][l] for p in per], dim=0).contiguous() for l in range(len(per[0][k]))] for k in per[0]}
        if self.features_layout == "channels_last":
            out["features"] = [f.contiguous(memory_format=torch.channels_last) for f in out["features"]]
        return out

    def band(self, y0: int, y1: int, device) -> torch.Tensor:
        """Mosaic pixel rows [y0, y1) as uint8 [y1-y0, W, 3]; strips of 128 rows seeded by the strip index, so
        every rank generates identical pixels for the rows it holds."""
        strips = []
        for s0 in range(y0 // 128 * 128, y1, 128):
            g = torch.Generator(device=device).manual_seed(104729 + s0 // 128)
            strip = torch.randint(0, 256, (128, self.width, 3), dtype=torch.uint8, generator=g, device=device)
            strips.append(strip[max(y0 - s0, 0):min(y1 - s0, 128)])
        return torch.cat(strips, dim=0).contiguous()

    def bytes_per_tile(self) -> int:
        s = self.base.shapes
        n = sum(15 * gh * gw for gh, gw in s.rpn_grids) + sum(s.channels * gh * gw for gh, gw in s.feature_grids)
        return 4 * (n + self.base.rpn.post_nms_top_n * 5 * s.num_classes)


def mosaic(height: int = 16384, width: int = 16384, tile: int = 1024, overlap: int = 128, resized: int = 800,
           channels: int = 256, num_classes: int = 3, post_nms_top_n: int = 1000, detections_per_img: int = 300,
           threshold: float = 0.5, features_layout: str = "channels_last") -> MosaicWorkload:
    base = faster_rcnn_batch(num_images=1, original=tile, resized=resized, channels=4, num_classes=num_classes,
                             post_nms_top_n=post_nms_top_n, detections_per_img=detections_per_img, threshold=threshold,
                             pin=False)
    base.shapes.channels = channels
    base.host = {}
    from .mosaic import tile_grid
    nt = len(tile_grid(height, width, tile, overlap))
    name = (f"config 5: {height}x{width} mosaic, {nt} tiles of {tile} px with {overlap} px overlap, post-head path per tile "
            f"({tile}^2 -> {resized}^2, {post_nms_top_n} RPN proposals, {channels} ch {features_layout} FPN maps, RoIAlign 7x7, "
            f"{detections_per_img} dets) + all-gather + seam NMS + crops")
    return MosaicWorkload(name, height, width, tile, overlap, base, features_layout)
