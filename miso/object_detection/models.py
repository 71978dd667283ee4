"""Model factory with the reference's two entry points (ref:miso/object_detection/models.py:7-25):
a torchvision ResNet-50-FPN detector whose prediction heads are re-sized for the project's label set.

`weights` defaults to None because this build has no network access (random initialisation); pass
the torchvision weights enums to start from the COCO checkpoints as the reference does. The
detectors are built by name from one table so that both entry points share the head surgery.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torchvision.models.detection as tvd
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor
from torchvision.models.detection.mask_rcnn import MaskRCNNPredictor

MASK_HIDDEN_CHANNELS = 256
# name -> (torchvision constructor, extra constructor arguments, has a mask head)
_ARCHITECTURES: Dict[str, Tuple[Callable, dict, bool]] = {
    "fasterrcnn_resnet50": (tvd.fasterrcnn_resnet50_fpn, {"box_detections_per_img": 300}, False),
    "maskrcnn_resnet50": (tvd.maskrcnn_resnet50_fpn, {}, True),
}


def _build(model_name: str, num_classes: int, weights, weights_backbone, want_masks: bool):
    if model_name not in _ARCHITECTURES or _ARCHITECTURES[model_name][2] != want_masks:
        return None                      # the reference returns None for a name the entry point does not know
    ctor, extra, with_masks = _ARCHITECTURES[model_name]
    detector = ctor(weights=weights, weights_backbone=weights_backbone, **extra)
    heads = detector.roi_heads
    heads.box_predictor = FastRCNNPredictor(heads.box_predictor.cls_score.in_features, num_classes)
    if with_masks:
        heads.mask_predictor = MaskRCNNPredictor(heads.mask_predictor.conv5_mask.in_channels, MASK_HIDDEN_CHANNELS,
                                                 num_classes)
    return detector


def get_object_detection_model(num_classes, model_name="fasterrcnn_resnet50", weights=None, weights_backbone=None):
    """Faster R-CNN R50-FPN, 300 detections per image, box predictor for `num_classes` (background included)."""
    return _build(model_name, num_classes, weights, weights_backbone, want_masks=False)


def get_instance_segmentation_model(num_classes, model_name="maskrcnn_resnet50", weights=None, weights_backbone=None):
    """Mask R-CNN R50-FPN with box and mask predictors for `num_classes` (background included)."""
    return _build(model_name, num_classes, weights, weights_backbone, want_masks=True)
