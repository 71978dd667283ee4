"""Model factory (ref:miso/object_detection/models.py:7-25). `weights` defaults to None because
this build has no network access; pass the torchvision weights enum to reproduce the reference."""
from torchvision.models.detection import maskrcnn_resnet50_fpn
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor, fasterrcnn_resnet50_fpn
from torchvision.models.detection.mask_rcnn import MaskRCNNPredictor


def get_object_detection_model(num_classes, model_name="fasterrcnn_resnet50", weights=None, weights_backbone=None):
    if model_name == "fasterrcnn_resnet50":
        model = fasterrcnn_resnet50_fpn(weights=weights, weights_backbone=weights_backbone, box_detections_per_img=300)
        in_features = model.roi_heads.box_predictor.cls_score.in_features
        model.roi_heads.box_predictor = FastRCNNPredictor(in_features, num_classes)
        return model


def get_instance_segmentation_model(num_classes, model_name="maskrcnn_resnet50", weights=None, weights_backbone=None):
    if model_name == "maskrcnn_resnet50":
        model = maskrcnn_resnet50_fpn(weights=weights, weights_backbone=weights_backbone)
        in_features = model.roi_heads.box_predictor.cls_score.in_features
        model.roi_heads.box_predictor = FastRCNNPredictor(in_features, num_classes)
        in_features_mask = model.roi_heads.mask_predictor.conv5_mask.in_channels
        model.roi_heads.mask_predictor = MaskRCNNPredictor(in_features_mask, 256, num_classes)
        return model
