"""Drop-in for the dehazing -> detection hand-off of models/detection.py (reference detection.py:75-139).

`IntegratedDetectionSystem.forward` dehazes a batch and feeds the result to the detector after an ImageNet
normalisation.  The reference does that per image with `.clone().sub(cpu_tensor).div(cpu_tensor)` (detection.py:109-121:
three passes and two CPU-resident constants per image, which on a CUDA batch is a device mismatch).  Here the whole
dehazed batch is normalised by ONE libadb200 launch (`adb_image_affine`: y = x*(1/std[c]) - mean[c]/std[c], NCHW fp32,
12 B read + 12 B written per pixel) and handed to the detector as the list of per-image [3,H,W] views the reference builds.

The detector itself (torchvision Faster / Mask R-CNN, detection.py:7-73) is library code outside the hot path: it is
constructed exactly as the reference constructs it and runs as torchvision runs it.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from .. import engine as _engine

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_DETECTORS = ("faster_rcnn_resnet50_fpn", "faster_rcnn_mobilenet_v3_large_fpn", "mask_rcnn_resnet50_fpn")


def _affine3(x, scale3, shift3):
    n, _, h, w = x.shape
    out = torch.empty_like(x)
    scale = (C.c_float * 3)(*scale3)
    shift = (C.c_float * 3)(*shift3)
    _lib.call("adb_image_affine", _lib.ptr(x), n, h, w, scale, shift, _lib.ptr(out), _lib.current_stream())
    return out


class _NormalizeFn(torch.autograd.Function):
    """(x - mean[c]) / std[c]; the reference's sub / div are differentiable (detection.py:109-121), so is this: dx = g / std[c]."""

    @staticmethod
    def forward(ctx, x):
        return _affine3(x.contiguous(), [1.0 / s for s in IMAGENET_STD], [-m / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)])

    @staticmethod
    def backward(ctx, g):
        return _affine3(g.contiguous().float(), [1.0 / s for s in IMAGENET_STD], [0.0, 0.0, 0.0])


def normalize_for_detection(images):
    """(x - mean[c]) / std[c] over an NCHW fp32 CUDA batch in one launch (detection.py:109-121)."""
    _engine.require_cuda(images, "normalize_for_detection")
    if torch.is_grad_enabled() and images.requires_grad:
        return _NormalizeFn.apply(images)
    return _affine3(images.contiguous(), [1.0 / s for s in IMAGENET_STD], [-m / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)])


class DetectionModel(nn.Module):
    """torchvision detector with the reference's head replacement (detection.py:7-73).  Library code, not a libadb200 path."""

    def __init__(self, num_classes=91, model_name="faster_rcnn_resnet50_fpn", pretrained=True):
        super().__init__()
        if model_name not in _DETECTORS:
            raise ValueError(f"Unsupported detection model: {model_name}")
        import torchvision.models.detection as detection
        from torchvision.models.detection.faster_rcnn import FastRCNNPredictor
        self.model_name, self.num_classes = model_name, num_classes
        # pretrained=False must not reach for the backbone's ImageNet weights either (no network on the GPU box)
        kw = {"weights": "DEFAULT"} if pretrained else {"weights": None, "weights_backbone": None}
        if model_name == "faster_rcnn_resnet50_fpn":
            self.model = detection.fasterrcnn_resnet50_fpn(**kw)
        elif model_name == "faster_rcnn_mobilenet_v3_large_fpn":
            self.model = detection.fasterrcnn_mobilenet_v3_large_fpn(**kw)
        else:
            self.model = detection.maskrcnn_resnet50_fpn(**kw)
        in_features = self.model.roi_heads.box_predictor.cls_score.in_features
        self.model.roi_heads.box_predictor = FastRCNNPredictor(in_features, num_classes)
        if model_name == "mask_rcnn_resnet50_fpn":
            from torchvision.models.detection.mask_rcnn import MaskRCNNPredictor
            in_mask = self.model.roi_heads.mask_predictor.conv5_mask.in_channels
            self.model.roi_heads.mask_predictor = MaskRCNNPredictor(in_mask, 256, num_classes)

    def forward(self, images, targets=None):
        return self.model(images, targets) if targets is not None else self.model(images)


class IntegratedDetectionSystem(nn.Module):
    """Dehazing router + frozen detector (detection.py:75-127).  Returns (detection_results, dehazed_images)."""

    def __init__(self, dehazing_model, detection_model):
        super().__init__()
        self.dehazing_model = dehazing_model
        self.detection_model = detection_model
        for param in self.detection_model.parameters():
            param.requires_grad = False

    def forward(self, images, targets=None):
        if isinstance(images, (list, tuple)):          # the detection loader yields a list of equally sized images
            images = torch.stack(list(images))
        dehazed_images, _ = self.dehazing_model(images)
        normalized = normalize_for_detection(dehazed_images)
        detection_results = self.detection_model(list(normalized.unbind(0)), targets)
        return detection_results, dehazed_images


def create_detection_model(config):
    return DetectionModel(num_classes=91, model_name=config["detection"]["model"], pretrained=config["detection"]["pretrained"])


def create_integrated_system(dehazing_router, detection_model):
    return IntegratedDetectionSystem(dehazing_model=dehazing_router, detection_model=detection_model)
