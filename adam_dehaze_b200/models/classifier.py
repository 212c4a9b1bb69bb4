"""Drop-in for models/classifier.py — the fog-intensity classifier ("HDEN").

Same constructor, attributes (model_name, num_classes, feature_dim), head layout (classifier.{1,4}.*) and return
structure `(logits, features)` as /root/reference/models/classifier.py:6-103, 139-145.  The backbone object is the
torchvision model the reference instantiates (so state_dict keys `backbone.*` and random init are identical); it is
used as a parameter container only — the forward runs in libadb200 kernels (engine.ResNetEngine / DenseNetEngine).

Arms on the B200 path: resnet18 (reference default), resnet34, and densenet121 (the north_star's HDEN; an extension —
the reference raises ValueError("Unsupported model") for it).  resnet50 / mobilenet / efficientnet arms keep the
reference's constructor behaviour but have no kernels yet and raise NotImplementedError on forward.
"""
import torch
import torch.nn as nn
import torchvision.models as tvm

from .. import engine as _engine


def _tv(name, pretrained):
    # torchvision >= 0.13 spells pretrained=True as weights="DEFAULT"; there is no network on the GPU box, so a
    # request for pretrained weights that are not cached raises from torchvision exactly as the reference would.
    return getattr(tvm, name)(weights="DEFAULT" if pretrained else None)


class FogIntensityClassifier(nn.Module):
    def __init__(self, model_name="resnet18", num_classes=3, pretrained=True):
        super().__init__()
        self.model_name = model_name
        self.num_classes = num_classes
        if model_name.startswith("resnet"):
            dims = {"resnet18": 512, "resnet34": 512, "resnet50": 2048}
            if model_name not in dims:
                raise ValueError(f"Unsupported ResNet variant: {model_name}")
            self.backbone = _tv(model_name, pretrained)
            self.feature_dim = dims[model_name]
            self.backbone.fc = nn.Identity()
        elif model_name == "densenet121":
            self.backbone = _tv("densenet121", pretrained)
            self.feature_dim = self.backbone.classifier.in_features
            self.backbone.classifier = nn.Identity()
        elif model_name.startswith("efficientnet"):
            raise NotImplementedError("efficientnet backbones need `timm`, which the B200 image does not ship")
        elif model_name.startswith("mobilenet"):
            dims = {"mobilenet_v2": 1280, "mobilenet_v3_small": 576, "mobilenet_v3_large": 960}
            if model_name not in dims:
                raise ValueError(f"Unsupported MobileNet variant: {model_name}")
            self.backbone = _tv(model_name, pretrained)
            self.feature_dim = dims[model_name]
            self.backbone.classifier = nn.Identity()
        else:
            raise ValueError(f"Unsupported model: {model_name}")
        self.classifier = nn.Sequential(
            nn.Dropout(0.3),
            nn.Linear(self.feature_dim, 256),
            nn.ReLU(),
            nn.Dropout(0.2),
            nn.Linear(256, self.num_classes),
        )

    def _hden_engine(self):
        eng = self.__dict__.get("_adb_engine")
        if eng is None:
            if self.model_name in ("resnet18", "resnet34"):
                eng = _engine.ResNetEngine(self)
            elif self.model_name == "densenet121":
                eng = _engine.DenseNetEngine(self)
            else:
                raise NotImplementedError(f"backbone '{self.model_name}' has no B200 kernels yet (resnet18/34, densenet121 do)")
            self.__dict__["_adb_engine"] = eng
        return eng

    def forward(self, x):
        """x [B,3,H,W] fp32 CUDA -> (logits [B,num_classes] fp32, features [B,feature_dim] fp32)."""
        return self._hden_engine().forward(x)

    def route_guard(self, **kw):
        """The fp32 re-evaluation path for near-tie rows (adam_dehaze_b200/route_guard.py); created on first use.
        Keyword arguments (eps, cap, use_graph) rebuild it."""
        g = self.__dict__.get("_adb_route_guard")
        if g is None or kw:
            from ..route_guard import RouteGuard
            g = RouteGuard(self, **kw)
            self.__dict__["_adb_route_guard"] = g
        return g

    def refine_logits(self, x, logits):
        """Patch, in place and on the device, the logits of the rows whose top-2 gap is within the bf16 error of the trunk with
        their fp32 re-evaluation, so that argmax(logits) equals the fp32 reference's (models/routing.py:41-43)."""
        return self.route_guard().refine(x, logits)

    def extract_features(self, x):
        with torch.no_grad():
            return self._hden_engine().forward(x)[1]


def create_classifier(config):
    """Factory with the reference's config keys (classifier.py:139-145)."""
    cfg = config["classifier"]
    return FogIntensityClassifier(model_name=cfg["model"], num_classes=cfg["num_classes"], pretrained=cfg["pretrained"])
