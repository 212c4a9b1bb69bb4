"""Drop-in for training/loss.py (reference loss.py:7-241): ContentLoss, PerceptualLoss, DehazingLoss, JointLoss and their
factories.

Every term runs on libadb200 kernels, forward AND backward, wrapped in torch.autograd.Function so the reference's
`loss.backward()` keeps working:
  * L1 / cross-entropy: fused warp-shuffle reductions (adb_l1_mse_fwd / adb_l1_bwd / adb_ce_fwd_bwd);
  * ContentLoss (VGG16 features at indices 9/16/23) and PerceptualLoss (LPIPS-alex): frozen conv trunks on the tcgen05
    conv kernel, data gradients back to `pred` (training/perceptual.py).
Pretrained weights.  VGG16 / AlexNet come from the torch hub cache; the LPIPS 'lin' layers from the installed `lpips`
package (lpips/weights/v0.1/alex.pth) or the file named by $ADB_LPIPS_WEIGHTS.  Nothing can be downloaded on the GPU box.
When a set is missing the module falls back to a FIXED-SEED random init (the same on every rank and in every process,
independent of the caller's RNG state), prints a one-line notice, and records it in `.pretrained = False`;
`allow_random_weights=False` turns the fallback into an error, and evaluation/metrics.py labels an 'lpips' number computed
from such weights as not comparable.  The arithmetic is the same either way (sub-module names follow torchvision / lpips,
so real checkpoints load with load_state_dict).
"""
import torch
import torch.nn as nn

import os
import sys
from collections import OrderedDict

from .. import ops
from .. import _lib
from . import perceptual as _perc


class _L1Mean(torch.autograd.Function):
    """mean |pred - target| (nn.L1Loss, loss.py:121) — one pass forward, one pass backward."""

    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous().float(), target.contiguous().float()
        ctx.save_for_backward(pred, target)
        return ops.l1_mse(pred, target)[0]

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        grad = torch.empty_like(pred)
        # dL/dpred = g * sign(pred - target) / numel; g is a device scalar, folded in afterwards without a host sync
        _lib.call("adb_l1_bwd", _lib.ptr(pred), _lib.ptr(target), pred.numel(), 1.0, _lib.ptr(grad), _lib.current_stream())
        return grad * g, None


class _CrossEntropy(torch.autograd.Function):
    """nn.CrossEntropyLoss (mean reduction, loss.py:177) with the gradient produced in the same launch."""

    @staticmethod
    def forward(ctx, logits, labels):
        loss, grad = ops.cross_entropy(logits.contiguous().float(), labels.contiguous(), 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected CUDA tensors — this package runs on B200 (sm_100a) only and has no CPU path")


def _hub_file(name):
    path = os.path.join(torch.hub.get_dir(), "checkpoints", name)
    return path if os.path.exists(path) else None


def _lpips_lin_file():
    """The LPIPS-alex linear-layer weights: $ADB_LPIPS_WEIGHTS, or the file the `lpips` package ships."""
    env = os.environ.get("ADB_LPIPS_WEIGHTS")
    if env and os.path.exists(env):
        return env
    try:
        import importlib.util
        spec = importlib.util.find_spec("lpips")
        if spec and spec.origin:
            path = os.path.join(os.path.dirname(spec.origin), "weights", "v0.1", "alex.pth")
            return path if os.path.exists(path) else None
    except Exception:  # noqa: BLE001
        pass
    return None


def _seeded_reinit(module, seed):
    """Re-draw a module's random init from a private fixed-seed generator: identical on every rank, whatever the caller's RNG
    state is (replicas seeded differently would otherwise optimise different losses)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (6.0 / fan_in) ** 0.5)     # He-uniform: keeps the activation scale through the ReLU trunk
            else:
                p.zero_()


class ContentLoss(nn.Module):
    """VGG16 feature MSE at `features` indices 9 / 16 / 23, mean of the three (reference loss.py:7-84)."""

    def __init__(self, pretrained_model="vgg16", content_layers=None, allow_random_weights=True):
        super().__init__()
        import torchvision.models as tvm
        if pretrained_model not in ("vgg16",):
            raise ValueError(f"Unsupported model: {pretrained_model}" if pretrained_model != "vgg19" else
                             "vgg19 content loss is not built on the B200 path")
        self.content_layers = content_layers or ["relu2_2", "relu3_3", "relu4_3"]
        if self.content_layers != ["relu2_2", "relu3_3", "relu4_3"]:
            raise NotImplementedError("ContentLoss on the B200 path taps the reference's default layers (indices 9/16/23)")
        vgg = tvm.vgg16(weights=None)
        ck = _hub_file("vgg16-397923af.pth")
        self.pretrained = bool(ck)
        if ck:
            vgg.load_state_dict(torch.load(ck, map_location="cpu"))
        else:
            if not allow_random_weights:
                raise FileNotFoundError("ContentLoss: vgg16-397923af.pth is not in the torch hub cache (loss.py:20 needs the pretrained VGG16)")
            _seeded_reinit(vgg.features, 16)
            print("ContentLoss: pretrained VGG16 weights are not in the torch hub cache — using fixed-seed random-init features", file=sys.stderr)
        self.model = vgg.features.eval()
        for p in self.model.parameters():
            p.requires_grad = False
        self.__dict__["_net"] = None

    def forward(self, x, target):
        _need_cuda(x, "ContentLoss")
        if self.__dict__["_net"] is None:
            self.__dict__["_net"] = _perc.vgg16_content_net(self.model)
        return _perc.content_loss(self.__dict__["_net"], x, target)


class _LPIPSAlex(nn.Module):
    """Parameter container with lpips.LPIPS(net='alex')'s state_dict names (net.slice{1..5}.{0,3,6,8,10}.*,
    lin{0..4}.model.1.weight, scaling_layer.{shift,scale})."""

    def __init__(self, allow_random_weights=True):
        super().__init__()
        import torchvision.models as tvm
        alex = tvm.alexnet(weights=None)
        ck = _hub_file("alexnet-owt-7be5be79.pth")
        lin_ck = _lpips_lin_file()
        self.pretrained = bool(ck) and bool(lin_ck)
        if not self.pretrained and not allow_random_weights:
            raise FileNotFoundError("PerceptualLoss: LPIPS needs alexnet-owt-7be5be79.pth in the torch hub cache and the lpips package's "
                                    "weights/v0.1/alex.pth (or $ADB_LPIPS_WEIGHTS); neither can be downloaded here")
        if ck:
            alex.load_state_dict(torch.load(ck, map_location="cpu"))
        else:
            _seeded_reinit(alex.features, 17)
            print("PerceptualLoss: pretrained AlexNet / LPIPS weights are not available offline — using fixed-seed random-init weights", file=sys.stderr)
        f = alex.features
        self.net = nn.Module()
        for i, idx in enumerate((0, 3, 6, 8, 10)):
            setattr(self.net, f"slice{i + 1}", nn.Sequential(OrderedDict([(str(idx), f[idx])])))
        self.scaling_layer = nn.Module()
        self.scaling_layer.register_buffer("shift", torch.tensor(_perc.LPIPS_SHIFT).view(1, 3, 1, 1))
        self.scaling_layer.register_buffer("scale", torch.tensor(_perc.LPIPS_SCALE).view(1, 3, 1, 1))
        lin_sd = torch.load(lin_ck, map_location="cpu") if lin_ck else None
        g = torch.Generator().manual_seed(18)
        for i, c in enumerate((64, 192, 384, 256, 256)):
            lin = nn.Module()
            conv = nn.Conv2d(c, 1, 1, bias=False)
            with torch.no_grad():
                if lin_sd is not None:
                    conv.weight.copy_(lin_sd[f"lin{i}.model.1.weight"])
                else:
                    conv.weight.copy_(torch.rand(conv.weight.shape, generator=g) / c)     # LPIPS' lin layers are non-negative
            lin.model = nn.Sequential(OrderedDict([("0", nn.Dropout()), ("1", conv)]))
            setattr(self, f"lin{i}", lin)
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def convs(self):
        return [getattr(self.net, f"slice{i + 1}")[0] for i in range(5)]

    def lin_weights(self):
        return [getattr(self, f"lin{i}").model[1].weight for i in range(5)]


class PerceptualLoss(nn.Module):
    """LPIPS(net='alex') on inputs mapped to [-1, 1]; returns [B,1,1,1] like lpips (reference loss.py:86-108)."""

    def __init__(self, net="alex", allow_random_weights=True):
        super().__init__()
        if net != "alex":
            raise NotImplementedError("PerceptualLoss on the B200 path implements LPIPS(net='alex')")
        self.loss_fn = _LPIPSAlex(allow_random_weights)
        self.pretrained = self.loss_fn.pretrained
        self.__dict__["_net"] = None

    def forward(self, x, target):
        _need_cuda(x, "PerceptualLoss")
        if self.__dict__["_net"] is None:
            self.__dict__["_net"] = _perc.alexnet_lpips_net(self.loss_fn.convs())
        return _perc.lpips_distance(self.__dict__["_net"], self.loss_fn.lin_weights(), x, target)


class DehazingLoss(nn.Module):
    def __init__(self, lambda_l1=1.0, lambda_content=0.1, lambda_perceptual=0.1):
        super().__init__()
        self.lambda_l1, self.lambda_content, self.lambda_perceptual = lambda_l1, lambda_content, lambda_perceptual
        # the reference always builds both sub-losses (loss.py:121-123); a zero lambda skips building the trunk here
        self.content_loss = ContentLoss() if lambda_content != 0 else None
        self.perceptual_loss = PerceptualLoss() if lambda_perceptual != 0 else None

    def forward(self, pred, target):
        """Returns (total, {'l1','content','perceptual','total'}) like loss.py:125-162."""
        _need_cuda(pred, "DehazingLoss")
        l1 = _L1Mean.apply(pred, target)
        zero = torch.zeros((), device=pred.device)
        content = self.content_loss(pred, target) if self.content_loss is not None else zero
        perceptual = self.perceptual_loss(pred, target) if self.perceptual_loss is not None else zero
        if perceptual.dim() > 0:
            perceptual = perceptual.mean()
        total = self.lambda_l1 * l1 + self.lambda_content * content + self.lambda_perceptual * perceptual
        return total, {"l1": l1, "content": content, "perceptual": perceptual, "total": total}


class JointLoss(nn.Module):
    def __init__(self, lambda_dehazing=1.0, lambda_classification=0.2, lambda_detection=0.5, config=None,
                 dehazing_loss=None):
        super().__init__()
        self.lambda_dehazing = lambda_dehazing
        self.lambda_classification = lambda_classification
        self.lambda_detection = lambda_detection
        self.dehazing_loss = dehazing_loss if dehazing_loss is not None else DehazingLoss()

    def forward(self, pred, target_clear, pred_intensity=None, target_intensity=None, detection_loss=None):
        """Returns (total, {'dehazing','classification','detection','total','dehazing_components'}), loss.py:179-224."""
        dehazing, parts = self.dehazing_loss(pred, target_clear)
        if pred_intensity is not None and target_intensity is not None:
            ce = _CrossEntropy.apply(pred_intensity, target_intensity)
        else:
            ce = torch.zeros((), device=pred.device)      # (torch.tensor(0.0, device=...) is a synchronising H2D copy)
        det = detection_loss if detection_loss is not None else torch.zeros((), device=pred.device)
        total = self.lambda_dehazing * dehazing + self.lambda_classification * ce + self.lambda_detection * det
        return total, {"dehazing": dehazing, "classification": ce, "detection": det, "total": total,
                       "dehazing_components": parts}


def get_dehazing_loss(config):
    """Factory (loss.py:226-232): fixed lambdas 1.0 / 0.1 / 0.1 as in the reference."""
    return DehazingLoss(lambda_l1=1.0, lambda_content=0.1, lambda_perceptual=0.1)


def get_joint_loss(config):
    """Factory (loss.py:234-241)."""
    jt = config["joint_training"]
    return JointLoss(lambda_dehazing=jt["lambda_dehazing"], lambda_classification=jt["lambda_classification"],
                     lambda_detection=jt["lambda_detection"], config=config)
