"""Frozen feature networks of the perceptual loss terms (reference training/loss.py:7-108) on libadb200 kernels:
VGG16 `features[:24]` for ContentLoss and the AlexNet trunk + LPIPS head for PerceptualLoss.

Forward: every conv is one adb_conv2d launch with bias + ReLU fused (3-channel stems through adb_stem_pack); pred and
target ride in ONE batch (pred first).  Backward (w.r.t. pred only, the networks are frozen — loss.py:27-28): ReLU mask,
adb_maxpool_bwd, the data-gradient form of adb_conv2d on the pred half of the batch, and adb_stem_unpack back to the
NCHW fp32 image.  The gradient is produced inside the loss forward (while the activations are live) and scaled by the
incoming autograd gradient later, so no feature map outlives the call.
"""
import torch

from .. import _lib, ops
from ..ops import ACT_NONE, ACT_RELU, ConvSpec
from . import autograd as _ag


def _f32(n, dev):
    return torch.empty(n, dtype=torch.float32, device=dev)


class _Layer:
    pass


class FrozenNet:
    """conv(+bias)+ReLU / max-pool chain with feature taps.  `arch`: list of
    ("stem", conv, dict(kh, kw, stride, pad, kp)) | ("conv", conv) | ("pool", k, stride, pad) | ("tap",)."""

    def __init__(self, arch, in_scale, in_shift):
        self.arch = arch
        self.in_scale = [float(v) for v in in_scale]
        self.in_shift = [float(v) for v in in_shift]
        self._cache = _ag._WeightCache()

    # ------------------------------------------------------------------ packing
    def _stem_specs(self, conv, g):
        w, b = conv.weight, conv.bias
        co, _, kh, kw = w.shape
        kp = g["kp"]
        if g["full"]:      # full im2col: the stem is a 1x1 conv over kp channels, K index (r*kw+s)*3 + c
            wp = torch.zeros(ops.pad16(co), kp, dtype=torch.float32, device=w.device)
            wp[:co, :kh * kw * 3] = w.detach().float().permute(0, 2, 3, 1).reshape(co, kh * kw * 3)
            scale, shift = ops.fold_bn(co, b, None, device=w.device)
            fwd = ConvSpec(ops.CONV_S1, 1, 1, 0, co, wp.to(torch.bfloat16).contiguous(), scale, shift, ACT_RELU)
            wd = wp[:co].t().reshape(kp, co, 1, 1).contiguous()          # dcols[j] = sum_co dz[co] * W[co][j]
            return fwd, ConvSpec.from_conv(wd, pad=0)
        fwd = ConvSpec.from_stem(w, kp, bias=b, act=ACT_RELU)
        wp = torch.zeros(co, kh, kp, dtype=torch.float32, device=w.device)
        wp[:, :, :3 * kw] = w.detach().float().permute(0, 2, 3, 1).reshape(co, kh, kw * 3)
        wd = wp.flip(1).permute(2, 0, 1).reshape(kp, co, kh, 1).contiguous()  # rows flipped, channels swapped
        return fwd, ConvSpec.from_conv(wd, pad=kh // 2)

    def specs(self):
        out = []
        for item in self.arch:
            if item[0] == "stem":
                conv, g = item[1], item[2]
                out.append(self._cache.get(("s", id(conv)), (conv.weight, conv.bias), lambda c=conv, g=g: self._stem_specs(c, g)))
            elif item[0] == "conv":
                conv = item[1]
                k = conv.kernel_size[0]
                out.append(self._cache.get(("c", id(conv)), (conv.weight, conv.bias), lambda c=conv, k=k: (
                    ConvSpec.from_conv(c.weight, bias=c.bias, act=ACT_RELU, pad=c.padding[0]), _ag._dgrad_spec_s1(c.weight))))
            else:
                out.append(None)
        return out

    # ------------------------------------------------------------------ execution
    def forward(self, images):
        """images: NCHW fp32 [N,3,H,W] (pred rows first).  Returns (taps, saved)."""
        n, _, h, w = images.shape
        dev = images.device
        for item in self.arch:
            if item[0] in ("stem", "conv") and item[1].weight.device != dev:
                raise RuntimeError(f"perceptual loss trunk lives on {item[1].weight.device} but the images are on {dev}: "
                                   "move the criterion with .to(device) (train_dehazing.py:46) — there is no CPU path")
        st = _lib.current_stream()
        sc = (_lib.C.c_float * 3)(*self.in_scale)
        sh = (_lib.C.c_float * 3)(*self.in_shift)
        xin = torch.empty_like(images)
        _lib.call("adb_image_affine", _lib.ptr(images), n, h, w, sc, sh, _lib.ptr(xin), st)
        specs = self.specs()
        taps, saved = [], []
        cur = None
        for item, sp in zip(self.arch, specs):
            if item[0] == "stem":
                g = item[2]
                cols = ops.stem_pack(xin, g["kw"], g["pad"], g["kp"], stride=g["stride"], kh=g["kh"] if g["full"] else 1)
                cur = ops.conv2d(sp[0], cols)
                saved.append(("stem", cols.shape, cur))
                del cols
            elif item[0] == "conv":
                nxt = ops.conv2d(sp[0], cur)
                saved.append(("conv", cur, nxt))
                cur = nxt
            elif item[0] == "pool":
                _, k, s, p = item
                nb, hh, ww, c = cur.shape
                ho, wo = (hh + 2 * p - k) // s + 1, (ww + 2 * p - k) // s + 1
                nxt = torch.empty((nb, ho, wo, c), dtype=torch.bfloat16, device=dev)
                _lib.call("adb_maxpool_fwd", _lib.ptr(cur), nb, hh, ww, c, k, s, p, _lib.ptr(nxt), st)
                saved.append(("pool", cur, nxt))
                cur = nxt
            else:
                taps.append(cur)
                saved.append(("tap", len(taps) - 1))
        return taps, saved, (h, w)

    def backward(self, dtaps, saved, hw, nb):
        """dtaps[i]: gradient (NHWC bf16, first nb images valid) of tap i or None.  Returns d loss / d images[:nb] (NCHW fp32)."""
        specs = self.specs()
        st = _lib.current_stream()
        g = None
        dx = None
        for item, sp, sv in zip(reversed(self.arch), reversed(specs), reversed(saved)):
            if item[0] == "tap":
                d = dtaps[sv[1]]
                if d is not None:
                    if g is None:
                        g = d
                    else:
                        _, hh, ww, c = g.shape
                        _lib.call("adb_add_bf16", _lib.ptr(g), c, _lib.ptr(d), c, nb * hh * ww, c, st)
                continue
            if g is None:
                continue
            if item[0] == "pool":
                _, k, s, p = item
                x, y = sv[1], sv[2]
                _, hh, ww, c = x.shape
                dxp = torch.empty((nb, hh, ww, c), dtype=torch.bfloat16, device=x.device)
                _lib.call("adb_maxpool_bwd", _lib.ptr(g), _lib.ptr(x), _lib.ptr(y), nb, hh, ww, c, k, s, p, _lib.ptr(dxp), st)
                g = dxp
                continue
            # conv / stem: ReLU mask (in place), then the data gradient
            y = sv[2]
            _, hh, ww, c = y.shape
            px = nb * hh * ww
            scratch = _f32(int(_lib.load().adb_bn_scratch_floats(px, c)), y.device)
            _lib.call("adb_bn_bwd", _lib.ptr(g), c, _lib.ptr(y), c, None, 0, px, c, ACT_RELU, None, None, None, _lib.ptr(scratch),
                      _lib.ptr(g), c, None, 0, None, None, 0, st)
            if item[0] == "conv":
                g = ops.conv2d(sp[1], g, n=nb)
            else:
                geo = item[2]
                dcols = ops.conv2d(sp[1], g, n=nb)
                h, w = hw
                dx = torch.empty((nb, 3, h, w), dtype=torch.float32, device=y.device)
                sc = (_lib.C.c_float * 3)(*self.in_scale)
                _lib.call("adb_stem_unpack", _lib.ptr(dcols), nb, h, w, geo["kh"] if geo["full"] else 1, geo["kw"], geo["pad"],
                          geo["stride"], geo["kp"], sc, 0, _lib.ptr(dx), st)
                g = None
        return dx


# ---------------------------------------------------------------------- the two networks
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
LPIPS_SHIFT = (-0.030, -0.088, -0.188)
LPIPS_SCALE = (0.458, 0.448, 0.450)


def vgg16_content_net(features):
    """torchvision vgg16().features up to index 23 with taps after the pools at 9 / 16 / 23 (loss.py:31-45,72-78: the
    layer map sends 'relu2_2/3_3/4_3' to indices that are the MaxPool layers)."""
    arch = []
    first = True
    for idx in range(24):
        m = features[idx]
        if isinstance(m, torch.nn.Conv2d):
            if first:
                arch.append(("stem", m, dict(kh=3, kw=3, stride=1, pad=1, kp=16, full=False)))
                first = False
            else:
                arch.append(("conv", m))
        elif isinstance(m, torch.nn.MaxPool2d):
            arch.append(("pool", 2, 2, 0))
        if idx in (9, 16, 23):
            arch.append(("tap",))
    return FrozenNet(arch, [1.0 / s for s in IMAGENET_STD], [-m / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)])


def alexnet_lpips_net(convs):
    """AlexNet trunk as LPIPS slices it (five ReLU taps); input affine = LPIPS' [-1,1] mapping (loss.py:104-105) followed
    by its ScalingLayer."""
    c1, c2, c3, c4, c5 = convs
    arch = [("stem", c1, dict(kh=11, kw=11, stride=4, pad=2, kp=384, full=True)), ("tap",),
            ("pool", 3, 2, 0), ("conv", c2), ("tap",),
            ("pool", 3, 2, 0), ("conv", c3), ("tap",), ("conv", c4), ("tap",), ("conv", c5), ("tap",)]
    return FrozenNet(arch, [2.0 / s for s in LPIPS_SCALE], [(-1.0 - sh) / s for sh, s in zip(LPIPS_SHIFT, LPIPS_SCALE)])


class _PrecomputedGrad(torch.autograd.Function):
    """value with d value / d pred already evaluated (per sample when value is [B,...])."""

    @staticmethod
    def forward(ctx, pred, value, dpred):
        ctx.save_for_backward(dpred)
        ctx.per_sample = value.dim() > 0
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        if ctx.per_sample:
            return dpred * g.reshape(-1, 1, 1, 1), None, None
        return dpred * g, None, None


def content_loss(net, pred, target):
    """ContentLoss.forward (loss.py:47-84): mean over the three taps of mse(features(pred), features(target))."""
    b = pred.shape[0]
    want_grad = torch.is_grad_enabled() and pred.requires_grad
    with torch.no_grad():
        both = torch.cat([pred.detach().float(), target.detach().float()], 0).contiguous()
        taps, saved, hw = net.forward(both)
        val = torch.zeros(1, dtype=torch.float32, device=pred.device)
        dtaps = []
        for f in taps:
            half = f[:b].numel()
            fb = f[b:]
            da = torch.empty_like(f[:b]) if want_grad else None
            _lib.call("adb_mse_feat", _lib.ptr(f), _lib.ptr(fb), half, 1.0 / len(taps), _lib.ptr(val), _lib.ptr(da), _lib.current_stream())
            dtaps.append(da)
        value = (val / len(taps)).reshape(())
        if not want_grad:
            return value
        dpred = net.backward(dtaps, saved, hw, b)
    return _PrecomputedGrad.apply(pred, value, dpred)


def lpips_distance(net, lin_weights, pred, target):
    """lpips.LPIPS(net='alex').forward(2x-1, 2t-1) -> [B,1,1,1] (loss.py:86-108), eval mode (no dropout)."""
    b = pred.shape[0]
    want_grad = torch.is_grad_enabled() and pred.requires_grad
    with torch.no_grad():
        both = torch.cat([pred.detach().float(), target.detach().float()], 0).contiguous()
        taps, saved, hw = net.forward(both)
        val = torch.zeros(b, dtype=torch.float32, device=pred.device)
        dtaps = []
        for f, lw in zip(taps, lin_weights):
            _, hh, ww, c = f.shape
            da = torch.empty_like(f[:b]) if want_grad else None
            _lib.call("adb_lpips_tap", _lib.ptr(f), _lib.ptr(f[b:]), b, hh, ww, c, _lib.ptr(lw.detach().float().reshape(-1).contiguous()),
                      1.0, _lib.ptr(val), _lib.ptr(da), _lib.current_stream())
            dtaps.append(da)
        value = val.reshape(b, 1, 1, 1)
        if not want_grad:
            return value
        dpred = net.backward(dtaps, saved, hw, b)
    return _PrecomputedGrad.apply(pred, value, dpred)
