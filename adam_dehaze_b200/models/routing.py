"""Drop-in for models/routing.py — Hard / Soft / Gated routers (reference routing.py:5-252).

HardRouter: one libadb200 launch turns the HDEN logits into `intensity` + three ascending bucket index lists + their
counts, all on the device; each branch then walks its bucket with the count read on the device, so the batch is never
synchronised with the host (the reference pays >= 9 syncs per batch, routing.py:55-61) and the gathered sub-batches /
scattered outputs are never materialised (the stem reads through the index list, the last conv writes through it).
"""
import torch
import torch.nn as nn

from .. import ops
from .. import engine as _engine

_NAMES = ("low", "medium", "high")


class HardRouter(nn.Module):
    def __init__(self, models, classifier=None, device="cuda"):
        super().__init__()
        self.models = nn.ModuleDict(models)
        self.classifier = classifier
        self.device = device
        # re-evaluate near-tie rows of the bf16 HDEN in fp32 before the argmax (route decisions bit-exact vs the fp32 reference)
        self.route_guard = True

    def forward(self, x, intensity=None):
        """Returns (outputs, {'intensity', 'low_mask', 'medium_mask', 'high_mask'}) like routing.py:23-68.

        NB: the reference's drivers call router(x, logits) positionally, which binds logits to
        `intensity`; like the reference we take `intensity` literally (int64 class ids)."""
        _engine.require_cuda(x, "HardRouter")
        x = x.contiguous()
        training = any(m.training for m in self.models.values())
        # routing.py:31 starts from zeros_like(x); in eval mode every routed row is overwritten in full by its branch's image
        # epilogue, so only the rows no branch owns (class id outside {0,1,2}) are cleared — on the device, after routing
        outputs = torch.zeros_like(x) if training else torch.empty_like(x)
        if intensity is None and self.classifier is not None:
            with torch.no_grad():
                logits, _ = self.classifier(x)
                if self.route_guard and not self.classifier.training and hasattr(self.classifier, "refine_logits"):
                    logits = self.classifier.refine_logits(x, logits.contiguous())
            inten, masks, bidx, bcnt = ops.route(logits=logits)
        elif intensity is not None:
            if intensity.dim() != 1 or intensity.shape[0] != x.shape[0] or intensity.is_floating_point():
                # the reference compares `intensity == k` elementwise; a [B,3] float logits tensor yields masks that
                # select nothing consistently — we refuse instead of returning zeros silently
                raise ValueError("HardRouter.forward(x, intensity): intensity must be an integer tensor of shape [B]")
            inten, masks, bidx, bcnt = ops.route(intensity=intensity.to(x.device))
        else:
            raise ValueError("HardRouter needs a classifier or an explicit intensity tensor")
        if not training:
            from .. import _lib
            _lib.call("adb_zero_unrouted", _lib.ptr(outputs), _lib.ptr(inten), x.shape[0], x[0].numel(), _lib.current_stream())
        if training:
            # train() mode (routing.py:55-61 under model.train()): each branch sees its own sub-batch (its BatchNorm statistics
            # are the sub-batch's, as in the reference) and autograd must reach it, so the buckets are gathered / scattered
            # with differentiable index ops; the bucket sizes are read on the host (one sync per batch, training only).
            counts = bcnt.tolist()
            for name, model in self.models.items():
                k = _NAMES.index(name)
                if counts[k] == 0:
                    continue
                idx = bidx[k, :counts[k]].long()
                outputs = outputs.index_copy(0, idx, model(x.index_select(0, idx)))
            return outputs, {"intensity": inten, "low_mask": masks[0], "medium_mask": masks[1], "high_mask": masks[2]}
        for name, model in self.models.items():
            k = _NAMES.index(name)
            model.forward_bucket(x, outputs, bidx[k], bcnt[k:k + 1], count=x.shape[0])
        return outputs, {"intensity": inten, "low_mask": masks[0], "medium_mask": masks[1], "high_mask": masks[2]}


class _Blend3Fn(torch.autograd.Function):
    """sum_k softmax(logits/T)[:,k] * y_k with a kernel-built backward (dy_k and dlogits), routing.py:111-127."""

    @staticmethod
    def forward(ctx, y0, y1, y2, logits, temperature):
        y0, y1, y2 = y0.contiguous(), y1.contiguous(), y2.contiguous()
        blend, weights = ops.blend3(y0, y1, y2, logits, temperature)
        ctx.save_for_backward(y0, y1, y2, weights)
        ctx.temperature = temperature
        ctx.mark_non_differentiable(weights)
        return blend, weights

    @staticmethod
    def backward(ctx, dout, _dw):
        from .. import _lib
        y0, y1, y2, weights = ctx.saved_tensors
        dout = dout.contiguous().float()
        b, chw = y0.shape[0], y0[0].numel()
        d0, d1, d2 = torch.empty_like(y0), torch.empty_like(y1), torch.empty_like(y2)
        dwt = torch.empty((b, 3), dtype=torch.float32, device=y0.device)
        dlog = torch.empty((b, 3), dtype=torch.float32, device=y0.device)
        _lib.call("adb_blend3_bwd", _lib.ptr(dout), _lib.ptr(y0), _lib.ptr(y1), _lib.ptr(y2), _lib.ptr(weights),
                  float(ctx.temperature), b, chw, _lib.ptr(d0), _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(dwt), _lib.ptr(dlog),
                  _lib.current_stream())
        return d0, d1, d2, dlog, None


def _blend(y0, y1, y2, logits, temperature):
    if torch.is_grad_enabled() and any(t.requires_grad for t in (y0, y1, y2, logits)):
        return _Blend3Fn.apply(y0, y1, y2, logits.contiguous().float(), temperature)
    return ops.blend3(y0, y1, y2, logits, temperature)


class _GateMLPFn(torch.autograd.Function):
    """GatedRouter.gate_network up to the softmax (routing.py:155-163): Linear-ReLU-Dropout-Linear-ReLU-Linear, forward
    and backward on adb_linear / adb_linear_bwd / adb_mul_f32; the dropout mask comes from torch's RNG in train() mode."""

    @staticmethod
    def forward(ctx, feats, w0, b0, w3, b3, w5, b5, p_drop, training):
        from .. import _lib
        feats = feats.contiguous().float()
        st = _lib.current_stream()
        h0 = ops.linear(feats, w0, b0, relu=True)
        if training and p_drop > 0:
            mask = (torch.rand_like(h0) >= p_drop).float() / (1.0 - p_drop)
            h0d = torch.empty_like(h0)
            _lib.call("adb_mul_f32", _lib.ptr(h0), _lib.ptr(mask), None, h0.numel(), _lib.ptr(h0d), st)
        else:
            mask, h0d = None, h0
        h1 = ops.linear(h0d, w3, b3, relu=True)
        logits = ops.linear(h1, w5, b5, relu=False)
        ctx.save_for_backward(feats, h0, h0d, h1, w0, w3, w5)
        ctx.mask = mask
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        from .. import _lib
        feats, h0, h0d, h1, w0, w3, w5 = ctx.saved_tensors
        st = _lib.current_stream()
        n = feats.shape[0]
        dlogits = dlogits.contiguous().float()

        def lin_bwd(x, w, dy):
            dx = torch.empty_like(x)
            dw, db = torch.empty_like(w), torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
            _lib.call("adb_linear_bwd", _lib.ptr(x), _lib.ptr(w.detach()), _lib.ptr(dy), n, w.shape[1], w.shape[0], _lib.ptr(dx),
                      _lib.ptr(dw), _lib.ptr(db), st)
            return dx, dw, db

        def gate(d, m, act):      # d * m, zeroed where the ReLU output `act` is not positive
            out = torch.empty_like(d)
            _lib.call("adb_mul_f32", _lib.ptr(d), _lib.ptr(m), _lib.ptr(act), d.numel(), _lib.ptr(out), st)
            return out
        dh1, dw5, db5 = lin_bwd(h1, w5, dlogits)
        dh0d, dw3, db3 = lin_bwd(h0d, w3, gate(dh1, None, h1))
        dfeat, dw0, db0 = lin_bwd(feats, w0, gate(dh0d, ctx.mask, h0))
        return dfeat, dw0, db0, dw3, db3, dw5, db5, None, None


class SoftRouter(nn.Module):
    def __init__(self, models, classifier=None, temperature=1.0, device="cuda"):
        super().__init__()
        self.models = nn.ModuleDict(models)
        self.classifier = classifier
        self.temperature = temperature
        self.device = device

    def forward(self, x, classifier_logits=None):
        """Returns (blend, {'weights', 'individual_outputs'}) like routing.py:90-132."""
        _engine.require_cuda(x, "SoftRouter")
        if classifier_logits is None and self.classifier is not None:
            logits, _ = self.classifier(x)
        else:
            logits = classifier_logits
        if logits is None:
            raise ValueError("SoftRouter needs a classifier or precomputed logits")
        outs = {name: self.models[name](x) for name in _NAMES if name in self.models}
        if len(outs) != 3:
            raise NotImplementedError("SoftRouter on the B200 path blends exactly the three branches low/medium/high")
        blend, weights = _blend(outs["low"], outs["medium"], outs["high"], logits, self.temperature)
        return blend, {"weights": weights, "individual_outputs": outs}


class GatedRouter(nn.Module):
    def __init__(self, models, classifier=None, feature_dim=512, device="cuda"):
        super().__init__()
        self.models = nn.ModuleDict(models)
        self.classifier = classifier
        self.device = device
        self.gate_network = nn.Sequential(
            nn.Linear(feature_dim, 256), nn.ReLU(inplace=True), nn.Dropout(0.3),
            nn.Linear(256, 128), nn.ReLU(inplace=True),
            nn.Linear(128, len(models)), nn.Softmax(dim=1),
        )
        self.use_feature_fusion = False

    def forward(self, x):
        """Returns (blend, {'gate_weights', 'individual_outputs'}) like routing.py:173-226 (eval mode)."""
        _engine.require_cuda(x, "GatedRouter")
        if self.classifier is None:
            raise NotImplementedError("GatedRouter without a classifier (uniform weights) is not built on the B200 path")
        _, feats = self.classifier(x)
        g = self.gate_network
        if torch.is_grad_enabled() and (self.training or feats.requires_grad):
            gate_logits = _GateMLPFn.apply(feats, g[0].weight, g[0].bias, g[3].weight, g[3].bias, g[5].weight, g[5].bias,
                                           float(g[2].p), self.training)
        else:
            hid = ops.linear(feats, g[0].weight, g[0].bias, relu=True)
            gate_logits = ops.head_mlp(hid, g[3].weight.detach(), g[3].bias.detach(), g[5].weight.detach(), g[5].bias.detach())
        outs = {name: self.models[name](x) for name in _NAMES if name in self.models}
        if len(outs) != 3:
            raise NotImplementedError("GatedRouter on the B200 path blends exactly the three branches low/medium/high")
        blend, weights = _blend(outs["low"], outs["medium"], outs["high"], gate_logits, 1.0)
        return blend, {"gate_weights": weights, "individual_outputs": outs}


def create_router(models, classifier, config):
    """Factory with the reference's config keys (routing.py:228-252)."""
    kind = config["routing"]["type"]
    if kind == "hard":
        return HardRouter(models=models, classifier=classifier, device=config["device"])
    if kind == "soft":
        return SoftRouter(models=models, classifier=classifier, temperature=config["routing"]["temperature"],
                          device=config["device"])
    if kind == "gated":
        feature_dim = getattr(classifier, "feature_dim", 512)
        return GatedRouter(models=models, classifier=classifier, feature_dim=feature_dim, device=config["device"])
    raise ValueError(f"Unsupported routing type: {kind}")
