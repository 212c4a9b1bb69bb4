"""Execution engine: turns the drop-in nn.Modules (parameter containers with the reference's state_dict layout) into
sequences of fused libadb200 launches on NHWC bf16 feature maps.

One engine per module instance.  It (re)packs weights whenever a parameter or BN buffer changes (torch `_version`
counters), owns the activation buffers of a micro-batch, and walks a routed bucket micro-batch by micro-batch with the
bucket's live count kept on the device (`n_dev`) — no host synchronisation anywhere on the path.

eval(): BatchNorm uses running statistics folded into the conv epilogues (the launches below).  train(): the same engines
hand the forward to training/autograd.py (batch-statistics BatchNorm on a tape, kernel-built backward).  There is no
fallback to torch ops in either mode; CPU tensors raise.
"""
import os

import torch

from . import ops
from .ops import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, EPI_DOT, EPI_IMAGE, IMG_BLEND, IMG_GUIDED, IMG_RESIDUAL,
                  AttnParams, ConvSpec)

# activation bytes one image may take before the micro-batch is cut (per branch pass)
_MICRO_BATCH_BYTES = 24 << 30
_MAX_MICRO_BATCH = 16


def bn_args(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


def require_inference(module, what):
    """Stand-alone ConvBlock / ResidualBlock calls (BlockEngine) exist for eval-mode parity checks only; training goes through
    the branch / classifier engines' train() path (training/autograd.py)."""
    if module.training:
        raise NotImplementedError(
            f"{what}: a stand-alone block call runs in eval mode only (BatchNorm with running statistics); call .eval() first, "
            "or train the block as part of a branch model (LightweightDehazeModel etc.), whose train() mode is kernel-built. "
            "There is no torch fallback.")


def require_cuda(x, what):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError(f"{what}: expected a CUDA tensor — this package runs on B200 (sm_100a) only and has no CPU path")
    if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"{what}: expected an NCHW float32 image batch with 3 channels, got {tuple(x.shape)} {x.dtype}")


def require_cuda_any(x, what):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4):
        raise RuntimeError(f"{what}: expected a 4-D CUDA tensor — this package runs on B200 (sm_100a) only and has no CPU path")


def block_engine(module):
    eng = module.__dict__.get("_adb_engine")
    if eng is None:
        eng = BlockEngine(module)
        module.__dict__["_adb_engine"] = eng
    return eng


class _Versioned:
    """Re-pack when any tensor of the module changed in place (optimizer step, load_state_dict, .to())."""

    def __init__(self, module):
        self.module = module
        self._sig = None

    def signature(self):
        items = list(self.module.named_parameters()) + list(self.module.named_buffers())
        return tuple((k, t.data_ptr(), t._version, str(t.device)) for k, t in items)

    def stale(self):
        sig = self.signature()
        if sig != self._sig:
            self._sig = sig
            return True
        return False


def _conv_block_spec(cb, act_override=None, stem_kp=None, stride=1, pad=None):
    """ConvBlock (Conv2d [+BN] [+ReLU]) -> one fused launch spec."""
    conv = cb.block[0]
    bn = None
    act = ACT_NONE
    for m in list(cb.block)[1:]:
        if isinstance(m, torch.nn.BatchNorm2d):
            bn = bn_args(m)
        elif isinstance(m, torch.nn.ReLU):
            act = ACT_RELU
    if act_override is not None:
        act = act_override
    if stem_kp:
        return ConvSpec.from_stem(conv.weight, stem_kp, bias=conv.bias, bn=bn, act=act)
    return ConvSpec.from_conv(conv.weight, bias=conv.bias, bn=bn, act=act, stride=conv.stride[0], pad=conv.padding[0])


def _res_specs(rb):
    return (_conv_block_spec(rb.conv1), _conv_block_spec(rb.conv2, act_override=ACT_RELU))  # ReLU after the residual add


def _attn_params(ab):
    return AttnParams(ab.fc[0].weight, ab.fc[2].weight, ab.conv_spatial.weight)


# AttentionBlock pools folded from the producing conv's epilogue partials (False: a separate pass over the map)
POOL_FOLD = os.environ.get("ADB_NO_POOL_FOLD", "") == ""


class BranchEngine:
    """Runs LightweightDehazeModel / MediumIntensityDehazeModel / HighIntensityDehazeModel forwards."""

    def __init__(self, model, kind):
        assert kind in ("light", "unet", "unet_attn", "low_unet", "corun", "dual")
        self.model, self.kind = model, kind
        self._ver = _Versioned(model)
        self.S = None
        self._bufs = {}
        self.tune = None
        self.train_cache = None   # bf16 weight packings of the training path (training/autograd.py)

    # ------------------------------------------------------------------ packing
    def specs(self):
        stale = self._ver.stale()   # always evaluated: it records the signature the packed specs correspond to
        if self.S is not None and not stale:
            return self.S
        m = self.model
        S = {}
        if self.kind == "light":
            S["init"] = _conv_block_spec(m.init_conv, stem_kp=16)
            S["res"] = [_res_specs(rb) for rb in m.residual_blocks]
            S["out0"] = _conv_block_spec(m.output_conv[0])
            S["out1"] = ConvSpec.from_conv(m.output_conv[1].weight, bias=m.output_conv[1].bias, act=ACT_SIGMOID)
        elif self.kind == "low_unet":      # LowIntensityDehazeModel, low_intensity.py:56-125
            S["init"] = _conv_block_spec(m.init_conv, stem_kp=16)
            S["down"] = _conv_block_spec(m.down1[0])
            S["res"] = [_res_specs(m.down1[1])] + [_res_specs(rb) for rb in m.bottleneck]
            S["up"] = ConvSpec.from_convT(m.up1[0].weight, bias=m.up1[0].bias, bn=bn_args(m.up1[1]), act=ACT_RELU)
            S["out0"] = _conv_block_spec(m.output_conv[0])
            S["out1"] = _conv_block_spec(m.output_conv[1])
            # clamp(x + (sigmoid(z) - 0.5) * 2) == clamp(x + tanh(z / 2)): halve the epilogue affine, use the tanh/residual epilogue
            sp = ConvSpec.from_conv(m.output_conv[2].weight, bias=m.output_conv[2].bias, act=ACT_TANH)
            sp.scale, sp.shift = (sp.scale * 0.5).contiguous(), (sp.shift * 0.5).contiguous()
            S["out2"] = sp
        elif self.kind == "corun":         # COrunInspiredModel, medium_intensity.py:128-199
            S["init"] = _conv_block_spec(m.init_conv, stem_kp=32)
            S["scale1"] = _conv_block_spec(m.scale1_conv)
            S["scale2"] = _conv_block_spec(m.scale2_conv[1])
            S["scale3"] = _conv_block_spec(m.scale3_conv[1])
            S["fusion"] = _conv_block_spec(m.fusion_conv)
            S["res"] = [_res_specs(rb) for rb in m.residual_blocks]
            S["out0"] = _conv_block_spec(m.output_conv[0])
            S["out1"] = ConvSpec.from_conv(m.output_conv[1].weight, bias=m.output_conv[1].bias, act=ACT_TANH)
        elif self.kind == "dual":          # DualBranchAttentionModel, high_intensity.py:149-223
            gb, lb, tb, fc = m.global_branch, m.local_branch, m.transmission_branch, m.fusion_conv
            S["g_init"] = _conv_block_spec(gb[0], stem_kp=32)
            S["g_res"] = [_res_specs(gb[2]), _res_specs(gb[5]), _res_specs(gb[7]), _res_specs(gb[9])]
            S["g_attn"] = [_attn_params(gb[3]), _attn_params(gb[6])]
            S["g_out"] = _conv_block_spec(gb[11])
            S["l_init"] = _conv_block_spec(lb[0], stem_kp=16)
            S["l_res"] = [_res_specs(lb[1]), _res_specs(lb[2])]
            S["l_out"] = _conv_block_spec(lb[3])
            S["t0"] = _conv_block_spec(tb[0])
            S["t1"] = _conv_block_spec(tb[1])
            # (1 - sigmoid(w.f + b)) == sigmoid(-(w.f + b)): the 1x1 head rides in the DOT epilogue with negated weights
            cp = S["t1"].cout_pad
            w = tb[2].weight.detach().float().reshape(-1)
            dot_w = torch.zeros(cp, dtype=torch.float32, device=w.device)
            dot_w[:w.numel()] = -w
            S["t_dot"] = (dot_w.contiguous(), -float(tb[2].bias.detach().float().item()))
            S["f0"] = _conv_block_spec(fc[0])
            S["f1"] = ConvSpec.from_conv(fc[1].weight, bias=fc[1].bias, act=ACT_TANH)
        else:
            attn = self.kind == "unet_attn"
            S["init"] = _conv_block_spec(m.init_conv, stem_kp=32)
            S["enc"] = []
            for e in m.encoder:
                d = {"down": _conv_block_spec(e[0]), "res": [_res_specs(e[1]), _res_specs(e[2])]}
                if attn:
                    d["attn"] = _attn_params(e[3])
                S["enc"].append(d)
            if attn:
                S["bott"] = [(_res_specs(m.bottleneck[0]), _attn_params(m.bottleneck[1])),
                             (_res_specs(m.bottleneck[2]), _attn_params(m.bottleneck[3]))]
            else:
                S["bott"] = [(_res_specs(m.bottleneck[0]), None), (_res_specs(m.bottleneck[1]), None)]
            S["dec"] = []
            for dmod in m.decoder:
                d = {"up": ConvSpec.from_convT(dmod[0].weight, bias=dmod[0].bias, bn=bn_args(dmod[1]), act=ACT_RELU),
                     "res": _res_specs(dmod[3])}
                if attn:
                    d["attn"] = _attn_params(dmod[4])
                S["dec"].append(d)
            S["out0"] = _conv_block_spec(m.output_conv[0])
            S["out1"] = _conv_block_spec(m.output_conv[1])
            S["out2"] = ConvSpec.from_conv(m.output_conv[2].weight, bias=m.output_conv[2].bias, act=ACT_TANH)
            if attn:
                S["det0"] = _conv_block_spec(m.detail_branch[0], stem_kp=16)
                S["det1"] = _conv_block_spec(m.detail_branch[1])
                w = m.detail_branch[2].weight.detach().float().reshape(-1)
                dot_w = torch.zeros(16, dtype=torch.float32, device=w.device)
                dot_w[:w.numel()] = w
                # the 1x1 bias rides in the launch descriptor: one host read per re-pack, none per forward
                S["det_dot"] = (dot_w.contiguous(), float(m.detail_branch[2].bias.detach().float().item()))
        self.S = S
        return S

    # ------------------------------------------------------------------ buffers
    def bytes_per_image(self, h, w):
        c = self.model.base_channels
        px = h * w
        if self.kind == "light":
            return px * (16 + 3 * c) * 2
        if self.kind == "low_unet":
            return px * (16 + 3 * c) * 2 + (px // 4) * 2 * 2 * c * 2
        if self.kind == "corun":
            return px * (32 + c + 7 * c + 2 * 2 * c + c) * 2 + (px // 4) * 3 * c * 2 + (px // 16) * 5 * c * 2
        if self.kind == "dual":
            return px * (32 + 16 + c + 4 * (c // 2) + 2 * (c // 2)) * 2 + px * 4 + (px // 4) * 3 * c * 2 + (px // 16) * 2 * c * 2
        # stem operand + full-res maps (f0, x2, tmp, head) + half/quarter-res pyramids + guidance path
        full = px * (32 + 4 * c + 16 + 2 * 16) * 2 + px * 4
        return full + (px // 4) * 3 * 2 * c * 2 + (px // 16) * 2 * 4 * c * 2

    def micro_batch(self, h, w, limit):
        mb = max(1, min(_MAX_MICRO_BATCH, _MICRO_BATCH_BYTES // max(1, self.bytes_per_image(h, w))))
        return int(min(mb, limit))

    def _buf(self, name, shape, device, dtype=torch.bfloat16):
        key = (name, tuple(shape), str(device), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    def release_buffers(self):
        self._bufs.clear()

    # ------------------------------------------------------------------ execution
    def forward(self, x, out=None, index=None, n_dev=None, count=None):
        """x: NCHW fp32 CUDA batch.  Without routing arguments the whole batch is processed in order.  With
        (index int32[>=count], n_dev device int) only bucket rows index[0:*n_dev] are read and their outputs scattered
        to the same rows of `out`; `count` is the host-side upper bound of the bucket size (default: batch size)."""
        require_cuda(x, type(self.model).__name__)
        if self.model.training:
            return self._forward_train(x, out, index)
        x = x.contiguous()
        b, _, h, w = x.shape
        need = {"light": 1, "low_unet": 2}.get(self.kind, 4)
        if h % need or w % need:
            raise ValueError(f"{type(self.model).__name__}: H and W must be multiples of {need} on the B200 path (got {h}x{w})")
        if out is None:
            out = torch.empty_like(x)
        upper = b if count is None else int(count)
        S = self.specs()
        mb = self.micro_batch(h, w, upper)
        run = {"light": self._run_light, "low_unet": self._run_low_unet, "corun": self._run_corun,
               "dual": self._run_dual}.get(self.kind, self._run_unet)
        for start in range(0, upper, mb):
            run(S, x, out, index, n_dev, start, min(mb, upper - start), mb)
        return out

    def _forward_train(self, x, out, index):
        """train() mode: batch-statistics BatchNorm forward on a tape + kernel-built backward (training/autograd.py)."""
        from .training import autograd as _ag
        name = type(self.model).__name__
        if index is not None or out is not None:
            raise NotImplementedError(f"{name}: routed buckets are an inference path; train() mode takes a plain batch")
        b, _, h, w = x.shape
        need, wmin = (1, 16) if self.kind == "light" else ((2, 32) if self.kind == "low_unet" else (4, 64))
        if h % need or w % need or w < wmin:
            raise ValueError(f"{name}: train() mode needs H, W multiples of {need} and W >= {wmin} (got {h}x{w})")
        if self.train_cache is None:
            self.train_cache = _ag._WeightCache()
        return _ag.train_forward(self, x.contiguous())

    def _kw(self, n_dev, n_start, n):
        kw = {"n": n, "n_dev": n_dev, "n_start": n_start}
        if self.tune:
            kw["tune"] = self.tune
        return kw

    def _res(self, specs, f, tmp, kw, pool=False):
        """ResidualBlock (base_model.py:36-41): conv-bn-relu -> conv-bn -> += f -> relu, in place on f.
        pool=True (the block feeds an AttentionBlock): the second conv's epilogue also leaves the per-channel (sum, max) partials of
        the block output, returned (None when that launch does not produce them) — the attention's pool pass folds them instead of
        re-reading f."""
        ops.conv2d(specs[0], f, dst=tmp, **kw)
        if not pool or not POOL_FOLD:
            ops.conv2d(specs[1], tmp, dst=f, residual=f, **kw)
            return None
        dev = f.device
        _, part = ops.conv2d(specs[1], tmp, dst=f, residual=f, stats="pool",
                             stat_alloc=lambda shape: self._buf("pool_part", shape, dev, torch.float32), **kw)
        return part

    def _run_light(self, S, x, out, index, n_dev, n_start, n, cap):
        dev = x.device
        _, _, h, w = x.shape
        c = S["init"].cout_pad
        kw = self._kw(n_dev, n_start, n)
        x3 = self._buf("x3", (cap, h, w, 16), dev)
        f = self._buf("f", (cap, h, w, c), dev)
        t = self._buf("t", (cap, h, w, c), dev)
        ops.stem_pack(x, 3, 1, 16, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x3)
        ops.conv2d(S["init"], x3, dst=f, **kw)
        for specs in S["res"]:
            self._res(specs, f, t, kw)
        ops.conv2d(S["out0"], f, dst=t, **kw)
        ops.conv2d(S["out1"], t, epi=EPI_IMAGE,
                   image=dict(mode=IMG_BLEND, x=x, out=out, index=index, alpha=self.model.skip_alpha), **kw)

    def _run_low_unet(self, S, x, out, index, n_dev, n_start, n, cap):
        """LowIntensityDehazeModel.forward, low_intensity.py:96-115."""
        dev = x.device
        _, _, h, w = x.shape
        c, c2 = S["init"].cout_pad, S["down"].cout_pad
        kw = self._kw(n_dev, n_start, n)
        x3 = self._buf("x3", (cap, h, w, 16), dev)
        f0 = self._buf("f0", (cap, h, w, c), dev)
        f1 = self._buf("f1", (cap, h // 2, w // 2, c2), dev)
        t1 = self._buf("t1", (cap, h // 2, w // 2, c2), dev)
        up = self._buf("up", (cap, h, w, c), dev)
        t0 = self._buf("t0", (cap, h, w, c), dev)
        ops.stem_pack(x, 3, 1, 16, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x3)
        ops.conv2d(S["init"], x3, dst=f0, **kw)
        ops.conv2d(S["down"], f0, dst=f1, **kw)
        for specs in S["res"]:
            self._res(specs, f1, t1, kw)
        ops.conv2d(S["up"], f1, dst=up, **kw)
        ops.conv2d(S["out0"], up, f0, dst=t0, **kw)            # cat([up, init_features]) never materialised
        ops.conv2d(S["out1"], t0, dst=up, **kw)
        ops.conv2d(S["out2"], up, epi=EPI_IMAGE, image=dict(mode=IMG_RESIDUAL, x=x, out=out, index=index), **kw)

    def _run_corun(self, S, x, out, index, n_dev, n_start, n, cap):
        """COrunInspiredModel.forward, medium_intensity.py:170-190."""
        dev = x.device
        _, _, h, w = x.shape
        c = S["init"].cout_pad
        c2, c4 = S["scale2"].cout_pad, S["scale3"].cout_pad
        cf = S["fusion"].cout_pad
        kw = self._kw(n_dev, n_start, n)
        nd = dict(n=n, n_dev=n_dev, n_start=n_start)
        x7 = self._buf("x7", (cap, h, w, 32), dev)
        f0 = self._buf("f0", (cap, h, w, c), dev)
        fused = self._buf("fused", (cap, h, w, c + c2 + c4), dev)
        ops.stem_pack(x, 7, 3, 32, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x7)
        ops.conv2d(S["init"], x7, dst=f0, **kw)
        ops.conv2d(S["scale1"], f0, dst=fused, dst_c_off=0, **kw)
        p2 = ops.maxpool_kxk(f0, 2, out=self._buf("p2", (cap, h // 2, w // 2, c), dev), **nd)
        s2 = ops.conv2d(S["scale2"], p2, dst=self._buf("s2", (cap, h // 2, w // 2, c2), dev), **kw)
        ops.upsample_bilinear(s2, 2, out=fused, c_off=c, **nd)
        p4 = ops.maxpool_kxk(f0, 4, out=self._buf("p4", (cap, h // 4, w // 4, c), dev), **nd)
        s4 = ops.conv2d(S["scale3"], p4, dst=self._buf("s4", (cap, h // 4, w // 4, c4), dev), **kw)
        ops.upsample_bilinear(s4, 4, out=fused, c_off=c + c2, **nd)
        g = self._buf("g", (cap, h, w, cf), dev)
        t = self._buf("t", (cap, h, w, cf), dev)
        ops.conv2d(S["fusion"], fused, dst=g, **kw)
        for specs in S["res"]:
            self._res(specs, g, t, kw)
        ops.conv2d(S["out0"], g, dst=f0, **kw)
        ops.conv2d(S["out1"], f0, epi=EPI_IMAGE, image=dict(mode=IMG_RESIDUAL, x=x, out=out, index=index), **kw)

    def _run_dual(self, S, x, out, index, n_dev, n_start, n, cap):
        """DualBranchAttentionModel.forward, high_intensity.py:203-223."""
        dev = x.device
        _, _, h, w = x.shape
        c = S["g_init"].cout_pad
        ch = S["g_out"].cout_pad
        kw = self._kw(n_dev, n_start, n)
        nd = dict(n=n, n_dev=n_dev, n_start=n_start)
        scratch = self._bufs.setdefault(("attn_scratch", cap, str(dev)), {})

        def attend(ap, f):
            return ops.attention(f, ap, n=n, n_dev=n_dev, n_start=n_start, out=f, scratch=scratch)

        # global branch
        x7 = self._buf("x7", (cap, h, w, 32), dev)
        gf = self._buf("gf", (cap, h, w, c), dev)
        ops.stem_pack(x, 7, 3, 32, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x7)
        ops.conv2d(S["g_init"], x7, dst=gf, **kw)
        g1 = ops.maxpool_kxk(gf, 2, out=self._buf("g1", (cap, h // 2, w // 2, c), dev), **nd)
        t1 = self._buf("gt1", tuple(g1.shape), dev)
        self._res(S["g_res"][0], g1, t1, kw)
        attend(S["g_attn"][0], g1)
        g2 = ops.maxpool_kxk(g1, 2, out=self._buf("g2", (cap, h // 4, w // 4, c), dev), **nd)
        t2 = self._buf("gt2", tuple(g2.shape), dev)
        self._res(S["g_res"][1], g2, t2, kw)
        attend(S["g_attn"][1], g2)
        self._res(S["g_res"][2], g2, t2, kw)
        ops.upsample_bilinear(g2, 2, out=g1, **nd)
        self._res(S["g_res"][3], g1, t1, kw)
        ops.upsample_bilinear(g1, 2, out=gf, **nd)
        G = self._buf("G", (cap, h, w, ch), dev)
        ops.conv2d(S["g_out"], gf, dst=G, **kw)
        # local branch
        x3 = self._buf("x3", (cap, h, w, 16), dev)
        lf = self._buf("lf", (cap, h, w, ch), dev)
        lt = self._buf("lt", (cap, h, w, ch), dev)
        ops.stem_pack(x, 3, 1, 16, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x3)
        ops.conv2d(S["l_init"], x3, dst=lf, **kw)
        for specs in S["l_res"]:
            self._res(specs, lf, lt, kw)
        L = self._buf("L", (cap, h, w, ch), dev)
        ops.conv2d(S["l_out"], lf, dst=L, **kw)
        # transmission: (1 - t) as fp32 [n,h,w]; fusion: residual, out = clamp(x + (1 - t) * residual)
        ops.conv2d(S["t0"], G, L, dst=lt, **kw)                # cat([global, local]) never materialised
        one_minus_t = self._buf("omt", (cap, h, w), dev, torch.float32)
        dot_w, dot_b = S["t_dot"]
        ops.conv2d(S["t1"], lt, epi=EPI_DOT, dot=(dot_w, dot_b, one_minus_t), **kw)
        ops.conv2d(S["f0"], G, L, dst=lf, **kw)
        ops.conv2d(S["f1"], lf, epi=EPI_IMAGE, image=dict(mode=IMG_GUIDED, x=x, out=out, index=index, guidance=one_minus_t), **kw)

    def _run_unet(self, S, x, out, index, n_dev, n_start, n, cap):
        dev = x.device
        _, _, h, w = x.shape
        attn = self.kind == "unet_attn"
        c = S["init"].cout_pad
        kw = self._kw(n_dev, n_start, n)
        scratch = self._bufs.setdefault(("attn_scratch", cap, str(dev)), {})

        def attend(ap, f, part=None):
            return ops.attention(f, ap, n=n, n_dev=n_dev, n_start=n_start, out=f, scratch=scratch, pool_partials=part)

        guidance = None
        if attn:
            x3 = self._buf("x3", (cap, h, w, 16), dev)
            g0 = self._buf("g0", (cap, h, w, 16), dev)
            guidance = self._buf("guidance", (cap, h, w), dev, torch.float32)
            ops.stem_pack(x, 3, 1, 16, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x3)
            ops.conv2d(S["det0"], x3, dst=g0, **kw)
            dot_w, dot_b = S["det_dot"]
            ops.conv2d(S["det1"], g0, epi=EPI_DOT, dot=(dot_w, dot_b, guidance), **kw)

        x7 = self._buf("x7", (cap, h, w, 32), dev)
        f0 = self._buf("f0", (cap, h, w, c), dev)
        ops.stem_pack(x, 7, 3, 32, index=index, n_dev=n_dev, n_start=n_start, n=n, out=x7)
        ops.conv2d(S["init"], x7, dst=f0, **kw)

        feats = [f0]
        hh, ww, cc = h, w, c
        for li, e in enumerate(S["enc"]):
            hh, ww, cc = hh // 2, ww // 2, e["down"].cout_pad
            f = self._buf(f"f{li + 1}", (cap, hh, ww, cc), dev)
            t = self._buf(f"t{li + 1}", (cap, hh, ww, cc), dev)
            ops.conv2d(e["down"], feats[-1], dst=f, **kw)
            for ri, specs in enumerate(e["res"]):
                part = self._res(specs, f, t, kw, pool=attn and ri == len(e["res"]) - 1)
            if attn:
                attend(e["attn"], f, part)
            feats.append(f)

        # bottleneck works on a copy-free alias: feats[-1] is only needed as the bottleneck input
        bt = feats[-1]
        t2 = self._buf("t2", tuple(bt.shape), dev)
        for specs, ap in S["bott"]:
            part = self._res(specs, bt, t2, kw, pool=ap is not None)
            if ap is not None:
                attend(ap, bt, part)

        # decoder 0: up(bottleneck) -> res (-> attn); then cat with feats[1]
        d0 = S["dec"][0]
        x1 = self._buf("x1", tuple(feats[1].shape), dev)
        t1 = self._buf("t1", tuple(feats[1].shape), dev)
        ops.conv2d(d0["up"], bt, dst=x1, **kw)
        part = self._res(d0["res"], x1, t1, kw, pool=attn)
        if attn:
            attend(d0["attn"], x1, part)
        # decoder 1 reads cat([x1, feats[1]]) without materialising it
        d1 = S["dec"][1]
        x2 = self._buf("x2", tuple(f0.shape), dev)
        tf = self._buf("tf", tuple(f0.shape), dev)
        ops.conv2d(d1["up"], x1, feats[1], dst=x2, **kw)
        part = self._res(d1["res"], x2, tf, kw, pool=attn)
        if attn:
            attend(d1["attn"], x2, part)
        # head reads cat([x2, f0])
        ops.conv2d(S["out0"], x2, f0, dst=tf, **kw)
        r1 = self._buf("r1", (cap, h, w, S["out1"].cout_pad), dev)
        ops.conv2d(S["out1"], tf, dst=r1, **kw)
        img = dict(mode=IMG_GUIDED if attn else IMG_RESIDUAL, x=x, out=out, index=index, guidance=guidance)
        ops.conv2d(S["out2"], r1, epi=EPI_IMAGE, image=img, **kw)


class BlockEngine:
    """Stand-alone forwards of the building blocks (ConvBlock / ResidualBlock / AttentionBlock) on NCHW fp32 CUDA
    tensors — layout conversion on both sides, used for unit-level parity tests and ad-hoc composition."""

    def __init__(self, module):
        self.module = module
        self._ver = _Versioned(module)
        self._spec = None

    def _get(self, build):
        stale = self._ver.stale()
        if self._spec is None or stale:
            self._spec = build()
        return self._spec

    def conv_block(self, x):
        require_inference(self.module, "ConvBlock")
        spec = self._get(lambda: _conv_block_spec(self.module))
        cin = x.shape[1]
        if cin % 16:
            raise ValueError("ConvBlock on the B200 path needs in_channels to be a multiple of 16 (3-channel stems run "
                             "inside the branch models through adb_stem_pack)")
        y = ops.conv2d(spec, ops.nchw_to_nhwc(x, cin))
        return ops.nhwc_to_nchw(y, spec.cout)

    def residual_block(self, x):
        require_inference(self.module, "ResidualBlock")
        specs = self._get(lambda: _res_specs(self.module))
        f = ops.nchw_to_nhwc(x, x.shape[1])
        t = ops.conv2d(specs[0], f)
        y = ops.conv2d(specs[1], t, residual=f)
        return ops.nhwc_to_nchw(y, specs[1].cout)

    def attention_block(self, x):
        ap = self._get(lambda: _attn_params(self.module))
        y = ops.attention(ops.nchw_to_nhwc(x, x.shape[1]), ap)
        return ops.nhwc_to_nchw(y, x.shape[1])


class ResNetEngine:
    """torchvision resnet18/34 trunk (BasicBlock) + the reference head, models/classifier.py:24-36,72-97."""

    def __init__(self, classifier):
        self.clf = classifier
        self._ver = _Versioned(classifier)
        self.S = None
        self.train_cache = None

    def _forward_train(self, x):
        """train() mode (train_joint.py:117-121,141): batch-statistics BatchNorm, dropout in the head, kernel-built backward."""
        from .training import autograd as _ag
        b, _, h, w = x.shape
        if h % 32 or w % 32:
            raise ValueError(f"FogIntensityClassifier (B200 path): H and W must be multiples of 32, got {h}x{w}")
        if self.train_cache is None:
            self.train_cache = _ag._WeightCache()
        return _ag.train_forward_classifier(self, x.contiguous())

    def specs(self):
        stale = self._ver.stale()   # always evaluated: it records the signature the packed specs correspond to
        if self.S is not None and not stale:
            return self.S
        bb = self.clf.backbone
        S = {"stem": None, "layers": []}
        S["stem"] = _stem7x7s2_spec(bb.conv1, bb.bn1)
        for layer in (bb.layer1, bb.layer2, bb.layer3, bb.layer4):
            blocks = []
            for blk in layer:
                stride = blk.conv1.stride[0]
                c1 = ConvSpec.from_conv(blk.conv1.weight, bn=bn_args(blk.bn1), act=ACT_RELU, stride=stride, pad=1)
                c2 = ConvSpec.from_conv(blk.conv2.weight, bn=bn_args(blk.bn2), act=ACT_RELU, pad=1)
                ds = None
                if blk.downsample is not None:
                    ds = ConvSpec.from_conv(blk.downsample[0].weight, bn=bn_args(blk.downsample[1]), act=ACT_NONE,
                                            stride=blk.downsample[0].stride[0], pad=0)
                blocks.append((c1, c2, ds))
            S["layers"].append(blocks)
        head = self.clf.classifier
        S["head"] = tuple(t.detach().float().contiguous() for t in (head[1].weight, head[1].bias, head[4].weight, head[4].bias))
        self.S = S
        return S

    @staticmethod
    def chunk_size(b, h, w):
        """Images per trunk pass: bounds the activation footprint at 24 GB."""
        per_img = (h // 2) * (w // 2) * (32 + 2 * 128) + (h // 4) * (w // 4) * 3 * 128
        return max(1, min(b, (24 << 30) // max(1, per_img)))

    def forward(self, x, chunk=None):
        require_cuda(x, "FogIntensityClassifier")
        if self.clf.training:
            return self._forward_train(x)
        x = x.contiguous()
        b, _, h, w = x.shape
        if h % 32 or w % 32:
            raise ValueError(f"FogIntensityClassifier (B200 path): H and W must be multiples of 32, got {h}x{w}")
        S = self.specs()
        chunk = chunk or self.chunk_size(b, h, w)
        feats = torch.empty((b, self.clf.feature_dim), dtype=torch.float32, device=x.device)
        for s in range(0, b, chunk):
            n = min(chunk, b - s)
            xs = x[s:s + n]
            cols = _stem_operand(xs)
            f = ops.conv2d(S["stem"], cols)
            del cols
            f = ops.maxpool3x3s2(f)
            for blocks in S["layers"]:
                for (c1, c2, ds) in blocks:
                    t = ops.conv2d(c1, f)
                    idn = ops.conv2d(ds, f) if ds is not None else f
                    f = ops.conv2d(c2, t, residual=idn)
            feats[s:s + n] = ops.global_avgpool(f)
        w1, b1, w2, b2 = S["head"]
        return ops.head_mlp(feats, w1, b1, w2, b2), feats


def _stem7x7s2_spec(conv, bn):
    """7x7 stride-2 pad-3 3->C stem (torchvision conv1 / conv0 + norm + ReLU) as a 4x4-tap conv over the space-to-depth image
    (ADB_CONV_K4_S2D): the operand is 32 bytes per output pixel instead of the 320 of a full im2col (K = 147 -> 160), which
    was 18 % of HDEN's DRAM traffic."""
    return ConvSpec.from_stem_s2d(conv.weight, bn=bn_args(bn), act=ACT_RELU)


def _stem_operand(x):
    """[n,3,H,W] fp32 -> [n,H/2,W/2,16] bf16 space-to-depth operand (channel (py*2+px)*3 + c)."""
    return ops.stem_pack(x, 2, 0, 16, stride=2, kh=2)


class DenseNetEngine:
    """torchvision densenet121 trunk + the reference head (north_star HDEN; the reference itself ships no DenseNet arm).

    Dense blocks are concat-free: every layer's 3x3 conv stores its 32 new channels straight into the block's
    [n,h,w,C_total] buffer at its channel offset.  norm2/relu2 ride in conv1's epilogue; norm1/relu1 (a different affine
    of the same concat per layer) is applied to conv1's input operand in shared memory (adb_conv_desc.pre_scale/shift),
    so no normalised copy of the concat is ever written."""

    def __init__(self, classifier):
        self.clf = classifier
        self._ver = _Versioned(classifier)
        self.S = None
        self.train_cache = None

    _forward_train = ResNetEngine._forward_train

    @staticmethod
    def chunk_size(b, h, w):
        """Images per trunk pass: bounds the activation footprint at 24 GB."""
        per_img = (h // 2) * (w // 2) * (32 + 128) + (h // 4) * (w // 4) * 2 * (256 + 256 + 128)
        return max(1, min(b, (24 << 30) // max(1, per_img)))

    @staticmethod
    def _affine(bn):
        s, b = ops.fold_bn(bn.num_features, None, bn_args(bn), cout_pad=bn.num_features, device=bn.weight.device)
        return s, b

    def specs(self):
        stale = self._ver.stale()   # always evaluated: it records the signature the packed specs correspond to
        if self.S is not None and not stale:
            return self.S
        ft = self.clf.backbone.features
        S = {"stem": _stem7x7s2_spec(ft.conv0, ft.norm0), "blocks": [], "trans": []}
        for bi in range(4):
            layers = []
            for layer in getattr(ft, f"denseblock{bi + 1}").children():
                pre = self._affine(layer.norm1)
                c1 = ConvSpec.from_conv(layer.conv1.weight, bn=bn_args(layer.norm2), act=ACT_RELU, pad=0)
                c2 = ConvSpec.from_conv(layer.conv2.weight, act=ACT_NONE, pad=1)
                layers.append((pre, c1, c2))
            S["blocks"].append(layers)
            if bi < 3:
                tr = getattr(ft, f"transition{bi + 1}")
                S["trans"].append((self._affine(tr.norm), ConvSpec.from_conv(tr.conv.weight, act=ACT_NONE, pad=0)))
        S["final"] = self._affine(ft.norm5)
        head = self.clf.classifier
        S["head"] = tuple(t.detach().float().contiguous() for t in (head[1].weight, head[1].bias, head[4].weight, head[4].bias))
        self.S = S
        return S

    def forward(self, x, chunk=None):
        require_cuda(x, "FogIntensityClassifier")
        if self.clf.training:
            return self._forward_train(x)
        x = x.contiguous()
        b, _, h, w = x.shape
        if h % 32 or w % 32:
            raise ValueError(f"FogIntensityClassifier (B200 path): H and W must be multiples of 32, got {h}x{w}")
        S = self.specs()
        chunk = chunk or self.chunk_size(b, h, w)
        feats = torch.empty((b, self.clf.feature_dim), dtype=torch.float32, device=x.device)
        dev = x.device
        for s in range(0, b, chunk):
            n = min(chunk, b - s)
            cols = _stem_operand(x[s:s + n])
            f = ops.conv2d(S["stem"], cols)
            del cols
            c_in = f.shape[3]
            hh, ww = (f.shape[1] - 1) // 2 + 1, (f.shape[2] - 1) // 2 + 1
            pending = ("max", f)   # the op that produces the next block's input writes straight into its buffer
            for bi, layers in enumerate(S["blocks"]):
                c_total = c_in + 32 * len(layers)
                buf = torch.empty((n, hh, ww, c_total), dtype=torch.bfloat16, device=dev)
                if pending[0] == "max":
                    ops.maxpool3x3s2(pending[1], out=buf)
                else:
                    ops.avgpool2x2(pending[1], c=c_in, out=buf)
                pending = None
                c = c_in
                for (pre, c1, c2) in layers:
                    # norm1/relu1 ride inside conv1's operand path (adb_conv_desc.pre_*): the block buffer is read once
                    t = ops.conv2d(c1, buf, c0=c, pre=pre)
                    ops.conv2d(c2, t, dst=buf, dst_c_off=c)
                    c += 32
                if bi < 3:
                    pre, conv = S["trans"][bi]
                    pending = ("avg", ops.conv2d(conv, buf, c0=c, pre=pre))
                    c_in = conv.cout
                    hh, ww = hh // 2, ww // 2
                else:
                    a = ops.affine_relu(buf, c, S["final"][0], S["final"][1])
                    feats[s:s + n] = ops.global_avgpool(a)
        w1, b1, w2, b2 = S["head"]
        return ops.head_mlp(feats, w1, b1, w2, b2), feats
