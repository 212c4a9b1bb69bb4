"""Training-mode execution of the branch models: forward with batch-statistics BatchNorm, and a reverse-mode tape whose
backward is built from libadb200 kernels only (reference: model.train(); loss.backward() in
training/train_dehazing.py:66-92 and training/train_joint.py:129-150).

Data gradients re-use the tcgen05 implicit-GEMM forward kernel (adb_conv2d) with transformed weights:
  * k x k stride-1 conv  -> the same conv with the kernel rotated by 180 degrees and in/out channels swapped;
  * 4x4 stride-2 conv    -> ConvTranspose2d(4,2,1) with the weight as is (the kernel's four sub-pixel phases);
  * ConvTranspose2d(4,2,1) -> 4x4 stride-2 conv with the weight as is.
Weight gradients run on adb_wgrad (pixel-reduction GEMM), BatchNorm/activation/attention/head backward on the
HBM-bound kernels of csrc/train.cu.  Activations and their gradients are NHWC bf16; parameter gradients fp32 in the
parameter's own layout, handed to autograd by `BranchTrainFn` so optimizers/DDP hooks see ordinary `.grad`s.
"""
import ctypes as C

import os

import torch

from .. import _lib, ops
from ..ops import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, CONV_S1, CONV_S2, CONVT_4X4S2, IMG_BLEND, IMG_GUIDED,
                   IMG_RESIDUAL, WG_OIHW, WG_STEM, ConvSpec)


def _f32(n, dev):
    return torch.empty(n, dtype=torch.float32, device=dev)


class Node:
    """An activation on the tape: NHWC bf16 tensor (or fp32 [n,h,w] for the guidance map) + its accumulated gradient."""
    __slots__ = ("t", "c", "grad")

    def __init__(self, t, c=None):
        self.t, self.c, self.grad = t, (t.shape[-1] if c is None else c), None

    def accumulate(self, g):
        if self.grad is None:
            self.grad = g
        else:
            n, h, w, p = self.grad.shape
            _lib.call("adb_add_bf16", _lib.ptr(self.grad), p, _lib.ptr(g), g.shape[3], n * h * w, self.c, _lib.current_stream())

    def accumulate_conv(self, spec, src):
        """grad += conv(src) with the accumulation fused into the conv epilogue where the kernel allows it.
        `spec` may be a list of (channel offset, spec) pieces when the gradient has more channels than one launch's N range
        (DenseNet 1x1 convs over up to 1024 concatenated channels)."""
        if isinstance(spec, list):
            if self.grad is None:
                nb, h, w, _ = src.shape
                self.grad = torch.empty((nb, h, w, self.c), dtype=torch.bfloat16, device=src.device)
                for off, sp in spec:
                    ops.conv2d(sp, src, dst=self.grad, dst_c_off=off)
            else:
                for off, sp in spec:
                    ops.conv2d(sp, src, dst=self.grad, dst_c_off=off, residual=self.grad[..., off:off + sp.cout_pad])
            return
        if self.grad is None:
            self.grad = ops.conv2d(spec, src)
        elif spec.kind == CONVT_4X4S2:       # the sub-pixel store path has no residual input
            self.accumulate(ops.conv2d(spec, src))
        else:
            ops.conv2d(spec, src, dst=self.grad, residual=self.grad)


# Test hook (tests/test_gpu_train.py, same-mask gradient parity): when a dict is installed here, every conv of a train()-mode
# forward leaves its raw bf16 output under id(weight parameter), so the fp32 oracle can be differentiated AT the activations
# this implementation actually produced (identical ReLU / clamp masks).
_RECORD = None


def _rec(w, z):
    if _RECORD is not None:
        _RECORD[id(w)] = z


def _touch_bn_buffers(bn):
    """adb_bn_train_stats updates running_mean / running_var / num_batches_tracked through raw pointers; bump their torch
    version counters so the eval-path packings keyed on `_version` (engine._Versioned) re-fold the new statistics even when
    no optimizer step follows (train()-mode forward under no_grad, frozen modules, calibration passes)."""
    from .optim import _bump_versions
    _bump_versions([t for t in (bn.running_mean, bn.running_var, bn.num_batches_tracked) if t is not None])


class BlockBuffer:
    """DenseNet block buffer: the raw (pre-norm) features of a dense block, [n,h,w,C_total] bf16, written in place by the
    layers' 3x3 convs at their channel offsets (the concat is never materialised), plus its lazily zeroed gradient."""
    __slots__ = ("t", "c", "grad")

    def __init__(self, t):
        self.t, self.c, self.grad = t, t.shape[3], None

    def g(self):
        if self.grad is None:
            self.grad = torch.zeros_like(self.t)
        return self.grad


def _flat_specs(obj):
    if isinstance(obj, ConvSpec):
        return [obj]
    if isinstance(obj, (list, tuple)):
        out = []
        for o in obj:
            out += _flat_specs(o)
        return out
    return []


def derive_index_maps(build, weight):
    """Build a packing and, for each of its bf16 tensors, the int32 map `packed.flat[i] = weight.flat[idx[i]]` (-1 = zero
    padding): the same builder applied to an arange under ops.pack_as(float32) (pack functions only permute and pad)."""
    val = build(weight)
    with torch.no_grad(), ops.pack_as(torch.float32):
        ramp = torch.arange(1, weight.numel() + 1, dtype=torch.float32, device=weight.device).view(weight.shape)
        ival = build(ramp)
    maps = [(sp.w_packed, (isp.w_packed.reshape(-1).to(torch.int32) - 1).contiguous())
            for sp, isp in zip(_flat_specs(val), _flat_specs(ival))]
    maps += [(sp.w_fold, (isp.w_fold.reshape(-1).to(torch.int32) - 1).contiguous())
             for sp, isp in zip(_flat_specs(val), _flat_specs(ival)) if sp.w_fold is not None]
    return val, maps


class _WeightCache:
    """bf16 packings of a parameter for the forward and the data-gradient launches.  When parameters change (optimizer step)
    every packing built with `weight=` is refreshed through an index map derived once (the same builder applied to an arange
    under ops.pack_as(float32)) — ALL stale packings of the cache in ONE adb_gather_cast_multi launch, the first time any
    of them is asked for after the step; other entries are rebuilt."""

    def __init__(self):
        self._c = {}
        self._jobs = None          # (key tuple, device job table, total blocks) of the last multi-gather

    @staticmethod
    def _sig(params):
        return tuple((p.data_ptr(), p._version) for p in params if p is not None)

    def _refresh_stale(self):
        """Re-pack every map-backed entry whose parameters changed, in one launch."""
        stale = [(k, e) for k, e in self._c.items() if e["maps"] is not None and e["sig"] != self._sig(e["params"])
                 and e["wptr"] == e["weight"].data_ptr()]
        if not stale:
            return
        keys = tuple(k for k, _ in stale)
        if self._jobs is None or self._jobs[0] != keys:
            rows, block = [], 0
            for _, e in stale:
                for packed, idx in e["maps"]:
                    n = idx.numel()
                    rows.append([e["weight"].data_ptr(), idx.data_ptr(), packed.data_ptr(), n, block])
                    block += (n + 2047) // 2048
            dev = stale[0][1]["weight"].device
            self._jobs = (keys, torch.tensor(rows, dtype=torch.int64).to(dev), block, len(rows))
        _, table, blocks, njobs = self._jobs
        _lib.call("adb_gather_cast_multi", _lib.ptr(table), njobs, blocks, _lib.current_stream())
        for _, e in stale:
            if e["bias"] is not None:
                for sp in _flat_specs(e["val"]):
                    sp.shift[:e["bias"].numel()].copy_(e["bias"].detach())
            e["sig"] = self._sig(e["params"])

    def get(self, key, params, build, weight=None, bias=None):
        """build(): returns the packing (a ConvSpec, or nested lists/tuples of them).  With `weight` given, build takes
        the weight tensor as its only argument."""
        sig = self._sig(params)
        hit = self._c.get(key)
        if hit is not None and hit["sig"] == sig:
            return hit["val"]
        if hit is not None and hit["maps"] is not None and weight is not None and hit["wptr"] == weight.data_ptr():
            self._refresh_stale()
            hit = self._c[key]
            if hit["sig"] == sig:
                return hit["val"]
        if weight is None:
            hit = {"sig": sig, "val": build(), "maps": None, "wptr": None, "params": params, "weight": None, "bias": None}
        else:
            val, maps = derive_index_maps(build, weight)
            hit = {"sig": sig, "val": val, "maps": maps, "wptr": weight.data_ptr(), "params": params, "weight": weight, "bias": bias}
            self._jobs = None
        self._c[key] = hit
        return hit["val"]


def _pad_rows16(w):
    co = w.shape[0]
    cp = ops.pad16(co)
    if cp == co:
        return w
    out = torch.zeros((cp,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[:co] = w
    return out


def _embed4x4(w, pad):
    """A k x k stride-2 conv with padding `pad` reads x[2i + r - pad]; as a 4x4 pad-1 kernel the tap sits at r + 1 - pad
    (4x4 p1 -> itself, 3x3 p1 -> rows/cols 0..2, 1x1 p0 -> the single tap at (1, 1))."""
    co, ci, kh, kw = w.shape
    if kh == 4 and kw == 4 and pad == 1:
        return w
    o = 1 - pad
    if o < 0 or o + kh > 4 or o + kw > 4:
        raise NotImplementedError(f"stride-2 {kh}x{kw} pad {pad} conv has no data-gradient form on the B200 path")
    out = torch.zeros((co, ci, 4, 4), dtype=w.dtype, device=w.device)
    out[:, :, o:o + kh, o:o + kw] = w
    return out


def _dgrad_spec_s1(w):
    """W [co][ci][k][k] -> spec of dX = conv(dZ, rot180(W)^T); dZ carries pad16(co) channels."""
    wd = _pad_rows16(w.detach().float()).flip(2, 3).permute(1, 0, 2, 3).contiguous()
    return ConvSpec.from_conv(wd, stride=1, pad=w.shape[2] // 2)


class _PaddedBN:
    """A BatchNorm2d whose width is not a multiple of 16, seen through buffers padded to the conv's N (gamma = beta = 0 and
    unit running variance in the padding, so padded channels normalise to exactly 0)."""

    def __init__(self, bn, c):
        self.bn, self.eps, self.momentum = bn, bn.eps, bn.momentum
        co = bn.num_features
        dev = bn.weight.device

        def pad(t, fill):
            out = torch.full((c,), fill, dtype=torch.float32, device=dev)
            out[:co] = t.detach().float()
            return out
        self.weight, self.bias = pad(bn.weight, 0.0), pad(bn.bias, 0.0)
        self.running_mean, self.running_var = pad(bn.running_mean, 0.0), pad(bn.running_var, 1.0)
        self.num_batches_tracked = bn.num_batches_tracked

    def write_back(self):
        co = self.bn.num_features
        with torch.no_grad():
            self.bn.running_mean.copy_(self.running_mean[:co])
            self.bn.running_var.copy_(self.running_var[:co])


# BatchNorm statistics from the conv epilogue's partials where the launch produces them (False: always a pass over z)
EPILOGUE_STATS = os.environ.get("ADB_NO_EPILOGUE_STATS", "") == ""


class Tape:
    def __init__(self, model_cache):
        self.back = []
        self.pg = {}
        self.wc = model_cache

    # ------------------------------------------------------------------ helpers
    def _bn_forward(self, z, c, bn, act, residual=None, partials=None):
        """partials: the producing conv's epilogue statistics (ops.conv2d(stats=True)); None -> one pass over z."""
        n, h, w, pitch = z.shape
        dev = z.device
        px = n * h * w
        stats = _f32(4 * c, dev)
        mean, rstd, scale, shift = stats[:c], stats[c:2 * c], stats[2 * c:3 * c], stats[3 * c:]
        st = _lib.current_stream()
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        if partials is not None:
            slots = partials.shape[0] * partials.shape[1]
            scratch = _f32(int(_lib.load().adb_bn_stat_scratch_floats(slots, c)), dev)
            _lib.call("adb_bn_finalize_stats", _lib.ptr(partials), slots, partials.shape[3], px, c, _lib.ptr(bn.weight), _lib.ptr(bn.bias),
                      float(bn.eps), mom, _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.ptr(bn.num_batches_tracked),
                      _lib.ptr(scratch), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(scale), _lib.ptr(shift), st)
        else:
            scratch = _f32(int(_lib.load().adb_bn_scratch_floats(px, c)), dev)
            _lib.call("adb_bn_train_stats", _lib.ptr(z), px, c, pitch, _lib.ptr(bn.weight), _lib.ptr(bn.bias), float(bn.eps), mom,
                      _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.ptr(bn.num_batches_tracked), _lib.ptr(scratch),
                      _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(scale), _lib.ptr(shift), st)
        _touch_bn_buffers(bn)
        y = torch.empty_like(z)
        _lib.call("adb_affine_act", _lib.ptr(z), pitch, px, c, _lib.ptr(scale), _lib.ptr(shift),
                  _lib.ptr(residual), 0 if residual is None else residual.shape[3], act, _lib.ptr(y), pitch, st)
        return y, stats

    def _bn_backward(self, dy, y, z, c, act, bn, stats, keep_g):
        """Returns (g, dz).  g = dy*act'(y) (written over dy); dz aliases g unless keep_g.  `stats` = the forward's
        [mean | rstd | scale | shift].  ReLU layers whose g nobody else needs take adb_bn_relu_bwd (mask recomputed from z:
        neither y is read nor g written) and return (None, dz)."""
        n, h, w, pitch = dy.shape
        px = n * h * w
        dev = dy.device
        mean, rstd, scale, shift = stats[:c], stats[c:2 * c], stats[2 * c:3 * c], stats[3 * c:]
        scratch = _f32(int(_lib.load().adb_bn_scratch_floats(px, c)), dev)
        dgamma, dbeta = _f32(c, dev), _f32(c, dev)
        if act == ACT_RELU and not keep_g:
            _lib.call("adb_bn_relu_bwd", _lib.ptr(dy), pitch, _lib.ptr(z), z.shape[3], px, c, _lib.ptr(scale), _lib.ptr(shift),
                      _lib.ptr(bn.weight), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(scratch), _lib.ptr(dy), pitch, 0,
                      _lib.ptr(dgamma), _lib.ptr(dbeta), 0, _lib.current_stream())
            self.pg[bn.weight], self.pg[bn.bias] = dgamma, dbeta
            return None, dy
        dz = torch.empty_like(dy) if keep_g else dy
        _lib.call("adb_bn_bwd", _lib.ptr(dy), pitch, _lib.ptr(y), 0 if y is None else y.shape[3], _lib.ptr(z), z.shape[3], px, c, act,
                  _lib.ptr(bn.weight), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(scratch), _lib.ptr(dy), pitch, _lib.ptr(dz), pitch,
                  _lib.ptr(dgamma), _lib.ptr(dbeta), 0, _lib.current_stream())
        self.pg[bn.weight], self.pg[bn.bias] = dgamma, dbeta
        return dy, dz

    # ------------------------------------------------------------------ layers
    def stem(self, x, kw, pad, kp):
        return Node(ops.stem_pack(x, kw, pad, kp))

    def conv_bn_act(self, conv, bn, act, srcs, residual=None, stem_kp=None):
        """nn.Conv2d (stride 1 'same' or 4x4 stride 2) [+bias] -> BatchNorm2d(train) -> act, optional residual add before
        the activation (ResidualBlock, base_model.py:36-41).  srcs: one or two Nodes read as a channel concat."""
        w, b = conv.weight, conv.bias
        stride = conv.stride[0]
        if stem_kp:
            fspec = self.wc.get(("f", id(conv)), (w, b), lambda wt: ConvSpec.from_stem(wt, stem_kp, bias=b), weight=w, bias=b)
        else:
            fspec = self.wc.get(("f", id(conv)), (w, b), lambda wt: ConvSpec.from_conv(wt, bias=b, stride=stride, pad=conv.padding[0]),
                                weight=w, bias=b)
        a = srcs[0]
        bsrc = srcs[1] if len(srcs) > 1 else None
        r = ops.conv2d(fspec, a.t, None if bsrc is None else bsrc.t, c0=a.c, c1=None if bsrc is None else bsrc.c, stats=EPILOGUE_STATS)
        z, part = r if EPILOGUE_STATS else (r, None)
        _rec(w, z)
        c = fspec.cout_pad
        bnp = bn if c == fspec.cout else _PaddedBN(bn, c)     # e.g. 24 channels inside a 32-channel map: padded affine
        y, stats = self._bn_forward(z, c, bnp, act, None if residual is None else residual.t, partials=part)
        if bnp is not bn:
            bnp.write_back()
        out = Node(y, c)

        def backward():
            dy = out.grad
            out.grad = None
            g, dz = self._bn_backward(dy, y, z, c, act, bnp, stats, keep_g=residual is not None)
            if bnp is not bn:
                co = fspec.cout
                self.pg[bn.weight], self.pg[bn.bias] = self.pg.pop(bnp.weight)[:co].contiguous(), self.pg.pop(bnp.bias)[:co].contiguous()
            if residual is not None:
                residual.accumulate(g)
            self._conv_backward(conv, dz, c, srcs, stem_kp, cz_true=None if c == fspec.cout else fspec.cout)
        self.back.append(backward)
        return out

    def _conv_backward(self, conv, dz, cz, srcs, stem_kp=None, cz_true=None):
        """Weight gradient + data gradients of a Conv2d given dL/d(conv output) `dz` (NHWC bf16, cz channels)."""
        w = conv.weight
        co, ci, kh, kw = w.shape
        stride = conv.stride[0]
        a = srcs[0]
        bsrc = srcs[1] if len(srcs) > 1 else None
        if stem_kp:
            self.pg[w] = ops.wgrad(dz, a.t, kh=kh, kw=1, pad=kh // 2, cs=cz, cs_true=cz_true, layout=WG_STEM, stem_kw=kw)
            return
        kind = CONV_S1 if stride == 1 else CONV_S2
        self.pg[w] = ops.wgrad(dz, a.t, None if bsrc is None else bsrc.t, kind=kind, kh=kh, kw=kw, pad=conv.padding[0], cs=cz,
                               cs_true=cz_true, c0=a.c, c1=None if bsrc is None else bsrc.c)
        off = 0
        for i, s in enumerate(srcs):
            lo, hi = off, off + s.c
            off = hi
            if stride == 1 and hi - lo > 512:      # more gradient channels than one launch's N range: 256-channel pieces
                spec = self.wc.get(("d", id(conv), i), (w,), lambda wt, lo=lo, hi=hi: [
                    (o - lo, _dgrad_spec_s1(wt[:, o:min(o + 256, hi)])) for o in range(lo, hi, 256)], weight=w)
            elif stride == 1:
                spec = self.wc.get(("d", id(conv), i), (w,), lambda wt, lo=lo, hi=hi: _dgrad_spec_s1(wt[:, lo:hi]), weight=w)
            else:   # stride-2 conv: dX = ConvTranspose2d(dZ, W) — the 4x4/pad-1 sub-pixel kernel, smaller filters embedded
                spec = self.wc.get(("d", id(conv), i), (w,), lambda wt, lo=lo, hi=hi: ConvSpec.from_convT(
                    _embed4x4(_pad_rows16(wt.detach()[:, lo:hi]), conv.padding[0])), weight=w)
            s.accumulate_conv(spec, dz)

    def convT_bn_act(self, convT, bn, act, srcs):
        """nn.ConvTranspose2d(4,2,1) + bias -> BatchNorm2d(train) -> act (decoder `up`, medium:53-55,63-65; high:57-59,68-70)."""
        w, b = convT.weight, convT.bias
        fspec = self.wc.get(("f", id(convT)), (w, b), lambda wt: ConvSpec.from_convT(wt, bias=b), weight=w, bias=b)
        a = srcs[0]
        bsrc = srcs[1] if len(srcs) > 1 else None
        z = ops.conv2d(fspec, a.t, None if bsrc is None else bsrc.t, c0=a.c, c1=None if bsrc is None else bsrc.c)
        _rec(w, z)
        c = fspec.cout_pad
        y, stats = self._bn_forward(z, c, bn, act)
        out = Node(y, c)

        def backward():
            dy = out.grad
            out.grad = None
            _, dz = self._bn_backward(dy, y, z, c, act, bn, stats, keep_g=False)
            gw = torch.empty_like(w, dtype=torch.float32)
            off = 0
            for i, s in enumerate(srcs):
                lo, hi = off, off + s.c
                off = hi
                # dWt[ci][co][r][s]: the stride-2 form with the maps swapped (small = layer input, large = dz)
                ops.wgrad(s.t, dz, kind=CONV_S2, kh=4, kw=4, pad=1, cs=s.c, c0=c, out=gw[lo:hi])
                spec = self.wc.get(("d", id(convT), i), (w,), lambda wt, lo=lo, hi=hi: ConvSpec.from_conv(wt.detach()[lo:hi], stride=2, pad=1),
                                   weight=w)
                s.accumulate_conv(spec, dz)
            self.pg[w] = gw
            if b is not None:
                # the bias feeds a batch-statistics BatchNorm: its gradient is sum(dz) == 0 identically
                self.pg[b] = torch.zeros_like(b, dtype=torch.float32)
        self.back.append(backward)
        return out

    def stem_full_bn_act(self, conv, bn, act, x):
        """7x7 stride-2 3->C stem (torchvision resnet conv1/bn1/relu) over the full-im2col operand (K = 147 -> 160)."""
        w = conv.weight
        co = w.shape[0]

        def build(wt):
            wp = torch.zeros(ops.pad16(co), 160, dtype=ops.pack_dtype(), device=wt.device)
            wp[:co, :147] = wt.detach().float().permute(0, 2, 3, 1).reshape(co, 147).to(ops.pack_dtype())
            scale, shift = ops.fold_bn(co, None, None, device=wt.device)
            return ConvSpec(CONV_S1, 1, 1, 0, co, wp.contiguous(), scale, shift, ACT_NONE)
        fspec = self.wc.get(("f", id(conv)), (w,), build, weight=w)
        cols = ops.stem_pack(x, 7, 3, 160, stride=2, kh=7)
        z = ops.conv2d(fspec, cols)
        c = fspec.cout_pad
        y, stats = self._bn_forward(z, c, bn, act)
        out = Node(y, c)

        def backward():
            dy = out.grad
            out.grad = None
            _, dz = self._bn_backward(dy, y, z, c, act, bn, stats, keep_g=False)
            self.pg[w] = ops.wgrad(dz, cols, kh=1, kw=1, pad=0, cs=c, layout=WG_STEM, stem_kw=49).view_as(w)
        self.back.append(backward)
        return out

    def maxpool(self, x, k, stride, pad):
        n, h, w, c = x.t.shape
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        y = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.t.device)
        _lib.call("adb_maxpool_fwd", _lib.ptr(x.t), n, h, w, c, k, stride, pad, _lib.ptr(y), _lib.current_stream())
        out = Node(y, x.c)

        def backward():
            dy = out.grad
            out.grad = None
            dx = torch.empty_like(x.t)
            _lib.call("adb_maxpool_bwd", _lib.ptr(dy), _lib.ptr(x.t), _lib.ptr(y), n, h, w, c, k, stride, pad, _lib.ptr(dx),
                      _lib.current_stream())
            x.accumulate(dx)
        self.back.append(backward)
        return out

    def global_avgpool(self, x):
        """[n,h,w,c] bf16 -> fp32 [n,c]; its gradient arrives through `head_backward` (dfeat)."""
        n, h, w, c = x.t.shape
        feats = ops.global_avgpool(x.t)
        out = Node(feats, c)

        def backward():
            df = out.grad
            out.grad = None
            dx = torch.empty_like(x.t)
            _lib.call("adb_broadcast_hw", _lib.ptr(df), n, h, w, c, 1.0 / (h * w), _lib.ptr(dx), _lib.current_stream())
            x.accumulate(dx)
        self.back.append(backward)
        return out

    def head_mlp(self, head, feats):
        """Dropout -> Linear -> ReLU -> Dropout -> Linear (classifier.py:72-78) in train() mode; masks from torch's RNG."""
        f = feats.t
        n, fin = f.shape
        dev = f.device
        st = _lib.current_stream()
        p1, p2 = float(head[0].p), float(head[3].p)
        l1, l2 = head[1], head[4]
        m1 = (torch.rand((n, fin), device=dev) >= p1).float() / (1.0 - p1)
        m2 = (torch.rand((n, l1.out_features), device=dev) >= p2).float() / (1.0 - p2)
        x0 = torch.empty_like(f)
        _lib.call("adb_mul_f32", _lib.ptr(f), _lib.ptr(m1), None, f.numel(), _lib.ptr(x0), st)
        h = ops.linear(x0, l1.weight, l1.bias, relu=True)
        h2 = torch.empty_like(h)
        _lib.call("adb_mul_f32", _lib.ptr(h), _lib.ptr(m2), None, h.numel(), _lib.ptr(h2), st)
        logits = ops.linear(h2, l2.weight, l2.bias, relu=False)

        def backward(dlogits, dfeat_extra):
            dh2 = torch.empty_like(h2)
            dw2, db2 = torch.empty_like(l2.weight), torch.empty_like(l2.bias)
            _lib.call("adb_linear_bwd", _lib.ptr(h2), _lib.ptr(l2.weight.detach()), _lib.ptr(dlogits), n, l1.out_features, l2.out_features,
                      _lib.ptr(dh2), _lib.ptr(dw2), _lib.ptr(db2), st)
            dh = torch.empty_like(h)
            _lib.call("adb_mul_f32", _lib.ptr(dh2), _lib.ptr(m2), _lib.ptr(h), h.numel(), _lib.ptr(dh), st)     # dropout + ReLU gate
            dx0 = torch.empty_like(x0)
            dw1, db1 = torch.empty_like(l1.weight), torch.empty_like(l1.bias)
            _lib.call("adb_linear_bwd", _lib.ptr(x0), _lib.ptr(l1.weight.detach()), _lib.ptr(dh), n, fin, l1.out_features,
                      _lib.ptr(dx0), _lib.ptr(dw1), _lib.ptr(db1), st)
            df = torch.empty_like(f)
            _lib.call("adb_mul_f32", _lib.ptr(dx0), _lib.ptr(m1), None, f.numel(), _lib.ptr(df), st)
            if dfeat_extra is not None:
                df = df + dfeat_extra          # gradient through the returned features (GatedRouter)
            self.pg[l1.weight], self.pg[l1.bias], self.pg[l2.weight], self.pg[l2.bias] = dw1, db1, dw2, db2
            feats.grad = df
        self.head_backward = backward
        return logits

    # ------------------------------------------------------------------ DenseNet pieces (torchvision densenet121 as HDEN)
    def bn_act_prefix(self, B, c, bn, act):
        """norm/relu over the first c channels of a block buffer -> a dense [n,h,w,c] map (norm1/relu1 of a dense layer,
        the transition norm, norm5).  Backward adds dz into the buffer gradient's channel prefix."""
        n, h, w, pitch = B.t.shape
        px = n * h * w
        dev = B.t.device
        st = _lib.current_stream()
        scratch = _f32(int(_lib.load().adb_bn_scratch_floats(px, c)), dev)
        stats = _f32(4 * c, dev)
        mean, rstd, scale, shift = stats[:c], stats[c:2 * c], stats[2 * c:3 * c], stats[3 * c:]
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        _lib.call("adb_bn_train_stats", _lib.ptr(B.t), px, c, pitch, _lib.ptr(bn.weight), _lib.ptr(bn.bias), float(bn.eps), mom,
                  _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.ptr(bn.num_batches_tracked), _lib.ptr(scratch),
                  _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(scale), _lib.ptr(shift), st)
        _touch_bn_buffers(bn)
        y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dev)
        _lib.call("adb_affine_act", _lib.ptr(B.t), pitch, px, c, _lib.ptr(scale), _lib.ptr(shift), None, 0, act, _lib.ptr(y), c, st)
        out = Node(y, c)

        def backward():
            dy = out.grad
            out.grad = None
            sc2 = _f32(int(_lib.load().adb_bn_scratch_floats(px, c)), dev)
            dgamma, dbeta = _f32(c, dev), _f32(c, dev)
            if act == ACT_RELU:     # mask from z, dz added straight into the buffer gradient's prefix
                _lib.call("adb_bn_relu_bwd", _lib.ptr(dy), c, _lib.ptr(B.t), pitch, px, c, _lib.ptr(scale), _lib.ptr(shift),
                          _lib.ptr(bn.weight), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(sc2), _lib.ptr(B.g()), pitch, 1,
                          _lib.ptr(dgamma), _lib.ptr(dbeta), 0, st)
                self.pg[bn.weight], self.pg[bn.bias] = dgamma, dbeta
                return
            _lib.call("adb_bn_bwd", _lib.ptr(dy), c, _lib.ptr(y), c, _lib.ptr(B.t), pitch, px, c, act, _lib.ptr(bn.weight), _lib.ptr(mean),
                      _lib.ptr(rstd), _lib.ptr(sc2), _lib.ptr(dy), c, _lib.ptr(dy), c, _lib.ptr(dgamma), _lib.ptr(dbeta), 0, st)
            self.pg[bn.weight], self.pg[bn.bias] = dgamma, dbeta
            _lib.call("adb_add_bf16", _lib.ptr(B.g()), pitch, _lib.ptr(dy), c, px, c, st)
        self.back.append(backward)
        return out

    def conv_plain(self, conv, src, into=None, c_off=0):
        """A bias-free conv with no norm / activation after it (DenseNet conv2 / transition conv).  `into`: write the output
        at channel offset c_off of a BlockBuffer (its gradient is read from the buffer gradient's slice)."""
        w = conv.weight
        co = w.shape[0]
        fspec = self.wc.get(("f", id(conv)), (w,), lambda wt: ConvSpec.from_conv(wt, stride=conv.stride[0], pad=conv.padding[0]), weight=w)
        if into is None:
            out = Node(ops.conv2d(fspec, src.t, c0=src.c), co)
        else:
            ops.conv2d(fspec, src.t, c0=src.c, dst=into.t, dst_c_off=c_off)
            out = None

        def backward():
            if into is None:
                dz = out.grad
                out.grad = None
            else:
                dz = into.g()[..., c_off:c_off + co]
            self._conv_backward(conv, dz, co, [src])
        self.back.append(backward)
        return out

    def avgpool_into(self, x, B, c):
        n, h, w, _ = x.t.shape
        ops.avgpool2x2(x.t, c=c, out=B.t)

        def backward():
            dx = torch.empty_like(x.t)
            _lib.call("adb_avgpool2x2_bwd", _lib.ptr(B.g()), B.t.shape[3], n, h, w, c, _lib.ptr(dx), x.t.shape[3], _lib.current_stream())
            x.accumulate(dx)
        self.back.append(backward)

    def maxpool3x3s2_into(self, x, B, c):
        n, h, w, _ = x.t.shape
        ops.maxpool3x3s2(x.t, out=B.t)

        def backward():
            dy = B.g()[..., :c].contiguous()
            y = B.t[..., :c].contiguous()
            dx = torch.empty_like(x.t)
            _lib.call("adb_maxpool_bwd", _lib.ptr(dy), _lib.ptr(x.t), _lib.ptr(y), n, h, w, c, 3, 2, 1, _lib.ptr(dx), _lib.current_stream())
            x.accumulate(dx)
        self.back.append(backward)

    def upsample(self, x, scale):
        """nn.UpsamplingBilinear2d(scale_factor=scale) (align_corners=True)."""
        n, h, w, c = x.t.shape
        y = ops.upsample_bilinear(x.t, scale)
        out = Node(y, x.c)

        def backward():
            dy = out.grad
            out.grad = None
            dx = torch.empty_like(x.t)
            _lib.call("adb_upsample_bilinear_bwd", _lib.ptr(dy), dy.shape[3], 0, n, h, w, c, scale, _lib.ptr(dx), _lib.current_stream())
            x.accumulate(dx)
        self.back.append(backward)
        return out

    def concat(self, nodes):
        """torch.cat(..., dim=1) of more than two maps (COrunInspiredModel's three scales, medium_intensity.py:184)."""
        y = torch.cat([nd.t[..., :nd.c] for nd in nodes], dim=3).contiguous()
        out = Node(y, y.shape[3])

        def backward():
            dy = out.grad
            out.grad = None
            off = 0
            for nd in nodes:
                nd.accumulate(dy[..., off:off + nd.c].contiguous())
                off += nd.c
        self.back.append(backward)
        return out

    def res_block(self, rb, x):
        t = self.conv_bn_act(rb.conv1.block[0], rb.conv1.block[1], ACT_RELU, [x])
        return self.conv_bn_act(rb.conv2.block[0], rb.conv2.block[1], ACT_RELU, [t], residual=x)

    def conv_block(self, cb, srcs, stem_kp=None):
        mods = list(cb.block)
        act = ACT_RELU if any(isinstance(m, torch.nn.ReLU) for m in mods) else ACT_NONE
        if len(mods) < 2 or not isinstance(mods[1], torch.nn.BatchNorm2d):
            raise NotImplementedError("training a ConvBlock without BatchNorm is not part of the default branch models")
        return self.conv_bn_act(mods[0], mods[1], act, srcs, stem_kp=stem_kp)

    def attention(self, ab, x):
        ap = self.wc.get(("attn", id(ab)), (ab.fc[0].weight, ab.fc[2].weight, ab.conv_spatial.weight),
                         lambda: ops.AttnParams(ab.fc[0].weight, ab.fc[2].weight, ab.conv_spatial.weight))
        keep = {}
        y = ops.attention(x.t, ap, scratch=keep)
        out = Node(y, x.c)
        n, h, w, c = x.t.shape

        def backward():
            dy = out.grad
            out.grad = None
            dev = dy.device
            pool = keep[("pool", n, h, w, c)]
            gate, stats, spatial = keep[("gate", n, c)], keep[("stats", n, h, w)], keep[("spatial", n, h, w)]
            scratch = _f32(int(_lib.load().adb_attn_bwd_scratch_floats(n, h, w, c)), dev)
            dx = torch.empty_like(dy)
            dw1, dw2, dws = _f32(ap.c_red * c, dev), _f32(c * ap.c_red, dev), _f32(98, dev)
            _lib.call("adb_attn_bwd", _lib.ptr(dy), _lib.ptr(x.t), n, h, w, c, _lib.ptr(pool), _lib.ptr(gate), _lib.ptr(stats),
                      _lib.ptr(spatial), _lib.ptr(ap.w1), _lib.ptr(ap.w2), ap.c_red, _lib.ptr(ap.wsp), _lib.ptr(scratch), _lib.ptr(dx),
                      _lib.ptr(dw1), _lib.ptr(dw2), _lib.ptr(dws), _lib.current_stream())
            self.pg[ab.fc[0].weight] = dw1.view_as(ab.fc[0].weight)
            self.pg[ab.fc[2].weight] = dw2.view_as(ab.fc[2].weight)
            self.pg[ab.conv_spatial.weight] = dws.view_as(ab.conv_spatial.weight)
            x.accumulate(dx)
        self.back.append(backward)
        return out

    def dot_head(self, conv1x1, src, complement=False):
        """nn.Conv2d(c, 1, 1) + Sigmoid -> fp32 [n,h,w] guidance map (detail_branch tail, high:87-89).
        complement: returns 1 - sigmoid(.) = sigmoid(-.) (DualBranchAttentionModel's (1 - transmission), high:217-221)."""
        n, h, w, pitch = src.t.shape
        c = src.c
        sign = -1.0 if complement else 1.0
        cw = conv1x1.weight.numel()
        wv = torch.zeros(c, dtype=torch.float32, device=src.t.device)
        wv[:cw] = sign * conv1x1.weight.detach().float().reshape(-1)
        bv = (sign * conv1x1.bias.detach().float()).reshape(-1).contiguous()
        g = torch.empty((n, h, w), dtype=torch.float32, device=src.t.device)
        _lib.call("adb_dot_head_fwd", _lib.ptr(src.t), pitch, c, _lib.ptr(wv), _lib.ptr(bv), n * h * w, _lib.ptr(g), _lib.current_stream())
        out = Node(g, 1)

        def backward():
            dg = out.grad
            out.grad = None
            dy = torch.empty_like(src.t)
            red = _f32(c + 1, g.device)
            _lib.call("adb_dot_head_bwd", _lib.ptr(dg), _lib.ptr(g), _lib.ptr(src.t), pitch, c, _lib.ptr(wv), n * h * w, _lib.ptr(dy),
                      pitch, _lib.ptr(red), _lib.current_stream())
            self.pg[conv1x1.weight] = (sign * red[:cw]).view_as(conv1x1.weight)
            self.pg[conv1x1.bias] = (sign * red[c:c + 1]).view_as(conv1x1.bias)
            src.accumulate(dy)
        self.back.append(backward)
        return out

    def image_head(self, conv, src, x, mode, act, guidance=None, alpha=None):
        """Final nn.Conv2d(c, 3, 3, padding=1) + act + the output arithmetic -> NCHW fp32 (low:45, medium:117, high:135-138)."""
        w, b = conv.weight, conv.bias
        fspec = self.wc.get(("f", id(conv)), (w, b), lambda wt: ConvSpec.from_conv(wt, bias=b, pad=conv.padding[0]), weight=w, bias=b)
        z = ops.conv2d(fspec, src.t, c0=src.c)
        _rec(w, z)
        n, h, wd, pitch = z.shape
        out = torch.empty_like(x)
        gptr = None if guidance is None else guidance.t
        aval = None if alpha is None else alpha.detach().float().reshape(1)
        _lib.call("adb_img_head_fwd", _lib.ptr(z), pitch, _lib.ptr(x), _lib.ptr(gptr), _lib.ptr(aval), mode, act, n, h, wd,
                  _lib.ptr(out), _lib.current_stream())

        def backward(dout):
            dz = torch.empty_like(z)
            red = _f32(4, z.device)
            dgd = torch.empty_like(guidance.t) if guidance is not None else None
            _lib.call("adb_img_head_bwd", _lib.ptr(dout), _lib.ptr(z), pitch, _lib.ptr(x), _lib.ptr(gptr), _lib.ptr(aval), mode, act,
                      n, h, wd, _lib.ptr(dz), _lib.ptr(dgd), _lib.ptr(red), _lib.current_stream())
            self.pg[b] = red[:3].view_as(b)
            if alpha is not None:
                self.pg[alpha] = red[3:4].reshape(alpha.shape)
            if guidance is not None:
                guidance.grad = dgd
            self._conv_backward(conv, dz, pitch, [src], cz_true=3)
        self.head_backward = backward
        return out

    # ------------------------------------------------------------------ reverse sweep
    def _sweep(self):
        """Run the recorded closures last-to-first, dropping each one (and with it the activations it saved) as soon as it
        has run, and leave no reference cycle behind (closures hold the tape): the step's memory goes back to the caching
        allocator by reference count, in the same order every step, instead of whenever the cycle collector next runs —
        with cycles left over, the allocator's free lists differ from step to step and a step now and then stalls for
        hundreds of ms in cudaMalloc."""
        self.head_backward = None
        back, self.back = self.back, []
        while back:
            back.pop()()
        pg, self.pg = self.pg, {}
        return pg

    def backward_classifier(self, dlogits, dfeats):
        self.head_backward(dlogits, dfeats)
        return self._sweep()

    def backward(self, dout):
        self.head_backward(dout)
        return self._sweep()


# ---------------------------------------------------------------------- branch forwards (train mode)
def forward_light(t, m, x):
    """LightweightDehazeModel.forward, low_intensity.py:33-45."""
    f = t.conv_block(m.init_conv, [t.stem(x, 3, 1, 16)], stem_kp=16)
    for rb in m.residual_blocks:
        f = t.res_block(rb, f)
    f = t.conv_block(m.output_conv[0], [f])
    return t.image_head(m.output_conv[1], f, x, IMG_BLEND, ACT_SIGMOID, alpha=m.skip_alpha)


def forward_low_unet(t, m, x):
    """LowIntensityDehazeModel.forward (non-default Light variant), low_intensity.py:96-115."""
    f0 = t.conv_block(m.init_conv, [t.stem(x, 3, 1, 16)], stem_kp=16)
    f = t.conv_block(m.down1[0], [f0])
    f = t.res_block(m.down1[1], f)
    for rb in m.bottleneck:
        f = t.res_block(rb, f)
    up = t.convT_bn_act(m.up1[0], m.up1[1], ACT_RELU, [f])
    r = t.conv_block(m.output_conv[0], [up, f0])
    r = t.conv_block(m.output_conv[1], [r])
    # clamp(x + (sigmoid(z) - 0.5) * 2): the head's activation is 2*sigmoid(z) - 1
    return t.image_head(m.output_conv[2], r, x, IMG_RESIDUAL, _lib.ACT_SIGMOID2)


def forward_corun(t, m, x):
    """COrunInspiredModel.forward (non-default Medium variant), medium_intensity.py:170-190."""
    f0 = t.conv_block(m.init_conv, [t.stem(x, 7, 3, 32)], stem_kp=32)
    s1 = t.conv_block(m.scale1_conv, [f0])
    s2 = t.upsample(t.conv_block(m.scale2_conv[1], [t.maxpool(f0, 2, 2, 0)]), 2)
    s3 = t.upsample(t.conv_block(m.scale3_conv[1], [t.maxpool(f0, 4, 4, 0)]), 4)
    g = t.conv_block(m.fusion_conv, [t.concat([s1, s2, s3])])
    for rb in m.residual_blocks:
        g = t.res_block(rb, g)
    r = t.conv_block(m.output_conv[0], [g])
    return t.image_head(m.output_conv[1], r, x, IMG_RESIDUAL, ACT_TANH)


def forward_dual(t, m, x):
    """DualBranchAttentionModel.forward (non-default Complex variant), high_intensity.py:203-223."""
    gb, lb, tb, fc = m.global_branch, m.local_branch, m.transmission_branch, m.fusion_conv
    g = t.conv_block(gb[0], [t.stem(x, 7, 3, 32)], stem_kp=32)
    g = t.attention(gb[3], t.res_block(gb[2], t.maxpool(g, 2, 2, 0)))
    g = t.attention(gb[6], t.res_block(gb[5], t.maxpool(g, 2, 2, 0)))
    g = t.upsample(t.res_block(gb[7], g), 2)
    g = t.upsample(t.res_block(gb[9], g), 2)
    G = t.conv_block(gb[11], [g])
    lf = t.conv_block(lb[0], [t.stem(x, 3, 1, 16)], stem_kp=16)
    lf = t.res_block(lb[2], t.res_block(lb[1], lf))
    L = t.conv_block(lb[3], [lf])
    tr = t.conv_block(tb[1], [t.conv_block(tb[0], [G, L])])
    one_minus_t = t.dot_head(tb[2], tr, complement=True)
    r = t.conv_block(fc[0], [G, L])
    return t.image_head(fc[1], r, x, IMG_GUIDED, ACT_TANH, guidance=one_minus_t)


def forward_unet(t, m, x, attn):
    """MediumIntensityDehazeModel.forward (medium_intensity.py:78-117) / HighIntensityDehazeModel.forward (high:92-138)."""
    guidance = None
    if attn:
        g = t.conv_block(m.detail_branch[0], [t.stem(x, 3, 1, 16)], stem_kp=16)
        g = t.conv_block(m.detail_branch[1], [g])
        guidance = t.dot_head(m.detail_branch[2], g)
    f0 = t.conv_block(m.init_conv, [t.stem(x, 7, 3, 32)], stem_kp=32)
    feats = [f0]
    for e in m.encoder:
        f = t.conv_block(e[0], [feats[-1]])
        f = t.res_block(e[1], f)
        f = t.res_block(e[2], f)
        if attn:
            f = t.attention(e[3], f)
        feats.append(f)
    b = feats[-1]
    for mod in m.bottleneck:
        b = t.attention(mod, b) if type(mod).__name__ == "AttentionBlock" else t.res_block(mod, b)
    d0, d1 = m.decoder[0], m.decoder[1]
    x1 = t.convT_bn_act(d0[0], d0[1], ACT_RELU, [b])
    x1 = t.res_block(d0[3], x1)
    if attn:
        x1 = t.attention(d0[4], x1)
    x2 = t.convT_bn_act(d1[0], d1[1], ACT_RELU, [x1, feats[1]])
    x2 = t.res_block(d1[3], x2)
    if attn:
        x2 = t.attention(d1[4], x2)
    r = t.conv_block(m.output_conv[0], [x2, f0])
    r = t.conv_block(m.output_conv[1], [r])
    return t.image_head(m.output_conv[2], r, x, IMG_GUIDED if attn else IMG_RESIDUAL, ACT_TANH, guidance=guidance)


def forward_resnet(t, clf, x):
    """FogIntensityClassifier.forward (classifier.py:80-97) with a torchvision resnet18/34 backbone in train() mode."""
    bb = clf.backbone
    f = t.stem_full_bn_act(bb.conv1, bb.bn1, ACT_RELU, x)
    f = t.maxpool(f, 3, 2, 1)
    for layer in (bb.layer1, bb.layer2, bb.layer3, bb.layer4):
        for blk in layer:
            u = t.conv_bn_act(blk.conv1, blk.bn1, ACT_RELU, [f])
            idn = f if blk.downsample is None else t.conv_bn_act(blk.downsample[0], blk.downsample[1], ACT_NONE, [f])
            f = t.conv_bn_act(blk.conv2, blk.bn2, ACT_RELU, [u], residual=idn)
    feats = t.global_avgpool(f)
    logits = t.head_mlp(clf.classifier, feats)
    return logits, feats.t


def forward_densenet(t, clf, x):
    """FogIntensityClassifier.forward with the torchvision densenet121 backbone (the north_star's HDEN) in train() mode:
    stem conv0/norm0/relu0/pool0, four dense blocks (norm1-relu1-conv1-norm2-relu2-conv2 per layer, concat by writing into
    the block buffer), three transitions (norm-relu-conv-avgpool), norm5-relu, global average pool, the reference head."""
    ft = clf.backbone.features
    f = t.stem_full_bn_act(ft.conv0, ft.norm0, ACT_RELU, x)
    n, h, w, c_in = f.t.shape
    hh, ww = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    dev = f.t.device
    pending = ("max", f)
    for bi in range(4):
        layers = list(getattr(ft, f"denseblock{bi + 1}").children())
        c_total = c_in + 32 * len(layers)
        B = BlockBuffer(torch.empty((n, hh, ww, c_total), dtype=torch.bfloat16, device=dev))
        if pending[0] == "max":
            t.maxpool3x3s2_into(pending[1], B, c_in)
        else:
            t.avgpool_into(pending[1], B, c_in)
        c = c_in
        for layer in layers:
            y1 = t.bn_act_prefix(B, c, layer.norm1, ACT_RELU)
            u = t.conv_bn_act(layer.conv1, layer.norm2, ACT_RELU, [y1])
            t.conv_plain(layer.conv2, u, into=B, c_off=c)
            c += 32
        if bi < 3:
            tr = getattr(ft, f"transition{bi + 1}")
            y = t.bn_act_prefix(B, c, tr.norm, ACT_RELU)
            z = t.conv_plain(tr.conv, y)
            pending = ("avg", z)
            c_in = tr.conv.weight.shape[0]
            hh, ww = hh // 2, ww // 2
        else:
            y = t.bn_act_prefix(B, c, ft.norm5, ACT_RELU)
    feats = t.global_avgpool(y)
    logits = t.head_mlp(clf.classifier, feats)
    return logits, feats.t


class ClassifierTrainFn(torch.autograd.Function):
    """autograd node of the HDEN forward in train() mode; returns (logits, features)."""

    @staticmethod
    def forward(ctx, engine, x, *params):
        tape = Tape(engine.train_cache)
        fwd = forward_densenet if engine.clf.model_name == "densenet121" else forward_resnet
        logits, feats = fwd(tape, engine.clf, x)
        ctx.tape, ctx.params = tape, params
        return logits, feats

    @staticmethod
    def backward(ctx, dlogits, dfeats):
        tape = ctx.tape
        ctx.tape = None
        if dlogits is None:
            dlogits = torch.zeros((dfeats.shape[0], tape_classes(ctx)), device=dfeats.device)
        pg = tape.backward_classifier(dlogits.contiguous().float(), None if dfeats is None else dfeats.contiguous().float())
        return (None, None) + tuple(pg.get(p) for p in ctx.params)


def tape_classes(ctx):
    return ctx.params[-1].shape[0]


def train_forward_classifier(engine, x):
    params = tuple(engine.clf.parameters())
    return ClassifierTrainFn.apply(engine, x, *params)


class BranchTrainFn(torch.autograd.Function):
    """autograd node of one branch forward in train() mode: the parameters are inputs so `.grad` accumulates as usual."""

    @staticmethod
    def forward(ctx, engine, x, *params):
        tape = Tape(engine.train_cache)
        if engine.kind == "light":
            out = forward_light(tape, engine.model, x)
        elif engine.kind == "low_unet":
            out = forward_low_unet(tape, engine.model, x)
        elif engine.kind == "corun":
            out = forward_corun(tape, engine.model, x)
        elif engine.kind == "dual":
            out = forward_dual(tape, engine.model, x)
        elif engine.kind in ("unet", "unet_attn"):
            out = forward_unet(tape, engine.model, x, engine.kind == "unet_attn")
        else:
            raise NotImplementedError(
                f"{type(engine.model).__name__}: training on the B200 path covers the default branch models "
                "(LightweightDehazeModel, MediumIntensityDehazeModel, HighIntensityDehazeModel); there is no torch fallback")
        ctx.tape, ctx.params = tape, params
        return out

    @staticmethod
    def backward(ctx, dout):
        pg = ctx.tape.backward(dout.contiguous().float())
        ctx.tape = None
        return (None, None) + tuple(pg.get(p) for p in ctx.params)


def train_forward(engine, x):
    params = tuple(engine.model.parameters())
    return BranchTrainFn.apply(engine, x, *params)
