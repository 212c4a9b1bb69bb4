"""Host-side operator layer over the C-ABI: weight packing, BN folding and one Python call per fused kernel.

Feature maps are NHWC bf16 torch tensors (`[n, h, w, c_pitch]`), images NCHW fp32.  PyTorch is used for device memory
and the stream only; every arithmetic op on the hot path is a libadb200 kernel.  Packing (a one-off, per weight
version) uses torch tensor ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, CONV_K4_S2D, CONV_S1, CONV_S2, CONVT_4X4S2, EPI_DOT, EPI_FEATURE,
                   EPI_IMAGE, IMG_BLEND, IMG_GUIDED, IMG_RESIDUAL, WG_OIHW, WG_STEM, ConvDesc, WgradDesc)

__all__ = [
    "pad16", "fold_bn", "pack_conv_weight", "pack_stem_s2d_weight", "pack_conv_weight_fold", "fold_eligible", "pack_convT_weight", "pack_stem_weight", "ConvSpec", "conv2d",
    "stem_pack", "nchw_to_nhwc", "nhwc_to_nchw", "attention", "maxpool3x3s2", "global_avgpool", "affine_relu", "avgpool2x2", "maxpool_kxk", "upsample_bilinear", "head_mlp",
    "linear", "route", "blend3", "l1_mse", "cross_entropy", "wgrad",
]


def pad16(c):
    return max(16, (c + 15) // 16 * 16)


_PACK_DTYPE = [torch.bfloat16]


class pack_as:
    """`with pack_as(torch.float32):` — the pack_* functions keep fp32 instead of rounding to bf16.  Used once per packing
    to derive its index map (pack an arange): every pack function is a pure permutation + zero padding."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        _PACK_DTYPE.append(self.dtype)

    def __exit__(self, *a):
        _PACK_DTYPE.pop()


def pack_dtype():
    return _PACK_DTYPE[-1]


# --------------------------------------------------------------------------- packing (host logic, CPU-testable)
def fold_bn(cout, bias=None, bn=None, cout_pad=None, device=None):
    """Epilogue affine of Conv(+bias) -> BatchNorm(eval): y = acc*scale + shift  (base_model.py:11-16).

    bn = (weight, bias, running_mean, running_var, eps) or None.  Returns fp32 (scale, shift) of length cout_pad
    (zero in the padding so padded channels come out as act(0)).
    """
    cout_pad = cout_pad or pad16(cout)
    dev = device if device is not None else ((bias if bias is not None else bn[0]).device
                                             if (bias is not None or bn is not None) else "cpu")
    scale = torch.ones(cout, dtype=torch.float32, device=dev)
    shift = torch.zeros(cout, dtype=torch.float32, device=dev)
    if bias is not None:
        shift = shift + bias.detach().float()
    if bn is not None:
        g, b, mean, var, eps = bn
        s = g.detach().float() / torch.sqrt(var.detach().float() + eps)
        shift = (shift - mean.detach().float()) * s + b.detach().float()
        scale = scale * s
    out_s = torch.zeros(cout_pad, dtype=torch.float32, device=dev)
    out_b = torch.zeros(cout_pad, dtype=torch.float32, device=dev)
    out_s[:cout] = scale
    out_b[:cout] = shift
    return out_s.contiguous(), out_b.contiguous()


def pack_conv_weight(w, cout_pad=None):
    """nn.Conv2d weight [co, ci, kh, kw] -> bf16 [cout_pad, kh*kw*ci], K index (r*kw + s)*ci + c."""
    co, ci, kh, kw = w.shape
    cout_pad = cout_pad or pad16(co)
    p = w.detach().float().permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
    out = torch.zeros(cout_pad, kh * kw * ci, dtype=pack_dtype(), device=w.device)
    out[:co] = p.to(pack_dtype())
    return out.contiguous()


def fold_eligible(co, kh, kw, stride, pad, cout_pad=None):
    """3x3 stride-1 'same' convs whose three filter rows fit one MMA's N (3*cout_pad <= 256): the rolling-row kernel."""
    cp = cout_pad or pad16(co)
    # (cout_pad 16: the 3-channel image heads and the 16-channel guidance / transmission heads, IMAGE / DOT epilogues only)
    return kh == 3 and kw == 3 and stride == 1 and pad == 1 and cp in (16, 32, 64)


def pack_conv_weight_fold(w, cout_pad=None):
    """nn.Conv2d 3x3 weight [co, ci, 3, 3] -> bf16 [3 (s), 3*cout_pad, ci]: row (2 - r)*cout_pad + co of column tap s
    (adb_conv_desc.w_fold, include/adb200.h) — filter row r lands in the N block of the output row it contributes to."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    cout_pad = cout_pad or pad16(co)
    out = torch.zeros(3, 3, cout_pad, ci, dtype=pack_dtype(), device=w.device)      # [s][blk][co][ci]
    p = w.detach().float().permute(3, 2, 0, 1)                                       # [s][r][co][ci]
    out[:, :, :co] = p.flip(1).to(pack_dtype())
    return out.reshape(3, 3 * cout_pad, ci).contiguous()


def pack_convT_weight(wt, cout_pad=None):
    """nn.ConvTranspose2d(4, 2, 1) weight [ci, co, 4, 4] -> bf16 [4 phases, cout_pad, 4*ci] (include/adb200.h)."""
    ci, co, kh, kw = wt.shape
    assert kh == 4 and kw == 4
    cout_pad = cout_pad or pad16(co)
    out = torch.zeros(4, cout_pad, 4 * ci, dtype=pack_dtype(), device=wt.device)
    w = wt.detach().float()
    for a in range(2):
        for b in range(2):
            for i in range(2):
                for j in range(2):
                    r = 2 * i if a else 1 + 2 * i
                    s = 2 * j if b else 1 + 2 * j
                    t = i * 2 + j
                    out[a * 2 + b, :co, t * ci:(t + 1) * ci] = w[:, :, r, s].t().to(pack_dtype())
    return out.contiguous()


def pack_stem_weight(w, kp, cout_pad=None):
    """Stem conv weight [co, 3, kh, kw] for the kh x 1 conv over the stem_pack operand: K index r*kp + s*3 + c."""
    co, ci, kh, kw = w.shape
    assert ci == 3 and kp >= 3 * kw
    cout_pad = cout_pad or pad16(co)
    out = torch.zeros(cout_pad, kh, kp, dtype=torch.float32, device=w.device)
    out[:co, :, :3 * kw] = w.detach().float().permute(0, 2, 3, 1).reshape(co, kh, kw * 3)
    return out.reshape(cout_pad, kh * kp).to(pack_dtype()).contiguous()


def pack_stem_s2d_weight(w, cout_pad=None):
    """7x7 stride-2 pad-3 stem weight [co, 3, 7, 7] -> the 4x4-tap form over the space-to-depth image (ADB_CONV_K4_S2D):
    [cout_pad, 16 taps * 16], column (R*4 + S)*16 + (py*2 + px)*3 + c = W[co][c][2R+py-1][2S+px-1], zero outside the 7x7."""
    co, ci, kh, kw = w.shape
    assert ci == 3 and kh == 7 and kw == 7
    cout_pad = cout_pad or pad16(co)
    out = torch.zeros(cout_pad, 4, 4, 16, dtype=torch.float32, device=w.device)
    wf = w.detach().float()
    for R in range(4):
        for py in range(2):
            u = 2 * R + py - 1
            if not 0 <= u < 7:
                continue
            for S in range(4):
                for px in range(2):
                    v = 2 * S + px - 1
                    if 0 <= v < 7:
                        q = (py * 2 + px) * 3
                        out[:co, R, S, q:q + 3] = wf[:, :, u, v]
    return out.reshape(cout_pad, 256).to(pack_dtype()).contiguous()


class ConvSpec:
    """Packed parameters + geometry of one fused conv launch."""

    def __init__(self, kind, kh, kw, pad, cout, w_packed, scale, shift, act, w_fold=None):
        self.kind, self.kh, self.kw, self.pad = kind, kh, kw, pad
        self.cout = cout
        self.cout_pad = scale.numel()
        self.w_packed, self.scale, self.shift, self.act = w_packed, scale, shift, act
        self.w_fold = w_fold      # row-folded packing for the rolling-row kernel (eligible 3x3 convs only)

    @staticmethod
    def from_conv(weight, bias=None, bn=None, act=ACT_NONE, stride=1, pad=None):
        co, ci, kh, kw = weight.shape
        pad = kh // 2 if pad is None else pad
        scale, shift = fold_bn(co, bias, bn, device=weight.device)
        fold = pack_conv_weight_fold(weight) if fold_eligible(co, kh, kw, stride, pad) else None
        return ConvSpec(CONV_S1 if stride == 1 else CONV_S2, kh, kw, pad, co, pack_conv_weight(weight), scale, shift, act, fold)

    @staticmethod
    def from_convT(weight, bias=None, bn=None, act=ACT_NONE):
        ci, co, kh, kw = weight.shape
        scale, shift = fold_bn(co, bias, bn, device=weight.device)
        return ConvSpec(CONVT_4X4S2, 4, 4, 1, co, pack_convT_weight(weight), scale, shift, act)

    @staticmethod
    def from_stem_s2d(weight, bn=None, act=ACT_NONE):
        """torchvision's 7x7 stride-2 stem over the space-to-depth operand of stem_pack(x, 2, 0, 16, stride=2, kh=2)."""
        co = weight.shape[0]
        scale, shift = fold_bn(co, None, bn, device=weight.device)
        return ConvSpec(CONV_K4_S2D, 4, 4, 2, co, pack_stem_s2d_weight(weight), scale, shift, act)

    @staticmethod
    def from_stem(weight, kp, bias=None, bn=None, act=ACT_NONE):
        co, ci, kh, kw = weight.shape
        scale, shift = fold_bn(co, bias, bn, device=weight.device)
        return ConvSpec(CONV_S1, kh, 1, kh // 2, co, pack_stem_weight(weight, kp), scale, shift, act)


# --------------------------------------------------------------------------- kernels
def _out_hw(spec, h, w):
    if spec.kind == CONV_S2:
        return h // 2, w // 2
    if spec.kind == CONVT_4X4S2:
        return h * 2, w * 2
    return h, w


def _pitch(t):
    """Channel pitch of an NHWC map that is dense in n/h/w: a contiguous tensor or a last-dim slice of one."""
    n, h, w, c = t.shape
    p = t.stride(2) if w > 1 else (t.stride(1) // max(1, w) if h > 1 else (t.stride(0) // max(1, h * w) if n > 1 else c))
    assert t.stride(3) == 1 and p >= c and p % 8 == 0 and (t.storage_offset() * 2) % 16 == 0, "not an NHWC map / channel slice"
    assert (w == 1 or t.stride(2) == p) and (h == 1 or t.stride(1) == w * p) and (n == 1 or t.stride(0) == h * w * p)
    return p


def conv2d(spec, src0, src1=None, *, c0=None, c1=None, dst=None, dst_c_off=0, residual=None, n=None, n_dev=None,
           n_start=0, epi=EPI_FEATURE, dot=None, image=None, tune=None, pre=None, stats=False, stat_alloc=None):
    """Launch one fused conv.  src*/dst/residual: NHWC bf16.  Returns dst (FEATURE), dot_out (DOT) or None (IMAGE).
    stats=True / "pool" (FEATURE): returns (dst, partials) — partials = fp32 [n, slots_per_image, 2, cout_pad] per-32-pixel (sum,
    sum of squares) / (sum, max) of the stored output written by the epilogue (adb_conv_desc.stat_out), or None when this
    launch does not produce them.  stat_alloc(shape) -> fp32 tensor supplies the partials buffer (engine buffer cache).

    dot   = (dot_w fp32[16], dot_b float, dot_out fp32[n,h,w])
    image = dict(mode=IMG_*, x=NCHW fp32, out=NCHW fp32, index=int32|None, guidance=fp32|None, alpha=fp32 scalar|None)
    """
    assert src0.dtype == torch.bfloat16 and src0.is_cuda
    nb, h, w, _ = src0.shape
    p0 = _pitch(src0)        # contiguous map, or a channel slice [..., a:b] of a wider one (DenseNet block buffer)
    n = nb if n is None else n
    d = ConvDesc()
    d.src0, d.c0, d.c0_pitch = src0.data_ptr(), (c0 or src0.shape[3]), p0
    if src1 is not None:
        assert src1.shape[:3] == src0.shape[:3] and src1.is_contiguous()
        d.src1, d.c1, d.c1_pitch = src1.data_ptr(), (c1 or src1.shape[3]), src1.shape[3]
    d.n, d.h_in, d.w_in = n, h, w
    d.kind, d.kh, d.kw, d.pad = spec.kind, spec.kh, spec.kw, spec.pad
    d.w_packed, d.scale, d.shift = spec.w_packed.data_ptr(), spec.scale.data_ptr(), spec.shift.data_ptr()
    if spec.w_fold is not None:
        d.w_fold = spec.w_fold.data_ptr()
    d.cout, d.cout_pad, d.act, d.epi = spec.cout, spec.cout_pad, spec.act, epi
    ho, wo = _out_hw(spec, h, w)
    ret = None
    if epi == EPI_FEATURE:
        if dst is None:
            dst = torch.empty((nb, ho, wo, spec.cout_pad), dtype=torch.bfloat16, device=src0.device)
        assert dst.dtype == torch.bfloat16 and dst.is_contiguous() and tuple(dst.shape[:3]) == (nb, ho, wo)
        d.dst, d.dst_pitch, d.dst_c_off = dst.data_ptr(), dst.shape[3], dst_c_off
        if residual is not None:
            assert tuple(residual.shape[:3]) == (nb, ho, wo)
            d.residual, d.res_pitch = residual.data_ptr(), _pitch(residual)
        ret = dst
    elif epi == EPI_DOT:
        dot_w, dot_b, dot_out = dot
        d.dot_w, d.dot_b, d.dot_out = dot_w.data_ptr(), float(dot_b), dot_out.data_ptr()
        ret = dot_out
    else:
        d.img_mode = image["mode"]
        d.img_x, d.img_out = image["x"].data_ptr(), image["out"].data_ptr()
        if image.get("index") is not None:
            d.img_index = image["index"].data_ptr()
        if image.get("guidance") is not None:
            d.img_guidance = image["guidance"].data_ptr()
        if image.get("alpha") is not None:
            d.img_alpha = image["alpha"].data_ptr()
    if n_dev is not None:
        d.n_dev = n_dev.data_ptr()
    d.n_start = n_start
    if pre is not None:        # (scale, shift) fp32 over the input channels: relu(x*scale + shift) fused into the operand
        d.pre_scale, d.pre_shift = pre[0].data_ptr(), pre[1].data_ptr()
    if tune:
        d.tune_mt, d.tune_stages, d.tune_acc_stages = tune.get("mt", 0), tune.get("stages", 0), tune.get("acc", 0)
        d.tune_flags = tune.get("flags", 0)
    if stats:
        assert epi == EPI_FEATURE
        slots = int(_lib.load().adb_conv2d_stat_slots(C.byref(d)))
        part = None
        if slots > 0:
            shape = (n, slots // n, 2, spec.cout_pad)
            part = stat_alloc(shape) if stat_alloc else torch.empty(shape, dtype=torch.float32, device=src0.device)
            d.stat_out, d.stat_mode = part.data_ptr(), (2 if stats == "pool" else 1)
        _lib.call("adb_conv2d", C.byref(d), _lib.current_stream())
        return ret, part
    _lib.call("adb_conv2d", C.byref(d), _lib.current_stream())
    return ret


_WG_WS = {}


def _wgrad_workspace(nbytes, device):
    """One grow-only fp32 scratch per device for the wgrad partial slabs (stream-ordered re-use)."""
    key = str(device)
    t = _WG_WS.get(key)
    if t is None or t.numel() * 4 < nbytes:
        t = _WG_WS[key] = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
    return t


def wgrad(small, large0, large1=None, *, kind=CONV_S1, kh=3, kw=3, pad=1, cs=None, cs_true=None, c0=None, c1=None, n=None,
          out=None, layout=WG_OIHW, stem_kw=0, accumulate=False, mode=0):
    """Weight gradient of one conv: out[m][c][r][s] (+)= sum small[.., m] * large[.. + tap, c]  (see adb_wgrad in adb200.h).

    small: NHWC bf16 [n, hs, ws, pitch]; large0/large1: NHWC bf16 [n, h, w, pitch] (concat sources).
    Returns the fp32 gradient in the parameter layout ([cs_true][c0+c1][kh][kw], or [cs_true][3][kh][stem_kw] for STEM)."""
    assert small.dtype == torch.bfloat16 and large0.dtype == torch.bfloat16 and large0.is_contiguous()
    nb, h, w, p0 = large0.shape
    d = WgradDesc()
    d.grad, d.cg, d.cg_pitch = small.data_ptr(), (cs or small.shape[3]), _pitch(small)
    d.cg_true = cs_true or 0
    d.act0, d.c0, d.c0_pitch = large0.data_ptr(), (c0 or p0), p0
    if large1 is not None:
        assert large1.is_contiguous() and large1.shape[:3] == large0.shape[:3]
        d.act1, d.c1, d.c1_pitch = large1.data_ptr(), (c1 or large1.shape[3]), large1.shape[3]
    d.n, d.h_in, d.w_in = (nb if n is None else n), h, w
    d.kind, d.kh, d.kw, d.pad = kind, kh, kw, pad
    d.layout, d.stem_kw, d.accumulate, d.mode = layout, stem_kw, int(bool(accumulate)), mode
    rows = d.cg_true or d.cg
    if out is None:
        shape = (rows, 3, kh, stem_kw) if layout == WG_STEM else (rows, d.c0 + d.c1, kh, kw)
        out = torch.empty(shape, dtype=torch.float32, device=small.device)
        assert not accumulate
    assert out.dtype == torch.float32 and out.is_contiguous()
    need = int(_lib.load().adb_wgrad_workspace_bytes(C.byref(d)))
    if need < 0:
        raise _lib.AdbError(f"adb_wgrad_workspace_bytes: {_lib.last_error()}")
    ws = _wgrad_workspace(min(need, 256 << 20), small.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    d.dw = out.data_ptr()
    _lib.call("adb_wgrad", C.byref(d), _lib.current_stream())
    return out


def stem_pack(x, kw, pad, kp, *, stride=1, kh=1, index=None, n_dev=None, n_start=0, n=None, out=None):
    """NCHW fp32 image batch -> [n, ho, wo, kp] bf16 stem operand (taps unrolled into channels, see adb200.h)."""
    assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous() and x.shape[1] == 3
    b, _, h, w = x.shape
    n = b if n is None else n
    wo = (w + 2 * pad - kw) // stride + 1
    ho = (h + 2 * pad - kh) // stride + 1 if kh > 1 else h
    if out is None:
        out = torch.empty((n, ho, wo, kp), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_stem_pack", _lib.ptr(x), _lib.ptr(index), _lib.ptr(n_dev), n_start, n, h, w, kh, kw, pad, stride, kp,
              _lib.ptr(out), _lib.current_stream())
    return out


def nchw_to_nhwc(x, c_pitch=None):
    n, c, h, w = x.shape
    c_pitch = c_pitch or (c + 7) // 8 * 8
    out = torch.empty((n, h, w, c_pitch), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_nchw_to_nhwc_bf16", _lib.ptr(x.contiguous().float()), n, c, h, w, c_pitch, _lib.ptr(out), _lib.current_stream())
    return out


def nhwc_to_nchw(x, c=None):
    n, h, w, p = x.shape
    c = c or p
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _lib.call("adb_nhwc_bf16_to_nchw", _lib.ptr(x), n, c, h, w, p, _lib.ptr(out), _lib.current_stream())
    return out


class AttnParams:
    """AttentionBlock weights (base_model.py:53-62): fc.0 [c/r, c, 1, 1], fc.2 [c, c/r, 1, 1], conv_spatial [1, 2, 7, 7]."""

    def __init__(self, fc0, fc2, conv_spatial):
        self.c_red, self.c = fc0.shape[0], fc0.shape[1]
        self.w1 = fc0.detach().float().reshape(self.c_red, self.c).contiguous()
        self.w2 = fc2.detach().float().reshape(self.c, self.c_red).contiguous()
        self.wsp = conv_spatial.detach().float().reshape(98).contiguous()


def pool_scratch_floats(n, h, w, c):
    return int(_lib.load().adb_pool_scratch_floats(n, h, w, c))


def attention(x, ap, *, n=None, n_dev=None, n_start=0, out=None, scratch=None, pool_partials=None):
    """AttentionBlock forward on an NHWC bf16 map: pool -> gate + channel stats -> spatial gate apply.
    pool_partials: the (sum, max) partials the conv that produced x wrote (conv2d(stats="pool")): the pool pass over x is
    replaced by a fold of them."""
    nb, h, w, c = x.shape
    assert c == ap.c and x.is_contiguous()
    n = nb if n is None else n
    dev = x.device
    if scratch is None:
        scratch = {}
    pool = scratch.get(("pool", nb, h, w, c))
    if pool is None:
        pool = scratch[("pool", nb, h, w, c)] = torch.empty(pool_scratch_floats(nb, h, w, c), dtype=torch.float32, device=dev)
    gate = scratch.setdefault(("gate", nb, c), torch.empty((nb, c), dtype=torch.float32, device=dev))
    stats = scratch.setdefault(("stats", nb, h, w), torch.empty((nb, h, w, 2), dtype=torch.float32, device=dev))
    spatial = scratch.setdefault(("spatial", nb, h, w), torch.empty((nb, h, w), dtype=torch.float32, device=dev))
    if out is None:
        out = torch.empty_like(x)
    st = _lib.current_stream()
    nd = _lib.ptr(n_dev)
    if pool_partials is not None:
        assert pool_partials.shape[0] >= n and pool_partials.shape[3] >= c and pool_partials.is_contiguous()
        spi = pool_partials.shape[1]
        key = ("pool_fold", nb, spi, c)
        fold = scratch.get(key)
        if fold is None:
            fold = scratch[key] = torch.empty(int(_lib.load().adb_attn_pool_stat_scratch_floats(nb, spi, c)), dtype=torch.float32, device=dev)
        _lib.call("adb_attn_pool_from_stats", _lib.ptr(pool_partials), n, spi, pool_partials.shape[3], c, nd, n_start, _lib.ptr(fold),
                  _lib.ptr(pool), st)
    else:
        _lib.call("adb_attn_pool", _lib.ptr(x), n, h, w, c, nd, n_start, _lib.ptr(pool), st)
    _lib.call("adb_attn_gate_stats", _lib.ptr(x), n, h, w, c, nd, n_start, _lib.ptr(pool), _lib.ptr(ap.w1),
              _lib.ptr(ap.w2), ap.c_red, _lib.ptr(gate), _lib.ptr(stats), st)
    _lib.call("adb_attn_apply", _lib.ptr(x), n, h, w, c, nd, n_start, _lib.ptr(gate), _lib.ptr(stats),
              _lib.ptr(ap.wsp), _lib.ptr(spatial), _lib.ptr(out), st)
    return out


def maxpool3x3s2(x, out=None):
    """3x3/2 max pool (pad 1); `out` may be a wider buffer (channel pitch >= c) whose first c channels are written."""
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_maxpool3x3s2", _lib.ptr(x), n, h, w, c, _lib.ptr(out), out.shape[3], _lib.current_stream())
    return out


def global_avgpool(x):
    n, h, w, c = x.shape
    scratch = torch.empty(pool_scratch_floats(n, h, w, c), dtype=torch.float32, device=x.device)
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    _lib.call("adb_global_avgpool", _lib.ptr(x), n, h, w, c, _lib.ptr(scratch), _lib.ptr(out), _lib.current_stream())
    return out


def affine_relu(x, c, scale, shift, out=None):
    """relu(x[..., :c]*scale + shift) -> NHWC bf16 with pitch c (DenseNet pre-activation)."""
    n, h, w, p = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_affine_relu", _lib.ptr(x), n * h * w, c, p, _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(out),
              out.shape[3], _lib.current_stream())
    return out


def avgpool2x2(x, c=None, out=None):
    n, h, w, p = x.shape
    c = c or p
    if out is None:
        out = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_avgpool2x2", _lib.ptr(x), n, h, w, c, p, _lib.ptr(out), out.shape[3], _lib.current_stream())
    return out


def maxpool_kxk(x, k, *, n=None, n_dev=None, n_start=0, out=None):
    """nn.MaxPool2d(k, k), k in {2, 4}, NHWC bf16."""
    nb, h, w, c = x.shape
    if out is None:
        out = torch.empty((nb, h // k, w // k, c), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_maxpool_kxk", _lib.ptr(x), nb if n is None else n, h, w, c, k, _lib.ptr(n_dev), n_start, _lib.ptr(out),
              _lib.current_stream())
    return out


def upsample_bilinear(x, scale, *, n=None, n_dev=None, n_start=0, out=None, c_off=0):
    """nn.UpsamplingBilinear2d(scale_factor=scale) (align_corners=True) into channels [c_off, c_off+c) of `out`."""
    nb, h, w, c = x.shape
    if out is None:
        out = torch.empty((nb, h * scale, w * scale, c), dtype=torch.bfloat16, device=x.device)
    _lib.call("adb_upsample_bilinear", _lib.ptr(x), nb if n is None else n, h, w, c, scale, _lib.ptr(n_dev), n_start,
              _lib.ptr(out), out.shape[3], c_off, _lib.current_stream())
    return out


def head_mlp(feat, w1, b1, w2, b2):
    n, f = feat.shape
    w1, b1, w2, b2 = (t.detach().float().contiguous() for t in (w1, b1, w2, b2))
    logits = torch.empty((n, w2.shape[0]), dtype=torch.float32, device=feat.device)
    _lib.call("adb_head_mlp", _lib.ptr(feat), n, f, _lib.ptr(w1), _lib.ptr(b1), w1.shape[0], _lib.ptr(w2), _lib.ptr(b2),
              w2.shape[0], _lib.ptr(logits), _lib.current_stream())
    return logits


def linear(x, w, b=None, relu=False):
    n, fin = x.shape
    w = w.detach().float().contiguous()
    b = None if b is None else b.detach().float().contiguous()
    y = torch.empty((n, w.shape[0]), dtype=torch.float32, device=x.device)
    _lib.call("adb_linear", _lib.ptr(x.contiguous()), n, fin, _lib.ptr(w), _lib.ptr(b), w.shape[0], int(relu), _lib.ptr(y),
              _lib.current_stream())
    return y


def route(logits=None, intensity=None, batch=None):
    """argmax + stable 3-way bucketing on the device.  Returns (intensity int64[b], masks bool[3,b],
    bucket_index int32[3,b], bucket_count int32[3]) — all device tensors, no host sync."""
    if logits is not None:
        logits = logits.contiguous().float()
        b, classes = logits.shape
        dev = logits.device
    else:
        b, classes, dev = intensity.shape[0], 3, intensity.device
    out_int = torch.empty(b, dtype=torch.int64, device=dev)
    masks = torch.empty((3, b), dtype=torch.uint8, device=dev)
    bidx = torch.empty((3, b), dtype=torch.int32, device=dev)
    bcnt = torch.empty(3, dtype=torch.int32, device=dev)
    inten = None if intensity is None else intensity.contiguous().to(torch.int64)
    _lib.call("adb_route", _lib.ptr(logits), _lib.ptr(inten), b, classes, _lib.ptr(out_int), _lib.ptr(masks),
              _lib.ptr(bidx), _lib.ptr(bcnt), _lib.current_stream())
    return out_int, masks.view(torch.bool), bidx, bcnt


def blend3(y0, y1, y2, logits_or_weights, temperature):
    b = y0.shape[0]
    chw = y0[0].numel()
    out = torch.empty_like(y0)
    wts = torch.empty((b, 3), dtype=torch.float32, device=y0.device)
    _lib.call("adb_blend3", _lib.ptr(y0), _lib.ptr(y1), _lib.ptr(y2), _lib.ptr(logits_or_weights.contiguous().float()),
              float(temperature), b, chw, _lib.ptr(wts), _lib.ptr(out), _lib.current_stream())
    return out, wts


def l1_mse(pred, target):
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    _lib.call("adb_l1_mse_fwd", _lib.ptr(pred), _lib.ptr(target), pred.numel(), _lib.ptr(out), _lib.current_stream())
    return out


def cross_entropy(logits, labels, grad_scale=1.0, want_grad=True):
    b, k = logits.shape
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    _lib.call("adb_ce_fwd_bwd", _lib.ptr(logits), _lib.ptr(labels), b, k, float(grad_scale), _lib.ptr(loss),
              _lib.ptr(grad), _lib.current_stream())
    return loss, grad
