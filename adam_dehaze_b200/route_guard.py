"""Route guard — bit-exact route decisions against the fp32 reference (models/routing.py:41-43) without a host round trip.

HDEN runs in bf16 (fp32 accumulate); its logits are within a few 1e-4 of the fp32 reference's, bounded at 2e-2 by the
parity tests.  `HardRouter` takes the argmax of those logits, so an image whose two largest logits are closer than that
error could be sent to another branch than the reference sends it to.  The guard closes the gap:

  1. `adb_guard_flags` lists, on the device, the batch rows whose top-2 gap is below `eps` (= twice the logit bound);
  2. the listed rows — and only those — are re-run through the classifier trunk in fp32 storage and fp32 FMA arithmetic
     (`adb_f32_conv2d` / `adb_f32_pool` / `adb_f32_global_avgpool` + the fp32 head), `cap` rows per pass;
  3. `adb_guard_scatter` writes the fp32 logits over the bf16 ones before `adb_route` takes the argmax.

The host never learns how many rows were listed: one CUDA graph per input shape holds a WHILE node whose body is one pass
(`adb_guard_graph_*`), so an empty list costs one flag kernel and one graph launch.  `use_graph=False` enqueues
ceil(B / cap) passes instead (every kernel of a pass exits at once when its rows are not live).
"""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import F32ConvDesc
from .engine import _Versioned, bn_args

DEFAULT_EPS = 4e-2      # 2 x the 2e-2 logit bound the parity tests hold the bf16 HDEN to
DEFAULT_CAP = 8         # rows per fp32 pass (fp32 maps of a 1024x2048 image take ~0.45 GB)


def _w32(w):
    """[co, ci, kh, kw] -> fp32 [kh*kw, ci, co] (adb_f32_conv_desc.w)."""
    co, ci, kh, kw = w.shape
    return w.detach().float().permute(2, 3, 1, 0).reshape(kh * kw, ci, co).contiguous()


def _affine(bn):
    c = bn.num_features
    return ops.fold_bn(c, None, bn_args(bn), cout_pad=c, device=bn.weight.device)


class _Conv:
    __slots__ = ("w", "cin", "cout", "kh", "kw", "stride", "pad", "pre", "post", "relu")

    def __init__(self, conv, pre=None, post=None, relu=False):
        self.w = _w32(conv.weight)
        self.cout, self.cin, self.kh, self.kw = conv.weight.shape
        self.stride, self.pad = conv.stride[0], conv.padding[0]
        self.pre, self.post, self.relu = pre, post, relu


class RouteGuard:
    def __init__(self, classifier, eps=DEFAULT_EPS, cap=DEFAULT_CAP, use_graph=True):
        if classifier.model_name not in ("resnet18", "resnet34", "densenet121"):
            raise NotImplementedError(f"route guard: backbone '{classifier.model_name}' has no fp32 path")
        self.clf, self.eps, self.cap, self.use_graph = classifier, float(eps), int(cap), bool(use_graph)
        self._ver = _Versioned(classifier)
        self._packs = None
        self._state = None          # device-side list / cursor / pointer slots
        self._programs = {}         # (h, w) -> list of launches
        self._graphs = {}           # (h, w) -> graph context
        self._side = None

    # ------------------------------------------------------------------ fp32 packings
    def _pack(self):
        stale = self._ver.stale()
        if self._packs is not None and not stale:
            return self._packs
        self._drop_graphs()
        self._programs.clear()
        clf, bb = self.clf, self.clf.backbone
        P = {}
        if clf.model_name.startswith("resnet"):
            P["stem"] = _Conv(bb.conv1, post=_affine(bb.bn1), relu=True)
            P["layers"] = []
            for layer in (bb.layer1, bb.layer2, bb.layer3, bb.layer4):
                for blk in layer:
                    ds = None
                    if blk.downsample is not None:
                        ds = _Conv(blk.downsample[0], post=_affine(blk.downsample[1]))
                    P["layers"].append((_Conv(blk.conv1, post=_affine(blk.bn1), relu=True),
                                        _Conv(blk.conv2, post=_affine(blk.bn2), relu=True), ds))
        else:
            ft = bb.features
            P["stem"] = _Conv(ft.conv0, post=_affine(ft.norm0), relu=True)
            P["blocks"], P["trans"] = [], []
            for bi in range(4):
                layers = []
                for layer in getattr(ft, f"denseblock{bi + 1}").children():
                    layers.append((_Conv(layer.conv1, pre=_affine(layer.norm1), post=_affine(layer.norm2), relu=True),
                                   _Conv(layer.conv2)))
                P["blocks"].append(layers)
                if bi < 3:
                    tr = getattr(ft, f"transition{bi + 1}")
                    P["trans"].append(_Conv(tr.conv, pre=_affine(tr.norm)))
            P["final"] = _affine(ft.norm5)
        head = clf.classifier
        P["head"] = tuple(t.detach().float().contiguous() for t in (head[1].weight, head[1].bias, head[4].weight, head[4].bias))
        self._packs = P
        return P

    # ------------------------------------------------------------------ device state
    def _get_state(self, b, dev):
        st = self._state
        if st is None or st["index"].numel() < b or st["index"].device != dev:
            self._drop_graphs()
            self._programs.clear()
            st = self._state = {
                "index": torch.zeros(max(b, 256), dtype=torch.int32, device=dev),
                "count": torch.zeros(1, dtype=torch.int32, device=dev),
                "cursor": torch.zeros(1, dtype=torch.int32, device=dev),
                "slots": torch.zeros(2, dtype=torch.int64, device=dev),
            }
        return st

    def _drop_graphs(self):
        for ctx in self._graphs.values():
            _lib.call("adb_guard_graph_destroy", ctx)
        self._graphs.clear()

    def __del__(self):
        try:
            self._drop_graphs()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    # ------------------------------------------------------------------ one pass as a list of prepared launches
    def _program(self, h, w, dev):
        key = (h, w)
        prog = self._programs.get(key)
        if prog is not None:
            return prog
        P, st, cap = self._pack(), self._state, self.cap
        live = (st["index"].data_ptr(), st["count"].data_ptr(), st["cursor"].data_ptr())
        prog, keep = [], []          # keep: tensors / descriptors the launches point at

        def buf(*shape):
            t = torch.empty(shape, dtype=torch.float32, device=dev)
            keep.append(t)
            return t

        def conv(cv, src, hh, ww, dst=None, dst_c_off=0, in_pitch=None, residual=None, nchw=False):
            ho = (hh + 2 * cv.pad - cv.kh) // cv.stride + 1
            wo = (ww + 2 * cv.pad - cv.kw) // cv.stride + 1
            if dst is None:
                dst = buf(cap, ho, wo, cv.cout)
            d = F32ConvDesc()
            d.flag_index, d.flag_count, d.cursor, d.cap = live[0], live[1], live[2], cap
            if nchw:
                d.x_slot, d.in_nchw = st["slots"].data_ptr(), 1
            else:
                d.x, d.in_pitch = src.data_ptr(), (in_pitch or src.shape[3])
            d.h_in, d.w_in, d.cin = hh, ww, cv.cin
            d.kh, d.kw, d.stride, d.pad = cv.kh, cv.kw, cv.stride, cv.pad
            d.w, d.cout = cv.w.data_ptr(), cv.cout
            if cv.pre is not None:
                d.pre_scale, d.pre_shift = cv.pre[0].data_ptr(), cv.pre[1].data_ptr()
            if cv.post is not None:
                d.post_scale, d.post_shift = cv.post[0].data_ptr(), cv.post[1].data_ptr()
            d.post_relu = int(cv.relu)
            if residual is not None:
                d.residual, d.res_pitch = residual.data_ptr(), residual.shape[3]
            d.y, d.out_pitch, d.out_c_off = dst.data_ptr(), dst.shape[3], dst_c_off
            keep.append(d)
            prog.append(("adb_f32_conv2d", (C.byref(d),)))
            return dst, ho, wo

        def pool(src, hh, ww, c, mode, dst=None):
            ho, wo = ((hh - 1) // 2 + 1, (ww - 1) // 2 + 1) if mode == 0 else (hh // 2, ww // 2)
            if dst is None:
                dst = buf(cap, ho, wo, c)
            prog.append(("adb_f32_pool", (live[0], live[1], live[2], cap, _lib.ptr(src), hh, ww, c, src.shape[3], mode,
                                          _lib.ptr(dst), dst.shape[3])))
            return dst, ho, wo

        feats = buf(cap, self.clf.feature_dim)
        logits_c = buf(cap, self.clf.num_classes)
        if "layers" in P:                                  # torchvision resnet18/34 (BasicBlock)
            f, hh, ww = conv(P["stem"], None, h, w, nchw=True)
            f, hh, ww = pool(f, hh, ww, P["stem"].cout, 0)
            for c1, c2, ds in P["layers"]:
                t, h2, w2 = conv(c1, f, hh, ww)
                idn = conv(ds, f, hh, ww)[0] if ds is not None else f
                f, hh, ww = conv(c2, t, h2, w2, residual=idn)
            prog.append(("adb_f32_global_avgpool", (live[0], live[1], live[2], cap, _lib.ptr(f), hh * ww, f.shape[3], f.shape[3],
                                                    None, None, _lib.ptr(feats))))
        else:                                              # torchvision densenet121
            f, hh, ww = conv(P["stem"], None, h, w, nchw=True)
            c_in = P["stem"].cout
            pending = ("max", f, hh, ww)
            for bi, layers in enumerate(P["blocks"]):
                kind, src, sh, sw = pending
                ho, wo = ((sh - 1) // 2 + 1, (sw - 1) // 2 + 1) if kind == "max" else (sh // 2, sw // 2)
                block = buf(cap, ho, wo, c_in + 32 * len(layers))
                pool(src, sh, sw, c_in, 0 if kind == "max" else 1, dst=block)
                hh, ww, c = ho, wo, c_in
                for c1, c2 in layers:
                    assert c1.cin == c
                    t, _, _ = conv(c1, block, hh, ww, in_pitch=block.shape[3])
                    conv(c2, t, hh, ww, dst=block, dst_c_off=c)
                    c += c2.cout
                if bi < 3:
                    tr = P["trans"][bi]
                    t, _, _ = conv(tr, block, hh, ww, in_pitch=block.shape[3])
                    pending, c_in = ("avg", t, hh, ww), tr.cout
                else:
                    prog.append(("adb_f32_global_avgpool", (live[0], live[1], live[2], cap, _lib.ptr(block), hh * ww, c,
                                                            block.shape[3], _lib.ptr(P["final"][0]), _lib.ptr(P["final"][1]),
                                                            _lib.ptr(feats))))
        w1, b1, w2, b2 = P["head"]
        prog.append(("adb_head_mlp", (_lib.ptr(feats), cap, feats.shape[1], _lib.ptr(w1), _lib.ptr(b1), w1.shape[0], _lib.ptr(w2),
                                      _lib.ptr(b2), w2.shape[0], _lib.ptr(logits_c))))
        prog.append(("adb_guard_scatter", (live[0], live[1], live[2], cap, _lib.ptr(logits_c), self.clf.num_classes,
                                           C.c_void_p(st["slots"].data_ptr() + 8))))
        self._programs[key] = (prog, keep)
        return self._programs[key]

    def _run_pass(self, prog, stream):
        for name, args in prog:
            _lib.call(name, *args, stream)

    # ------------------------------------------------------------------ public
    def refine(self, x, logits):
        """x: the NCHW fp32 CUDA batch the logits were computed from; logits: fp32 [B, classes] CUDA, patched IN PLACE for
        the rows whose top-2 gap is below eps.  Returns logits.  No host synchronisation."""
        if not (x.is_cuda and logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()):
            raise RuntimeError("route guard: expected CUDA fp32 tensors — this package runs on B200 (sm_100a) only and has no CPU path")
        x = x.contiguous()
        b, _, h, w = x.shape
        if self.clf.training:
            raise RuntimeError("route guard: the classifier must be in eval() mode (running BatchNorm statistics)")
        self._pack()
        st = self._get_state(b, x.device)
        stream = _lib.current_stream()
        _lib.call("adb_guard_flags", _lib.ptr(logits), b, logits.shape[1], self.eps, _lib.ptr(st["index"]), _lib.ptr(st["count"]),
                  _lib.ptr(st["cursor"]), stream)
        _lib.call("adb_guard_set_slots", _lib.ptr(st["slots"]), _lib.ptr(x), _lib.ptr(logits), stream)
        prog, _ = self._program(h, w, x.device)
        if not self.use_graph:
            for _ in range((b + self.cap - 1) // self.cap):
                self._run_pass(prog, stream)
                _lib.call("adb_guard_advance", _lib.ptr(st["count"]), _lib.ptr(st["cursor"]), self.cap, stream)
            return logits
        ctx = self._graphs.get((h, w))
        if ctx is None:
            if self._side is None:
                self._side = torch.cuda.Stream(device=x.device)
            side = C.c_void_p(self._side.cuda_stream)
            ctx = C.c_void_p()
            _lib.call("adb_guard_graph_begin", _lib.ptr(st["count"]), _lib.ptr(st["cursor"]), self.cap, side, C.byref(ctx))
            try:
                self._run_pass(prog, side)
            finally:
                _lib.call("adb_guard_graph_end", ctx)
            self._graphs[(h, w)] = ctx
        _lib.call("adb_guard_graph_launch", ctx, stream)
        return logits

    def flagged(self):
        """Number of rows the last refine() re-evaluated (host read: for tests and reports, not for the hot path)."""
        return int(self._state["count"].item()) if self._state is not None else 0
