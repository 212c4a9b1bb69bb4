"""Time (CUDA events) representative adb_conv2d launches of the three branches; optional tuning sweep.

    python tools/prof_conv.py [--sweep] [--reps 5] [--only NAME]
Prints one line per (shape, tune): ms, TFLOP/s (true FLOPs), fraction of the measured sustained bf16 peak.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adam_dehaze_b200 import ops  # noqa: E402

SHAPES = [
    # name, kind, cin(s), cout, k, n, h, w   (h, w = INPUT size)
    ("light_32_3x3", "s1", (32,), 32, 3, 2, 1024, 2048),
    ("med_64_3x3", "s1", (64,), 64, 3, 2, 1024, 2048),
    ("med_128_3x3", "s1", (128,), 128, 3, 4, 512, 1024),
    ("med_256_3x3", "s1", (256,), 256, 3, 8, 256, 512),
    ("med_cat_128_64", "s1", (64, 64), 64, 3, 2, 1024, 2048),
    ("med_down_64_128", "s2", (64,), 128, 4, 2, 1024, 2048),
    ("med_down_128_256", "s2", (128,), 256, 4, 4, 512, 1024),
    ("med_up_256_128", "t", (256,), 128, 4, 8, 256, 512),
    ("med_up_cat_256_64", "t", (128, 128), 64, 4, 4, 512, 1024),
    ("med_256_3x3_resdst", "resdst", (256,), 256, 3, 8, 256, 512),
    ("cpx_96_3x3", "s1", (96,), 96, 3, 2, 1024, 2048),
    ("cpx_192_3x3", "s1", (192,), 192, 3, 4, 512, 1024),
    ("cpx_384_3x3", "s1", (384,), 384, 3, 8, 256, 512),
    ("cpx_cat_192_96", "s1", (96, 96), 96, 3, 2, 1024, 2048),
    ("cpx_96_48", "s1", (96,), 48, 3, 2, 1024, 2048),
    ("cpx_down_96_192", "s2", (96,), 192, 4, 2, 1024, 2048),
    ("cpx_down_192_384", "s2", (192,), 384, 4, 4, 512, 1024),
    ("cpx_up_384_192", "t", (384,), 192, 4, 8, 256, 512),
    ("cpx_up_cat_384_96", "t", (192, 192), 96, 4, 4, 512, 1024),
    ("med_out_32_3_img", "img", (32,), 3, 3, 2, 1024, 2048),
    ("cpx_out_48_3_img", "img", (48,), 3, 3, 2, 1024, 2048),
    ("cpx_det_16_16_dot", "dot", (16,), 16, 3, 2, 1024, 2048),
    ("cpx_stem_7x1_32_96", "stem7", (32,), 96, 7, 2, 1024, 2048),
    ("light_stem_3x1_16_32", "stem3", (16,), 32, 3, 2, 1024, 2048),
    ("med_64_3x3_res", "res", (64,), 64, 3, 2, 1024, 2048),
    ("cpx_192_3x3_res", "res", (192,), 192, 3, 4, 512, 1024),
    ("med_64_3x3_resin", "resin", (64,), 64, 3, 2, 1024, 2048),
    ("cpx_192_3x3_resin", "resin", (192,), 192, 3, 4, 512, 1024),
    ("med_64_3x3_resdst", "resdst", (64,), 64, 3, 2, 1024, 2048),
    ("cpx_192_3x3_resdst", "resdst", (192,), 192, 3, 4, 512, 1024),
    ("dense_1x1_512_128", "s1", (512,), 128, 1, 8, 128, 256),
    ("dense_3x3_128_32", "s1", (128,), 32, 3, 8, 128, 256),
    # DenseNet conv1 with the consumer pre-activation fused into the operand: (c_in, buffer pitch)
    ("dense_pre_224_128", "pre", (224, 256), 128, 1, 8, 256, 512),
    ("dense_pre_480_128", "pre", (480, 512), 128, 1, 8, 128, 256),
    ("dense_pre_992_128", "pre", (992, 1024), 128, 1, 8, 64, 128),
]


STATS = None   # None | True (sum, sum of squares) | "pool" (sum, max): epilogue channel partials on the launches that offer them
_STAT_BUF = {}


def _stat_kwargs():
    if STATS is None:
        return {}

    def alloc(shape):
        if shape not in _STAT_BUF:
            _STAT_BUF[shape] = torch.empty(shape, dtype=torch.float32, device="cuda")
        return _STAT_BUF[shape]
    return dict(stats=STATS, stat_alloc=alloc)


def run(shape, tune, reps):
    name, kind, cins, cout, k, n, h, w = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    if kind == "pre":
        cin, pitch = cins
        buf = torch.randn((n, h, w, pitch), generator=g, device=dev).to(torch.bfloat16)
        wt = torch.randn((cout, cin, 1, 1), generator=g, device=dev) / cin ** 0.5
        spec = ops.ConvSpec.from_conv(wt, act=ops.ACT_RELU, pad=0)
        pre = (torch.rand(cin, generator=g, device=dev) + 0.5, torch.randn(cin, generator=g, device=dev) * 0.1)
        dst = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=dev)
        kw = dict(c0=cin, pre=pre, dst=dst, tune=tune)
        ops.conv2d(spec, buf, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            ops.conv2d(spec, buf, **kw)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        return ms, 2.0 * n * h * w * cin * cout / (ms * 1e-3) / 1e12
    srcs = [torch.randn((n, h, w, c), generator=g, device=dev).to(torch.bfloat16) for c in cins]
    cin = sum(cins)
    kwargs = {}
    if kind == "t":
        wt = torch.randn((cin, cout, 4, 4), generator=g, device=dev) / (cin * 4) ** 0.5
        spec = ops.ConvSpec.from_convT(wt, act=ops.ACT_RELU)
        flops = 2.0 * n * h * w * 16 * cin * cout
    elif kind in ("stem7", "stem3"):
        kk = 7 if kind == "stem7" else 3
        wt = torch.randn((cout, 3, kk, kk), generator=g, device=dev) / (3 * kk * kk) ** 0.5
        spec = ops.ConvSpec.from_stem(wt, cin, act=ops.ACT_RELU)
        flops = 2.0 * n * h * w * kk * kk * 3 * cout
    else:
        wt = torch.randn((cout, cin, k, k), generator=g, device=dev) / (cin * k * k) ** 0.5
        act = ops.ACT_TANH if kind == "img" else ops.ACT_RELU
        spec = ops.ConvSpec.from_conv(wt, act=act, stride=2 if kind == "s2" else 1, pad=(1 if kind == "s2" else k // 2))
        oh, ow = (h // 2, w // 2) if kind == "s2" else (h, w)
        flops = 2.0 * n * oh * ow * k * k * cin * cout
        if kind == "img":
            xb = torch.rand((n, 3, h, w), device=dev)
            kwargs = dict(epi=ops.EPI_IMAGE, image=dict(mode=ops.IMG_RESIDUAL, x=xb, out=torch.empty_like(xb),
                                                        index=torch.arange(n, dtype=torch.int32, device=dev)))
        elif kind == "dot":
            kwargs = dict(epi=ops.EPI_DOT, dot=(torch.randn(16, device=dev), 0.1, torch.empty((n, h, w), device=dev)))
        elif kind == "res":
            kwargs = dict(residual=torch.randn((n, h, w, cout), generator=g, device=dev).to(torch.bfloat16))
        elif kind == "resin":      # residual = the conv's own input (L2-hot rows): separates DRAM traffic from access cost
            kwargs = dict(residual=srcs[0])
        elif kind == "resdst":     # in place, like ResidualBlock.conv2 in the engine: dst == residual
            r = torch.randn((n, h, w, cout), generator=g, device=dev).to(torch.bfloat16)
            kwargs = dict(residual=r, dst=r)
    src1 = srcs[1] if len(srcs) > 1 else None
    if kwargs.get("epi") is None:
        dst = ops.conv2d(spec, srcs[0], src1, tune=tune, **kwargs)
        kwargs.setdefault("dst", dst)
        if kind != "t":
            kwargs.update(_stat_kwargs())
    else:
        ops.conv2d(spec, srcs[0], src1, tune=tune, **kwargs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.conv2d(spec, srcs[0], src1, tune=tune, **kwargs)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return ms, flops / (ms * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None)
    ap.add_argument("--ab-flags", type=int, default=None, help="time each shape with default tuning and with these tune flags (e.g. 128 = no ragged 64-channel boxes)")
    ap.add_argument("--flags", type=int, default=None, help="time each shape with exactly these tune flags (16 = CTA pair, 32 = 1-CTA, 512 = no rolling-row kernel)")
    ap.add_argument("--stats", default=None, choices=["bn", "pool"], help="also write the epilogue channel partials (adb_conv_desc.stat_out)")
    args = ap.parse_args()
    global STATS
    STATS = {None: None, "bn": True, "pool": "pool"}[args.stats]
    peak = 1414.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    tunes = [None]
    if args.ab_flags is not None:
        tunes = [None, {"flags": args.ab_flags}]
    if args.flags is not None:
        tunes = [{"flags": args.flags}]
    if args.sweep:
        tunes = [{"flags": 32}, {"flags": 16}, {"flags": 16, "mt": 1}, {"flags": 16, "mt": 2}, {"flags": 32, "mt": 1}, {"flags": 32, "mt": 2}]
    for shape in SHAPES:
        if args.only and args.only not in shape[0]:
            continue
        for tune in tunes:
            try:
                ms, tf = run(shape, tune, args.reps)
                print(f"{shape[0]:22s} tune={str(tune):28s} {ms:8.3f} ms  {tf:8.1f} TFLOP/s  {tf / peak:6.3f} of sustained", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{shape[0]:22s} tune={tune} FAILED: {e}", flush=True)


if __name__ == "__main__":
    main()
