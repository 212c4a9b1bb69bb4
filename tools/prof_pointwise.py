"""Time (CUDA events) the HBM-bound kernels at the shapes of the 1024x2048 pipeline; prints ms and achieved GB/s of the
ALGORITHMIC bytes (each logical tensor read once + written once) against the measured HBM peak.

    python tools/prof_pointwise.py [--reps 10] [--n 4]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adam_dehaze_b200 import ops  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--n", type=int, default=4)
    args = ap.parse_args()
    peak = 6450.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n, H, W = args.n, 1024, 2048
    dev = "cuda"
    x = torch.rand((n, 3, H, W), device=dev)
    rows = []

    def rec(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        rows.append((name, ms, gbs))
        print(f"{name:34s} {ms:8.3f} ms  {gbs:8.0f} GB/s  {gbs / peak:5.2f} of measured HBM peak", flush=True)

    for (kh, kw, pad, stride, kp) in [(1, 3, 1, 1, 16), (1, 7, 3, 1, 32), (7, 7, 3, 2, 160)]:
        ho, wo = (H // stride, W // stride)
        out = torch.empty((n, ho if kh > 1 else H, wo, kp), dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.stem_pack(x, kw, pad, kp, stride=stride, kh=kh, out=out), args.reps)
        rec(f"stem_pack k{kh}x{kw} s{stride} kp{kp}", ms, x.numel() * 4 + out.numel() * 2)
    for (c, h, w) in [(96, H, W), (192, H // 2, W // 2), (384, H // 4, W // 4)]:
        f = torch.randn((n, h, w, c), device=dev).to(torch.bfloat16)
        ap_ = ops.AttnParams(torch.randn(c // 16, c, 1, 1, device=dev), torch.randn(c, c // 16, 1, 1, device=dev),
                             torch.randn(1, 2, 7, 7, device=dev))
        scratch = {}
        y = torch.empty_like(f)
        ops.attention(f, ap_, out=y, scratch=scratch)
        from adam_dehaze_b200 import _lib
        st = _lib.current_stream()
        pool = scratch[("pool", n, h, w, c)]
        gate, stats, spatial = scratch[("gate", n, c)], scratch[("stats", n, h, w)], scratch[("spatial", n, h, w)]
        nb = f.numel() * 2
        ms = timeit(lambda: _lib.call("adb_attn_pool", _lib.ptr(f), n, h, w, c, None, 0, _lib.ptr(pool), st), args.reps)
        rec(f"attn_pool c{c} {h}x{w}", ms, nb)
        ms = timeit(lambda: _lib.call("adb_attn_gate_stats", _lib.ptr(f), n, h, w, c, None, 0, _lib.ptr(pool), _lib.ptr(ap_.w1),
                                      _lib.ptr(ap_.w2), ap_.c_red, _lib.ptr(gate), _lib.ptr(stats), st), args.reps)
        rec(f"attn_gate_stats c{c} {h}x{w}", ms, nb + n * h * w * 8)
        ms = timeit(lambda: _lib.call("adb_attn_apply", _lib.ptr(f), n, h, w, c, None, 0, _lib.ptr(gate), _lib.ptr(stats),
                                      _lib.ptr(ap_.wsp), _lib.ptr(spatial), _lib.ptr(y), st), args.reps)
        rec(f"attn_apply c{c} {h}x{w}", ms, 2 * nb + n * h * w * 16)
        del f, y
    for (c, pitch, h, w) in [(160, 256, 256, 512), (304, 512, 128, 256), (624, 1024, 64, 128)]:
        buf = torch.randn((n * 2, h, w, pitch), device=dev).to(torch.bfloat16)
        sc, sh = torch.rand(c, device=dev), torch.rand(c, device=dev)
        out = torch.empty((n * 2, h, w, c), dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.affine_relu(buf, c, sc, sh, out=out), args.reps)
        rec(f"affine_relu c{c}/{pitch} {h}x{w}", ms, out.numel() * 4)
    f = torch.randn((n * 2, 512, 1024, 64), device=dev).to(torch.bfloat16)
    out = torch.empty((n * 2, 256, 512, 256), dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.maxpool3x3s2(f, out=out), args.reps)
    rec("maxpool3x3s2 c64 512x1024", ms, f.numel() * 2 + n * 2 * 256 * 512 * 64 * 2)


if __name__ == "__main__":
    main()
