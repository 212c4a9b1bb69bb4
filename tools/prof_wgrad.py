"""Isolated weight-gradient launches for ncu / CUDA-event timing:  python tools/prof_wgrad.py [--only NAME] [--reps N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from adam_dehaze_b200 import ops  # noqa: E402

SHAPES = {   # name: (n, h, w, cin, cout, k, kind)   — the training bench's layers at 16 x 512 x 512
    "light_32_3x3": (16, 512, 512, 32, 32, 3, ops.CONV_S1),
    "med_64_3x3": (16, 512, 512, 64, 64, 3, ops.CONV_S1),
    "med_128_3x3": (16, 256, 256, 128, 128, 3, ops.CONV_S1),
    "med_256_3x3": (16, 128, 128, 256, 256, 3, ops.CONV_S1),
    "cpx_96_3x3": (16, 512, 512, 96, 96, 3, ops.CONV_S1),
    "cpx_192_3x3": (16, 256, 256, 192, 192, 3, ops.CONV_S1),
    "cpx_384_3x3": (16, 128, 128, 384, 384, 3, ops.CONV_S1),
    "med_down_64_128": (16, 512, 512, 64, 128, 4, ops.CONV_S2),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 channel-major kernel, 2 tap-packed kernel")
    args = ap.parse_args()
    print(f"{'shape':18s} {'ms':>8s} {'TFLOP/s':>9s}")
    for name, (n, h, w, ci, co, k, kind) in SHAPES.items():
        if args.only and name != args.only:
            continue
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn((n, h, w, ci), generator=g, device="cuda").to(torch.bfloat16)
        hs, ws = (h // 2, w // 2) if kind == ops.CONV_S2 else (h, w)
        dz = torch.randn((n, hs, ws, co), generator=g, device="cuda").to(torch.bfloat16)
        out = torch.empty((co, ci, k, k), dtype=torch.float32, device="cuda")
        kw = dict(kind=kind, kh=k, kw=k, pad=1 if k > 1 else 0, out=out, mode=args.mode)
        ops.wgrad(dz, x, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            ops.wgrad(dz, x, **kw)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.reps
        fl = 2.0 * n * hs * ws * k * k * ci * co
        print(f"{name:18s} {ms:8.3f} {fl / ms / 1e9:9.1f}")


if __name__ == "__main__":
    main()
