"""Per-launch conv table of one model at 1024x2048 (CUDA events around every adb_conv2d call).

    python tools/prof_model.py [low|medium|high|densenet121|resnet18] [--n 8]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from adam_dehaze_b200 import ops as ops_mod  # noqa: E402
import adam_dehaze_b200.engine as eng  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", nargs="?", default="high")
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    hden = args.model if args.model in ("densenet121", "resnet18") else "densenet121"
    cfg = dict(bench.CFG, classifier=dict(bench.CFG["classifier"], model=hden))
    branches, clf = bench.build_models(cfg, dev)
    m = clf if args.model in ("densenet121", "resnet18") else branches[args.model]
    x, _ = bench.synth_batch_on_device(args.n, args.height, args.width, dev, 42)
    rec = []
    orig = ops_mod.conv2d

    def timed_conv(spec, src0, src1=None, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig(spec, src0, src1, **kw)
        b.record()
        n = kw.get("n") or src0.shape[0]
        cin = src0.shape[3] + (src1.shape[3] if src1 is not None else 0)
        rec.append((a, b, bench.conv_flops(spec, src0, src1, n, kw), spec.kind, spec.kh, spec.kw, cin, spec.cout,
                    src0.shape[1], src0.shape[2], n))
        return r

    m(x)
    torch.cuda.synchronize()
    eng.ops.conv2d = timed_conv
    ops_mod.conv2d = timed_conv
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); m(x); e.record()
    torch.cuda.synchronize()
    tot = sum(a.elapsed_time(b) for a, b, *_ in rec)
    print(f"{args.model}: {s.elapsed_time(e) / args.n:.3f} ms/img total, conv {tot / args.n:.3f} ms/img over {len(rec)} launches")
    agg = {}
    for a, b, fl, kind, kh, kw, cin, cout, h, w, n in rec:
        key = (kind, kh, kw, cin, cout, h, w)
        t = agg.setdefault(key, [0, 0.0, 0.0])
        t[0] += 1; t[1] += a.elapsed_time(b); t[2] += fl
    print(f"{'kind':>4} {'k':>4} {'cin':>5} {'cout':>5} {'h':>5} {'w':>5} {'cnt':>4} {'ms/img':>8} {'share':>6} {'TFLOP/s':>8}")
    for key, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        kind, kh, kw, cin, cout, h, w = key
        print(f"{kind:>4} {kh}x{kw:<2} {cin:>5} {cout:>5} {h:>5} {w:>5} {cnt:>4} {ms / args.n:8.4f} {100 * ms / tot:5.1f}% {fl / ms / 1e9:8.1f}")


if __name__ == "__main__":
    main()
