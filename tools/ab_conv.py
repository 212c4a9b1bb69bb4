"""A/B timing of one conv shape under two tune settings in the SAME process, interleaved (run-to-run clock/power drift
cancels):  python tools/ab_conv.py NAME [--a '{}'] [--b '{"flags": 512}'] [--rounds 6]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import prof_conv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="+")
    ap.add_argument("--a", default="null")
    ap.add_argument("--b", default='{"flags": 512}')
    ap.add_argument("--rounds", type=int, default=5)
    args = ap.parse_args()
    ta, tb = json.loads(args.a), json.loads(args.b)
    for shape in prof_conv.SHAPES:
        if not any(n in shape[0] for n in args.names):
            continue
        ra, rb = [], []
        for _ in range(args.rounds):
            ra.append(prof_conv.run(shape, ta, 5)[0])
            rb.append(prof_conv.run(shape, tb, 5)[0])
        ma, mb = sorted(ra)[len(ra) // 2], sorted(rb)[len(rb) // 2]
        print(f"{shape[0]:22s} A={ta}: {ma:.4f} ms   B={tb}: {mb:.4f} ms   A/B = {ma / mb:.3f}", flush=True)


if __name__ == "__main__":
    main()
