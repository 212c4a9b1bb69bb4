"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

    python tools/summarize_launches.py gpurun_out/launches.csv [--range adb_timed] > profiles/rNN_launches.md

With --range only launches inside that NVTX push/pop range are counted (bench.py wraps its timed region in
`adb_timed`).  ncu serialises launches and runs them cold, so the absolute times are not the bench's — the SHARE of each
kernel is what must agree with the CUDA-event numbers bench.py prints.
"""
import argparse
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.match(r"([\w:]+(?:<[^(]{0,60})?)", name)
    return (m.group(1) if m else name)[:90]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--range", default=None)
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    nvtx_col = next((c for c in rd.fieldnames if "Push/Pop_Range" in c), None)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        if args.range and nvtx_col and args.range not in r.get(nvtx_col, ""):
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        elif r.get("Metric Unit") in ("ms", "msecond"):
            ns *= 1e6
        rows.append((short(r["Kernel Name"]), ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    tot = sum(ns for _, ns, _, _ in rows) or 1.0
    agg = collections.OrderedDict()
    for k, ns, g, b in rows:
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
    print(f"launches: {len(rows)}   total (serialised, cold): {tot / 1e6:.3f} ms" + (f"   NVTX range: {args.range}" if args.range else ""))
    print()
    print("| kernel | launches | total ms | share | avg us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, (n, ns, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / n / 1e3:.1f} | {mx / 1e3:.1f} |")


if __name__ == "__main__":
    sys.exit(main())
