"""DRAM bytes per conv launch of bench.py's roofline pass, from an ncu CSV.

    ncu --nvtx --nvtx-include "adb_roofline_low/" --nvtx-include "adb_roofline_medium/" --nvtx-include "adb_roofline_high/" \
        --nvtx-include "adb_roofline_densenet121/" -k regex:conv_ \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/ncu_traffic.csv python bench.py --steps 1 --warmup 3 --batch 12 --no-e2e --no-cpu-baseline
    python tools/conv_traffic.py gpurun_out/ncu_traffic.csv [bench line of the same command.json] > profiles/r2_conv_traffic.json

bench.py wraps each model's instrumented pass (8 images at 1024x2048) in an NVTX range `adb_roofline_<model>`; the average
uses bench.py's own mix weighting (every model's bytes AND launches enter with weight 1 for Light/Medium/Complex and 3 for
HDEN, which runs on every image) so that `roofline.traffic` and `roofline.flops_per_launch_avg` describe the same average launch.
"""
import csv
import json
import sys


def main():
    path = sys.argv[1]
    # images per model of the captured pass: from the bench line of the same command (roofline.per_model[*].images_per_launch)
    images = {}
    if len(sys.argv) > 2:
        with open(sys.argv[2]) as fh:
            line = json.load(fh)
        images = {k: v.get("images_per_launch") for k, v in line["roofline"]["per_model"].items()}
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    nvtx_col = next((c for c in rd.fieldnames if "Push/Pop_Range" in c or "NVTX" in c.upper()), None)
    per = {}
    for r in rd:
        key = r["ID"]
        d = per.setdefault(key, {"range": r.get(nvtx_col, "") if nvtx_col else "", "name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "")
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        elif r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1.0)
        d[r["Metric Name"]] = v
    models = {}
    for d in per.values():
        if "conv_igemm" not in d["name"] and "conv_roll" not in d["name"]:
            continue
        rng = d["range"]
        m = next((k for k in ("low", "medium", "high", "densenet121", "resnet18") if f"adb_roofline_{k}" in rng), None)
        if m is None:
            continue
        a = models.setdefault(m, {"launches": 0, "read": 0.0, "write": 0.0, "ns": 0.0})
        a["launches"] += 1
        a["read"] += d.get("dram__bytes_read.sum", 0.0)
        a["write"] += d.get("dram__bytes_write.sum", 0.0)
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    hden = "densenet121" if "densenet121" in models else "resnet18"
    launches = sum(a["launches"] for a in models.values())
    tot = sum((a["read"] + a["write"]) * (3 if k == hden else 1) for k, a in models.items())
    w_launches = sum(a["launches"] * (3 if k == hden else 1) for k, a in models.items())
    out = {"dram_bytes_per_launch_avg": tot / max(1, w_launches), "launches": launches, "images": 8, "height": 1024, "width": 2048,
           "hden": hden, "per_model": {k: {"launches": a["launches"], "images": images.get(k) or 8, "dram_read_bytes": a["read"],
                                            "dram_write_bytes": a["write"], "ncu_ms": a["ns"] / 1e6} for k, a in models.items()},
           "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the adb_conv2d (conv_igemm / conv_roll) launches of bench.py's roofline pass ({path}), "
                     "weighting as flops_per_launch_avg"}
    json.dump(out, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
