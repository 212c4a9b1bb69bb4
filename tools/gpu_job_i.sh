#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests; echo "pytest rc=$?" > gpurun_out/rc_i.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc_i.txt; tail -2 gpurun_out/smoke.log
ADB_NO_POOL_FOLD=1 timeout 900 python bench.py --no-eager --no-cpu-baseline --no-train > gpurun_out/bench_nofold.json 2> gpurun_out/bench_nofold.err; echo "bench_nofold rc=$?" >> gpurun_out/rc_i.txt
timeout 1200 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?" >> gpurun_out/rc_i.txt
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_i.json 2> gpurun_out/bench_ref_i.err; echo "bench_ref rc=$?" >> gpurun_out/rc_i.txt
python - <<'PY'
import json
for f in ("bench_nofold", "bench_i", "bench_ref_i"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"))
        if "per_branch_ms_per_image" in d:
            for m in ("low", "medium", "high", "densenet121"):
                print("  ", m, d["per_branch_ms_per_image"][m]["ms"], d["per_branch_ms_per_image"][m]["ms_by_entry_point"])
        if d.get("train"): print("  train", d["train"]["value"], d["train"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/rc_i.txt
