#!/bin/bash
# Round-2 closing check on one B200: GPU suite, smoke, default bench (both arms), as the driver runs them.
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests; echo "pytest rc=$?" > gpurun_out/rc_final.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc_final.txt; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench_ref rc=$?" >> gpurun_out/rc_final.txt
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?" >> gpurun_out/rc_final.txt
python - <<'PY'
import json
for f in ("bench_ref_final", "bench_final"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("gpu_launches"))
        if d.get("train"): print("  train", d["train"]["value"], d["train"]["ms_per_step"], d["train"].get("graphed_step", {}).get("samples_per_s"))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/rc_final.txt
