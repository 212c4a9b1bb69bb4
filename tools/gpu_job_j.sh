#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q --timeout=180 -p no:cacheprovider -rf -x > gpurun_out/pytest_j.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_j.log
{
for sel in _res med_256_3x3 cpx_192_3x3; do
  python tools/prof_conv.py --only $sel --reps 10
done
python tools/timeline.py --detail cpx_192_3x3_resdst
} 2>&1 | tee gpurun_out/prof_res_j.txt | cut -c1-900
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-train > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_j",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d["clocks"])
        for m in ("low", "medium", "high", "densenet121"):
            print("  ", m, d["per_branch_ms_per_image"][m]["ms"], d["per_branch_ms_per_image"][m]["ms_by_entry_point"])
    except Exception as e:
        print(f, "unreadable", e)
PY
