#!/bin/bash
# pool-fold check: conv / model tests, Complex per-model profile, inference-only bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_pointwise.py -m gpu -q --timeout=180 -p no:cacheprovider -rf > gpurun_out/pytest_b.log 2>&1
echo "pytest rc=$?" > gpurun_out/rc_b.txt
tail -15 gpurun_out/pytest_b.log
ADB_NO_EPILOGUE_STATS=1 timeout 600 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/train_nostat.json 2> gpurun_out/train_nostat.err; echo "train_nostat rc=$?" >> gpurun_out/rc_b.txt
timeout 900 python bench.py --no-eager --no-train --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?" >> gpurun_out/rc_b.txt
python - <<'PY'
import json
for f in ("train_nostat", "bench_b"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"))
        if f == "bench_b":
            print(json.dumps(d["per_branch_ms_per_image"]["high"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/rc_b.txt
