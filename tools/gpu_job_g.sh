#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests; echo "pytest rc=$?" > gpurun_out/rc_g.txt
{
for sel in med_ ; do
  python tools/prof_conv.py --only $sel --reps 10
  python tools/prof_conv.py --only $sel --reps 10 --flags 2048
done
} 2>&1 | tee gpurun_out/prof_med_split.txt
timeout 900 python bench.py --no-eager --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?" >> gpurun_out/rc_g.txt
ADB_NO_EPILOGUE_STATS=1 timeout 600 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/train_nostat_g.json 2> gpurun_out/train_nostat_g.err; echo "train_nostat rc=$?" >> gpurun_out/rc_g.txt
python - <<'PY'
import json
for f in ("bench_g", "train_nostat_g"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"))
        if f == "bench_g":
            for m in ("low", "medium", "high", "densenet121"):
                print("  ", m, d["per_branch_ms_per_image"][m]["ms"], d["per_branch_ms_per_image"][m]["ms_by_entry_point"])
            print("  train", d["train"]["value"], d["train"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/rc_g.txt
