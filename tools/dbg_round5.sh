mkdir -p gpurun_out
T=${1:-r1k}
timeout 300 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_${T}_train.json 2> gpurun_out/bench_${T}_train.err
tail -3 gpurun_out/bench_${T}_train.err
ADB_PROFILE_HOST=1 timeout 300 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_${T}_train_prof.json 2> gpurun_out/bench_${T}_train_prof.err
grep -v "^ContentLoss\|^PerceptualLoss" gpurun_out/bench_${T}_train_prof.err | head -90
python -c "
import json
d=json.load(open('gpurun_out/bench_${T}_train.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['train_detail']['ms_by_entry_point'])
"
