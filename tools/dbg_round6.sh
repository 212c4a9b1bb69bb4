mkdir -p gpurun_out
T=${1:-r1l}
timeout 900 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider -rf > gpurun_out/pytest_${T}.log 2>&1; tail -5 gpurun_out/pytest_${T}.log
timeout 300 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_${T}_train.json 2> gpurun_out/bench_${T}_train.err
tail -3 gpurun_out/bench_${T}_train.err
timeout 300 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_${T}b_train.json 2> gpurun_out/bench_${T}b_train.err
tail -2 gpurun_out/bench_${T}b_train.err
python -c "
import json
for f in ['gpurun_out/bench_${T}_train.json','gpurun_out/bench_${T}b_train.json']:
    d=json.load(open(f))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['train_detail']['ms_by_entry_point'])
"
