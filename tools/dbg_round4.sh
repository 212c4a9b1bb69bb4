mkdir -p gpurun_out
T=${1:-r1j}
timeout 600 python -m pytest tests/test_gpu_train.py -q --timeout=240 -p no:cacheprovider -rf > gpurun_out/pytest_train_${T}.log 2>&1; tail -5 gpurun_out/pytest_train_${T}.log
timeout 120 python tools/prof_bn.py > gpurun_out/prof_bn_${T}.txt 2>&1
timeout 300 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_${T}_train.json 2> gpurun_out/bench_${T}_train.err
cat gpurun_out/prof_bn_${T}.txt; cat gpurun_out/bench_${T}_train.json
