mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q --timeout=240 -p no:cacheprovider -rf > gpurun_out/pytest_train_r1h.log 2>&1; tail -8 gpurun_out/pytest_train_r1h.log
timeout 120 python tools/prof_wgrad.py > gpurun_out/prof_wgrad_r1h.txt 2>&1
timeout 120 python tools/prof_bn.py > gpurun_out/prof_bn_r1h.txt 2>&1
timeout 300 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_r1h_train.json 2> gpurun_out/bench_r1h_train.err
cat gpurun_out/prof_wgrad_r1h.txt gpurun_out/prof_bn_r1h.txt; cat gpurun_out/bench_r1h_train.json
