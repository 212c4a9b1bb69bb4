mkdir -p gpurun_out
T=${1:-r1i}
timeout 900 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider -rf > gpurun_out/pytest_${T}.log 2>&1; tail -5 gpurun_out/pytest_${T}.log
timeout 120 python tools/prof_bn.py > gpurun_out/prof_bn_${T}.txt 2>&1
timeout 300 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_${T}_train.json 2> gpurun_out/bench_${T}_train.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bn_|chan_reduce|affine_act|add_bf16" -c 60 --csv --log-file gpurun_out/ncu_bn_${T}.csv python tools/prof_bn.py > gpurun_out/ncu_bn_${T}.log 2>&1
cat gpurun_out/prof_bn_${T}.txt; cat gpurun_out/bench_${T}_train.json
