#!/bin/bash
# round-2 verification job: GPU suite, smoke, train A/B of the epilogue statistics, default bench
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests; echo "pytest rc=$?" > gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
tail -5 gpurun_out/smoke.log
ADB_NO_EPILOGUE_STATS=1 timeout 600 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/train_nostat.json 2> gpurun_out/train_nostat.err; echo "train_nostat rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/train_stat.json 2> gpurun_out/train_stat.err; echo "train_stat rc=$?" >> gpurun_out/rc.txt
tail -c 600 gpurun_out/train_nostat.json; echo; tail -c 600 gpurun_out/train_stat.json; echo
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/rc.txt
tail -c 1500 gpurun_out/bench_default.json
cat gpurun_out/rc.txt
