#!/bin/bash
# One GPU checkpoint (through gpurun): GPU tests, inference + reference + training bench lines, ncu launch lists of the
# same commands at a small batch, isolated wgrad timings and one `ncu --set full` of the wgrad kernel.
mkdir -p gpurun_out
TAG="${1:-r1b}"
timeout 900 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider -rf > gpurun_out/pytest_${TAG}.log 2>&1; tail -5 gpurun_out/pytest_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_train.json 2> gpurun_out/bench_${TAG}_train.err
SMALL="--mode train --steps 1 --warmup 3 --batch 4"
python bench.py $SMALL > gpurun_out/plain_train_small_${TAG}.json 2> gpurun_out/plain_train_small_${TAG}.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train_${TAG}.csv python bench.py $SMALL > gpurun_out/ncu_launches_train_${TAG}.log 2>&1
ISMALL="--steps 1 --warmup 3 --batch 12 --no-e2e --no-cpu-baseline"
python bench.py $ISMALL > gpurun_out/plain_small_${TAG}.json 2> gpurun_out/plain_small_${TAG}.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "adb_timed/" --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py $ISMALL > gpurun_out/ncu_launches_${TAG}.log 2>&1
# dram bytes of the conv launches of bench.py's roofline pass (NVTX ranges adb_roofline_<model>) -> roofline.traffic
timeout 600 ncu --nvtx --nvtx-include "adb_roofline_low/" --nvtx-include "adb_roofline_medium/" --nvtx-include "adb_roofline_high/" \
    --nvtx-include "adb_roofline_densenet121/" -k regex:conv_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/ncu_traffic_${TAG}.csv python bench.py $ISMALL > gpurun_out/ncu_traffic_${TAG}.log 2>&1
python tools/conv_traffic.py gpurun_out/ncu_traffic_${TAG}.csv > gpurun_out/conv_traffic_${TAG}.json 2> gpurun_out/conv_traffic_${TAG}.err
python tools/prof_wgrad.py > gpurun_out/prof_wgrad_${TAG}.txt 2>&1
python tools/prof_bn.py > gpurun_out/prof_bn_${TAG}.txt 2>&1
for shape in med_64_3x3 cpx_192_3x3 cpx_384_3x3; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_wgrad_${shape} \
      python tools/prof_wgrad.py --only $shape --reps 2 > gpurun_out/ncu_wgrad_${shape}.log 2>&1
done
cat gpurun_out/bench_${TAG}.json; cat gpurun_out/bench_${TAG}_train.json; cat gpurun_out/prof_wgrad_${TAG}.txt
