#!/bin/bash
# Round-2 profile pass (one gpurun call): ncu launch lists of the bench commands at a small batch, DRAM bytes of the conv
# launches of bench.py's roofline pass, `ncu --set full` captures of the rolling-row kernel and of conv_igemm (1-CTA vs
# CTA-pair on the same shape for the tensor-pipe counter reconciliation).  Every ncu run follows the same command's plain run.
mkdir -p gpurun_out
TAG="${1:-r2}"
EXTRA="sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg,sm__pipe_tensor_cycles_active.max,sm__pipe_tensor_cycles_active.min,sm__cycles_elapsed.avg,sm__cycles_elapsed.max,sm__cycles_active.avg,smsp__inst_executed_pipe_uniform.sum"
ISMALL="--steps 1 --warmup 3 --batch 12 --no-e2e --no-cpu-baseline --no-eager --no-train"
python bench.py $ISMALL > gpurun_out/plain_small_${TAG}.json 2> gpurun_out/plain_small_${TAG}.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "adb_timed/" --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py $ISMALL > gpurun_out/ncu_launches_${TAG}.log 2>&1
python bench.py $ISMALL > /dev/null 2>&1 &&
timeout 900 ncu --nvtx --nvtx-include "adb_roofline_low/" --nvtx-include "adb_roofline_medium/" --nvtx-include "adb_roofline_high/" \
    --nvtx-include "adb_roofline_densenet121/" -k regex:conv_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/ncu_traffic_${TAG}.csv python bench.py $ISMALL > gpurun_out/ncu_traffic_${TAG}.log 2>&1
python tools/conv_traffic.py gpurun_out/ncu_traffic_${TAG}.csv gpurun_out/plain_small_${TAG}.json > gpurun_out/conv_traffic_${TAG}.json 2> gpurun_out/conv_traffic_${TAG}.err
for shape in light_32_3x3 med_64_3x3 dense_3x3_128_32; do
  python tools/prof_conv.py --only $shape --reps 2 > gpurun_out/plain_${shape}.log 2>&1 &&
  timeout 300 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:conv_roll -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_roll_${shape} \
      python tools/prof_conv.py --only $shape --reps 2 > gpurun_out/ncu_${shape}.log 2>&1
done
for fl in 16 32; do
  python tools/prof_conv.py --only med_256_3x3 --flags $fl --reps 2 > gpurun_out/plain_med256_${fl}.log 2>&1 &&
  timeout 300 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:conv_igemm -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_med256_flags${fl} \
      python tools/prof_conv.py --only med_256_3x3 --flags $fl --reps 2 > gpurun_out/ncu_med256_${fl}.log 2>&1
done
python tools/prof_conv.py --only cpx_96_3x3 --reps 2 > gpurun_out/plain_cpx96.log 2>&1 &&
timeout 300 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:conv_igemm -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_cpx_96_3x3 \
    python tools/prof_conv.py --only cpx_96_3x3 --reps 2 > gpurun_out/ncu_cpx96.log 2>&1
SMALL="--mode train --steps 1 --warmup 3 --batch 4"
python bench.py $SMALL > gpurun_out/plain_train_small_${TAG}.json 2> gpurun_out/plain_train_small_${TAG}.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train_${TAG}.csv python bench.py $SMALL > gpurun_out/ncu_launches_train_${TAG}.log 2>&1
python tools/prof_conv.py > gpurun_out/prof_conv_${TAG}.txt 2>&1
ls -la gpurun_out | tail -30
