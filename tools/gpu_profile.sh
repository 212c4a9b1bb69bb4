#!/bin/bash
# Run on the GPU box (through gpurun): the bench line, the ncu launch list of the same command at a small batch, and one
# `ncu --set full` capture per representative conv shape.  Everything lands in gpurun_out/ (copy summaries to profiles/).
mkdir -p gpurun_out
TAG="${1:-r1}"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
SMALL="--steps 1 --warmup 3 --batch 12 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > gpurun_out/plain_small_${TAG}.json 2> gpurun_out/plain_small_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "adb_timed/" --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py $SMALL > gpurun_out/ncu_launches_${TAG}.log 2>&1
for shape in med_256_3x3 cpx_192_3x3 med_64_3x3; do
  python tools/prof_conv.py --only $shape --reps 2 > gpurun_out/plain_${shape}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_${shape} \
      python tools/prof_conv.py --only $shape --reps 2 > gpurun_out/ncu_${shape}.log 2>&1
done
cat gpurun_out/bench_${TAG}.json
