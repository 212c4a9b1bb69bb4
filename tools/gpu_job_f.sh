#!/bin/bash
mkdir -p gpurun_out
{
for sel in cpx_192_3x3 cpx_384_3x3 cpx_down cpx_up cpx_cat_192_96 med_256_3x3 med_128_3x3; do
  python tools/prof_conv.py --only $sel --reps 10
  python tools/prof_conv.py --only $sel --reps 10 --flags 1024
done
} 2>&1 | tee gpurun_out/prof_nsplit.txt
