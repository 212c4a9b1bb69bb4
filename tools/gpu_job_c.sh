#!/bin/bash
mkdir -p gpurun_out
for sel in cpx_96_3x3 cpx_192_3x3 cpx_384_3x3 med_128_3x3 med_256_3x3 cpx_192_3x3_resdst cpx_down_96_192; do
  python tools/prof_conv.py --only $sel --reps 10
  python tools/prof_conv.py --only $sel --reps 10 --stats pool
  python tools/prof_conv.py --only $sel --reps 10 --stats bn
done 2>&1 | grep -v "^$" | tee gpurun_out/prof_stats_ab.txt
