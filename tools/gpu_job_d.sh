#!/bin/bash
mkdir -p gpurun_out
{
python tools/timeline.py cpx_192_3x3 cpx_192_3x3_resdst cpx_96_3x3 med_128_3x3
python tools/timeline.py --detail cpx_192_3x3 cpx_192_3x3_resdst cpx_96_3x3 med_128_3x3
for sel in cpx_96_3x3 cpx_192_3x3 cpx_384_3x3; do
  python tools/prof_conv.py --only $sel --reps 10
  python tools/prof_conv.py --only $sel --reps 10 --stats pool
done
} 2>&1 | tee gpurun_out/timeline_epi.txt | cut -c1-1500
