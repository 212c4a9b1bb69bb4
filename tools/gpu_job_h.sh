#!/bin/bash
mkdir -p gpurun_out
{
for sel in cpx_192_3x3 cpx_384_3x3 cpx_96_3x3 cpx_cat_192_96 cpx_up_cat_384_96 cpx_96_48 med_up_cat_256_64 med_cat_128_64; do
  python tools/prof_conv.py --only $sel --reps 10 --sweep
done
python tools/timeline.py cpx_96_3x3 cpx_cat_192_96 med_up_cat_256_64
} 2>&1 | tee gpurun_out/prof_sweep_h.txt | cut -c1-400
