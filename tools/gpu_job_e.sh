#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -q --timeout=180 -p no:cacheprovider -rf -x > gpurun_out/pytest_e.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_e.log
{
for sel in cpx_96_3x3 cpx_192_3x3 cpx_384_3x3 med_128_3x3 med_256_3x3 cpx_cat_192_96 cpx_96_48 cpx_down cpx_up dense_1x1 med_64_3x3; do
  python tools/prof_conv.py --only $sel --reps 10
done
python tools/prof_conv.py --only cpx_192_3x3 --reps 10 --stats pool
python tools/timeline.py --detail cpx_192_3x3 cpx_192_3x3_resdst cpx_96_3x3
} 2>&1 | tee gpurun_out/prof_epi2.txt | cut -c1-1200
