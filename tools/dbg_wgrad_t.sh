set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_train.py -x -q -k "densenet121" > gpurun_out/dbg_dn.log 2>&1
rc=$?
echo "rc=$rc" >> gpurun_out/dbg_dn.log
if [ $rc -ne 0 ]; then
  timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_train.py -x -q -k "densenet121" > gpurun_out/dbg_dn_san.log 2>&1
fi
timeout 120 python tools/prof_wgrad.py --mode 1 > gpurun_out/prof_wgrad_m1.txt 2>&1
timeout 120 python tools/prof_wgrad.py --mode 2 > gpurun_out/prof_wgrad_m2.txt 2>&1
tail -5 gpurun_out/dbg_dn.log
