#!/bin/bash
# Run on the GPU box (through gpurun): GPU tests with per-test isolation of failures.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
sel="${1:-tests}"
timeout 900 python -m pytest $sel -m gpu -q --timeout=180 -p no:cacheprovider -rf > gpurun_out/pytest.log 2>&1
rc=$?
tail -40 gpurun_out/pytest.log
if [ $rc -ne 0 ]; then
  # a trapped kernel poisons the CUDA context: re-run each failed test in a fresh process
  grep -E "^FAILED " gpurun_out/pytest.log | awk '{print $2}' | head -12 > gpurun_out/failed.txt
  while read -r t; do
    echo "=== isolated: $t" >> gpurun_out/isolated.log
    timeout 240 python -m pytest "$t" -m gpu -q --timeout=180 -p no:cacheprovider -x 2>&1 | tail -25 >> gpurun_out/isolated.log
  done < gpurun_out/failed.txt
  tail -120 gpurun_out/isolated.log
fi
exit $rc
