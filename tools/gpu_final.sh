mkdir -p gpurun_out
T=${1:-r1o}
python tools/nccl_stdout_check.py > gpurun_out/nccl_stdout_${T}.txt 2> gpurun_out/nccl_stderr_${T}.txt; echo "nccl stdout:"; cat gpurun_out/nccl_stdout_${T}.txt; echo "nccl stderr lines: $(wc -l < gpurun_out/nccl_stderr_${T}.txt)"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${T}.log 2>&1; tail -2 gpurun_out/smoke_${T}.log
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_${T}.log 2>&1; tail -3 gpurun_out/pytest_${T}.log
python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; cat gpurun_out/bench_${T}.json
