#!/bin/bash
# Runs the GPU test files one process each (a trapped kernel kills only its own file), then
# smoke and a short bench.  Usage (on the GPU box): bash tools/gpu_suite.sh [bench-args]
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/gpu.txt 2>&1
for f in ${TSR_SUITE_FILES:-em spread edges code_parts detect pipeline full_size}; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout=600 -s > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/test_$f.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
