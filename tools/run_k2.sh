mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_code_parts.py -q -m gpu -x --timeout=300 > gpurun_out/test_code_parts.log 2>&1; echo "code_parts exit $?"; tail -4 gpurun_out/test_code_parts.log
timeout 300 python tools/time_kernels.py code_parts 2>&1 | tail -3
for f in 0 4; do timeout 120 python tools/trace_code_parts.py $f 2>&1 | tail -9; done
