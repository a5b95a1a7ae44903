mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detect.py -q -m gpu -x --timeout=600 > gpurun_out/test_detect.log 2>&1; echo "detect exit $?"; tail -4 gpurun_out/test_detect.log
timeout 300 python tools/time_kernels.py detect code_parts 2>&1 | tail -3
