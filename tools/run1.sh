mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_code_parts.py -q -m gpu --timeout=200 -x -k "int8 or 200_parts" -s > gpurun_out/test_cp_int8.log 2>&1; echo "int8 tests exit $?"; grep -v "timed out" gpurun_out/test_cp_int8.log | tail -6
timeout 120 python tools/refine_probe.py 2>&1 | grep -v "timed out" | tail -5
