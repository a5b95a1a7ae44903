mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_em.py tests/test_gpu_pipeline.py -q -m gpu -x --timeout=200 > gpurun_out/test_detect.log 2>&1; echo "detect+em+pipeline exit $?"; tail -4 gpurun_out/test_detect.log
timeout 200 python tools/time_kernels.py detect 2>&1 | tail -1
timeout 200 python bench.py --workload stream --no-cpu --steps 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stream ms', d['ms_per_step'])"
timeout 200 python bench.py --workload em --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('em ms', d['ms_per_step'])"
