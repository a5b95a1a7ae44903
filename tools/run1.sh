mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu_all.log
