mkdir -p gpurun_out
for f in edges em; do
timeout 600 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --timeout=300 > gpurun_out/test_$f.log 2>&1; echo "$f exit $?"; tail -15 gpurun_out/test_$f.log
done
