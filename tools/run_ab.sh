P=template-speech-recognition_b200
cp $P/libtsr_b200.so /tmp/keep.so
for v in old new old new; do
  cp tools/build/libtsr_${v}.so $P/libtsr_b200.so
  echo "== $v"; timeout 100 python tools/time_kernels.py code_parts 2>&1 | tail -1
done
cp /tmp/keep.so $P/libtsr_b200.so
echo "== nissue overrides (new lib)"
for n in 1 2; do
  echo "nissue $n"; TSR_DEBUG_DETECT_NISSUE=$n timeout 100 python tools/time_kernels.py detect 2>&1 | tail -1
  TSR_DEBUG_DETECT_NISSUE=$n timeout 150 python bench.py --workload stream --no-cpu --steps 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stream ms', d['ms_per_step'])"
done
timeout 600 python -m pytest tests/test_gpu_code_parts.py tests/test_gpu_pipeline.py -q -m gpu -x --timeout=200 2>&1 | tail -3
