mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x --timeout=300 > gpurun_out/test_pipeline.log 2>&1; echo "pipeline exit $?"; tail -4 gpurun_out/test_pipeline.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().split("\n")[-1])
print(d['value'], d['ms_per_step'], d['e2e'], {k:(v['ms'] if isinstance(v,dict) else v) for k,v in d['stages'].items()})
PY
