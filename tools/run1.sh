mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_code_parts.py -q -m gpu --timeout=60 -x -k "int8 or 200_parts" > gpurun_out/test_code_parts.log 2>&1; echo "code_parts tests exit $?"; grep -v "timed out" gpurun_out/test_code_parts.log | tail -2
timeout 60 python tools/refine_probe.py 2>&1 | grep -v "timed out" | tail -3
for f in "" "--fuse"; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra $f 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['stages'].items() if isinstance(v,dict) and 'ms' in v})
"; done
