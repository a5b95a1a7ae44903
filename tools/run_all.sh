bash tools/gpu_suite.sh > gpurun_out/suite.log 2>&1
cat gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:(round(v['ms'],4),round(v.get('frac',0),3)) if isinstance(v,dict) else v for k,v in d['stages'].items()})
print('cpu',d.get('cpu_baseline'))
print('roofline',d['roofline']['achieved'],d['roofline']['frac'])
"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 600 python bench.py --workload em > gpurun_out/em_n1.log 2>&1; echo "em exit $?"; tail -1 gpurun_out/em_n1.log | cut -c1-700
timeout 600 python bench.py --workload stream > gpurun_out/stream_n1.log 2>&1; echo "stream exit $?"; tail -1 gpurun_out/stream_n1.log | cut -c1-700
