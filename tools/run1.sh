mkdir -p gpurun_out
bash tools/gpu_suite.sh --steps 20 --warmup 5 > gpurun_out/suite.log 2>&1
cat gpurun_out/summary.txt
timeout 120 python tools/refine_probe.py 2>&1 | grep -v "timed out" | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
