#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) --set full captures of the
# dominant kernels.  Each ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
for k in detect code_parts spread_codes spread_edges threshold_blocks; do
  python tools/prof_stage.py $k --iters 2 > gpurun_out/prof_plain_$k.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${k} -s 1 -c 1 -f -o gpurun_out/r01_$k \
      python tools/prof_stage.py $k --iters 2 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
