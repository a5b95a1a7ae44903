#!/bin/bash
# Round-2 ncu evidence for profiles/: (1) launch list of the default bench command (pipeline + the em / stream extras),
# (2) --set full captures of the dominant kernels of configs 3 and 4.  Each ncu run follows a plain run of the same
# command that exited 0.  Usage (GPU box): bash tools/gpu_profiles_r02.sh
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
for k in detect code_parts spread_codes spread_edges detect_float threshold_blocks; do
  python tools/prof_stage.py $k --iters 2 > gpurun_out/prof_plain_$k.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${k} -s 1 -c 1 -f -o gpurun_out/r02_$k \
      python tools/prof_stage.py $k --iters 2 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
# EM (config 4): the E and M contractions (bitgemm*), the posterior kernel and the finalize
python bench.py --workload em --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_plain_em.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bitgemm|em_' -s 40 -c 12 -f -o gpurun_out/r02_em \
    python bench.py --workload em --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_em.log 2>&1
echo "ncu em exit $?"
ls -la gpurun_out/r02_*
# the exact 24-bit variant (INT8X3, detect_tc_kernel) for comparison
python tools/prof_stage.py detect --iters 2 --planes 4 > gpurun_out/prof_plain_detect_x3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:detect_tc -s 1 -c 1 -f -o gpurun_out/r02_detect_x3 \
    python tools/prof_stage.py detect --iters 2 --planes 4 > gpurun_out/ncu_detect_x3.log 2>&1
echo "ncu detect_x3 exit $?"
