#!/bin/bash
# ncu captures of the two tensor kernels (after a plain run of the same command exits 0)
mkdir -p gpurun_out
python tools/prof_stage.py all --iters 3 > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
cat gpurun_out/prof_plain.log
python tools/prof_stage.py code_parts --iters 2 > gpurun_out/prof_plain_code.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:code_parts_tc -s 1 -c 1 -f -o gpurun_out/prof_code python tools/prof_stage.py code_parts --iters 2 > gpurun_out/ncu_code.log 2>&1
echo "ncu code exit $?"
python tools/prof_stage.py detect --iters 2 > gpurun_out/prof_plain_detect.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:detect_tc -s 1 -c 1 -f -o gpurun_out/prof_detect python tools/prof_stage.py detect --iters 2 > gpurun_out/ncu_detect.log 2>&1
echo "ncu detect exit $?"
ls -la gpurun_out/*.ncu-rep
