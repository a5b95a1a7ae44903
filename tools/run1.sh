mkdir -p gpurun_out
bash tools/gpu_suite.sh --steps 20 --warmup 5 > gpurun_out/suite.log 2>&1
cat gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc $?"
prof() {  # suffix, prof_stage args
  local sfx=$1; shift
  python tools/prof_stage.py code_parts --iters 2 "$@" > gpurun_out/prof_plain_$sfx.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:code_parts -s 2 -c 2 -f -o gpurun_out/r02_$sfx \
      python tools/prof_stage.py code_parts --iters 2 "$@" > gpurun_out/ncu_$sfx.log 2>&1
  echo "ncu $sfx exit $?"
}
prof code_parts
prof code_parts_int8 --impl-code 2
prof code_parts_int8_200 --impl-code 2 --parts 200
