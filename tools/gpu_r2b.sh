#!/bin/bash
# developer tool, runs ON the GPU box: round-2 pass B -- whole GPU suite, device times of the rewritten kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
tail -25 gpurun_out/r2b_tests.log
for spec in "4 1" "3 1" "2 1" "4 2" "3 2" "2 2"; do
  set -- $spec
  timeout 120 python tools/profile_target.py --config $1 --mode $2 --arith f32 --steps 30 2>&1 | tail -1
done | tee gpurun_out/r2b_times.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench.err
for spec in "3 1 overlap cfg3" "4 2 fast cfg4fast"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o gpurun_out/r2b_$4 \
    python tools/profile_target.py --config $1 --mode $2 --arith f32 --steps 1 > gpurun_out/r2b_ncu_$4.log 2>&1
  echo "ncu $4 rc=$?"
done
