#!/bin/bash
# developer tool, runs ON the GPU box: first round-2 GPU pass -- tests, the new bench line, A/B of the staged experiments,
# ncu captures of the kernels that round 2 works on.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
timeout 500 tools/ab_run.sh > gpurun_out/r2a_ab.log 2>&1
cat gpurun_out/r2a_ab.log
for spec in "4 1 overlap cfg4" "3 1 overlap cfg3" "2 1 overlap cfg2" "4 2 fast cfg4fast"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o gpurun_out/r2a_$4 \
    python tools/profile_target.py --config $1 --mode $2 --arith f32 --steps 1 > gpurun_out/r2a_ncu_$4.log 2>&1
  echo "ncu $4 rc=$?"
done
