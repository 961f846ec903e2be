#!/bin/bash
# developer tool, runs ON the GPU box: round-2 pass D -- whole GPU suite, LDG vs TMA-staged A/B, band cost-model data
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
tail -22 gpurun_out/r2d_tests.log
for spec in "4 1 f32" "4 1 f32s" "2 1 f32" "2 1 f32s" "3 1 f32" "4 2 f32" "2 2 f32" "3 2 f32" "4 1 f32" "4 1 f32s"; do
  set -- $spec
  echo -n "mode $2 $3: "; timeout 120 python tools/profile_target.py --config $1 --mode $2 --arith $3 --steps 50 2>&1 | tail -1
done | tee gpurun_out/r2d_times.log
rm -f gpurun_out/r2d_bands.jsonl
timeout 300 python tools/dev_bands.py --config 4 --parts 2 4 8 16 --json gpurun_out/r2d_bands.jsonl > gpurun_out/r2d_bands_cfg4.log 2>&1
timeout 300 python tools/dev_bands.py --config 3 --parts 2 4 8 --json gpurun_out/r2d_bands.jsonl > gpurun_out/r2d_bands_cfg3.log 2>&1
tail -4 gpurun_out/r2d_bands_cfg4.log
for spec in "4 1 f32s overlap cfg4staged" "4 2 f32 fast cfg4fast" "3 1 f32 overlap cfg3"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -f -o gpurun_out/r2d_$5 \
    python tools/profile_target.py --config $1 --mode $2 --arith $3 --steps 1 > gpurun_out/r2d_ncu_$5.log 2>&1
  echo "ncu $5 rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2d_bench.err
