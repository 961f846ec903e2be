#!/bin/bash
# developer tool, runs ON the GPU box: round-2 pass H -- peer-group sync latency probe, FP64 A/B, final captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "2 512" "4 512" "2 4096"; do set -- $spec
  timeout 200 python tools/dev_peer_latency.py --world $1 --side $2 --steps 100 2>&1 | grep "rank" | sort
done | tee gpurun_out/r2h_peer_latency.log
for lib in "" "--lib area_average_interpolation_b200/csrc/gpurun_variants/f64seq.so" "" "--lib area_average_interpolation_b200/csrc/gpurun_variants/f64seq.so"; do
  echo -n "f64 [$lib]: "; timeout 120 python tools/profile_target.py --config 4 --arith f64 --steps 30 $lib 2>&1 | tail -1
done | tee gpurun_out/r2h_f64_ab.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2h_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_tests.log; tail -3 gpurun_out/r2h_tests.log
timeout 300 python tools/dev_bands.py --config 4 --parts 4 8 > gpurun_out/r2h_bands_cfg4.log 2>&1; tail -6 gpurun_out/r2h_bands_cfg4.log
for spec in "4 1 f32 overlap cfg4" "2 1 f32 overlap cfg2" "4 2 f32 fast cfg4fast" "4 1 f64 overlap cfg4f64"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -f -o gpurun_out/r2h_$5 \
    python tools/profile_target.py --config $1 --mode $2 --arith $3 --steps 1 > gpurun_out/r2h_ncu_$5.log 2>&1
  echo "ncu $5 rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2h_bench.err
