#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" "--lib area_average_interpolation_b200/csrc/gpurun_variants/fastnoskew.so" "" "--lib area_average_interpolation_b200/csrc/gpurun_variants/fastnoskew.so"; do
  for c in 4 2; do echo -n "fast cfg$c [$lib]: "; timeout 120 python tools/profile_target.py --config $c --mode 2 --arith f32 --dst float32 --steps 50 $lib 2>&1 | tail -1; done
done | tee gpurun_out/r2r_fast_skew_ab.log
timeout 900 python -m pytest tests -m gpu -q -k "fast or staged or batch or canvas_taller" > gpurun_out/r2r_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_tests.log; tail -4 gpurun_out/r2r_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast -s 1 -c 1 -f -o gpurun_out/r2r_cfg4fast \
    python tools/profile_target.py --config 4 --mode 2 --arith f32 --steps 1 > gpurun_out/r2r_ncu_cfg4fast.log 2>&1
