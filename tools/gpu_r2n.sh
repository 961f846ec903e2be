#!/bin/bash
# developer tool, runs ON the GPU box: round-2 pass N -- fast-mode predicated loads A/B, full suite, launch list, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" "--lib area_average_interpolation_b200/csrc/gpurun_variants/fastall.so" "" "--lib area_average_interpolation_b200/csrc/gpurun_variants/fastall.so"; do
  for c in 4 2 3; do echo -n "fast cfg$c [$lib]: "; timeout 120 python tools/profile_target.py --config $c --mode 2 --arith f32 --steps 50 $lib 2>&1 | tail -1; done
done | tee gpurun_out/r2n_fast_ab.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2n_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_tests.log; tail -3 gpurun_out/r2n_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast -s 1 -c 1 -f -o gpurun_out/r2n_cfg4fast \
    python tools/profile_target.py --config 4 --mode 2 --arith f32 --steps 1 > gpurun_out/r2n_ncu_cfg4fast.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2n_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_launches_bench.json 2> gpurun_out/r2n_launches.err
echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()"
