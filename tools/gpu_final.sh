#!/bin/bash
# developer tool, runs ON the GPU box: what the driver runs at round end -- GPU suite, smoke, default bench, reference arm --
# and then the ncu launch list of the headline bench command (a number printed under ncu is never a bench value)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "reference arm rc=$?"; cat gpurun_out/final_bench_reference.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/final_launches_bench.json 2> gpurun_out/final_launches.err
echo "launch list rc=$?"
