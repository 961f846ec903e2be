#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "peer_group or host_batch" > gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_tests.log; tail -3 gpurun_out/r2i_tests.log
for spec in "2 512" "4 512" "2 4096"; do set -- $spec
  timeout 200 python tools/dev_peer_latency.py --world $1 --side $2 --steps 100 2>&1 | grep "rank" | sort
done | tee gpurun_out/r2i_peer_latency.log | cut -c1-200
