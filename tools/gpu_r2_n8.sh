#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2m_scale_n8.json 2> gpurun_out/r2m_scale_n8.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2m_scale_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus 4 --steps 20 --warmup 3 --no-extra > gpurun_out/r2m_scale_n4.json 2> gpurun_out/r2m_scale_n4.err
echo "bench n4 rc=$?"
