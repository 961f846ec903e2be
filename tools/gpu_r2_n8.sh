#!/bin/bash
# developer tool, runs ON an 8-GPU box (gpurun --gpus 8): the 8-GPU bench line of the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3t_scale_n8.json 2> gpurun_out/r3t_scale_n8.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r3t_scale_n8.err
