#!/bin/bash
# developer tool, runs ON a multi-GPU box (gpurun --gpus N): the real multi-device checks + the N-GPU bench line
# usage: tools/gpu_r2_multi.sh N [tag]
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_n${N}_gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "several_devices or peer_group or multi_device" > gpurun_out/${TAG}_n${N}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_n${N}_tests.log; tail -4 gpurun_out/${TAG}_n${N}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_scale_n${N}.json 2> gpurun_out/${TAG}_scale_n${N}.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_scale_n${N}.err; head -c 1200 gpurun_out/${TAG}_scale_n${N}.json
