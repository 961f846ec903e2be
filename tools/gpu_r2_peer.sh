#!/bin/bash
# developer tool (gpurun --gpus N): end-to-end peer-group timing with 1 / 2 / 4 / 8 upload chunks per owner, headline only
cd "$(dirname "$0")/.."
N=${1:-8}; TAG=${2:-r2g}
mkdir -p gpurun_out
for c in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$c \
    bench.py --gpus $N --steps 20 --warmup 3 --no-extra --peer-chunks $c > gpurun_out/${TAG}_peer_n${N}_c$c.json 2> gpurun_out/${TAG}_peer_n${N}_c$c.err
  echo "chunks $c rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_peer_n${N}_c$c.json").read())
print("  e2e ms", round(d["e2e"]["ms_per_step"],3), "verified", d.get("verified"), d["e2e"].get("phases_ms_per_rank"))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-extra --no-peer > gpurun_out/${TAG}_nopeer_n${N}.json 2> gpurun_out/${TAG}_nopeer_n${N}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_nopeer_n${N}.json").read())
print("no peer: e2e ms", round(d["e2e"]["ms_per_step"],3), "verified", d.get("verified"))
PY
