#!/bin/bash
# 8-GPU C2: phases with the Gram / inverse chain on the side stream (default) and on the main stream
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for sg in ${SG_LIST:-1 0}; do
SMK_PHASES=1 SMK_SIDE_GRAM=$sg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/d_sg${sg}_n$N.json 2> gpurun_out/d_sg${sg}_n$N.err
python - <<PY
import json
j = json.loads(open("gpurun_out/d_sg${sg}_n$N.json").read().strip().splitlines()[-1])
print("side_gram=$sg", round(j["ms_per_step"], 4), "parity", j["parity"]["ok"], {k: round(v, 4) for k, v in j["phases_ms_per_step"].items()})
PY
done
