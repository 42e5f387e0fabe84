#!/bin/bash
# round 2, call F (8 GPUs): multi-rank parity on all 8, the bench line with extras, phase breakdowns for peer and NCCL exchange.
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/f_pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/f_pytest_multi_n$N.log
tail -3 gpurun_out/f_pytest_multi_n$N.log
run() { # name, bench args, env...
  name=$1; shift; bargs=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 $bargs > gpurun_out/f_${name}_n$N.json 2> gpurun_out/f_${name}_n$N.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/f_${name}_n$N.json").read().strip().splitlines()[-1])
    print("$name:", round(j["value"], 1), "it/s", round(j["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in j["roofline"]["launch_ms"].items()}, "parity", j["parity"], "e2e", j["e2e"] and round(j["e2e"]["value"], 1), {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()})
    for nm, e in (j.get("extra") or {}).items():
        print("   extra", nm, e.get("value"), e.get("ms_per_step"), e.get("error"), (e.get("roofline") or {}).get("launch_ms"), {k: round(v, 3) for k, v in (e.get("phases_ms_per_step") or {}).items()})
except Exception as ex:
    print("$name: failed", ex); print(open("gpurun_out/f_${name}_n$N.err").read()[-1500:])
PY
}
run peer "--extras c5,c3" SMK_PEER=1
run peer_phases "--extras c5,c3 --no-e2e" SMK_PEER=1 SMK_PHASES=1
run nccl_phases "--no-extras --no-e2e" SMK_PEER=0 SMK_PHASES=1
