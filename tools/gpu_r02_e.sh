#!/bin/bash
# round 2, call E (N GPUs): fused GEMM + reduce-scatter epilogue vs the two-kernel form; split-count sweep on the sharded products.
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/e_pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/e_pytest_multi_n$N.log
tail -3 gpurun_out/e_pytest_multi_n$N.log
run() { # name, env...
  name=$1; shift
  env SMK_PHASES=1 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-e2e > gpurun_out/e_${name}_n$N.json 2> gpurun_out/e_${name}_n$N.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/e_${name}_n$N.json").read().strip().splitlines()[-1])
    print("$name:", round(j["value"], 1), "it/s", round(j["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in j["roofline"]["launch_ms"].items()}, "parity", j["parity"]["ok"], {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()})
except Exception as ex:
    print("$name: failed", ex); print(open("gpurun_out/e_${name}_n$N.err").read()[-1500:])
PY
}
run fused SMK_PEER_FUSED=1
run unfused SMK_PEER_FUSED=0
run nccl SMK_PEER=0
for s in $SWEEP_NT; do run fused_nt$s SMK_GEMM_SPLITS_NT=$s; done
for s in $SWEEP_NN; do run fused_nn$s SMK_GEMM_SPLITS_NN=$s; done
