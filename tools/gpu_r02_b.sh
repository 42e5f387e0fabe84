#!/bin/bash
# round 2, call B (N GPUs, default 2): multi-rank parity (peer-memory exchange and NCCL), bench with phase breakdown for both.
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/b_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/b_pytest_multi.log
tail -5 gpurun_out/b_pytest_multi.log
run() { # name, env..., extra args
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/b_${name}_n$N.json 2> gpurun_out/b_${name}_n$N.err
  echo "$name rc=$?"; tail -c 600 gpurun_out/b_${name}_n$N.json; tail -3 gpurun_out/b_${name}_n$N.err
}
EXTRA="--extras c5,c3" run peer SMK_PEER=1
EXTRA="--no-extras --no-e2e" run peer_phases SMK_PEER=1 SMK_PHASES=1
EXTRA="--no-extras --no-e2e" run nccl_phases SMK_PEER=0 SMK_PHASES=1
EXTRA="--no-extras --no-e2e" run nccl SMK_PEER=0
ls -la gpurun_out | tail -12
