#!/bin/bash
# Multi-GPU check: 2-rank parity test, then bench.py at N = 1, 2, ... up to the GPUs present (C2), then C5 at the full GPU count.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1.log 2>&1
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
         bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
    fi
    tail -1 gpurun_out/bench_n$N.log | cut -c1-300
  fi
done
if [ $NG -ge 4 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
     tools/measure_c5_multi.py > gpurun_out/c5_n$NG.log 2>&1; tail -1 gpurun_out/c5_n$NG.log | cut -c1-500
fi
