#!/bin/bash
# Quick GPU pass: the parity tests named in $1 (pytest -k expression, "" = all), then the workloads named in $2.. (c3 c4 c3small c4small bench)
mkdir -p gpurun_out
K="$1"; shift
if [ -n "$K" ]; then timeout 1200 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu.log 2>&1
else timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; fi
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for w in "$@"; do
  if [ "$w" = bench ]; then timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-600
  else timeout 600 python tools/measure_c3_c4.py $w > gpurun_out/$w.log 2>&1; tail -2 gpurun_out/$w.log | cut -c1-900; fi
done
