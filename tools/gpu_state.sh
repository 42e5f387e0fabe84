#!/bin/bash
# State check on one GPU: smoke, C2 bench line, parity tests, C3/C4 measurements (most important first).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 400 python tools/measure_c3_c4.py c3 > gpurun_out/c3.log 2>&1; tail -2 gpurun_out/c3.log
timeout 400 python tools/measure_c3_c4.py c4 > gpurun_out/c4.log 2>&1; tail -2 gpurun_out/c4.log
