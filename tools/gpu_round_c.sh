#!/bin/bash
# C5 on one GPU (full 100000 x 50000, k = 256, BPP) with its launch list; gemm_skinny --set full on the big products of the C2 bench; C4 again.
mkdir -p gpurun_out
timeout 600 python tools/measure_dense.py 100000 50000 256 BPP 3 > gpurun_out/c5_n1.log 2>&1; tail -1 gpurun_out/c5_n1.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv \
   python tools/measure_dense.py 100000 50000 256 BPP 1 > gpurun_out/c5_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 2 -c 8 -o gpurun_out/prof_gemm -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_gemm.log 2>&1
timeout 600 python tools/measure_c3_c4.py c4 > gpurun_out/c4.log 2>&1; grep workload gpurun_out/c4.log | cut -c1-500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c4small.csv \
   python tools/measure_c3_c4.py c4small > gpurun_out/c4small_ncu.log 2>&1
ls gpurun_out | head -40
