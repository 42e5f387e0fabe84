#!/bin/bash
# Profiling pass: C3 launch list + ncu --set full of the two SpMM launches, gemm_skinny full capture on the C2 bench, C4 launch list.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv \
   python tools/measure_c3_c4.py c3 > gpurun_out/c3_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_seg_wide -s 4 -c 2 -o gpurun_out/prof_c3_spmm_wide -f \
   python tools/measure_c3_c4.py c3 > gpurun_out/prof_c3_spmm_wide.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 6 -c 2 -o gpurun_out/prof_gemm -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_gemm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_c4.csv \
   python tools/measure_c3_c4.py c4small > gpurun_out/c4_ncu.log 2>&1
ls -la gpurun_out
