#!/bin/bash
# C3 / C4 profiling pass: ncu launch lists (time only) and ncu --set full of the two SpMM launches of one C3 iteration.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv \
   python tools/measure_c3_c4.py c3 > gpurun_out/c3_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_seg -s 4 -c 2 -o gpurun_out/prof_c3_spmm_seg -f \
   python tools/measure_c3_c4.py c3 > gpurun_out/prof_c3_spmm_seg.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c4.csv \
   python tools/measure_c3_c4.py c4 > gpurun_out/c4_ncu.log 2>&1
ls -la gpurun_out
