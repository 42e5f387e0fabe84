#!/bin/bash
# ncu --set full captures of the C3 (sparse HALS) kernels: one launch each.  usage: gpu_prof_c3.sh "name:skip" ...
mkdir -p gpurun_out
for spec in "$@"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/prof_c3_$k -f \
     python tools/measure_c3_c4.py c3 > gpurun_out/prof_c3_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
