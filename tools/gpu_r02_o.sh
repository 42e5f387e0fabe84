#!/bin/bash
# gather-rate probe; a C2 column shard (20000 x 2500, the 8-GPU shape) on one GPU: phases and the ncu launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 tools/l2_gather_peak > gpurun_out/o_gather_peak.txt 2>&1; echo "gather rc=$?"; cat gpurun_out/o_gather_peak.txt
SMK_PHASES=1 timeout 300 python tools/measure_dense.py 20000 2500 64 BPP 20 > gpurun_out/o_shard_phases.json 2> gpurun_out/o_shard_phases.err; echo "shard rc=$?"; cut -c1-900 gpurun_out/o_shard_phases.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/o_launches_shard.csv python tools/measure_dense.py 20000 2500 64 BPP 3 > gpurun_out/o_ncu_shard.log 2>&1; echo "ncu rc=$?"
