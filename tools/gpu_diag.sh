#!/bin/bash
# Tree parity at scale: GPU hierclust (fused and generic rank-2) against the reference run with one thread (sequential initialiser).
mkdir -p gpurun_out
timeout 400 python tests/manual/diag_c4.py 40000 250000 16 1 > gpurun_out/diag_c4small_fused.log 2>&1; grep nodes gpurun_out/diag_c4small_fused.log | cut -c1-900
SMK_RANK2_FUSED=0 timeout 400 python tests/manual/diag_c4.py 40000 250000 16 1 > gpurun_out/diag_c4small_generic.log 2>&1; grep nodes gpurun_out/diag_c4small_generic.log | cut -c1-900
timeout 900 python tests/manual/diag_c4.py 320000 2000000 4 1 > gpurun_out/diag_c4_4_fused.log 2>&1; grep nodes gpurun_out/diag_c4_4_fused.log | cut -c1-900
