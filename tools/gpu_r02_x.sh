#!/bin/bash
# the 8-GPU column shard of C2 on one GPU: in-kernel split reduction for W'A too?
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for f in 0 1; do
SMK_GEMM_FIXUP=$f SMK_PHASES=1 timeout 300 python tools/measure_dense.py 20000 2500 64 BPP 30 > gpurun_out/x_shard_fix$f.json 2> gpurun_out/x_shard_fix$f.err; python -c "
import json; j=json.loads(open('gpurun_out/x_shard_fix$f.json').read().strip().splitlines()[-1]); print('shard fixup=$f', round(j['ms_per_iter'],4), {k: round(v,4) for k,v in j['phases_ms'].items()})"
SMK_GEMM_FIXUP=$f SMK_PHASES=1 timeout 300 python tools/measure_dense.py 20000 20000 64 BPP 10 > gpurun_out/x_full_fix$f.json 2> gpurun_out/x_full_fix$f.err; python -c "
import json; j=json.loads(open('gpurun_out/x_full_fix$f.json').read().strip().splitlines()[-1]); print('full fixup=$f', round(j['ms_per_iter'],4), {k: round(v,4) for k,v in j['phases_ms'].items()})"
done
