#!/bin/bash
# launch list (gpu__time_duration) of the W-side HALS kernels at C3
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hals_block" -s 34 -c 40 --csv --log-file gpurun_out/v_launches_hals.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/v_launches_hals.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
cur={}
for r in rows[hi+2:]:
    if len(r)<len(hdr): continue
    key=(r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0][-40:])
    cur.setdefault(key,{})[r[ix["Metric Name"]]]=r[ix["Metric Value"]]
for (i,n),m in list(cur.items())[:20]:
    print(i, n, m)
PY
