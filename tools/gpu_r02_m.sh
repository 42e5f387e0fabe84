#!/bin/bash
# multi-GPU validation of the current build: parity on all ranks, bench with phases (side-stream Grams on / off)
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/m_pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/m_pytest_multi_n$N.log
tail -3 gpurun_out/m_pytest_multi_n$N.log
run() { # name, bench args, env...
  name=$1; shift; bargs=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 $bargs > gpurun_out/m_${name}_n$N.json 2> gpurun_out/m_${name}_n$N.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/m_${name}_n$N.json").read().strip().splitlines()[-1])
    print("$name:", round(j["value"], 1), "it/s", round(j["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in j["roofline"]["launch_ms"].items()}, "parity", j["parity"]["ok"], j["parity"].get("max_rel_err_vs_reference"), "e2e", j["e2e"] and round(j["e2e"]["value"], 1), {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()})
    for nm, e in (j.get("extra") or {}).items():
        print("   extra", nm, e.get("value"), e.get("ms_per_step"), e.get("error"), (e.get("roofline") or {}).get("launch_ms"), {k: round(v, 3) for k, v in (e.get("phases_ms_per_step") or {}).items()})
except Exception as ex:
    print("$name: failed", ex); print(open("gpurun_out/m_${name}_n$N.err").read()[-1500:])
PY
}
run plain "$EXTRAS_PLAIN"
run phases "--no-extras --no-e2e" SMK_PHASES=1
[ -n "$SKIP_NOSIDEGRAM" ] || run phases_nosidegram "--no-extras --no-e2e" SMK_PHASES=1 SMK_SIDE_GRAM=0
