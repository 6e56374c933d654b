#!/bin/bash
# throughput vs proofs in flight per GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 3 4 5 6 8; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --inflight $N > gpurun_out/inflight_$N.json 2> gpurun_out/inflight_$N.err; echo "inflight $N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/inflight_$N.json").read().strip().splitlines()[-1])
    print("inflight $N value", round(d["value"],1), "ms/proof", round(d["ms_per_proof"],3), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("ERR", e)
PY
done | tee gpurun_out/inflight.log
