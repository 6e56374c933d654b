#!/bin/bash
# A/B of one environment switch: bash scripts/gpu_ab.sh VAR  (runs the GPU tests first, then bench with VAR=1 and VAR=0)
cd "$(dirname "$0")/.."
VAR=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/ab_tests.log
for V in 1 0; do
env $VAR=$V timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pool 2 > gpurun_out/ab_${VAR}_$V.json 2> gpurun_out/ab_${VAR}_$V.err; echo "bench $VAR=$V rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${VAR}_$V.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$VAR=$V value", round(d["value"],1), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches_per_proof"], "single-stream ms", round(r["whole_proof_kernel_ms_single_stream"],3))
    print("   ", " ".join(f"{k}={v['launches']:.0f}/{v['ms']:.3f}" for k,v in r["per_class"].items()))
except Exception as e: print("ERR", e)
PY
done
