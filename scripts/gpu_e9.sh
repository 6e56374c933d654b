#!/bin/bash
# GPU tests, the end-to-end phase breakdown with the run-based forward evaluation on / off (HG_FWD_RUNS), and the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/e9_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/e9_tests.log
for V in 1 0; do echo "HG_FWD_RUNS=$V"; env HG_FWD_RUNS=$V timeout 300 python scripts/dev_e2e_phases.py 2>&1 | tail -2; done | tee gpurun_out/e9_e2e_phases.txt
for V in 1 0; do
env HG_FWD_RUNS=$V timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/e9_fwd_$V.json 2> gpurun_out/e9_fwd_$V.err; echo "bench HG_FWD_RUNS=$V rc=$?"; tail -2 gpurun_out/e9_fwd_$V.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e9_fwd_$V.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("HG_FWD_RUNS=$V value", round(d["value"],1), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],1), "e2e lat", round(d["e2e"]["single_proof_latency_ms"],3), "launches", d["gpu_launches_per_proof"], "frac", round(r["frac"],4))
except Exception as e: print("ERR", e)
PY
done
