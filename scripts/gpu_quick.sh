#!/bin/bash
# GPU tests + one bench line (no CPU baseline): the check after a change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/quick_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/quick_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/quick_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", round(d["value"],1), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],1), "e2e lat", round(d["e2e"]["single_proof_latency_ms"],3), "launches", d["gpu_launches_per_proof"], "frac", round(r["frac"],4))
print("   ", " ".join(f"{k}={v['launches']:.0f}/{v['ms']:.3f}" for k,v in r["per_class"].items()))
print("   host", {k: round(v) for k,v in d["host_phases_us"].items()})
PY
