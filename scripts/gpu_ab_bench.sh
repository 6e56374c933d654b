#!/bin/bash
# GPU tests, then the whole-proof bench with the both product trees per builder launch on / off (HG_TREE_PAIR)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/e10_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/e10_tests.log
for V in 1 0; do
env HG_TREE_PAIR=$V timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pool 2 > gpurun_out/e10_pair_$V.json 2> gpurun_out/e10_pair_$V.err; echo "bench HG_TREE_PAIR=$V rc=$?"; tail -2 gpurun_out/e10_pair_$V.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e10_pair_$V.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("HG_TREE_PAIR=$V value", round(d["value"],1), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches_per_proof"], "single-stream ms", round(r["whole_proof_kernel_ms_single_stream"],3), "frac", round(r["frac"],4))
    print("   ", " ".join(f"{k}={v['launches']:.0f}/{v['ms']:.3f}" for k,v in r["per_class"].items()))
    print("   host", {k: round(v) for k,v in d["host_phases_us"].items()})
except Exception as e: print("ERR", e)
PY
done
