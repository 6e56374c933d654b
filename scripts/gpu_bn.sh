#!/bin/bash
# BN254: parity tests, then the config-5 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 -k "bn254 or fullsize_bn254" > gpurun_out/bn_tests.log 2>&1; echo "bn tests rc=$?"; tail -4 gpurun_out/bn_tests.log
timeout 500 python bench.py --steps 10 --warmup 3 --field bn254 --inflight 2 --pool 2 ${BN_ARGS} > gpurun_out/bench_bn.json 2> gpurun_out/bench_bn.err; echo "bn rc=$?"; tail -2 gpurun_out/bench_bn.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_bn.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bn254 value", round(d["value"],2), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],2), "e2e lat", round(d["e2e"]["single_proof_latency_ms"],2))
    print("   cpu", d["cpu_baseline"])
    print("   ", " ".join(f"{k}={v['launches']:.0f}/{v['ms']:.2f}" for k,v in r["per_class"].items()))
except Exception as e: print("ERR", e)
PY
