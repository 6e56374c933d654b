#!/bin/bash
# single-GPU bench lines: Goldilocks (default) and BN254 (config 5)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err; echo "gl rc=$?"; tail -2 gpurun_out/bench_gl.err
timeout 500 python bench.py --steps 10 --warmup 3 --field bn254 --inflight 2 > gpurun_out/bench_bn.json 2> gpurun_out/bench_bn.err; echo "bn rc=$?"; tail -2 gpurun_out/bench_bn.err
python - <<'PY'
import json
for n in ("gl","bn"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(n, "value", round(d["value"],2), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],2), "launches", d["gpu_launches_per_proof"], "witgen", d["witness_gen_ms_per_witness"]["generate_on_device_ms"])
        print("   dom", r["kernel"], "frac", round(r["frac"],3), "gp pipeline frac", round(r["grand_product_pipeline"]["frac"],3), "whole single-stream", round(r["whole_proof_frac_single_stream"],3), "inflight", round(r["whole_proof_frac_inflight"],3))
        print("   cpu", d["cpu_baseline"])
        for k,v in r["per_class"].items(): print("      ", k, v["launches"], round(v["ms"],3), round(v["alg_GB"],3))
    except Exception as e: print(n, "ERR", e)
PY
