#!/bin/bash
# round-2 GPU batch A: parity tests, then bench with the fused round 0 on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err; echo "bench fused rc=$?"
HG_GP_FUSE_R0=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_unfused.json 2> gpurun_out/r2a_bench_unfused.err; echo "bench unfused rc=$?"
python - <<'PY'
import json
for n in ("fused","unfused"):
    try:
        d=json.loads(open(f"gpurun_out/r2a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"],1), "ms/proof", round(d["ms_per_proof"],3), "lat", round(d["single_proof_latency_ms"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches_per_proof"])
        for k,v in d["roofline"]["per_class"].items(): print("   ", k, v["launches"], round(v["ms"],3), round(v["alg_GB"],3))
    except Exception as e: print(n, "ERR", e)
PY
