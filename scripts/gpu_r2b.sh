#!/bin/bash
# round-2 GPU batch B (2 GPUs): multi-rank bench with the sharded measurement in the line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; echo "bench ${N} gpu rc=$?"
tail -3 gpurun_out/r2b_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "lat", round(d["single_proof_latency_ms"],3))
    print("shard", json.dumps(d["shard"])[:600])
except Exception as e: print("ERR", e)
PY
