#!/bin/bash
# the driver's multi-GPU line: weak scaling (independent proofs) + the sharded-proof measurement in ONE bench line. bash scripts/gpu_scale.sh N
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_${N}gpu.json 2> gpurun_out/scale_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/scale_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_${N}gpu.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],3), "clocks", d.get("clocks"))
print("shard", d.get("shard"))
PY
