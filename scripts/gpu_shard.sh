#!/bin/bash
# sharded-proof measurement (ONE gkr::prove_gkr over N GPUs): bash scripts/gpu_shard.sh N [extra bench args]
cd "$(dirname "$0")/.."
N=${1:-2}; shift
mkdir -p gpurun_out
for T in 0 1; do
HG_SHARD_TIMING=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$T bench.py --gpus $N --steps 30 --warmup 3 --shard "$@" > gpurun_out/shard_${N}gpu_t$T.json 2> gpurun_out/shard_${N}gpu_t$T.err
grep "shard rank" gpurun_out/shard_${N}gpu_t$T.err
python - <<PY
import json
d=json.loads(open("gpurun_out/shard_${N}gpu_t$T.json").read().strip().splitlines()[-1])["shard"]
print("timing=$T gpus", d["gpus"], "sharded ms", round(d["ms_per_proof"],3), "one gpu ms", round(d["one_gpu_ms_per_proof"],3), "speedup", round(d["speedup_vs_1gpu"],3), "bytes_equal", d["bytes_equal"])
PY
done
