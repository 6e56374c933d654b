"""One full GKR proof of the bench workload bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import hyper_greco_b200
from hyper_greco_b200 import api
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_CONFIG
P, inp, bounds, segs, nv = bench.make_case(name, 0)
ins, ct0is = bench.LAST_WITNESS
ctx = api.Context(0)
prover = api.BfvSkEncryptProver(ctx, P)
flat = [ins["s"], ins["e"], ins["k1"]] + list(ins["ais"]) + list(ins["r1is"]) + [ins["r2is"]]
dev_inputs = [api.DeviceBuffer.from_numpy(ctx, np.array(v, dtype=np.uint64)) for v in flat]
d_ct = api.DeviceBuffer.from_numpy(ctx, np.array(ct0is, dtype=np.uint64))
prover.circuit.evaluate(dev_inputs)
tr0 = api.Keccak256Transcript()
point = tr0.squeeze_challenges(prover.ct0is_log2_size)
value = api.mle_eval_batch(ctx, d_ct, 1, prover.ct0is_log2_size, point)[0]
el = point.shape[1]
oc = [(np.zeros((0, el), np.uint64), np.zeros(el, np.uint64)), (point, value)]
def step():
    tr = api.Keccak256Transcript(); tr.squeeze_challenges(prover.ct0is_log2_size)
    prover.circuit.prove_gkr(oc, tr)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
