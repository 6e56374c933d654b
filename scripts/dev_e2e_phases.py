"""Where the end-to-end step (host vectors -> proof bytes) spends its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import hyper_greco_b200  # noqa
from hyper_greco_b200 import api
from hyper_greco_b200 import params
P = params.PARAMS[bench.DEFAULT_CONFIG]
ins, ct0is = bench.make_witness(bench.DEFAULT_CONFIG, 0)
ctx = api.Context(0)
prover = api.BfvSkEncryptProver(ctx, P)
flat = [ins["s"], ins["e"], ins["k1"]] + list(ins["ais"]) + list(ins["r1is"]) + [ins["r2is"]]
import torch
def pin(v):
    t = torch.from_numpy(np.array(v, dtype=np.uint64).view(np.int64)).pin_memory()
    return t.numpy().view(np.uint64), t
pinned = [pin(v) for v in flat]
host = [p[0] for p in pinned]
dev_inputs = [api.DeviceBuffer.from_numpy(ctx, v) for v in host]
d_ct = api.DeviceBuffer.from_numpy(ctx, np.array(ct0is, dtype=np.uint64))
el = 2
acc = {}
def tick(name, t0):
    ctx.synchronize(); t1 = time.perf_counter(); acc[name] = acc.get(name, 0) + (t1 - t0); return t1
def step(record):
    t = time.perf_counter()
    tr = api.Keccak256Transcript()
    prover.circuit.evaluate_host(host)
    if record: t = tick("upload+evaluate", t)
    pt = tr.squeeze_challenges(prover.ct0is_log2_size)
    val = api.mle_eval_batch(ctx, d_ct, 1, prover.ct0is_log2_size, pt)[0]
    if record: t = tick("out_claim", t)
    prover.circuit.prove_gkr([(np.zeros((0, el), np.uint64), np.zeros(el, np.uint64)), (pt, val)], tr, api.MODE_PREFETCH)
    if record: t = tick("prove_gkr", t)
    pr = tr.into_proof()
    if record: t = tick("into_proof", t)
    return pr
for _ in range(3): step(False)
N = 20
for _ in range(N): step(True)
print({k: round(1e3 * v / N, 3) for k, v in acc.items()}, "total", round(1e3 * sum(acc.values()) / N, 3))
ctx.profile(True)
prover.circuit.evaluate(dev_inputs)
ctx.synchronize()
prof = ctx.profile_read(); ctx.profile(False)
print({k: (v[0], round(v[1], 3)) for k, v in prof.items() if v[0]})
