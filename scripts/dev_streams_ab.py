"""A/B of HG_OPT_TWO_STREAMS on the full-size GKR proof (device-resident values)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import hyper_greco_b200  # noqa
from hyper_greco_b200 import api

P, inp, bounds, segs, nv = bench.make_case(bench.DEFAULT_CONFIG, 0)
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
out_claims = [(np.zeros((0, el), np.uint64), np.zeros(el, np.uint64)), (point, value)]

def step():
    tr = api.Keccak256Transcript()
    tr.squeeze_challenges(prover.ct0is_log2_size)
    prover.circuit.prove_gkr(out_claims, tr, api.MODE_PREFETCH)
    return tr.into_proof()

ref = None
for two in (0, 1, 0, 1):
    ctx.set_option(100, two)
    for _ in range(5):
        pr = step()
    ref = ref or pr
    assert pr == ref
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(30):
        step()
    dt = (time.perf_counter() - t0) / 30
    print(f"two_streams={two}: {dt*1e3:.3f} ms/proof", prover.circuit.timing())
