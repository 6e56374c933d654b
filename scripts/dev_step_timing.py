import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, hyper_greco_b200
from hyper_greco_b200 import api
import ctypes as C
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
oc = [(np.zeros((0, el), np.uint64), np.zeros(el, np.uint64)), (point, value)]
lens = np.array([0, len(point)], dtype=np.uint64)
pts = np.ascontiguousarray(point.reshape(-1)); vals = np.ascontiguousarray(np.concatenate([oc[0][1], value]))
for it in range(6):
    t0 = time.perf_counter()
    tr = api.Keccak256Transcript(); tr.squeeze_challenges(prover.ct0is_log2_size)
    t1 = time.perf_counter()
    rc = api.lib().hg_gkr_prove(prover.circuit.h, 2, api._p(lens), api._p(pts), api._p(vals), tr.h, 0)
    t2 = time.perf_counter()
    assert rc == 0
    claims = prover.circuit.prove_gkr(oc, api.Keccak256Transcript.from_proof(b"") if False else tr2) if False else None
    t3 = time.perf_counter()
    print("transcript+squeeze %.2f ms, C prove %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), prover.circuit.timing())
t0 = time.perf_counter()
tr = api.Keccak256Transcript(); tr.squeeze_challenges(prover.ct0is_log2_size)
claims = prover.circuit.prove_gkr(oc, tr)
print("python prove_gkr total %.2f ms" % ((time.perf_counter() - t0) * 1e3))
