import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hyper_greco_b200
from hyper_greco_b200 import api, params, witness
from oracle import hgo
P = params.by_n(32768)
args = witness.synth_witness(P, 0, p=witness.BN_R)
vals = witness.lasso_inputs(P, args, p=witness.BN_R)
inp = np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in vals], dtype=np.uint64)
bounds, segs, nv = witness.lasso_lookup_bounds(P), witness.lasso_lookup_segments(P), witness.lasso_num_vars(P)
ctx = api.Context(0, api.BN254)
pp = api.LassoPreprocessing.preprocess(bounds)
node = api.LassoNode(ctx, pp, nv, segs)
print("device bytes %.1f GB" % (node.device_bytes / 1e9))
buf = api.DeviceBuffer.from_field(ctx, inp)
for it in range(4):
    tr = api.Keccak256Transcript(api.BN254)
    t0 = time.perf_counter()
    node.prove_claim_reduction(buf, tr, 0, n_inputs=inp.shape[0])
    dt = time.perf_counter() - t0
    print("bn254 lasso node n=32768 k=16: %.2f ms" % (dt * 1e3), node.timing())
proof = tr.into_proof()
opp = hgo.Preprocessing(bounds)
t0 = time.perf_counter()
hgo.lasso_verify(1, opp, nv, proof)
print("oracle verifier accepts the BN254 GPU proof (%d bytes, %.1fs)" % (len(proof), time.perf_counter() - t0))
