"""BN254 Fr (E = F), n=32768 k=16: the whole BfvEncrypt::prove on the device, from host vectors; oracle verifier on the result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hyper_greco_b200  # noqa
from hyper_greco_b200 import api, params, witness
from oracle import hgo
P = params.by_n(32768)
args = witness.synth_witness(P, 0, p=witness.BN_R)
ins, ct0is = witness.get_inputs(P, args)
lm = lambda v: np.array([[(int(x) >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for x in v], dtype=np.uint64).reshape(-1)
flat = [lm(ins["s"]), lm(ins["e"]), lm(ins["k1"])] + [lm(a) for a in ins["ais"]] + [lm(a) for a in ins["r1is"]] + [lm(ins["r2is"])]
ct = lm(ct0is)
ctx = api.Context(0, api.BN254)
prover = api.BfvSkEncryptProver(ctx, P)
for it in range(5):
    t0 = time.perf_counter()
    proof, claims = prover.prove_host(flat, ct)
    dt = time.perf_counter() - t0
    print("bn254 BfvEncrypt::prove n=32768 k=16 from host vectors: %.2f ms, %d proof bytes" % (dt * 1e3, len(proof)), prover.circuit.timing())
ctx.profile(True)
prover.prove_host(flat, ct)
prof = ctx.profile_read(); ctx.profile(False)
print({k: round(v[1], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})
hgo.build()
hgo.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
hgo.bfv_verify(1, P, ins, ct0is, proof)
print("oracle verifier accepts the BN254 GPU proof (%.1fs)" % (time.perf_counter() - t0))
