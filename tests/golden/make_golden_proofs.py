"""Golden PROOF BYTES for the reference's own witnesses, produced by the CPU oracle (oracle/) with the default assumption
switches. They anchor both implementations to a committed artifact: tests check oracle == golden on the CPU and
device == golden on the GPU, and a maintainer with a Rust toolchain can diff `BfvEncrypt::prove` output against the same files
(until that is done parity with the Rust prover's bytes stays unpinned).

    python tests/golden/make_golden_proofs.py      # needs /root/reference only through the committed *.npz fixtures

Files: proof_<field>_<what>_<params>.bin (raw proof bytes) and golden_proofs.json (sha256, lengths, squeezed-challenge counts).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hyper_greco_b200  # noqa: E402,F401
from hyper_greco_b200 import params, witness  # noqa: E402
from oracle import hgo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
hgo.build()
name = "1024_1x27_65537"
P = params.PARAMS[name]
bounds, segs, nv = witness.lasso_lookup_bounds(P), witness.lasso_lookup_segments(P), witness.lasso_num_vars(P)
opp = hgo.Preprocessing(bounds)
rows = np.concatenate([np.full(l, opp.lookup_index(b), np.int32) for b, l in segs])
meta = {"params": name, "assumptions": {"A3_wire": 0, "A3_h1": 0, "A5_ascending": 1}, "files": {}}


def put(fname, proof, extra):
    open(os.path.join(OUT, fname), "wb").write(proof)
    meta["files"][fname] = dict(sha256=hashlib.sha256(proof).hexdigest(), bytes=len(proof), **extra)
    print(fname, len(proof))


ints = lambda a: [sum(int(r[j]) << (64 * j) for j in range(4)) for r in a.reshape(-1, 4)]
# Lasso node alone (LassoNode::prove_claim_reduction on a fresh transcript)
inp = np.load(os.path.join(OUT, f"lasso_inputs_{name}.npz"))["inputs"]
proof, r, s, nsq = hgo.lasso_prove(0, opp, nv, rows, inp)
put(f"proof_goldilocks_lasso_node_{name}.bin", proof, dict(base_squeezes=int(nsq)))
inp_bn = np.load(os.path.join(OUT, f"lasso_inputs_bn254_{name}.npz"))["inputs"]
proof, r, s, nsq = hgo.lasso_prove(1, opp, nv, rows, inp_bn)
put(f"proof_bn254_lasso_node_{name}.bin", proof, dict(base_squeezes=int(nsq)))
# BfvEncrypt::prove (whole circuit)
io = np.load(os.path.join(OUT, f"circuit_io_{name}.npz"))
ins = dict(s=list(map(int, io["s"])), e=list(map(int, io["e"])), k1=list(map(int, io["k1"])), ais=[list(map(int, a)) for a in io["ais"]],
           r1is=[list(map(int, a)) for a in io["r1is"]], r2is=list(map(int, io["r2is"])))
put(f"proof_goldilocks_bfv_encrypt_{name}.bin", hgo.bfv_prove(0, P, ins, list(map(int, io["ct0is"]))), {})
io = np.load(os.path.join(OUT, f"circuit_io_bn254_{name}.npz"))
ins = dict(s=ints(io["s"]), e=ints(io["e"]), k1=ints(io["k1"]), ais=[ints(a) for a in io["ais"]], r1is=[ints(a) for a in io["r1is"]], r2is=ints(io["r2is"]))
put(f"proof_bn254_bfv_encrypt_{name}.bin", hgo.bfv_prove(1, P, ins, ints(io["ct0is"])), {})
json.dump(meta, open(os.path.join(OUT, "golden_proofs.json"), "w"), indent=1)
