"""Hashes of FULL-SIZE proofs (BASELINE.json configs 3, 4 and 5), produced once by the CPU oracle (oracle/, default assumption
switches) on the synthetic witnesses of hyper-greco_b200/witness.py. The GPU tests prove the same witnesses and compare the
sha256 and the length, so byte equality with the CPU prover at n = 16384 / 32768 is checked on every `pytest -m gpu` run
without running the oracle prover there (it needs minutes on a few host cores).

    python tests/golden/make_golden_fullsize.py        # ~10 minutes on 8 cores; writes the "fullsize" section of golden_proofs.json

ORACLE-GENERATED, like every golden proof in this directory: parity with the Rust prover's bytes stays unpinned.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hyper_greco_b200  # noqa: E402,F401
from hyper_greco_b200 import params, witness  # noqa: E402
from oracle import hgo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
hgo.build()
hgo.set_num_threads(os.cpu_count() or 1)
full = {}


def lasso_case(field, n, seed):
    P = params.by_n(n)
    p = witness.BN_R if field == 1 else None
    args = witness.synth_witness(P, seed, p=p) if p else witness.synth_witness(P, seed)
    vals = witness.lasso_inputs(P, args, p=p) if p else witness.lasso_inputs(P, args)
    if field == 1:
        inp = np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in vals], dtype=np.uint64)
    else:
        inp = np.array(vals, dtype=np.uint64)
    bounds, segs, nv = witness.lasso_lookup_bounds(P), witness.lasso_lookup_segments(P), witness.lasso_num_vars(P)
    opp = hgo.Preprocessing(bounds)
    rows = np.concatenate([np.full(l, opp.lookup_index(b), np.int32) for b, l in segs])
    t0 = time.perf_counter()
    proof, r, s, nsq = hgo.lasso_prove(field, opp, nv, rows, inp)
    return proof, time.perf_counter() - t0


def bfv_case(field, n, seed):
    P = params.by_n(n)
    p = witness.BN_R if field == 1 else None
    args = witness.synth_witness(P, seed, p=p) if p else witness.synth_witness(P, seed)
    ins, ct0is = witness.get_inputs(P, args)   # python ints, already reduced mod p by synth_witness
    t0 = time.perf_counter()
    proof = hgo.bfv_prove(field, P, ins, ct0is, cap=1 << 26)
    return proof, time.perf_counter() - t0


CASES = [("lasso_node", 0, 32768, 0), ("bfv_encrypt", 0, 32768, 5), ("bfv_encrypt", 0, 16384, 3), ("lasso_node", 1, 32768, 0), ("bfv_encrypt", 1, 32768, 5)]
only = sys.argv[1:] and [int(x) for x in sys.argv[1:]]
path = os.path.join(OUT, "golden_proofs.json")
meta = json.load(open(path))
full = meta.get("fullsize", {})
for k, (what, field, n, seed) in enumerate(CASES):
    if only and k not in only:
        continue
    fn = lasso_case if what == "lasso_node" else bfv_case
    proof, dt = fn(field, n, seed)
    key = f"{'bn254' if field else 'goldilocks'}_{what}_n{n}_seed{seed}"
    full[key] = {"sha256": hashlib.sha256(proof).hexdigest(), "bytes": len(proof), "oracle_seconds": round(dt, 1)}
    print(key, full[key], flush=True)
    meta["fullsize"] = full
    json.dump(meta, open(path, "w"), indent=1)
