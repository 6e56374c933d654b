"""Developer check on a GPU box: parity of the CUDA Lasso node against the oracle on the golden fixtures."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hyper_greco_b200 as hg
from hyper_greco_b200 import api, params, witness
from oracle import hgo

def run(name, mode, opts=()):
    P = params.PARAMS[name]
    inp = np.load(os.path.join(ROOT, "tests/golden", f"lasso_inputs_{name}.npz"))["inputs"]
    bounds = witness.lasso_lookup_bounds(P); segs = witness.lasso_lookup_segments(P); nv = witness.lasso_num_vars(P)
    opp = hgo.Preprocessing(bounds)
    rows = np.concatenate([np.full(l, opp.lookup_index(b), np.int32) for b, l in segs])
    for w, v in opts: hgo.set_assumption(w, v)
    t0 = time.time(); oproof, orr, osum, nsq = hgo.lasso_prove(0, opp, nv, rows, inp); t1 = time.time()
    ctx = api.Context(0)
    for w, v in opts: ctx.set_option(w, v)
    pp = api.LassoPreprocessing(bounds)
    assert pp.memory_names() == opp.memory_names(), (pp.memory_names(), opp.memory_names())
    node = api.LassoNode(ctx, pp, nv, segs)
    tr = api.Keccak256Transcript()
    t2 = time.time(); pt, val = node.prove_claim_reduction(inp, tr, mode); t3 = time.time()
    proof = tr.into_proof()
    dims, rd, fc, e = node.download_polys()
    od, ord_, ofc, oe = hgo.lasso_polynomialize(0, opp, nv, rows, inp)
    print(name, "mode", mode, "opts", opts, "oracle %.3fs gpu %.3fs" % (t1 - t0, t3 - t2), "len", len(proof), len(oproof))
    print("  dims", (dims == od).all(), "E", (e[:, :, 0] == oe[:, :, 0]).all())
    chunks = sorted(set(opp.memory_to_dimension_index))
    for s, d in enumerate(chunks):
        print("  chunk", d, "read_cts", (rd[s] == ord_[d]).all(), "final", (fc[s] == ofc[d]).all())
    print("  claimed_sum", (val == osum).all(), "point", (pt.reshape(-1) == orr).all(), "squeezed", tr.num_squeezed, nsq)
    if proof != oproof:
        n = min(len(proof), len(oproof)); a = np.frombuffer(proof[:n], np.uint8); b = np.frombuffer(oproof[:n], np.uint8)
        bad = np.nonzero(a != b)[0]
        print("  PROOF MISMATCH first byte", bad[0] if len(bad) else n, "element", (bad[0] // 16) if len(bad) else -1)
    else:
        print("  PROOF BYTES EQUAL")
    hgo.lasso_verify(0, opp, nv, proof)
    print("  oracle verifier accepts GPU proof")
    for w, v in opts: hgo.set_assumption(w, {3: 0, 31: 0, 5: 1}[w])
    node.free(); ctx.close()

if __name__ == "__main__":
    run("1024_1x27_65537", api.MODE_PREFETCH)
    run("1024_1x27_65537", api.MODE_INTERACTIVE)
    run("4096_2x55_65537", api.MODE_PREFETCH)
    run("1024_1x27_65537", api.MODE_PREFETCH, ((3, 1), (31, 1)))
    run("1024_1x27_65537", api.MODE_PREFETCH, ((5, 0),))
