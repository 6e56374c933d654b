"""Throughput with several independent proofs in flight on ONE GPU: one host thread + one hg_ctx (own streams, own buffers) each."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import hyper_greco_b200  # noqa
from hyper_greco_b200 import api

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 2
STEPS = 40
P, inp, bounds, segs, nv = bench.make_case(bench.DEFAULT_CONFIG, 0)
ins, ct0is = bench.LAST_WITNESS
flat = [ins["s"], ins["e"], ins["k1"]] + list(ins["ais"]) + list(ins["r1is"]) + [ins["r2is"]]
host_np = [np.array(v, dtype=np.uint64) for v in flat]
ct_np = np.array(ct0is, dtype=np.uint64)

class Worker:
    def __init__(self):
        self.ctx = api.Context(0)
        self.prover = api.BfvSkEncryptProver(self.ctx, P)
        n = sum(v.size for v in host_np) + ct_np.size
        self.pin = torch.empty(n, dtype=torch.int64).pin_memory()
        h = self.pin.numpy().view(np.uint64)
        self.views, off = [], 0
        for v in host_np:
            h[off:off + v.size] = v; self.views.append(h[off:off + v.size]); off += v.size
        self.h_ct = h[off:]; self.h_ct[:] = ct_np
        self.proof = self.prover.prove_host(self.views, self.h_ct)[0]
        tr0 = api.Keccak256Transcript()
        pt = tr0.squeeze_challenges(self.prover.ct0is_log2_size)
        self.d_ct = api.DeviceBuffer.from_numpy(self.ctx, ct_np)
        val = api.mle_eval_batch(self.ctx, self.d_ct, 1, self.prover.ct0is_log2_size, pt)[0]
        self.claims = [(np.zeros((0, 2), np.uint64), np.zeros(2, np.uint64)), (pt, val)]
    def resident(self):
        tr = api.Keccak256Transcript(); tr.squeeze_challenges(self.prover.ct0is_log2_size)
        self.prover.circuit.prove_gkr(self.claims, tr, api.MODE_PREFETCH)
        return tr.into_proof()
    def e2e(self):
        return self.prover.prove_host(self.views, self.h_ct)[0]

workers = [Worker() for _ in range(NT)]
assert all(w.proof == workers[0].proof for w in workers)
for kind in ("resident", "e2e"):
    def run(w, n):
        f = getattr(w, kind)
        for _ in range(n):
            pr = f()
        assert pr == w.proof
    for w in workers: run(w, 25)
    ths = [threading.Thread(target=run, args=(w, STEPS)) for w in workers]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"in flight {NT} {kind}: {NT * STEPS / dt:.1f} proofs/s, {1e3 * dt / (NT * STEPS):.3f} ms/proof (latency {1e3 * dt / STEPS:.2f} ms)")
