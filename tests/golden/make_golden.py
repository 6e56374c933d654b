"""Generates tests/golden/*.npz from the reference's own witness fixtures. Run in the build container only
(/root/reference does not exist on the GPU box); the outputs are committed.

    python tests/golden/make_golden.py

For each Goldilocks fixture /root/reference/bfv-gkr/src/data/goldilocks/sk_enc_<n>_<k>x<bits>_65537.json it stores the
Lasso node's input vector (the `lasso_inputs_batched` layer of sk_encryption_circuit.rs:163-181 applied to the parsed
witness, sk_encryption_circuit.rs:365-415) as uint64, plus the result of the circuit identity check.
Also stores the known-answer values of SURVEY.md Appendix E (keccak256("") and the first challenge-chain values).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hyper_greco_b200  # noqa: E402,F401
from hyper_greco_b200 import params, witness  # noqa: E402

REF = "/root/reference/bfv-gkr/src/data/goldilocks"
OUT = os.path.dirname(os.path.abspath(__file__))

for name in ["1024_1x27_65537", "2048_1x52_65537", "4096_2x55_65537"]:
    P = params.PARAMS[name]
    args = witness.load_args_json(os.path.join(REF, f"sk_enc_{name}.json"))
    assert witness.check_circuit_identity(P, args), name
    inp = np.array(witness.lasso_inputs(P, args), dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, f"lasso_inputs_{name}.npz"), inputs=inp)
    print(name, inp.size, "rows")
    if name in ("1024_1x27_65537", "4096_2x55_65537"):
        # the circuit's input vectors (get_inputs, sk_encryption_circuit.rs:365-415) and its expected output ct0is
        ins, ct0is = witness.get_inputs(P, args)
        u = lambda v: np.array(v, dtype=np.uint64)
        np.savez_compressed(os.path.join(OUT, f"circuit_io_{name}.npz"), s=u(ins["s"]), e=u(ins["e"]), k1=u(ins["k1"]), ais=u(ins["ais"]),
                            r1is=u(ins["r1is"]), r2is=u(ins["r2is"]), ct0is=u(ct0is))

# BN254 witnesses of the reference (bfv-gkr/src/data/bn254): Lasso inputs as canonical 4 x u64 limbs
REF_BN = "/root/reference/bfv-gkr/src/data/bn254"
for name in ["1024_1x27_65537", "2048_1x52_65537"]:
    P = params.PARAMS[name]
    args = witness.load_args_json(os.path.join(REF_BN, f"sk_enc_{name}.json"))
    assert witness.check_circuit_identity(P, args, p=witness.BN_R), name
    vals = witness.lasso_inputs(P, args, p=witness.BN_R)
    limbs = np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in vals], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, f"lasso_inputs_bn254_{name}.npz"), inputs=limbs)
    print("bn254", name, limbs.shape)
    if name == "1024_1x27_65537":
        # circuit inputs / output of the BN254 witness (get_inputs), canonical 4 x u64 limbs per element
        ins, ct0is = witness.get_inputs(P, args)
        lm = lambda v: np.array([[(int(x) >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for x in v], dtype=np.uint64)
        np.savez_compressed(os.path.join(OUT, f"circuit_io_bn254_{name}.npz"), s=lm(ins["s"]), e=lm(ins["e"]), k1=lm(ins["k1"]),
                            ais=np.stack([lm(a) for a in ins["ais"]]), r1is=np.stack([lm(a) for a in ins["r1is"]]), r2is=lm(ins["r2is"]), ct0is=lm(ct0is))

kat = {
    "keccak256_empty": "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
    "chain_hashes": [
        "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
        "10ca3eff73ebec87d2394fc58560afeab86dac7a21f5e402ea0a55e5c8a6758f",
        "1cf8eebf67df4cc8de3bc92242c7a5691a7cdd7efe364b62c1b97063ed450b75",
        "5608ca83f9a41a423fa54d5a12c1dd1212e5c699b157bfbddbcb16ee12c07dce",
    ],
    "goldilocks_chain": ["15017384644633299356", "6854594310142832579", "9149254073876997563", "1396060396769822097"],
    "bn254_chain": [
        "7173236656320612194178997223602979818891828541827642103715116037219761443523",
        "21112123816342014025406352012828000244932007375891133415489782355590148704782",
        "9164035478753757386635631110257021653244911585343202658577009786749748967450",
        "5845656849544400234018505166145211271474548605642901147529119308656667134034",
    ],
    "source": "SURVEY.md Appendix E (derived from bfv-gkr/src/transcript.rs:149-154,199-203; keccak256('') is the well-known digest)",
}
json.dump(kat, open(os.path.join(OUT, "transcript_kat.json"), "w"), indent=1)
