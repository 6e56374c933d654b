"""CPU tests: the oracle against every known answer the reference offers for this path:
   * keccak256("") and the challenge chain of transcript.rs:149-154,199-203 (Appendix E)
   * the two subtable MLE identities of lasso/src/table/range.rs:293-331
   * preprocessing shapes / memory order (Appendix C)
   * the reference's own integration test: setup -> prove -> verify on its JSON witnesses (bfv-gkr/src/test.rs:1-48),
     restricted to the Lasso node, plus rejection of tampered proofs."""
import json
import os
import random

import numpy as np
import pytest

from conftest import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GL_P = 2**64 - 2**32 + 1
BN_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_keccak_and_chain_known_answers(oracle, golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "transcript_kat.json")))
    assert oracle.keccak256(b"").hex() == kat["keccak256_empty"]
    h = b""
    for want in kat["chain_hashes"]:
        h = oracle.keccak256(h)
        assert h.hex() == want
    assert [str(x) for x in oracle.challenges(oracle.GOLDILOCKS, 4)] == kat["goldilocks_chain"]
    assert [str(x) for x in oracle.challenges(oracle.BN254, 4)] == kat["bn254_chain"]
    # fe_mod_from_le_bytes: 256-bit little-endian integer mod p
    assert oracle.challenges(0, 1)[0] == int.from_bytes(bytes.fromhex(kat["keccak256_empty"]), "little") % GL_P


def _ext_mul_py(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % GL_P, (a[0] * b[1] + a[1] * b[0]) % GL_P)


def test_field_arithmetic_matches_python_ints(oracle):
    rnd = random.Random(1)
    edge = [0, 1, GL_P - 1, GL_P - 2, 2**32, 2**32 - 1, 2**63, 0xFFFFFFFF00000000]
    for _ in range(300):
        a = (rnd.choice(edge + [rnd.randrange(GL_P)]), rnd.choice(edge + [rnd.randrange(GL_P)]))
        b = (rnd.choice(edge + [rnd.randrange(GL_P)]), rnd.choice(edge + [rnd.randrange(GL_P)]))
        la, lb = np.array(a, np.uint64), np.array(b, np.uint64)
        assert tuple(int(x) for x in oracle.field_op(0, 0, la, lb)) == ((a[0] + b[0]) % GL_P, (a[1] + b[1]) % GL_P)
        assert tuple(int(x) for x in oracle.field_op(0, 1, la, lb)) == ((a[0] - b[0]) % GL_P, (a[1] - b[1]) % GL_P)
        assert tuple(int(x) for x in oracle.field_op(0, 2, la, lb)) == _ext_mul_py(a, b)
        if a != (0, 0):
            inv = tuple(int(x) for x in oracle.field_op(0, 3, la, lb))
            assert _ext_mul_py(a, inv) == (1, 0)
    for _ in range(100):
        a, b = rnd.randrange(BN_R), rnd.randrange(BN_R)
        la, lb = oracle.ints_to_limbs([a], 1), oracle.ints_to_limbs([b], 1)
        assert oracle.limbs_to_ints(oracle.field_op(1, 0, la, lb), 1)[0] == (a + b) % BN_R
        assert oracle.limbs_to_ints(oracle.field_op(1, 1, la, lb), 1)[0] == (a - b) % BN_R
        assert oracle.limbs_to_ints(oracle.field_op(1, 2, la, lb), 1)[0] == (a * b) % BN_R
        if a:
            assert oracle.limbs_to_ints(oracle.field_op(1, 3, la, lb), 1)[0] == pow(a, -1, BN_R)


@pytest.mark.parametrize("full,bound", [(True, 0), (False, (1 << 55) + 55), (False, 3), (False, 39), (False, 65537), (False, 2493),
                                        (False, 82638181), (False, 477501974462976257)])
def test_subtable_mle_identities(oracle, full, bound):
    """range.rs:293-309 (full_subtable_mle_eval_correct) and :311-331 (bound_subtable_mle_eval_correct): the dense
    table's MLE at a random point equals the closed-form evaluate_mle; here also for the bounds the circuits use."""
    rnd = random.Random(bound)
    for field in (0, 1):
        el = oracle.LIMBS[field] * oracle.DEGREE[field]
        mod = GL_P if field == 0 else BN_R
        pt_ints = [rnd.randrange(mod) for _ in range(16 * oracle.DEGREE[field])]
        pt = oracle.ints_to_limbs(pt_ints, field)
        tab, mle = oracle.subtable(field, full, bound, 16, pt)
        dense = oracle.mle_eval(field, tab, 16, pt)
        assert (dense == mle).all()
        # materialize: range.rs:15-17, :58-72
        t = oracle.limbs_to_ints(tab, field)
        cutoff = 65536 if full else (1 << ((bound.bit_length() - 1) % 16)) + bound % 65536
        assert t[:8] == [min(i, i if i < cutoff else 0) for i in range(8)]
        assert all(t[i] == (i if i < cutoff else 0) for i in (cutoff - 1, min(cutoff, 65535), 65535))
        assert el in (2, 4)


def test_preprocessing_matches_survey_appendix_c(oracle):
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params, witness
    expect = {1024: (5, 6, 6, 2, 14), 2048: (5, 6, 8, 4, 15), 4096: (6, 7, 9, 4, 16), 8192: (10, 11, 13, 4, 18), 16384: (19, 20, 22, 4, 19),
              32768: (22, 23, 25, 4, 21)}
    for n, (nl, ns, nm, nc, nv) in expect.items():
        P = params.by_n(n)
        pp = oracle.Preprocessing(witness.lasso_lookup_bounds(P))
        assert (pp.num_lookups, pp.num_subtables, pp.num_memories) == (nl, ns, nm)
        assert len(set(pp.memory_to_dimension_index)) == nc
        assert witness.lasso_num_vars(P) == nv
    pp = oracle.Preprocessing(witness.lasso_lookup_bounds(params.by_n(1024)))
    assert pp.memory_names() == ["bound_2493@0", "bound_3@0", "bound_39@0", "full@0", "bound_65537@1", "bound_82638181@1"]
    pp = oracle.Preprocessing(witness.lasso_lookup_bounds(params.by_n(4096)))
    assert pp.memory_names() == ["full@0", "full@1", "full@2", "bound_27424203952895201@3", "bound_3@0", "bound_39@0", "bound_39007@0",
                                 "bound_51933@0", "bound_65537@1"]
    pp = oracle.Preprocessing(witness.lasso_lookup_bounds(params.by_n(32768)))
    names = pp.memory_names()
    assert names[:10] == ["bound_3@0", "bound_34899@0", "bound_37227@0", "bound_39@0", "bound_40683@0", "bound_42261@0", "bound_46675@0",
                          "full@0", "full@1", "full@2"]
    assert names[10] == "bound_477501974462976257@3" and names[16] == "bound_65537@1" and names[-1] == "bound_94357@1"
    # chunk-bit examples of Appendix C
    cb = {b: pp.chunk_bits[pp.lookup_index(b)] for b in pp.lookup_bounds}
    assert cb[3] == [2] and cb[39] == [6] and cb[65537] == [16, 1] and cb[477501974462976257] == [16, 16, 16, 15]


@pytest.mark.parametrize("name", ["1024_1x27_65537", "2048_1x52_65537", "4096_2x55_65537"])
def test_reference_fixture_prove_verify_roundtrip(oracle, golden_dir, name):
    """generate_sk_enc_test! (bfv-gkr/src/test.rs:1-48) restricted to the Lasso node: prove, verify, no error; the returned
    claim is the MLE of the node's input at the squeezed point (sk_encryption_circuit.rs:512-516)."""
    P, inp, bounds, segs, nv, opp, rows = load_case(name, oracle, golden_dir)
    proof, r, s, nsq = oracle.lasso_prove(0, opp, nv, rows, inp)
    r2, s2, used = oracle.lasso_verify(0, opp, nv, proof)
    assert used == len(proof) and (r == r2).all() and (s == s2).all()
    padded = np.zeros(1 << nv, np.uint64)
    padded[: inp.size] = inp
    assert (oracle.mle_eval(0, padded, nv, r) == s).all()
    # deterministic
    assert oracle.lasso_prove(0, opp, nv, rows, inp)[0] == proof
    # Appendix D length: 1 + 2v (collation) + 2 GPs + openings, in Ext2 elements of 16 bytes
    m, lm = opp.num_memories, 16
    gp = lambda k: 2 * m + sum(4 * m + 3 * j for j in range(k))
    nchunks = len(set(opp.memory_to_dimension_index))
    assert len(proof) == 16 * (1 + 2 * nv + gp(nv) + gp(lm) + 3 * nchunks + m)


def test_tampered_proofs_are_rejected(oracle, golden_dir):
    P, inp, bounds, segs, nv, opp, rows = load_case("1024_1x27_65537", oracle, golden_dir)
    proof, *_ = oracle.lasso_prove(0, opp, nv, rows, inp)
    m = opp.num_memories
    # positions the reference verifier actually checks: grand-product roots / layer-0 evals, final openings
    coll = 16 * (1 + 2 * nv)
    for pos in (coll + 3, coll + 16 * 2 * m + 5, len(proof) - 1, len(proof) - 16 * 3):
        bad = bytearray(proof)
        bad[pos] ^= 1
        with pytest.raises(oracle.OracleError):
            oracle.lasso_verify(0, opp, nv, bytes(bad))
    with pytest.raises(oracle.OracleError):
        oracle.lasso_verify(0, opp, nv, proof[:-8])
    # a non-canonical field element is an error (transcript.rs:168)
    bad = bytearray(proof)
    bad[0:8] = b"\xff" * 8
    with pytest.raises(oracle.OracleError):
        oracle.lasso_verify(0, opp, nv, bytes(bad))


def test_out_of_range_witness_fails_memory_check(oracle, golden_dir):
    """A value outside its range changes combine(E) != input; the proof still verifies structurally (the reference's
    verifier does not tie claimed_sum to the input here) but the returned claim no longer matches the input MLE, which is
    what sk_encryption_circuit.rs:512-516 would catch."""
    P, inp, bounds, segs, nv, opp, rows = load_case("1024_1x27_65537", oracle, golden_dir)
    bad = inp.copy()
    bad[-1] = 200000  # k1 + K1_BOUND must be < 65537; 18 bits get truncated to sum(chunk_bits) = 17 (Q8)
    proof, r, s, _ = oracle.lasso_prove(0, opp, nv, rows, bad)
    oracle.lasso_verify(0, opp, nv, proof)
    padded = np.zeros(1 << nv, np.uint64)
    padded[: bad.size] = bad
    assert not (oracle.mle_eval(0, padded, nv, r) == s).all()


@pytest.mark.parametrize("opts", [((3, 1),), ((31, 1),), ((5, 0),), ((3, 1), (31, 1), (5, 0))])
def test_assumption_switches_roundtrip(oracle, golden_dir, opts):
    P, inp, bounds, segs, nv, opp, rows = load_case("1024_1x27_65537", oracle, golden_dir)
    base, *_ = oracle.lasso_prove(0, opp, nv, rows, inp)
    try:
        for w, v in opts:
            oracle.set_assumption(w, v)
        proof, *_ = oracle.lasso_prove(0, opp, nv, rows, inp)
        oracle.lasso_verify(0, opp, nv, proof)
        assert proof != base
    finally:
        for w, v in ((3, 0), (31, 0), (5, 1)):
            oracle.set_assumption(w, v)


def test_sumcheck_true_evals_sum_rule(oracle):
    """The traced TRUE round polynomial satisfies h_j(0)+h_j(1) = h_{j-1}(r_{j-1}) (a property of any sumcheck), while the
    claim itself is NOT the hypercube sum (F4) -- documents why A3' matters."""
    rnd = np.random.default_rng(3)
    nv, nterms = 5, 3
    tables = rnd.integers(0, GL_P, size=(2 * nterms, 1 << nv), dtype=np.uint64)
    coeffs = rnd.integers(0, GL_P, size=(nterms, 2), dtype=np.uint64)
    claim = np.array([5, 7], np.uint64)
    proof, te, r, fe = oracle.sumcheck_prove(0, 2, coeffs, tables, nv, claim)
    add = lambda a, b: oracle.field_op(0, 0, a, b)
    mul = lambda a, b: oracle.field_op(0, 2, a, b)
    sub = lambda a, b: oracle.field_op(0, 1, a, b)
    for j in range(1, nv):
        y = [te[j - 1][k] for k in range(4)]
        # Lagrange at r over points 0..3
        x = r[j - 1]
        acc = np.zeros(2, np.uint64)
        for i in range(4):
            num, den = np.array([1, 0], np.uint64), 1
            for k in range(4):
                if k != i:
                    num = mul(num, sub(x, np.array([k, 0], np.uint64)))
                    den = den * (i - k)
            inv = np.array([pow(den % GL_P, -1, GL_P), 0], np.uint64)
            acc = add(acc, mul(mul(num, inv), y[i]))
        assert (add(te[j][0], te[j][1]) == acc).all()
    # final evals are the MLEs at r
    for t in range(2 * nterms):
        assert (oracle.mle_eval(0, tables[t], nv, r) == fe[t]).all()


def test_synthetic_witness_follows_reference_shape():
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params, witness
    for n in (1024, 4096):
        P = params.by_n(n)
        a = witness.synth_witness(P, seed=7)
        assert witness.check_circuit_identity(P, a)
        assert len(a.s) == n and len(a.e) == n and len(a.k1) == n
        assert all(len(v) == n - 1 for v in a.r2is) and all(len(v) == 2 * n - 1 for v in a.r1is)
        inp = witness.lasso_inputs(P, a)
        segs = witness.lasso_lookup_segments(P)
        assert len(inp) == sum(l for _, l in segs)
        # every shifted value is inside its range (the reason the lookups succeed)
        pos = 0
        for (b, l), shift_b in zip(segs, [b for b, _ in segs]):
            seg = inp[pos:pos + l]
            pos += l
            assert max(seg) < b or b != shift_b


def test_forward_evaluation_matches_reference_fixtures(oracle, golden_dir):
    """The reference's witnesses satisfy ct0i = s*ai + e + k1*k0i + r1i*qi + r2i*(x^n+1) under the get_inputs layout; the
    oracle's circuit evaluation (NTT -> dot product -> INTT, relay/scale/sum layers) must reproduce ct0is exactly."""
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params
    for name in ("1024_1x27_65537", "4096_2x55_65537"):
        P = params.PARAMS[name]
        io = np.load(os.path.join(golden_dir, f"circuit_io_{name}.npz"))
        ins = {k: [int(v) for v in io[k]] for k in ("s", "e", "k1", "r2is")}
        ins["ais"] = [[int(v) for v in row] for row in io["ais"]]
        ins["r1is"] = [[int(v) for v in row] for row in io["r1is"]]
        lasso, summ = oracle.bfv_eval(0, P, ins)
        assert (summ == io["ct0is"]).all()
        assert (lasso == np.load(os.path.join(golden_dir, f"lasso_inputs_{name}.npz"))["inputs"]).all()


def test_ntt_is_the_dft_over_the_2_adic_root(oracle):
    """NTT KAT from the field definition: root = 7^((p-1)/2^32) (goldilocks crate ROOT_OF_UNITY), out[k] = sum_j in[j] w^(jk)."""
    log_n, n = 4, 16
    w = pow(pow(7, (GL_P - 1) >> 32, GL_P), 1 << (32 - log_n), GL_P)
    rnd = random.Random(4)
    x = [rnd.randrange(GL_P) for _ in range(n)]
    want = [sum(x[j] * pow(w, j * k, GL_P) for j in range(n)) % GL_P for k in range(n)]
    got = oracle.ntt(0, np.array([x], dtype=np.uint64), log_n)[0]
    assert [int(v) for v in got] == want
    back = oracle.ntt(0, got.reshape(1, -1), log_n, True)[0]
    assert [int(v) for v in back] == x


@pytest.mark.parametrize("name", ["1024_1x27_65537", "2048_1x52_65537"])
def test_bn254_reference_fixture_prove_verify_roundtrip(oracle, golden_dir, name):
    """test_sk_enc_valid_bn254_* (sk_encryption_circuit.rs:616-620) restricted to the Lasso node: F = E = bn256::Fr."""
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params, witness
    P = params.PARAMS[name]
    inp = np.load(os.path.join(golden_dir, f"lasso_inputs_bn254_{name}.npz"))["inputs"]
    opp = oracle.Preprocessing(witness.lasso_lookup_bounds(P))
    rows = np.concatenate([np.full(l, opp.lookup_index(b), np.int32) for b, l in witness.lasso_lookup_segments(P)])
    nv = witness.lasso_num_vars(P)
    proof, r, s, nsq = oracle.lasso_prove(1, opp, nv, rows, inp)
    r2, s2, used = oracle.lasso_verify(1, opp, nv, proof)
    assert used == len(proof) and (r == r2).all() and (s == s2).all()
    padded = np.zeros((1 << nv, 4), np.uint64)
    padded[: inp.shape[0]] = inp
    assert (oracle.mle_eval(1, padded, nv, r) == s).all()
    bad = bytearray(proof)
    bad[-1] ^= 1
    with pytest.raises(oracle.OracleError):
        oracle.lasso_verify(1, opp, nv, bytes(bad))


def _golden_proof(golden_dir, fname):
    import hashlib
    meta = json.load(open(os.path.join(golden_dir, "golden_proofs.json")))
    data = open(os.path.join(golden_dir, fname), "rb").read()
    assert hashlib.sha256(data).hexdigest() == meta["files"][fname]["sha256"] and len(data) == meta["files"][fname]["bytes"]
    return data, meta["files"][fname]


def test_oracle_reproduces_its_committed_golden_proofs(oracle, golden_dir):
    """The committed proof bytes (tests/golden/make_golden_proofs.py) are what the oracle produces today: any change of the
    restatement shows up here, and the same files are what the GPU tests and an off-box Rust run compare against."""
    name = "1024_1x27_65537"
    P, inp, bounds, segs, nv, opp, rows = load_case(name, oracle, golden_dir)
    want, info = _golden_proof(golden_dir, f"proof_goldilocks_lasso_node_{name}.bin")
    proof, r, s, nsq = oracle.lasso_prove(0, opp, nv, rows, inp)
    assert proof == want and nsq == info["base_squeezes"]
    oracle.lasso_verify(0, opp, nv, want)
    inp_bn = np.load(os.path.join(golden_dir, f"lasso_inputs_bn254_{name}.npz"))["inputs"]
    want, info = _golden_proof(golden_dir, f"proof_bn254_lasso_node_{name}.bin")
    proof, r, s, nsq = oracle.lasso_prove(1, opp, nv, rows, inp_bn)
    assert proof == want and nsq == info["base_squeezes"]
    io = np.load(os.path.join(golden_dir, f"circuit_io_{name}.npz"))
    ins = dict(s=[int(v) for v in io["s"]], e=[int(v) for v in io["e"]], k1=[int(v) for v in io["k1"]], ais=[[int(v) for v in a] for a in io["ais"]],
               r1is=[[int(v) for v in a] for a in io["r1is"]], r2is=[int(v) for v in io["r2is"]])
    want, _ = _golden_proof(golden_dir, f"proof_goldilocks_bfv_encrypt_{name}.bin")
    assert oracle.bfv_prove(0, P, ins, [int(v) for v in io["ct0is"]]) == want
    oracle.bfv_verify(0, P, ins, [int(v) for v in io["ct0is"]], want)


# ------------------------------------------------------------------------------------------------ interchange dumps
def _dump_tools():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("hg_dump", os.path.join(ROOT, "scripts", "hg_dump.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_interchange_dumps_are_consistent_with_the_golden_proofs(oracle, golden_dir):
    """tests/golden/dumps/*.hgdump (ORACLE-GENERATED, one per A3 / A3' / A5 setting): the written elements of the default-setting
    dump are the committed golden proof, the squeezed elements are the Appendix-E challenge chain, and the oracle reproduces the
    file today."""
    import hashlib
    import os
    hd = _dump_tools()
    name = "1024_1x27_65537"
    default = hd.SETTINGS[0]
    assert default == dict(A3_wire=0, A3_h1=0, A5_ascending=1)
    for field, tag in (("goldilocks", "goldilocks"), ("bn254", "bn254")):
        header, events = hd.read_dump(os.path.join(hd.DUMP_DIR, f"bfv_encrypt_{field}_{name}_{hd.setting_tag(default)}.hgdump"))
        proof = open(os.path.join(golden_dir, f"proof_{tag}_bfv_encrypt_{name}.bin"), "rb").read()
        assert hd.proof_of(events) == proof and header["proof_sha256"] == hashlib.sha256(proof).hexdigest()
        sq = [v for k, v in events if k == "S"]
        fid = 0 if field == "goldilocks" else 1
        want = oracle.challenges(fid, len(sq))
        assert [int.from_bytes(v, "big") for v in sq] == [int(x) for x in want]
    h, ev = hd.oracle_dump("goldilocks", name, hd.SETTINGS[3])
    _, committed = hd.read_dump(os.path.join(hd.DUMP_DIR, f"bfv_encrypt_goldilocks_{name}_{hd.setting_tag(hd.SETTINGS[3])}.hgdump"))
    assert [(ev[i:i + 1].decode(), ev[i + 1:i + 9]) for i in range(0, len(ev), 9)] == committed


def test_compare_dump_names_the_matching_setting(golden_dir, tmp_path):
    """scripts/compare_dump.py: every committed dump is recognised as its own setting; a dump with one flipped written byte matches
    none and the report points at the element."""
    import os
    import subprocess
    import sys
    hd = _dump_tools()
    name = "1024_1x27_65537"
    tool = os.path.join(ROOT, "scripts", "compare_dump.py")
    for s in (hd.SETTINGS[0], hd.SETTINGS[1], hd.SETTINGS[5]):
        path = os.path.join(hd.DUMP_DIR, f"bfv_encrypt_goldilocks_{name}_{hd.setting_tag(s)}.hgdump")
        r = subprocess.run([sys.executable, tool, path], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0 and ("MATCH" in r.stdout) and hd.setting_tag(s) in r.stdout.splitlines()[-1], r.stdout
    header, events = hd.read_dump(os.path.join(hd.DUMP_DIR, f"bfv_encrypt_goldilocks_{name}_{hd.setting_tag(hd.SETTINGS[0])}.hgdump"))
    wi = [i for i, (k, _) in enumerate(events) if k == "W"][100]
    events[wi] = ("W", bytes([events[wi][1][0] ^ 1]) + events[wi][1][1:])
    bad = tmp_path / "bad.hgdump"
    hd.write_dump(str(bad), {k: v for k, v in header.items() if k not in ("n_events",)}, b"".join(k.encode() + v for k, v in events))
    r = subprocess.run([sys.executable, tool, str(bad)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "NO MATCH" in r.stdout and "#100" in r.stdout, r.stdout
