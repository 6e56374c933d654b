"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes driver for oracle/_build/libhg_oracle.so (the CPU restatement of the reference's Lasso node, see
oracle/protocol.hpp). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (hyper-greco_b200/) never does.

Parity status: PINNED for keccak256(""), the challenge chain, the range.rs subtable MLE
identities and verifier acceptance; UNPINNED for everything that lives in the un-vendored gkr crate (round-message
format, distribute_powers order, ...), each behind a switch in `set_assumption`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "libhg_oracle.so")

GOLDILOCKS, BN254 = 0, 1
GL_P = 2**64 - 2**32 + 1
BN_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
LIMBS = {GOLDILOCKS: 1, BN254: 4}
DEGREE = {GOLDILOCKS: 2, BN254: 1}


def build(force=False):
    """Compile the oracle (g++, OpenMP). Building the checker is not using it."""
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _DIR, "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.hgo_last_error.restype = C.c_char_p
        L.hgo_pp_new.restype = C.c_void_p
        L.hgo_pp_new.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.hgo_pp_free.argtypes = [C.c_void_p]
        L.hgo_pp_info.argtypes = [C.c_void_p, C.c_void_p]
        L.hgo_pp_maps.argtypes = [C.c_void_p] * 6
        L.hgo_pp_lookup_memories.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hgo_pp_chunk_bits.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hgo_challenges.argtypes = [C.c_int, C.c_size_t, C.c_void_p]
        L.hgo_keccak256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.hgo_lasso_prove.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                      C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hgo_lasso_verify.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hgo_lasso_polynomialize.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hgo_sumcheck_prove.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hgo_mle_eval.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hgo_subtable.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hgo_ntt.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.hgo_bfv_eval.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_uint64] * 3 + [C.c_void_p] * 9
        L.hgo_bfv_prove.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_uint64] * 3 + [C.c_void_p] * 8 + [C.c_size_t, C.c_void_p, C.c_int]
        L.hgo_bfv_session_new.restype = C.c_void_p
        L.hgo_bfv_session_new.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_uint64] * 3 + [C.c_void_p] * 7
        L.hgo_bfv_session_free.argtypes = [C.c_void_p]
        L.hgo_bfv_session_prove.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.hgo_field_op.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise OracleError(lib().hgo_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_assumption(which, value):
    """which: 3 = A3 wire (0 coeffs / 1 evals), 31 = A3' h(1) (0 from claim / 1 true), 5 = A5 ascending (1/0)."""
    lib().hgo_set_assumption(which, value)


def events_begin():
    """Interchange dump: start recording every base-field squeeze ('S') / write ('W') of the oracle's transcripts."""
    lib().hgo_events_begin()


def events_end() -> bytes:
    """Stop recording; returns the event bytes: kind byte + element in proof encoding (big-endian), per event."""
    lib().hgo_events_end.restype = C.c_size_t
    lib().hgo_events_end.argtypes = [C.c_void_p, C.c_size_t]
    cap = 1 << 26
    buf = np.zeros(cap, np.uint8)
    n = lib().hgo_events_end(_p(buf), cap)
    if n > cap:
        raise OracleError("event log larger than 64 MiB")
    return buf[:n].tobytes()


def set_num_threads(n):
    lib().hgo_set_num_threads(n)


def num_threads():
    return lib().hgo_num_threads()


def keccak256(data: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    buf = np.frombuffer(data, np.uint8) if data else np.zeros(0, np.uint8)
    lib().hgo_keccak256(_p(np.ascontiguousarray(buf)), len(data), _p(out))
    return out.tobytes()


def challenges(field, n):
    """First n base-field values of the Keccak challenge chain as python ints."""
    out = np.zeros(n * LIMBS[field], np.uint64)
    lib().hgo_challenges(field, n, _p(out))
    return limbs_to_ints(out, field)


def ints_to_limbs(vals, field):
    k = LIMBS[field]
    if k == 1:
        return np.array([int(v) for v in vals], dtype=np.uint64)
    out = np.zeros(len(vals) * k, np.uint64)
    for i, v in enumerate(vals):
        v = int(v)
        for j in range(k):
            out[i * k + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr, field):
    k = LIMBS[field]
    arr = np.asarray(arr, np.uint64).reshape(-1, k)
    return [sum(int(row[j]) << (64 * j) for j in range(k)) for row in arr]


class Preprocessing:
    """LassoPreprocessing::preprocess (lasso.rs:527-627) for RangeLookup types given by their bounds."""

    def __init__(self, bounds, C_=4, log2M=16):
        b = np.array([int(x) for x in bounds], np.uint64)
        self.h = lib().hgo_pp_new(_p(b), len(b), C_, log2M)
        self.C, self.log2M = C_, log2M
        info = np.zeros(3, np.int32)
        lib().hgo_pp_info(self.h, _p(info))
        self.num_lookups, self.num_subtables, self.num_memories = (int(x) for x in info)
        lb = np.zeros(self.num_lookups, np.uint64)
        sf = np.zeros(self.num_subtables, np.int32)
        sb = np.zeros(self.num_subtables, np.uint64)
        ms = np.zeros(self.num_memories, np.int32)
        md = np.zeros(self.num_memories, np.int32)
        lib().hgo_pp_maps(self.h, _p(lb), _p(sf), _p(sb), _p(ms), _p(md))
        self.lookup_bounds = [int(x) for x in lb]
        self.subtables = [("full", 0) if f else ("bound", int(b_)) for f, b_ in zip(sf, sb)]
        self.memory_to_subtable_index = [int(x) for x in ms]
        self.memory_to_dimension_index = [int(x) for x in md]
        self.lookup_to_memory_indices = []
        self.chunk_bits = []
        tmp = np.zeros(64, np.int32)
        for l in range(self.num_lookups):
            n = lib().hgo_pp_lookup_memories(self.h, l, _p(tmp))
            self.lookup_to_memory_indices.append([int(x) for x in tmp[:n]])
            n = lib().hgo_pp_chunk_bits(self.h, l, _p(tmp))
            self.chunk_bits.append([int(x) for x in tmp[:n]])

    def lookup_index(self, bound):
        return self.lookup_bounds.index(int(bound))

    def memory_names(self):
        out = []
        for m in range(self.num_memories):
            k, b = self.subtables[self.memory_to_subtable_index[m]]
            out.append(("full" if k == "full" else f"bound_{b}") + f"@{self.memory_to_dimension_index[m]}")
        return out

    def __del__(self):
        try:
            lib().hgo_pp_free(self.h)
        except Exception:
            pass


def lasso_prove(field, pp, num_vars, rows, inputs_limbs, skip=0):
    """LassoNode::prove_claim_reduction (lasso.rs:57-114). rows = per-row lookup index (into pp.lookup_bounds).
    Returns (proof bytes, r limbs, claimed_sum limbs, number of base squeezes performed incl. skip)."""
    rows = np.ascontiguousarray(rows, np.int32)
    inputs_limbs = np.ascontiguousarray(inputs_limbs, np.uint64)
    n_inputs = inputs_limbs.size // LIMBS[field]
    el = LIMBS[field] * DEGREE[field]
    m = pp.num_memories
    cap = (64 + 8 * m * (num_vars + pp.log2M + 4) + 4 * (num_vars + pp.log2M) ** 2) * el * 8 * 4
    proof = np.zeros(cap, np.uint8)
    ln = C.c_size_t(0)
    nsq = C.c_size_t(0)
    r = np.zeros(num_vars * el, np.uint64)
    s = np.zeros(el, np.uint64)
    _chk(lib().hgo_lasso_prove(field, pp.h, num_vars, _p(rows), rows.size, _p(inputs_limbs), n_inputs, skip, _p(proof), cap,
                               C.byref(ln), _p(r), _p(s), C.byref(nsq)))
    return proof[: ln.value].tobytes(), r, s, nsq.value


def lasso_verify(field, pp, num_vars, proof: bytes, skip=0):
    """LassoNode::verify_claim_reduction (lasso.rs:116-139). Raises OracleError when the reference would Err/panic."""
    buf = np.frombuffer(proof, np.uint8)
    el = LIMBS[field] * DEGREE[field]
    r = np.zeros(num_vars * el, np.uint64)
    s = np.zeros(el, np.uint64)
    used = C.c_size_t(0)
    _chk(lib().hgo_lasso_verify(field, pp.h, num_vars, _p(np.ascontiguousarray(buf)), buf.size, skip, _p(r), _p(s), C.byref(used)))
    return r, s, used.value


def lasso_polynomialize(field, pp, num_vars, rows, inputs_limbs):
    rows = np.ascontiguousarray(rows, np.int32)
    inputs_limbs = np.ascontiguousarray(inputs_limbs, np.uint64)
    n_inputs = inputs_limbs.size // LIMBS[field]
    R, M, m = 1 << num_vars, 1 << pp.log2M, pp.num_memories
    dims = np.zeros((pp.C, R), np.uint64)
    rd = np.zeros((m, R), np.uint64)
    fc = np.zeros((m, M), np.uint64)
    e = np.zeros((m, R, LIMBS[field]), np.uint64)
    _chk(lib().hgo_lasso_polynomialize(field, pp.h, num_vars, _p(rows), rows.size, _p(inputs_limbs), n_inputs, _p(dims), _p(rd), _p(fc), _p(e)))
    return dims, rd, fc, e


def sumcheck_prove(field, arity, coeffs_limbs, tables_limbs, num_vars, claim_limbs, skip=0):
    """prove_sum_check for g = poly(0) * sum_i coeffs[i] * prod_{k<arity} poly(arity*i+k).
    Returns (proof bytes, true round evaluations [num_vars, degree+1, el], r, final evals)."""
    el = LIMBS[field] * DEGREE[field]
    coeffs_limbs = np.ascontiguousarray(coeffs_limbs, np.uint64)
    nterms = coeffs_limbs.size // el
    tables_limbs = np.ascontiguousarray(tables_limbs, np.uint64)
    d = arity + 1
    cap = num_vars * d * el * 8 + 64
    proof = np.zeros(cap, np.uint8)
    ln = C.c_size_t(0)
    te = np.zeros((num_vars, d + 1, el), np.uint64)
    r = np.zeros((num_vars, el), np.uint64)
    fe = np.zeros((nterms * arity, el), np.uint64)
    _chk(lib().hgo_sumcheck_prove(field, arity, nterms, num_vars, _p(coeffs_limbs), _p(tables_limbs), _p(np.ascontiguousarray(claim_limbs, np.uint64)),
                                  skip, _p(proof), cap, C.byref(ln), _p(te), _p(r), _p(fe)))
    return proof[: ln.value].tobytes(), te, r, fe


def mle_eval(field, table_limbs, num_vars, point_limbs):
    el = LIMBS[field] * DEGREE[field]
    out = np.zeros(el, np.uint64)
    _chk(lib().hgo_mle_eval(field, _p(np.ascontiguousarray(table_limbs, np.uint64)), num_vars, _p(np.ascontiguousarray(point_limbs, np.uint64)), _p(out)))
    return out


def subtable(field, full, bound, log2M, point_limbs=None, want_table=True):
    el = LIMBS[field] * DEGREE[field]
    tab = np.zeros((1 << log2M) * LIMBS[field], np.uint64) if want_table else None
    mle = np.zeros(el, np.uint64) if point_limbs is not None else None
    pt = np.ascontiguousarray(point_limbs, np.uint64) if point_limbs is not None else None
    _chk(lib().hgo_subtable(field, 1 if full else 0, int(bound), log2M, _p(pt), _p(tab), _p(mle)))
    return tab, mle


def field_op(field, op, a, b):
    """op: 0 add, 1 sub, 2 mul, 3 inv(a) on E elements given as limb arrays."""
    el = LIMBS[field] * DEGREE[field]
    out = np.zeros(el, np.uint64)
    _chk(lib().hgo_field_op(field, op, _p(np.ascontiguousarray(a, np.uint64)), _p(np.ascontiguousarray(b, np.uint64)), _p(out)))
    return out


def ntt(field, data_limbs, log_n, inverse=False):
    """Batched radix-2 NTT in natural order (assumption A9); data: [batch, 2^log_n(, limbs)]."""
    a = np.ascontiguousarray(data_limbs, np.uint64).copy()
    batch = a.size // ((1 << log_n) * LIMBS[field])
    _chk(lib().hgo_ntt(field, _p(a), log_n, 1 if inverse else 0, batch))
    return a


def bfv_eval(field, P, ins):
    """Forward evaluation of the BFV circuit on get_inputs() vectors -> (lasso inputs, `sum` node output), limb arrays."""
    k = LIMBS[field]
    K, L = P.K, P.log2_size
    N2 = 1 << L
    tl = lambda v: ints_to_limbs(v, field)
    s, e, k1 = tl(ins["s"]), tl(ins["e"]), tl(ins["k1"])
    ais = np.concatenate([tl(v) for v in ins["ais"]])
    r1is = np.concatenate([tl(v) for v in ins["r1is"]])
    r2is = tl(ins["r2is"])
    nch = max(1, (len(ins["r2is"]) + N2 - 1) // N2)
    lasso = np.zeros((K + nch + 3) * N2 * k, np.uint64)
    n_l = C.c_size_t(0)
    summ = np.zeros(K * N2 * k, np.uint64)
    _chk(lib().hgo_bfv_eval(field, L, K, _p(tl(P.QIS)), _p(tl(P.K0IS)), _p(np.array(P.R1_BOUNDS, np.uint64)), _p(np.array(P.R2_BOUNDS, np.uint64)),
                            P.S_BOUND, P.E_BOUND, P.K1_BOUND, _p(s), _p(e), _p(k1), _p(ais), _p(r1is), _p(r2is), _p(lasso), C.byref(n_l), _p(summ)))
    return lasso[: n_l.value * k], summ


def _bfv_args(field, P, ins, ct0is):
    tl = lambda v: ints_to_limbs(v, field)
    return [_p(x) for x in ()], dict(
        q=tl(P.QIS), k0=tl(P.K0IS), r1b=np.array(P.R1_BOUNDS, np.uint64), r2b=np.array(P.R2_BOUNDS, np.uint64),
        s=tl(ins["s"]), e=tl(ins["e"]), k1=tl(ins["k1"]), ais=np.concatenate([tl(v) for v in ins["ais"]]),
        r1is=np.concatenate([tl(v) for v in ins["r1is"]]), r2is=tl(ins["r2is"]), ct=tl(ct0is))


def bfv_prove(field, P, ins, ct0is, cap=1 << 24):
    """BfvEncrypt::prove (sk_encryption_circuit.rs:417-460) with the restated GKR engine (oracle/gkr.hpp). Returns proof bytes."""
    _, a = _bfv_args(field, P, ins, ct0is)
    proof = np.zeros(cap, np.uint8)
    ln = C.c_size_t(0)
    _chk(lib().hgo_bfv_prove(field, P.log2_size, P.K, _p(a["q"]), _p(a["k0"]), _p(a["r1b"]), _p(a["r2b"]), P.S_BOUND, P.E_BOUND, P.K1_BOUND,
                             _p(a["s"]), _p(a["e"]), _p(a["k1"]), _p(a["ais"]), _p(a["r1is"]), _p(a["r2is"]), _p(a["ct"]), _p(proof), cap, C.byref(ln), 0))
    return proof[: ln.value].tobytes()


def bfv_verify(field, P, ins, ct0is, proof: bytes):
    """BfvEncrypt::verify (sk_encryption_circuit.rs:462-517). Raises OracleError where the reference would panic / Err."""
    _, a = _bfv_args(field, P, ins, ct0is)
    buf = np.frombuffer(proof, np.uint8).copy()
    ln = C.c_size_t(buf.size)
    _chk(lib().hgo_bfv_prove(field, P.log2_size, P.K, _p(a["q"]), _p(a["k0"]), _p(a["r1b"]), _p(a["r2b"]), P.S_BOUND, P.E_BOUND, P.K1_BOUND,
                             _p(a["s"]), _p(a["e"]), _p(a["k1"]), _p(a["ais"]), _p(a["r1is"]), _p(a["r2is"]), _p(a["ct"]), _p(buf), buf.size, C.byref(ln), 1))


class BfvSession:
    """Circuit built and evaluated once (the reference's witness gen); prove() times only the `GKR prove` span."""

    def __init__(self, field, P, ins, ct0is):
        _, a = _bfv_args(field, P, ins, ct0is)
        self.h = lib().hgo_bfv_session_new(field, P.log2_size, P.K, _p(a["q"]), _p(a["k0"]), _p(a["r1b"]), _p(a["r2b"]), P.S_BOUND, P.E_BOUND, P.K1_BOUND,
                                           _p(a["s"]), _p(a["e"]), _p(a["k1"]), _p(a["ais"]), _p(a["r1is"]), _p(a["r2is"]), _p(a["ct"]))
        if not self.h:
            raise OracleError(lib().hgo_last_error().decode())
        self.buf = np.zeros(1 << 24, np.uint8)

    def prove(self) -> bytes:
        ln = C.c_size_t(0)
        _chk(lib().hgo_bfv_session_prove(self.h, _p(self.buf), self.buf.size, C.byref(ln)))
        return self.buf[: ln.value].tobytes()

    def __del__(self):
        try:
            lib().hgo_bfv_session_free(self.h)
        except Exception:
            pass


def bfv_prepare(field, P, ins, ct0is):
    return BfvSession(field, P, ins, ct0is)


def bfv_prove_prepared(session) -> bytes:
    return session.prove()
