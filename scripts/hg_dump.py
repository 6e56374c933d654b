"""Interchange dump of one proof run: what a REAL run of the Rust prover has to emit so that parity with it
(P2) can be decided off-box, and the same thing emitted by the CPU oracle here under every upstream-assumption switch.

File format `hg_dump` v1
    line 1   a JSON object (UTF-8, terminated by '\n'):
             {"format": "hg_dump", "version": 1, "field": "goldilocks" | "bn254", "params": "1024_1x27_65537",
              "what": "bfv_encrypt", "elem_bytes": 8 | 32, "n_events": N,
              optional: "assumptions": {"A3_wire": 0|1, "A3_h1": 0|1, "A5_ascending": 0|1}, "proof_sha256": "...", "producer": "..."}
    rest     N events, each 1 + elem_bytes bytes:  kind 'S' (a BASE-field challenge squeezed, transcript.rs:199-203) or
             'W' (a BASE-field element written, transcript.rs:183-188), then the element as the proof stores it (to_repr reversed,
             big-endian). The concatenation of the 'W' payloads is the proof; the S/W pattern is the protocol structure (node order,
             where alpha / gamma / mu are squeezed), the 'S' payloads pin the challenge derivation (A1, A2, A11).

The Rust side: patches/hyper-greco-dump.diff makes `HG_DUMP=/tmp/x.hgdump cargo test -r test_sk_enc_valid_goldilocks_1024_1x27_65537`
write this file. Here:
    python scripts/hg_dump.py make            # regenerate tests/golden/dumps/*.hgdump (CPU oracle, all 8 switch settings)
    python scripts/compare_dump.py FILE       # which switch setting (if any) does FILE agree with; where is the first difference
"""
import hashlib
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP_DIR = os.path.join(ROOT, "tests", "golden", "dumps")
SETTINGS = [dict(A3_wire=a, A3_h1=b, A5_ascending=c) for a, b, c in itertools.product((0, 1), (0, 1), (1, 0))]
ELEM_BYTES = {"goldilocks": 8, "bn254": 32}


def setting_tag(s):
    return f"a3w{s['A3_wire']}_a3h{s['A3_h1']}_a5{s['A5_ascending']}"


def write_dump(path, header, events: bytes):
    eb = header["elem_bytes"]
    assert len(events) % (1 + eb) == 0
    header = dict(header, format="hg_dump", version=1, n_events=len(events) // (1 + eb))
    with open(path, "wb") as f:
        f.write((json.dumps(header, sort_keys=True) + "\n").encode())
        f.write(events)


def read_dump(path):
    raw = open(path, "rb").read()
    nl = raw.index(b"\n")
    header = json.loads(raw[:nl].decode())
    if header.get("format") != "hg_dump" or header.get("version") != 1:
        raise ValueError(f"{path}: not an hg_dump v1 file")
    eb = int(header["elem_bytes"])
    body = raw[nl + 1:]
    if len(body) != header["n_events"] * (1 + eb):
        raise ValueError(f"{path}: {len(body)} event bytes, header says {header['n_events']} events of {1 + eb}")
    events = [(body[i:i + 1].decode(), body[i + 1:i + 1 + eb]) for i in range(0, len(body), 1 + eb)]
    return header, events


def proof_of(events):
    return b"".join(v for k, v in events if k == "W")


def oracle_dump(field, name, setting):
    """(header, event bytes) of BfvEncrypt::prove on the reference's own witness for `name`, by the CPU oracle under `setting`."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params
    from oracle import hgo
    hgo.build()
    P = params.PARAMS[name]
    g = os.path.join(ROOT, "tests", "golden")
    if field == "goldilocks":
        io = np.load(os.path.join(g, f"circuit_io_{name}.npz"))
        ints = lambda a: [int(v) for v in a.reshape(-1)]
    else:
        io = np.load(os.path.join(g, f"circuit_io_bn254_{name}.npz"))
        ints = lambda a: [sum(int(r[j]) << (64 * j) for j in range(4)) for r in a.reshape(-1, 4)]
    ins = dict(s=ints(io["s"]), e=ints(io["e"]), k1=ints(io["k1"]), ais=[ints(a) for a in io["ais"]], r1is=[ints(a) for a in io["r1is"]], r2is=ints(io["r2is"]))
    fid = 0 if field == "goldilocks" else 1
    hgo.set_assumption(3, setting["A3_wire"]); hgo.set_assumption(31, setting["A3_h1"]); hgo.set_assumption(5, setting["A5_ascending"])
    try:
        hgo.events_begin()
        proof = hgo.bfv_prove(fid, P, ins, ints(io["ct0is"]))
        ev = hgo.events_end()
    finally:
        hgo.set_assumption(3, 0); hgo.set_assumption(31, 0); hgo.set_assumption(5, 1)
    header = dict(field=field, params=name, what="bfv_encrypt", elem_bytes=ELEM_BYTES[field], assumptions=setting,
                  proof_sha256=hashlib.sha256(proof).hexdigest(), producer="hyper-greco_b200 CPU oracle (oracle/, ORACLE-GENERATED: not the Rust prover)")
    return header, ev


def main():
    if sys.argv[1:2] != ["make"]:
        print(__doc__)
        return
    os.makedirs(DUMP_DIR, exist_ok=True)
    for s in SETTINGS:
        h, ev = oracle_dump("goldilocks", "1024_1x27_65537", s)
        p = os.path.join(DUMP_DIR, f"bfv_encrypt_goldilocks_1024_1x27_65537_{setting_tag(s)}.hgdump")
        write_dump(p, h, ev)
        print(os.path.basename(p), len(ev), h["proof_sha256"][:16])
    h, ev = oracle_dump("bn254", "1024_1x27_65537", SETTINGS[0])
    p = os.path.join(DUMP_DIR, f"bfv_encrypt_bn254_1024_1x27_65537_{setting_tag(SETTINGS[0])}.hgdump")
    write_dump(p, h, ev)
    print(os.path.basename(p), len(ev), h["proof_sha256"][:16])


if __name__ == "__main__":
    main()
