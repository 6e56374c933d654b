"""Which upstream-assumption setting (A3 wire format, A3' source of h(1), A5 distribute_powers order) does a
proof dump agree with?

    python scripts/compare_dump.py path/to/run.hgdump [--regen]

The dump comes from a real run of the Rust prover (patches/hyper-greco-dump.diff) or from this repository (scripts/hg_dump.py).
It is compared, event by event, with the committed dumps of the same field / parameter set under all 8 switch settings
(tests/golden/dumps/, produced by the CPU oracle; --regen recomputes them instead of reading the files). For every setting the
report says whether the squeeze / write PATTERN agrees (protocol structure: assumptions A6, A7, A8, A9), whether the challenges
agree (A1, A2, A11), and where the first written element differs (A3, A3', A5 and everything the engine computes).
Exit status 0 iff some setting matches the dump completely.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hg_dump  # noqa: E402


def compare(events, ref):
    """-> dict(pattern_equal, first_pattern_diff, challenges_equal, first_write_diff (index among writes) or None, equal)"""
    pat_a, pat_b = "".join(k for k, _ in events), "".join(k for k, _ in ref)
    fp = next((i for i, (x, y) in enumerate(zip(pat_a, pat_b)) if x != y), None)
    if fp is None and len(pat_a) != len(pat_b):
        fp = min(len(pat_a), len(pat_b))
    sa, sb = [v for k, v in events if k == "S"], [v for k, v in ref if k == "S"]
    wa, wb = [v for k, v in events if k == "W"], [v for k, v in ref if k == "W"]
    n = min(len(sa), len(sb))
    ch_equal = sa[:n] == sb[:n]
    fw = next((i for i, (x, y) in enumerate(zip(wa, wb)) if x != y), None)
    if fw is None and len(wa) != len(wb):
        fw = min(len(wa), len(wb))
    return dict(pattern_equal=fp is None, first_pattern_diff=fp, challenges_equal=ch_equal and len(sa) == len(sb), first_write_diff=fw,
                equal=fp is None and ch_equal and fw is None and len(sa) == len(sb))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if len(args) != 1:
        print(__doc__)
        return 2
    header, events = hg_dump.read_dump(args[0])
    field, name = header["field"], header["params"]
    print(f"{args[0]}: {header.get('what', '?')} {field} {name}, {len(events)} events "
          f"({sum(1 for k, _ in events if k == 'S')} squeezes, {sum(1 for k, _ in events if k == 'W')} writes), producer: {header.get('producer', '?')}")
    matched = []
    for s in hg_dump.SETTINGS:
        tag = hg_dump.setting_tag(s)
        path = os.path.join(hg_dump.DUMP_DIR, f"bfv_encrypt_{field}_{name}_{tag}.hgdump")
        if "--regen" in sys.argv or not os.path.exists(path):
            if "--regen" not in sys.argv and field != "goldilocks":
                print(f"  {tag}: no committed dump for this field / parameter set (use --regen)")
                continue
            h, ev = hg_dump.oracle_dump(field, name, s)
            eb = h["elem_bytes"]
            ref = [(ev[i:i + 1].decode(), ev[i + 1:i + 1 + eb]) for i in range(0, len(ev), 1 + eb)]
        else:
            _, ref = hg_dump.read_dump(path)
        r = compare(events, ref)
        if r["equal"]:
            matched.append(tag)
        where = "identical" if r["equal"] else (
            f"pattern differs at event {r['first_pattern_diff']}" if not r["pattern_equal"] else
            ("challenges differ; " if not r["challenges_equal"] else "") + (f"first differing written element: #{r['first_write_diff']}" if r["first_write_diff"] is not None else "writes equal"))
        print(f"  A3_wire={s['A3_wire']} A3_h1={s['A3_h1']} A5_ascending={s['A5_ascending']}: {where}")
    if matched:
        print("MATCH: the dump agrees byte for byte with setting(s) " + ", ".join(matched))
        return 0
    print("NO MATCH: the dump agrees with none of the switch settings; the first differing event above localises the disagreement")
    return 1


if __name__ == "__main__":
    sys.exit(main())
