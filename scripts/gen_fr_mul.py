"""Generator + bit-exact simulator of the even/odd-column 8x32-bit Montgomery multiplication of BN254 Fr used on the device
(hyper-greco_b200/csrc/bn254.cuh: fr_mul_dev32). `python scripts/gen_fr_mul.py` builds the instruction list, executes it in Python with
exact carry-flag semantics on 20 000 random / extreme operand pairs and checks a * b * R^-1 mod r; the PTX in bn254.cuh is this list
printed one instruction per line (moduli and -r^-1 mod 2^32 as immediates)."""
import random
P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
W = 1 << 32
N = 8
INV32 = (-pow(P, -1, W)) % W
PW = [(P >> (32 * i)) & (W - 1) for i in range(N)]

class Prog:
    def __init__(self):
        self.ins = []
    def emit(self, op, d, *src):
        self.ins.append((op, d, src))

def build():
    p = Prog()
    A = [f"a{i}" for i in range(N)]; B = [f"b{i}" for i in range(N)]; M = [f"p{i}" for i in range(N)]
    ev = [f"e{i}" for i in range(N)]; od = [f"o{i}" for i in range(N)]
    def mul_n(acc, a, off, bi):
        for j in range(0, N, 2):
            p.emit("mul.lo", acc[j], a[off + j], bi)
            p.emit("mul.hi", acc[j + 1], a[off + j], bi)
    def cmad_n(acc, a, off, bi):
        p.emit("mad.lo.cc", acc[0], a[off], bi, acc[0])
        p.emit("madc.hi.cc", acc[1], a[off], bi, acc[1])
        for j in range(2, N, 2):
            p.emit("madc.lo.cc", acc[j], a[off + j], bi, acc[j])
            p.emit("madc.hi.cc", acc[j + 1], a[off + j], bi, acc[j + 1])
    def madc_n_rshift(odd, a, off, bi):
        for j in range(0, N - 2, 2):
            p.emit("madc.lo.cc", odd[j], a[off + j], bi, odd[j + 2])
            p.emit("madc.hi.cc", odd[j + 1], a[off + j], bi, odd[j + 3])
        j = N - 2
        p.emit("madc.lo.cc", odd[j], a[off + j], bi, "0")
        p.emit("madc.hi", odd[j + 1], a[off + j], bi, "0")
    def mad_n_redc(even, odd, bi, first):
        if first:
            mul_n(odd, A, 1, bi)
            mul_n(even, A, 0, bi)
        else:
            p.emit("add.cc", even[0], even[0], odd[1])
            madc_n_rshift(odd, A, 1, bi)
            cmad_n(even, A, 0, bi)
            p.emit("addc", odd[N - 1], odd[N - 1], "0")
        p.emit("mul.lo", "mi", even[0], "inv")
        cmad_n(odd, M, 1, "mi")
        cmad_n(even, M, 0, "mi")
        p.emit("addc", odd[N - 1], odd[N - 1], "0")
    for i in range(0, N, 2):
        mad_n_redc(ev, od, B[i], i == 0)
        mad_n_redc(od, ev, B[i + 1], False)
    # merge: result = even>>32 + odd  (even[0] == 0)
    p.emit("add.cc", ev[0], ev[0], od[1])
    for i in range(1, N - 1):
        p.emit("addc.cc", ev[i], ev[i], od[i + 1])
    p.emit("addc", ev[N - 1], ev[N - 1], "0")
    return p

def cmad_operand(a_list_name):
    return a_list_name

def simulate(prog, a, b):
    reg = {"0": 0, "inv": INV32}
    for i in range(N):
        reg[f"a{i}"] = (a >> (32 * i)) & (W - 1); reg[f"b{i}"] = (b >> (32 * i)) & (W - 1); reg[f"p{i}"] = PW[i]
    # a[8] / p[8] accesses (off + j with off = 1, j = 6 -> index 7): fine
    cf = 0
    for op, d, src in prog.ins:
        v = [reg[s] for s in src]
        if op == "mul.lo": reg[d] = (v[0] * v[1]) % W
        elif op == "mul.hi": reg[d] = (v[0] * v[1]) >> 32
        elif op in ("mad.lo.cc", "madc.lo.cc", "madc.lo"):
            t = (v[0] * v[1]) % W + v[2] + (cf if op.startswith("madc") else 0)
            reg[d] = t % W
            if op.endswith(".cc"): cf = t >> 32
        elif op in ("mad.hi.cc", "madc.hi.cc", "madc.hi"):
            t = ((v[0] * v[1]) >> 32) + v[2] + (cf if op.startswith("madc") else 0)
            reg[d] = t % W
            if op.endswith(".cc"): cf = t >> 32
            elif t >> 32: raise OverflowError("carry lost in " + op)
        elif op in ("add.cc", "addc.cc", "addc"):
            t = v[0] + v[1] + (cf if op.startswith("addc") else 0)
            reg[d] = t % W
            if op.endswith(".cc"): cf = t >> 32
            elif t >> 32: raise OverflowError("carry lost in addc")
        else: raise ValueError(op)
    # result: after the merge even[] holds words of T (aligned: even[0] is word... ) see below
    return reg

if __name__ == "__main__":
    prog = build()
    print(len(prog.ins), "instructions")
    R = 1 << 256
    Rinv = pow(R, -1, P)
    rnd = random.Random(1)
    bad = 0
    for t in range(20000):
        a = rnd.randrange(P) if t > 3 else [0, P - 1, 1, P - 1][t]
        b = rnd.randrange(P) if t > 3 else [0, P - 1, P - 1, 1][t]
        reg = simulate(prog, a, b)
        # candidates for where the result lives
        ev = sum(reg[f"e{i}"] << (32 * i) for i in range(N))
        want = a * b * Rinv % P
        if ev % P != want or ev >= 2 * P:
            bad += 1
            if bad < 5: print("MISMATCH", hex(ev), hex(want), ev >= 2 * P)
    print("bad", bad)
