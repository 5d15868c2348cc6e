"""Generates sipp_b200/csrc/line_programs.h: the lane-parallel straight-line programs of the G2 side of the Miller loop.

The latency-bound rounds of the prover (n <= 2^13 pairs per GPU) leave most of the GPU idle when one thread walks one
G2 point (k_lines: ~28 dependent Fq2 products per doubling).  `k_lines_wide` instead gives every pair a GROUP of 16
lanes that executes a static schedule: each level is either
    MUL   dst = s0 * s1 (+/-) s2 * s3        one lazy-reduction inner product (fq_dot<2>) per lane
    LIN   dst = sum_k c_k * s_k              small integer coefficients, one reduction per lane
over Fq values held in a shared-memory slot file, so a doubling step is 2 MUL levels instead of ~28 sequential products.
This script builds the schedules from the formulas below, checks them by executing them over Python integers against
the affine model in tests/golden/sipp_model.py (line values after final exponentiation), and emits the tables.

Coordinates: T = (X, Y, Z) homogeneous projective on the twist E': y^2 = x^3 + 3/xi, kept as (X, Y, Zh) with Z = xi * Zh
so that 3 b' Z^2 = 9 xi Zh^2 is a LIN op (small coefficients) instead of a product with the constant b'.  Halvings of the
usual formulas are removed by scaling the new T by 4 (projective coordinates; the lines are homogeneous in T, and any Fq2
factor of a line is killed by the final exponentiation).

    python tools/gen_line_programs.py            # writes the header
    python tools/gen_line_programs.py --check    # also runs the self-check against the model (slow-ish, ~10 s)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import sipp_model as m  # noqa: E402

P = m.P
LANES = 16
MAX_COEF_SUM = 2047


class Slots:
    def __init__(self):
        self.idx = {}

    def __getitem__(self, name):
        if name not in self.idx:
            self.idx[name] = len(self.idx)
            assert len(self.idx) <= 250, "slot file too large"
        return self.idx[name]


S = Slots()
# fixed slots (the kernel fills these before SETUP)
FIXED = ["ZERO", "XP", "YP", "QX0", "QX1", "QY0", "QY1", "XIINV0", "XIINV1", "G12_0", "G12_1", "G13_0", "G13_1"]
for name in FIXED:
    S[name]
OUT = ["OUT%d" % i for i in range(10)]  # l0 yP | l1 xP | xi l1 xP | l3 | xi l3   (c0, c1 each)
for name in OUT:
    S[name]


class Program:
    def __init__(self, name):
        self.name, self.levels = name, []

    def mul(self, ops):
        """ops: (dst, s0, s1, s2, s3, neg): dst = s0*s1 + (-1)^neg s2*s3"""
        assert len(ops) <= LANES, (self.name, len(ops))
        self.levels.append(("MUL", ops))

    def mul1(self, ops):
        """ops: (dst, s0, s1): dst = s0 * s1 (plain Fq product; the G1 programs)"""
        assert len(ops) <= LANES, (self.name, len(ops))
        self.levels.append(("MUL1", ops))

    def lin(self, ops):
        """ops: (dst, [(coef, src), ...]) with at most 4 terms"""
        assert len(ops) <= LANES, (self.name, len(ops))
        for dst, terms in ops:
            assert 1 <= len(terms) <= 4 and sum(abs(c) for c, _ in terms) <= MAX_COEF_SUM, (self.name, dst, terms)
        self.levels.append(("LIN", ops))


def f2(v):
    """an Fq2 operand: a base name ("X" -> slots X0, X1) or an explicit pair of slot names"""
    return v if isinstance(v, tuple) else (v + "0", v + "1")


def mul2(dst, a, b, conj_a=False):
    """Fq2 product dst = a * b (a optionally conjugated): two MUL ops"""
    d, a, b = f2(dst), f2(a), f2(b)
    if not conj_a:
        return [(d[0], a[0], b[0], a[1], b[1], 1), (d[1], a[0], b[1], a[1], b[0], 0)]
    return [(d[0], a[0], b[0], a[1], b[1], 0), (d[1], a[0], b[1], a[1], b[0], 1)]


def scale2(dst, a, k):
    """Fq2 * Fq: dst = a * k"""
    d, a = f2(dst), f2(a)
    return [(d[0], a[0], k, "ZERO", "ZERO", 0), (d[1], a[1], k, "ZERO", "ZERO", 0)]


class XI:
    """coefficient k * xi (xi = 9 + u)"""
    def __init__(self, k=1): self.k = k


class XI2:
    """coefficient k * xi^2 (xi^2 = 80 + 18 u)"""
    def __init__(self, k=1): self.k = k


def lin2(dst, terms):
    """Fq2 linear combination dst = sum coef * value; coef is an int, XI(k) or XI2(k)"""
    d = f2(dst)
    t0, t1 = {}, {}

    def acc(t, c, s):
        t[s] = t.get(s, 0) + c
    for c, v in terms:
        a = f2(v)
        if isinstance(c, (XI, XI2)):
            re, im = (9 * c.k, c.k) if isinstance(c, XI) else (80 * c.k, 18 * c.k)
            acc(t0, re, a[0]); acc(t0, -im, a[1])
            acc(t1, im, a[0]); acc(t1, re, a[1])
        else:
            acc(t0, c, a[0]); acc(t1, c, a[1])
    return [(d[0], [(c, s) for s, c in t0.items() if c]), (d[1], [(c, s) for s, c in t1.items() if c])]


OUT_L0, OUT_L1, OUT_XL1, OUT_L3, OUT_XL3 = ("OUT0", "OUT1"), ("OUT2", "OUT3"), ("OUT4", "OUT5"), ("OUT6", "OUT7"), ("OUT8", "OUT9")
G12, G13, XIINV = ("G12_0", "G12_1"), ("G13_0", "G13_1"), ("XIINV0", "XIINV1")


def build_dbl():
    """tangent at T and T <- 2T (x4):  b = Y^2, c = Zh^2, e = 3 b' Z^2 = 9 xi c, f = 3e, h = 2 Y Z = 2 xi (Y Zh)
         l0 = -h, l1 = 3 X^2, l3 = e - b ;  X' = 2 XY (b - f), Y' = (b + f)^2 - 12 e^2, Z' = 4 b h  =>  Zh' = 8 b (Y Zh)"""
    p = Program("DBL")
    p.mul(mul2("XY", "X", "Y") + mul2("B", "Y", "Y") + mul2("C", "ZH", "ZH") + mul2("YZ", "Y", "ZH") + mul2("J", "X", "X"))
    p.lin(lin2("E", [(XI(9), "C")]) + lin2("BMF", [(1, "B"), (XI(-27), "C")]) + lin2("BPF", [(1, "B"), (XI(27), "C")])
          + lin2(OUT_L3, [(XI(9), "C"), (-1, "B")]) + lin2(OUT_XL3, [(XI2(9), "C"), (XI(-1), "B")])
          + lin2("L0", [(XI(-2), "YZ")]) + lin2("L1", [(3, "J")]))
    p.mul(mul2("E2", "E", "E") + mul2("G2", "BPF", "BPF") + mul2("XN", "XY", "BMF") + mul2("ZN", "B", "YZ")
          + scale2(OUT_L0, "L0", "YP") + scale2(OUT_L1, "L1", "XP"))
    p.lin(lin2("X", [(2, "XN")]) + lin2("Y", [(1, "G2"), (-12, "E2")]) + lin2("ZH", [(8, "ZN")]) + lin2(OUT_XL1, [(XI(1), OUT_L1)]))
    return p


def build_add(name, q, sign):
    """chord through T and sign * Q (Q affine, hats qH = xi q) and T <- T + sign * Q:
         theta = Y - qy Z, lambda = X - qx Z ; l0 = lambda, l1 = -theta, l3 = theta qx - lambda qy
         c = theta^2, d = lambda^2, e = lambda d, f = Z c, g = X d, h = e + f - 2g
         X' = lambda h, Y' = theta (g - h) - e Y, Z' = Z e"""
    qx, qy, qhx, qhy = q + "X", q + "Y", q + "HX", q + "HY"
    p = Program(name)
    p.mul(mul2("A", qhy, "ZH") + mul2("BQ", qhx, "ZH"))
    p.lin(lin2("TH", [(1, "Y"), (-sign, "A")]) + lin2("LA", [(1, "X"), (-1, "BQ")]))
    p.mul(mul2("CC", "TH", "TH") + mul2("DD", "LA", "LA") + mul2("TQ", "TH", qx) + mul2("LQ", "LA", qy)
          + scale2(OUT_L0, "LA", "YP") + scale2("S1", "TH", "XP"))
    p.lin(lin2(OUT_L1, [(-1, "S1")]) + lin2(OUT_XL1, [(XI(-1), "S1")])
          + lin2(OUT_L3, [(1, "TQ"), (-sign, "LQ")]) + lin2(OUT_XL3, [(XI(1), "TQ"), (XI(-sign), "LQ")]))
    p.mul(mul2("EV", "LA", "DD") + mul2("FH", "ZH", "CC") + mul2("GV", "X", "DD"))
    p.lin(lin2("H", [(1, "EV"), (XI(1), "FH"), (-2, "GV")]) + lin2("GMH", [(3, "GV"), (-1, "EV"), (XI(-1), "FH")]))
    p.mul(mul2("X", "LA", "H") + mul2("YA", "TH", "GMH") + mul2("YB", "EV", "Y") + mul2("ZH", "ZH", "EV"))
    p.lin(lin2("Y", [(1, "YA"), (-1, "YB")]))
    return p


def build_dbl1():
    """G1 (y^2 = x^3 + 3 over Fq), T <- 2T (x4) in homogeneous projective coordinates (X1, Y1, Z1):
         b = Y^2, c = Z^2, e = 3 b Z^2 = 9c, f = 3e ; X' = 2 XY (b - f), Y' = (b + f)^2 - 12 e^2, Z' = 8 b (YZ)
       expanded so that no linear level sits between the two product levels (the G1 components are the longest chain of
       k_fold_wide: 128 doublings):  X' = 2 (XY b) - 54 (XY c),  Y' = b^2 + 54 bc - 243 c^2,  Z' = 8 b (YZ)"""
    p = Program("DBL1")
    p.mul1([("XY1", "X1", "Y1"), ("B1", "Y1", "Y1"), ("C1", "Z1", "Z1"), ("YZ1", "Y1", "Z1")])
    p.mul1([("BB1", "B1", "B1"), ("BC1", "B1", "C1"), ("CC1", "C1", "C1"), ("XYB1", "XY1", "B1"), ("XYC1", "XY1", "C1"), ("BYZ1", "B1", "YZ1")])
    p.lin([("X1", [(2, "XYB1"), (-54, "XYC1")]), ("Y1", [(1, "BB1"), (54, "BC1"), (-243, "CC1")]), ("Z1", [(8, "BYZ1")])])
    return p


def build_add1(name, sign):
    """G1, T <- T + sign * Q with Q = (QX0, QY0) affine (same formulas as build_add, over Fq)"""
    p = Program(name)
    p.mul1([("A1", "QY0", "Z1"), ("BQ1", "QX0", "Z1")])
    p.lin([("TH1", [(1, "Y1"), (-sign, "A1")]), ("LA1", [(1, "X1"), (-1, "BQ1")])])
    p.mul1([("CC1", "TH1", "TH1"), ("DD1", "LA1", "LA1")])
    p.mul1([("EV1", "LA1", "DD1"), ("FH1", "Z1", "CC1"), ("GV1", "X1", "DD1")])
    p.lin([("H1", [(1, "EV1"), (1, "FH1"), (-2, "GV1")]), ("GMH1", [(3, "GV1"), (-1, "EV1"), (-1, "FH1")])])
    p.mul1([("X1", "LA1", "H1"), ("YA1", "TH1", "GMH1"), ("YB1", "EV1", "Y1"), ("Z1", "Z1", "EV1")])
    p.lin([("Y1", [(1, "YA1"), (-1, "YB1")])])
    return p


def build_setup():
    """Q1 = pi(Q), Q2 = pi(Q1) (used negated), hats, and T = Q (Z = 1  =>  Zh = xi^-1)"""
    p = Program("SETUP")
    p.mul(mul2("Q1X", "QX", G12, conj_a=True) + mul2("Q1Y", "QY", G13, conj_a=True))
    p.lin(lin2("QHX", [(XI(1), "QX")]) + lin2("QHY", [(XI(1), "QY")]) + lin2("Q1HX", [(XI(1), "Q1X")]) + lin2("Q1HY", [(XI(1), "Q1Y")])
          + lin2("X", [(1, "QX")]) + lin2("Y", [(1, "QY")]) + lin2("ZH", [(1, XIINV)]))
    p.mul(mul2("Q2X", "Q1X", G12, conj_a=True) + mul2("Q2Y", "Q1Y", G13, conj_a=True))
    p.lin(lin2("Q2HX", [(XI(1), "Q2X")]) + lin2("Q2HY", [(XI(1), "Q2Y")]))
    return p


# ---------------------------------------------------------------------------------------------------------------------
# integer interpreter (plain residues: MUL is a * b mod p) -- the device runs the same tables in Montgomery form
# ---------------------------------------------------------------------------------------------------------------------
def run(prog, slots):
    for kind, ops in prog.levels:
        new = {}
        for op in ops:
            if kind == "MUL":
                dst, s0, s1, s2, s3, neg = op
                v = slots[s0] * slots[s1] + (-1 if neg else 1) * slots[s2] * slots[s3]
            elif kind == "MUL1":
                dst, s0, s1 = op
                v = slots[s0] * slots[s1]
            else:
                dst, terms = op
                v = sum(c * slots[s] for c, s in terms)
            new[dst] = v % P
        slots.update(new)  # all reads of a level happen before its writes


def model_lines_check(programs, seed=5, pairs=2):
    """full Miller loop through the programs -> f, compared with the affine model after the final exponentiation"""
    A, B = m.seeded_inputs(seed, pairs)
    digits = m.ATE_DIGITS
    for a, b in zip(A, B):
        slots = {"ZERO": 0, "XP": a[0], "YP": a[1], "QX0": b[0][0], "QX1": b[0][1], "QY0": b[1][0], "QY1": b[1][1]}
        xiinv = m.f2_inv(m.XI)
        slots.update({"XIINV0": xiinv[0], "XIINV1": xiinv[1], "G12_0": m.GAMMA[1][2][0], "G12_1": m.GAMMA[1][2][1],
                      "G13_0": m.GAMMA[1][3][0], "G13_1": m.GAMMA[1][3][1]})
        run(programs["SETUP"], slots)
        f = m.F12_ONE
        nlines = 0

        def take_line(f):
            l0 = (slots["OUT0"], slots["OUT1"]); l1 = (slots["OUT2"], slots["OUT3"]); l3 = (slots["OUT6"], slots["OUT7"])
            assert (slots["OUT4"], slots["OUT5"]) == m.f2_mul(m.XI, l1) and (slots["OUT8"], slots["OUT9"]) == m.f2_mul(m.XI, l3)
            return m.f12_mul(f, [l0, l1, m.F2_ZERO, l3, m.F2_ZERO, m.F2_ZERO])
        for i in range(63, -1, -1):
            if i != 63:
                f = m.f12_sqr(f)
            run(programs["DBL"], slots); f = take_line(f); nlines += 1
            if digits[i]:
                run(programs["ADD_P" if digits[i] == 1 else "ADD_M"], slots); f = take_line(f); nlines += 1
        run(programs["ADD_Q1"], slots); f = take_line(f); nlines += 1
        run(programs["ADD_Q2"], slots); f = take_line(f); nlines += 1
        assert nlines == 91
        assert m.final_exp(f) == m.pairing(a, b), "program Miller loop disagrees with the model"
    return True


def model_scalar_mul_check(programs, seed=9):
    """the fold kernels' use of the programs: [k]Q by signed digits through DBL / ADD_P / ADD_M (G2) and DBL1 / ADD1_* (G1)"""
    import random
    rng = random.Random(seed)
    A, B = m.seeded_inputs(seed, 1)
    for trial in range(4):
        k = rng.randrange(2, 1 << 66)
        digits = []  # NAF, little endian
        kk = k
        while kk:
            d = 0
            if kk & 1:
                d = 2 - (kk & 3)
                kk -= d
            digits.append(d)
            kk >>= 1
        assert digits[-1] == 1
        # G2
        q = B[0]
        xiinv = m.f2_inv(m.XI)
        slots = {"ZERO": 0, "XP": 0, "YP": 0, "QX0": q[0][0], "QX1": q[0][1], "QY0": q[1][0], "QY1": q[1][1], "XIINV0": xiinv[0], "XIINV1": xiinv[1],
                 "G12_0": m.GAMMA[1][2][0], "G12_1": m.GAMMA[1][2][1], "G13_0": m.GAMMA[1][3][0], "G13_1": m.GAMMA[1][3][1]}
        run(programs["SETUP"], slots)
        for d in reversed(digits[:-1]):
            run(programs["DBL"], slots)
            if d:
                run(programs["ADD_P" if d == 1 else "ADD_M"], slots)
        z = m.f2_mul(m.XI, (slots["ZH0"], slots["ZH1"]))
        zi = m.f2_inv(z)
        got = (m.f2_mul((slots["X0"], slots["X1"]), zi), m.f2_mul((slots["Y0"], slots["Y1"]), zi))
        assert got == m.g2_mul(q, k), "G2 scalar multiplication through the programs"
        # G1
        p1 = A[0]
        slots = {"ZERO": 0, "QX0": p1[0], "QY0": p1[1], "X1": p1[0], "Y1": p1[1], "Z1": 1}
        for d in reversed(digits[:-1]):
            run(programs["DBL1"], slots)
            if d:
                run(programs["ADD1_P" if d == 1 else "ADD1_M"], slots)
        zi = pow(slots["Z1"], P - 2, P)
        assert (slots["X1"] * zi % P, slots["Y1"] * zi % P) == m.g1_mul(p1, k), "G1 scalar multiplication through the programs"
    return True


def emit(programs, path):
    order = ["SETUP", "DBL", "ADD_P", "ADD_M", "ADD_Q1", "ADD_Q2", "DBL1", "ADD1_P", "ADD1_M"]
    code, types, index = [], [], {}
    for name in order:
        prog = programs[name]
        index[name] = (len(types), len(prog.levels))
        for kind, ops in prog.levels:
            types.append({"MUL": 0, "LIN": 1, "MUL1": 2}[kind])
            for lane in range(LANES):
                if lane < len(ops):
                    if kind == "MUL":
                        dst, s0, s1, s2, s3, neg = ops[lane]
                        srcs, coefs, flags = [S[s0], S[s1], S[s2], S[s3]], [0, 0, 0, 0], (1 if neg else 0)
                        dsti = S[dst]
                    elif kind == "MUL1":
                        dst, s0, s1 = ops[lane]
                        srcs, coefs, flags, dsti = [S[s0], S[s1], S["ZERO"], S["ZERO"]], [0, 0, 0, 0], 0, S[dst]
                    else:
                        dst, terms = ops[lane]
                        terms = terms + [(0, "ZERO")] * (4 - len(terms))
                        srcs, coefs, flags = [S[s] for _, s in terms], [c for c, _ in terms], 0
                        dsti = S[dst]
                else:
                    dsti, srcs, coefs, flags = S["DUMP%d" % lane], [S["ZERO"]] * 4, [0, 0, 0, 0], 0
                w0 = dsti | srcs[0] << 8 | srcs[1] << 16 | srcs[2] << 24
                w1 = srcs[3] | flags << 8
                w2 = (coefs[0] & 0xFFFF) | (coefs[1] & 0xFFFF) << 16
                w3 = (coefs[2] & 0xFFFF) | (coefs[3] & 0xFFFF) << 16
                code.append((w0, w1, w2, w3))
    out = ["// GENERATED by tools/gen_line_programs.py -- do not edit.  Lane-parallel schedules of the G2 side of the Miller loop.",
           "#pragma once", "#include <stdint.h>",
           "#define SIPP_LP_LANES %d" % LANES,
           "#define SIPP_LP_SLOTS %d" % len(S.idx),
           "#define SIPP_LP_LEVELS %d" % len(types)]
    for name in FIXED + OUT[:1] + ["X0", "X1", "Y0", "Y1", "ZH0", "ZH1", "Z1"]:
        out.append("#define SIPP_LP_SLOT_%s %d" % (name, S[name]))
    for name in order:
        out.append("#define SIPP_LP_%s_FIRST %d" % (name, index[name][0]))
        out.append("#define SIPP_LP_%s_LEVELS %d" % (name, index[name][1]))
    out.append("// level type: 0 = MUL (dst = s0 s1 +- s2 s3), 1 = LIN (dst = sum c_k s_k), 2 = MUL1 (dst = s0 s1)")
    out.append("#define SIPP_LP_TYPES_INIT { " + ", ".join(str(t) for t in types) + " }")
    out.append("// one 128-bit instruction per (level, lane): w0 = dst | s0<<8 | s1<<16 | s2<<24, w1 = s3 | flags<<8, w2 = c0 | c1<<16, w3 = c2 | c3<<16")
    out.append("#define SIPP_LP_CODE_INIT { \\")
    for i in range(0, len(code), 4):
        out.append("  " + " ".join("{0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}," % w for w in code[i:i + 4]) + " \\")
    out.append("}")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    return len(types), len(S.idx)


def main():
    programs = {"SETUP": build_setup(), "DBL": build_dbl(), "ADD_P": build_add("ADD_P", "Q", 1), "ADD_M": build_add("ADD_M", "Q", -1),
                "ADD_Q1": build_add("ADD_Q1", "Q1", 1), "ADD_Q2": build_add("ADD_Q2", "Q2", -1),
                "DBL1": build_dbl1(), "ADD1_P": build_add1("ADD1_P", 1), "ADD1_M": build_add1("ADD1_M", -1)}
    if "--check" in sys.argv:
        model_lines_check(programs)
        model_scalar_mul_check(programs)
        print("self-check against the affine model: ok")
    path = os.path.join(ROOT, "sipp_b200", "csrc", "line_programs.h")
    nlev, nslots = emit(programs, path)
    for name, p in programs.items():
        print("%-7s levels %s" % (name, " ".join("%s%d" % (k[0], len(o)) for k, o in p.levels)))
    print("wrote %s: %d levels, %d slots" % (path, nlev, nslots))


if __name__ == "__main__":
    main()
