"""Generates sipp_b200/csrc/fq12_programs.h: lane-parallel programs for Fq12 arithmetic on a 32-lane machine.

Same executor idea as tools/gen_line_programs.py (engine.cuh), one warp per machine: in a level every lane executes one
instruction of the same kind over a shared-memory slot file of Fq values,
    DOTn  dst = sum_{t<n} (+/-) a_t * b_t     n in {2, 4, 6}: ONE lazy-reduction inner product (fq_dot<n>)
    LIN   dst = c0 s0 + c1 s1 + c2 s2 + c3 s3
    INV   dst = s0^-1                          (binary Montgomery inversion; one lane)
so a dense Fq12 product is ONE DOT6 level on 24 lanes (+ two cheap LIN levels) instead of 4 dependent fq_dot<6> per lane
in the 6-lane version (coop.cuh).  This is what the single final exponentiation per product -- a strictly sequential chain
of ~190 cyclotomic squarings and ~100 products -- spends its time on in the latency-bound rounds.

Operands are "registers" of 24 slots (12 coefficients g_k.c at offset 2k + c, their xi-multiples at 12 + 2k + c) named at
run time: an instruction addresses (bank, offset) with bank 0 = globals (constants, scratch), 1 = D, 2 = A, 3 = B.
The programs are checked here by running them over Python integers against tests/golden/sipp_model.py.

    python tools/gen_fq12_programs.py [--check]
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import sipp_model as m  # noqa: E402

P = m.P
LANES = 32
MAX_COEF_SUM = 2047
BANKS = {"G": 0, "D": 1, "A": 2, "B": 3}

# ---- global bank layout ------------------------------------------------------------------------------------------------
ZERO = ("G", 0)
def GAM(k, i, c): return ("G", 1 + ((k - 1) * 6 + i) * 2 + c)      # gamma[k][i], k = 1..3, i = 0..5   (slots 1..36)
def H(j): return ("G", 37 + j)                                       # 24 scratch slots (37..60)
def X(j): return ("G", 61 + j)                                       # 3 more scratch slots (61..63)
def val(r, k, c): return (r, 2 * k + c)
def xiv(r, k, c): return (r, 12 + 2 * k + c)


class Program:
    def __init__(self, name):
        self.name, self.levels = name, []

    def dot(self, n, ops):
        """ops: (dst, [(a, b, neg), ...]) with at most n terms"""
        assert n in (2, 4, 6) and len(ops) <= LANES, (self.name, n, len(ops))
        for dst, terms in ops:
            assert len(terms) <= n
        self.levels.append(("DOT%d" % n, ops))

    def lin(self, ops):
        assert len(ops) <= LANES, (self.name, len(ops))
        for dst, terms in ops:
            assert 1 <= len(terms) <= 4 and sum(abs(c) for c, _ in terms) <= MAX_COEF_SUM, (self.name, dst, terms)
        self.levels.append(("LIN", ops))

    def inv(self, ops):
        self.levels.append(("INV", ops))  # (dst, src)


def fq2_mul_terms(a, b, c, conj_a=False):
    """terms of component c of the Fq2 product a * b; a, b are pairs of slots"""
    s = -1 if conj_a else 1
    if c == 0:
        return [(a[0], b[0], 0), (a[1], b[1], 1 if s == 1 else 0)]
    return [(a[0], b[1], 0), (a[1], b[0], 0 if s == 1 else 1)]


def xi_lin(dst, src, k=1):
    """dst = k * xi * src over Fq2 slot pairs: two LIN ops"""
    return [(dst[0], [(9 * k, src[0]), (-k, src[1])]), (dst[1], [(k, src[0]), (9 * k, src[1])])]


def reg2(r, k, xi=False):
    f = xiv if xi else val
    return (f(r, k, 0), f(r, k, 1))


def xi_of_sum(dst, srcs):
    """dst (an Fq2 slot pair) = xi * sum_j c_j src_j for at most two (c_j, src_j) Fq2 sources: two LIN ops of four terms"""
    assert len(srcs) <= 2
    return [(dst[0], [(9 * c, s[0]) for c, s in srcs] + [(-c, s[1]) for c, s in srcs]),
            (dst[1], [(c, s[0]) for c, s in srcs] + [(9 * c, s[1]) for c, s in srcs])]


def emit_mul12(p, d, a, b, pre=True, post_xi=False):
    """d = a * b (dense):  c_k = sum_{i<=k} a_i b_{k-i} + xi sum_{i>k} a_i b_{k-i+6}.
    pre = False: the xi-multiples of b are already in its xi slots; post_xi: also leave xi * d_3..5 in d's xi slots (what a
    following cyclotomic squaring of d reads), so that chains of squarings and products need no LIN level of their own"""
    ops = []
    if pre:
        for k in range(6):
            ops += xi_lin(reg2(b, k, True), reg2(b, k))
        p.lin(ops)
    ops = []
    for h in range(2):
        for k in range(6):
            for c in range(2):
                terms = []
                for t in range(3):
                    i = 3 * h + t
                    j = k - i
                    wrapped = j < 0
                    if wrapped:
                        j += 6
                    terms += fq2_mul_terms(reg2(a, i), reg2(b, j, wrapped), c)
                ops.append((H(h * 12 + 2 * k + c), terms))
    p.dot(6, ops)
    ops = [(val(d, k, c), [(1, H(2 * k + c)), (1, H(12 + 2 * k + c))]) for k in range(6) for c in range(2)]
    if post_xi:
        for k in range(0 if post_xi == "all" else 3, 6):
            ops += xi_of_sum(reg2(d, k, True), [(1, (H(2 * k), H(2 * k + 1))), (1, (H(12 + 2 * k), H(12 + 2 * k + 1)))])
    p.lin(ops)


def build_mul12():
    p = Program("MUL12")
    emit_mul12(p, "D", "A", "B")
    return p


def build_mul12x():
    """link of an exponentiation chain: D = A * B with B's xi slots valid on entry, D's xi slots 3..5 valid on exit"""
    p = Program("MUL12X")
    emit_mul12(p, "D", "A", "B", pre=False, post_xi=True)
    return p


def build_mul12y():
    """link of a GENERIC power chain (the verifier's Z_L^x, verifier_native.rs:59-61 -- proof elements need not be cyclotomic):
    D = A * B with all of B's xi slots valid on entry and all of D's valid on exit, so squarings D = D * D and products by a
    table entry chain without LIN levels of their own"""
    p = Program("MUL12Y")
    emit_mul12(p, "D", "A", "B", pre=False, post_xi="all")
    return p


def build_xi6():
    """xi slots of D <- xi * D (all six coefficients): what MUL12X expects of its second operand"""
    p = Program("XI6")
    ops = []
    for k in range(6):
        ops += xi_lin(reg2("D", k, True), reg2("D", k))
    p.lin(ops)
    return p


def build_copyx():
    """D <- A including the xi slots"""
    p = Program("COPYX")
    p.lin([(("D", j), [(1, ("A", j))]) for j in range(24)])
    return p


def build_csqr(fused=False):
    """fused = True (CSQRX): link of an exponentiation chain -- A's xi slots 3..5 are valid on entry (no LIN level before the
    products) and D's are valid on exit (six more outputs of the closing LIN level).
    Granger-Scott squaring in the cyclotomic subgroup, D = A^2 (tower.cuh fq12_cyc_sqr):
       pairs (a, b) = (g_k, g_{k+3}): t0 = a^2 + xi b^2, ab = a b;  r0 = 3 t0(0) - 2 g0, r3 = 6 ab(0) + 2 g3,
       r1 = 6 xi ab(2) + 2 g1, r4 = 3 t0(2) - 2 g4, r2 = 3 t0(1) - 2 g2, r5 = 6 ab(1) + 2 g5"""
    p = Program("CSQRX" if fused else "CSQR")
    ops = []
    if not fused:
        for k in range(3):
            ops += xi_lin(reg2("A", k + 3, True), reg2("A", k + 3))
        p.lin(ops)
    ops = []
    for k in range(3):
        a, b, xb = reg2("A", k), reg2("A", k + 3), reg2("A", k + 3, True)
        for c in range(2):
            ops.append((H(4 * k + c), fq2_mul_terms(a, a, c) + fq2_mul_terms(xb, b, c)))       # t0
            ops.append((H(4 * k + 2 + c), fq2_mul_terms(a, b, c)))                               # ab
    p.dot(4, ops)
    t0 = lambda k: (H(4 * k), H(4 * k + 1))
    ab = lambda k: (H(4 * k + 2), H(4 * k + 3))
    ops = []
    for c in range(2):
        ops.append((val("D", 0, c), [(3, t0(0)[c]), (-2, val("A", 0, c))]))
        ops.append((val("D", 3, c), [(6, ab(0)[c]), (2, val("A", 3, c))]))
        ops.append((val("D", 4, c), [(3, t0(2)[c]), (-2, val("A", 4, c))]))
        ops.append((val("D", 2, c), [(3, t0(1)[c]), (-2, val("A", 2, c))]))
        ops.append((val("D", 5, c), [(6, ab(1)[c]), (2, val("A", 5, c))]))
    x = xi_lin((val("D", 1, 0), val("D", 1, 1)), ab(2), 6)
    ops.append((x[0][0], x[0][1] + [(2, val("A", 1, 0))]))
    ops.append((x[1][0], x[1][1] + [(2, val("A", 1, 1))]))
    if fused:
        ops += xi_of_sum(reg2("D", 3, True), [(6, ab(0)), (2, reg2("A", 3))])
        ops += xi_of_sum(reg2("D", 4, True), [(3, t0(2)), (-2, reg2("A", 4))])
        ops += xi_of_sum(reg2("D", 5, True), [(6, ab(1)), (2, reg2("A", 5))])
    p.lin(ops)
    return p


def build_sparse():
    """D = A * (l0 + l1 w + l3 w^3) with the line in register B as the line kernels store it (10 slots):
       l0 | l1 | xi l1 | l3 | xi l3.   c_k = g_k l0 + g_{k-1} l1 [xi l1 if k < 1] + g_{k-3} l3 [xi l3 if k < 3]"""
    p = Program("SPARSE")
    line = lambda j: (("B", 2 * j), ("B", 2 * j + 1))
    ops = []
    for k in range(6):
        for c in range(2):
            terms = fq2_mul_terms(reg2("A", k), line(0), c)
            terms += fq2_mul_terms(reg2("A", (k + 5) % 6), line(2 if k < 1 else 1), c)
            terms += fq2_mul_terms(reg2("A", (k + 3) % 6), line(4 if k < 3 else 3), c)
            ops.append((val("D", k, c), terms))
    p.dot(6, ops)
    return p


def build_frob(kf):
    p = Program("FROB%d" % kf)
    ops = []
    for i in range(6):
        g = (GAM(kf, i, 0), GAM(kf, i, 1))
        for c in range(2):
            ops.append((val("D", i, c), fq2_mul_terms(reg2("A", i), g, c, conj_a=bool(kf & 1))))
    p.dot(2, ops)
    return p


def build_conj():
    p = Program("CONJ")
    p.lin([(val("D", k, c), [(-1 if k & 1 else 1, val("A", k, c))]) for k in range(6) for c in range(2)])
    return p


def build_copy():
    p = Program("COPY")
    p.lin([(val("D", k, c), [(1, val("A", k, c))]) for k in range(6) for c in range(2)])
    return p


def build_inv12():
    """D = A^-1 with B as a scratch register:  T = conj(A), n = A T in Fq6 = span{1, w^2, w^4}, n^-1 by the cubic-extension
    adjugate (one Fq inversion), D = T n^-1"""
    p = Program("INV12")
    p.lin([(val("B", k, c), [(-1 if k & 1 else 1, val("A", k, c))]) for k in range(6) for c in range(2)])
    emit_mul12(p, "D", "A", "B")
    n = [reg2("D", 0), reg2("D", 2), reg2("D", 4)]
    pr = lambda j: (H(2 * j), H(2 * j + 1))
    # products: 0 n0^2, 1 n1 n2, 2 n2^2, 3 n0 n1, 4 n1^2, 5 n0 n2
    pairs = [(0, 0), (1, 2), (2, 2), (0, 1), (1, 1), (0, 2)]
    p.dot(2, [(pr(j)[c], fq2_mul_terms(n[x], n[y], c)) for j, (x, y) in enumerate(pairs) for c in range(2)])
    t = [(H(12), H(13)), (H(14), H(15)), (H(16), H(17))]
    ops = []
    x = xi_lin(t[0], pr(1), -1)                                          # t0 = n0^2 - xi n1 n2
    ops += [(x[0][0], x[0][1] + [(1, pr(0)[0])]), (x[1][0], x[1][1] + [(1, pr(0)[1])])]
    x = xi_lin(t[1], pr(2), 1)                                           # t1 = xi n2^2 - n0 n1
    ops += [(x[0][0], x[0][1] + [(-1, pr(3)[0])]), (x[1][0], x[1][1] + [(-1, pr(3)[1])])]
    ops += [(t[2][c], [(1, pr(4)[c]), (-1, pr(5)[c])]) for c in range(2)]  # t2 = n1^2 - n0 n2
    p.lin(ops)
    n0t0, s = (H(18), H(19)), (H(20), H(21))
    p.dot(4, [(n0t0[c], fq2_mul_terms(n[0], t[0], c)) for c in range(2)]
          + [(s[c], fq2_mul_terms(n[2], t[1], c) + fq2_mul_terms(n[1], t[2], c)) for c in range(2)])
    d = (H(22), H(23))
    x = xi_lin(d, s, 1)                                                  # d = n0 t0 + xi (n2 t1 + n1 t2)
    p.lin([(x[0][0], x[0][1] + [(1, n0t0[0])]), (x[1][0], x[1][1] + [(1, n0t0[1])])])
    p.dot(2, [(X(0), [(d[0], d[0], 0), (d[1], d[1], 0)])])               # norm = d0^2 + d1^2
    p.inv([(X(1), X(0))])
    dinv = (H(0), H(1))
    p.dot(2, [(dinv[0], [(d[0], X(1), 0)]), (dinv[1], [(d[1], X(1), 1)])])  # d^-1 = conj(d) / norm
    ops = []
    for j in range(3):
        for c in range(2):
            ops.append((val("D", 2 * j, c), fq2_mul_terms(t[j], dinv, c)))     # n^-1 (even coefficients)
            ops.append((val("D", 2 * j + 1, c), [(ZERO, ZERO, 0)]))            # odd coefficients = 0
    p.dot(2, ops)
    emit_mul12(p, "D", "B", "D")
    return p


# ---------------------------------------------------------------------------------------------------------------------
# integer interpreter + self-check
# ---------------------------------------------------------------------------------------------------------------------
class Machine:
    def __init__(self):
        self.g = [0] * 64
        for k in (1, 2, 3):
            for i in range(6):
                for c in range(2):
                    self.g[GAM(k, i, c)[1]] = m.GAMMA[k][i][c]
        self.regs = {}

    def reg(self, name):
        return self.regs.setdefault(name, [0] * 24)

    def set12(self, name, v):
        r = self.reg(name)
        for k in range(6):
            r[2 * k], r[2 * k + 1] = v[k]

    def get12(self, name):
        r = self.reg(name)
        return [(r[2 * k], r[2 * k + 1]) for k in range(6)]

    def run(self, prog, D, A, B=None):
        bank = {"G": self.g, "D": self.reg(D), "A": self.reg(A), "B": self.reg(B) if B else None}
        rd = lambda s: bank[s[0]][s[1]]
        for kind, ops in prog.levels:
            new = []
            for op in ops:
                if kind.startswith("DOT"):
                    dst, terms = op
                    v = sum((-1 if neg else 1) * rd(a) * rd(b) for a, b, neg in terms)
                elif kind == "LIN":
                    dst, terms = op
                    v = sum(c * rd(s) for c, s in terms)
                else:
                    dst, src = op
                    v = pow(rd(src), P - 2, P)
                new.append((dst, v % P))
            for dst, v in new:
                bank[dst[0]][dst[1]] = v


def self_check(progs):
    rng = random.Random(3)
    rnd12 = lambda: [(rng.randrange(P), rng.randrange(P)) for _ in range(6)]
    mc = Machine()
    a, b = rnd12(), rnd12()
    mc.set12("R1", a); mc.set12("R2", b)
    mc.run(progs["MUL12"], "R0", "R1", "R2")
    assert mc.get12("R0") == m.f12_mul(a, b)
    mc.run(progs["MUL12"], "R1", "R1", "R1")                       # aliasing: in-place square
    assert mc.get12("R1") == m.f12_sqr(a)
    for k in (1, 2, 3):
        mc.set12("R1", a)
        mc.run(progs["FROB%d" % k], "R0", "R1")
        assert mc.get12("R0") == m.f12_frob(a, k)
    mc.run(progs["CONJ"], "R0", "R1"); assert mc.get12("R0") == m.f12_conj(a)
    mc.run(progs["COPY"], "R3", "R1"); assert mc.get12("R3") == a
    mc.run(progs["INV12"], "R0", "R1", "R4"); assert mc.get12("R0") == m.f12_inv(a)
    mc.set12("R5", a); mc.run(progs["INV12"], "R5", "R5", "R4"); assert mc.get12("R5") == m.f12_inv(a)   # in place
    l0, l1, l3 = (rng.randrange(P), rng.randrange(P)), (rng.randrange(P), rng.randrange(P)), (rng.randrange(P), rng.randrange(P))
    line = mc.reg("R6")
    for j, v in enumerate((l0, l1, m.f2_mul(m.XI, l1), l3, m.f2_mul(m.XI, l3))):
        line[2 * j], line[2 * j + 1] = v
    mc.set12("R1", a)
    mc.run(progs["SPARSE"], "R1", "R1", "R6")                      # in place
    assert mc.get12("R1") == m.f12_mul(a, [l0, l1, m.F2_ZERO, l3, m.F2_ZERO, m.F2_ZERO])
    # cyclotomic element: easy part of the final exponentiation of a random element
    t = m.f12_mul(m.f12_conj(a), m.f12_inv(a))
    cyc = m.f12_mul(m.f12_frob(t, 2), t)
    mc.set12("R1", cyc)
    mc.run(progs["CSQR"], "R1", "R1")
    assert mc.get12("R1") == m.f12_sqr(cyc)
    # exponentiation chain links: XI6 / COPYX once, then CSQRX / MUL12X without LIN levels of their own
    mc.set12("R7", cyc)
    mc.run(progs["XI6"], "R7", "R7")
    mc.run(progs["COPYX"], "R8", "R7")
    want = cyc
    for step in range(9):
        mc.run(progs["CSQRX"], "R8", "R8")
        want = m.f12_sqr(want)
        assert mc.get12("R8") == want, step
        if step % 3 != 1:
            mc.run(progs["MUL12X"], "R8", "R8", "R7")
            want = m.f12_mul(want, cyc)
            assert mc.get12("R8") == want, step
    assert mc.get12("R7") == cyc
    # generic power chain: MUL12Y squares in place and multiplies by a table entry prepared once by XI6
    mc.set12("R1", a)
    mc.run(progs["XI6"], "R1", "R1")
    mc.run(progs["COPYX"], "R2", "R1")
    want = a
    for step in range(7):
        mc.run(progs["MUL12Y"], "R2", "R2", "R2")
        want = m.f12_sqr(want)
        assert mc.get12("R2") == want, step
        if step % 2 == 0:
            mc.run(progs["MUL12Y"], "R2", "R2", "R1")
            want = m.f12_mul(want, a)
            assert mc.get12("R2") == want, step
    mc.run(progs["MUL12Y"], "R3", "R2", "R1")                      # not in place
    assert mc.get12("R3") == m.f12_mul(want, a) and mc.get12("R1") == a
    return True


TYPE_CODE = {"DOT2": 0, "LIN": 1, "DOT4": 2, "DOT6": 3, "INV": 4}


def enc_slot(s):
    bank, off = s
    assert 0 <= off < 64
    return BANKS[bank] << 6 | off


def emit(progs, path):
    order = ["MUL12", "CSQR", "FROB1", "FROB2", "FROB3", "CONJ", "COPY", "INV12", "SPARSE", "MUL12X", "CSQRX", "XI6", "COPYX", "MUL12Y"]
    code, types, index = [], [], {}
    dump = ("G", 63)  # idle lanes write X(2), which no program reads
    for name in order:
        prog = progs[name]
        index[name] = (len(types), len(prog.levels))
        for kind, ops in prog.levels:
            types.append(TYPE_CODE[kind])
            for lane in range(LANES):
                b = [0] * 16
                if lane >= len(ops):
                    b[0] = enc_slot(dump)
                    if kind == "INV":
                        b[14] = 1  # skip flag
                elif kind.startswith("DOT"):
                    dst, terms = ops[lane]
                    b[0] = enc_slot(dst)
                    mask = 0
                    for t in range(6):
                        if t < len(terms):
                            a, bb, neg = terms[t]
                            b[1 + 2 * t], b[2 + 2 * t] = enc_slot(a), enc_slot(bb)
                            mask |= (1 if neg else 0) << t
                    b[13] = mask
                elif kind == "LIN":
                    dst, terms = ops[lane]
                    b[0] = enc_slot(dst)
                    for t in range(4):
                        if t < len(terms):
                            c, s = terms[t]
                            b[1 + t] = enc_slot(s)
                            b[8 + 2 * t], b[9 + 2 * t] = c & 0xFF, (c >> 8) & 0xFF
                else:
                    dst, src = ops[lane]
                    b[0], b[1] = enc_slot(dst), enc_slot(src)
                w = [b[4 * i] | b[4 * i + 1] << 8 | b[4 * i + 2] << 16 | b[4 * i + 3] << 24 for i in range(4)]
                code.append(tuple(w))
    out = ["// GENERATED by tools/gen_fq12_programs.py -- do not edit.  Lane-parallel Fq12 programs for the 32-lane machine.",
           "#pragma once", "#include <stdint.h>",
           "#define SIPP_F12_LANES %d" % LANES, "#define SIPP_F12_LEVELS %d" % len(types),
           "#define SIPP_F12_REG_SLOTS 24", "#define SIPP_F12_GLOBAL_SLOTS 64", "#define SIPP_F12_GAMMA_SLOT0 1"]
    for name in order:
        out.append("#define SIPP_F12_%s_FIRST %d" % (name, index[name][0]))
        out.append("#define SIPP_F12_%s_LEVELS %d" % (name, index[name][1]))
    out.append("// level type: 0 = DOT2, 1 = LIN, 2 = DOT4, 3 = DOT6, 4 = INV")
    out.append("#define SIPP_F12_TYPES_INIT { " + ", ".join(str(t) for t in types) + " }")
    out.append("// 16 bytes per (level, lane): [0] dst; DOT: [1..12] a0 b0 .. a5 b5, [13] sign mask; LIN: [1..4] sources, [8..15] four int16")
    out.append("// coefficients; INV: [1] source, [14] skip.  A slot byte is bank << 6 | offset (bank 0 globals, 1 D, 2 A, 3 B).")
    out.append("#define SIPP_F12_CODE_INIT { \\")
    for i in range(0, len(code), 4):
        out.append("  " + " ".join("{0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}," % w for w in code[i:i + 4]) + " \\")
    out.append("}")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    return len(types)


def main():
    progs = {"MUL12": build_mul12(), "CSQR": build_csqr(), "FROB1": build_frob(1), "FROB2": build_frob(2), "FROB3": build_frob(3),
             "CONJ": build_conj(), "COPY": build_copy(), "INV12": build_inv12(), "SPARSE": build_sparse(),
             "MUL12X": build_mul12x(), "CSQRX": build_csqr(fused=True), "XI6": build_xi6(), "COPYX": build_copyx(), "MUL12Y": build_mul12y()}
    if "--check" in sys.argv:
        self_check(progs)
        print("self-check against the model: ok")
    path = os.path.join(ROOT, "sipp_b200", "csrc", "fq12_programs.h")
    n = emit(progs, path)
    for name, p in progs.items():
        print("%-6s levels %s" % (name, " ".join("%s:%d" % (k, len(o)) for k, o in p.levels)))
    print("wrote %s: %d levels" % (path, n))


if __name__ == "__main__":
    main()
