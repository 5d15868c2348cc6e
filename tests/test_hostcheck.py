"""The kernels' ALGORITHMS on the CPU: tests/hostcheck compiles the device headers (sipp_b200/csrc/*.cuh) for the host
with g++ -- the PTX carry chains replaced by the plain-C++ emulation in fq.cuh, warp shuffles by array reads -- and every
layer is compared with the oracle.  This is test infrastructure only (the product never computes on the host); it lets
algorithm changes be checked without a GPU, the `-m gpu` tests then check the real kernels."""
import ctypes
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "tests", "hostcheck")
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def le(v): return v.to_bytes(32, "little")


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC, "hostcheck.so")
    srcs = [os.path.join(HC, "hostcheck.cpp"), os.path.join(ROOT, "sipp_b200", "csrc", "glv.cc")]
    deps = srcs + [os.path.join(HC, "hostcheck_coop.inc")] + [os.path.join(ROOT, "sipp_b200", "csrc", f)
                                                               for f in os.listdir(os.path.join(ROOT, "sipp_b200", "csrc")) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", so] + srcs)
    return ctypes.CDLL(so)


def _buf(n): return ctypes.create_string_buffer(n)


def test_fq_and_dot(hc, oracle):
    rng = random.Random(1)
    edge = [0, 1, 2, P - 1, P - 2, 2**256 % P, (P - 1) // 2, 2**32 - 1]
    xs = edge + [rng.randrange(P) for _ in range(300)]
    ys = list(reversed(edge)) + [rng.randrange(P) for _ in range(300)]
    a, b = b"".join(map(le, xs)), b"".join(map(le, ys))
    for op, name in ((0, "FQ_MUL"), (1, "FQ_MUL"), (2, "FQ_ADD"), (3, "FQ_SUB")):
        out = _buf(len(a))
        assert hc.hc_fq_op(op, a, b, out, len(xs)) == 0
        assert out.raw == oracle.field_op(name, a, b)
    out = _buf(len(a))
    hc.hc_fq_op(4, a, None, out, len(xs))  # binary (Kaliski) inversion, all iteration counts incl. the edge values
    assert out.raw[:32 * 20] == oracle.field_op("FQ_INV", a[:32 * 20])
    assert out.raw == b"".join(le(pow(x, P - 2, P)) for x in xs)
    for n in (1, 2, 3, 4, 5, 6, 8, 12):
        for trial in range(40):
            big = trial == 0  # all operands p - 1: the largest lazy sum
            u = [P - 1 if big else rng.randrange(P) for _ in range(n)]
            v = [P - 1 if big else rng.randrange(P) for _ in range(n)]
            out = _buf(32)
            assert hc.hc_fq_dot(n, b"".join(map(le, u)), b"".join(map(le, v)), out) == 0
            assert int.from_bytes(out.raw, "little") == sum(x * y for x, y in zip(u, v)) % P


def test_fq12_and_coop(hc, oracle):
    rng = random.Random(2)
    a = b"".join(le(rng.randrange(P)) for _ in range(12))
    b = b"".join(le(rng.randrange(P)) for _ in range(12))
    out = _buf(384)
    for op, name in ((0, "FQ12_MUL"), (1, "FQ12_SQR"), (2, "FQ12_INV"), (3, "FQ12_FROB1"), (4, "FQ12_FROB2"), (5, "FQ12_FROB3"), (6, "FQ12_CONJ")):
        assert hc.hc_fq12_op(op, a, b if op == 0 else None, out) == 0
        assert out.raw == oracle.field_op(name, a, b if op == 0 else None), name
    hc.hc_coop_op(0, a, b, out)
    assert out.raw == oracle.field_op("FQ12_MUL", a, b)
    for op, name in ((3, "FQ12_FROB1"), (4, "FQ12_FROB2"), (5, "FQ12_FROB3"), (6, "FQ12_CONJ")):
        hc.hc_coop_op(op, a, None, out)
        assert out.raw == oracle.field_op(name, a)
    A, B = oracle.seeded_inputs(3, 1)
    gt = oracle.pairing(A, B)
    hc.hc_coop_op(1, gt, None, out)
    assert out.raw == oracle.field_op("FQ12_SQR", gt)
    hc.hc_fq12_op(7, gt, None, out)
    assert out.raw == oracle.field_op("FQ12_SQR", gt)


def test_pairing_and_inner_product(hc, oracle):
    A, B = oracle.seeded_inputs(21, 3)
    out = _buf(384)
    hc.hc_pairing(A[:64], B[:128], out, 0)
    assert out.raw == oracle.pairing(A[:64], B[:128])
    hc.hc_inner_product(A, B, 3, out)
    assert out.raw == oracle.inner_product(A, B)
    out2 = _buf(384)
    assert hc.hc_inner_product_machine(A, B, 3, out2) == 0  # lines -> 32-lane accumulate (shared squarings) -> machine FE
    assert out2.raw == out.raw


def test_fold_split(hc, oracle):
    """the lane-split GLV / GLS fold == the oracle's plain double-and-add fold, incl. exceptional points and scalars"""
    rng = random.Random(5)
    A, B = oracle.seeded_inputs(33, 4)
    a1, a2, b1, b2 = A[:64], A[64:128], B[:128], B[128:256]
    scalars = [1, 2, 3, R - 1, R - 2, 6 * 4965661367192848881**2] + [rng.randrange(1, R) for _ in range(6)]
    for k in scalars:
        x = le(k)
        cases1 = [(a1, a2), (bytes(64), a2), (a1, bytes(64)), (oracle.g1_mul(a2, x), a2), (oracle.g1_mul(a2, le(R - k)), a2)]
        for p1, p2 in cases1:
            for fn in (hc.hc_fold_split_g1, hc.hc_fold_wide_g1, hc.hc_fold_straus_g1):  # per-thread components / lane engine / shared doublings
                out = _buf(64)
                assert fn(p1, p2, x, x, out) == 0
                assert out.raw == oracle.fold_g1(p1 + p2, x)
        cases2 = [(b1, b2), (bytes(128), b2), (b1, bytes(128)), (oracle.g2_mul(b2, x), b2), (oracle.g2_mul(b2, le(R - k)), b2)]
        for p1, p2 in cases2:
            for fn in (hc.hc_fold_split_g2, hc.hc_fold_wide_g2, hc.hc_fold_straus_g2):
                out = _buf(128)
                assert fn(p1, p2, x, x, out) == 0
                assert out.raw == oracle.fold_g2(p1 + p2, x)


def test_line_engine(hc, oracle):
    """lane-parallel line programs (engine.cuh + line_programs.h, the code k_lines_wide runs): fq_lincomb4 against plain
    integers, and the whole Miller loop through the 16-lane schedule == the oracle's pairing"""
    rng = random.Random(9)
    for trial in range(200):
        if trial < 4:
            s = [P - 1] * 4
            c = [[2047, 0, 0, 0], [-2047, 0, 0, 0], [512, 512, 512, 511], [-512, -512, -512, -511]][trial]
        else:
            s = [rng.choice([0, 1, P - 1, rng.randrange(P)]) for _ in range(4)]
            c = [rng.randrange(-511, 512) for _ in range(4)]
        out = _buf(32)
        hc.hc_lincomb4(b"".join(map(le, s)), (ctypes.c_int * 4)(*c), out)
        assert int.from_bytes(out.raw, "little") == sum(x * y for x, y in zip(s, c)) % P, (s, c)
    A, B = oracle.seeded_inputs(77, 3)
    for i in range(3):
        out = _buf(384)
        assert hc.hc_pairing_wide(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128], out) == 0
        assert out.raw == oracle.pairing(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128])


def test_line_programs_generator_is_current():
    """line_programs.h is what tools/gen_line_programs.py emits, and its self-check against the pure-Python model passes"""
    path = os.path.join(ROOT, "sipp_b200", "csrc", "line_programs.h")
    before = open(path).read()
    subprocess.check_call([os.environ.get("PYTHON", "python"), os.path.join(ROOT, "tools", "gen_line_programs.py"), "--check"],
                          stdout=subprocess.DEVNULL)
    assert open(path).read() == before


def test_fq12_machine_final_exp(hc, oracle):
    """final exponentiation through the generated 32-lane programs (the code k_reduce_fe_eng runs) == the generic device
    code == the oracle's pairing value, for both normalisations"""
    hc.hc_final_exp.restype = ctypes.c_long
    rng = random.Random(11)
    for trial in range(3):
        f = b"".join(le(rng.randrange(P)) for _ in range(12))
        for ark in (0, 1):
            a, b = _buf(384), _buf(384)
            hc.hc_final_exp(0, f, a, ark)
            levels = hc.hc_final_exp(1, f, b, ark)
            assert a.raw == b.raw and levels > 500
    subprocess.check_call([os.environ.get("PYTHON", "python"), os.path.join(ROOT, "tools", "gen_fq12_programs.py"), "--check"], stdout=subprocess.DEVNULL)


def test_fq12_machine_generic_power(hc):
    """the verifier's Z_L^x (verifier_native.rs:59-61) through the MUL12Y chain links (the code k_gt_fold_eng runs) == plain
    square-and-multiply in the pure-Python model, on elements OUTSIDE the cyclotomic subgroup, incl. edge exponents"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import sipp_model as m
    hc.hc_pow_machine.restype = ctypes.c_long
    rng = random.Random(12)
    for k in (0, 1, 2, 3, 4, R - 1, rng.randrange(R), rng.randrange(R), (1 << 253) + 1, 3 << 200):
        f = [(rng.randrange(P), rng.randrange(P)) for _ in range(6)]
        out = _buf(384)
        levels = hc.hc_pow_machine(m.f12_bytes(f), le(k), out)
        want = list(m.F12_ONE)
        for bit in bin(k)[2:] if k else "":
            want = m.f12_sqr(want)
            if bit == "1":
                want = m.f12_mul(want, f)
        assert out.raw == m.f12_bytes(want), hex(k)
        assert levels <= 2 * (256 + 128) + 8


def test_matrix_fold_entry(hc):
    """k_mat_fold's entry formula on the emulated machines against the Python model: E'[i][j] = e(A_i + x A_{i+h}, B_j + x^-1 B_{j+h})
    (prover_native.rs:60-69 carried over to GT by bilinearity), incl. x = 1, x = r - 1 and an identity point"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import sipp_model as m
    hc.hc_mat_fold_entry.restype = ctypes.c_long
    rng = random.Random(21)
    A, B = m.seeded_inputs(11, 2)
    for x in (rng.randrange(1, R), 1, R - 1, rng.randrange(1, R)):
        xi = pow(x, -1, R)
        a = [A[0], None if x == 1 else A[1]]
        e = [[m.pairing(a[i], B[j]) for j in range(2)] for i in range(2)]
        want = m.pairing(m.g1_add(a[0], m.g1_mul(a[1], x)), m.g2_add(B[0], m.g2_mul(B[1], xi)))
        out = _buf(384)
        levels = hc.hc_mat_fold_entry(m.f12_bytes(e[0][0]), m.f12_bytes(e[1][1]), m.f12_bytes(e[1][0]), m.f12_bytes(e[0][1]), le(x), le(xi), out)
        assert 0 <= levels <= 2 * (67 + 34) + 8
        assert out.raw == m.f12_bytes(want), hex(x)
