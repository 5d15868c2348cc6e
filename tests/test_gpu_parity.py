"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.
Bit-exact at every level: Fq -> tower -> pairing -> inner products -> folds -> whole proofs -> verifier."""
import ctypes
import hashlib
import random

import pytest

pytestmark = pytest.mark.gpu

H = bytes.fromhex
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ONE12 = (1).to_bytes(32, "little") + bytes(352)


def le(v): return v.to_bytes(32, "little")


@pytest.fixture(scope="module")
def sipp():
    import sipp_b200
    from sipp_b200 import _lib
    _lib.require_gpu_once()
    return sipp_b200


@pytest.fixture(scope="module")
def lib(sipp):
    from sipp_b200 import _lib
    return _lib.load()


def _fq_op(lib, op, a, b=None):
    out = ctypes.create_string_buffer(len(a))
    assert lib.sipp_test_fq_op(op, a, b, out, len(a) // 32) == 0, lib.sipp_last_error()
    return out.raw


def _fq12_op(lib, op, a, b=None):
    out = ctypes.create_string_buffer(len(a))
    assert lib.sipp_test_fq12_op(op, a, b, out, len(a) // 384) == 0, lib.sipp_last_error()
    return out.raw


def test_fq_ops(lib, oracle):
    """K0: 254-bit Montgomery arithmetic, PTX carry chains and the portable version, vs the oracle"""
    rng = random.Random(1)
    edge = [0, 1, 2, P - 1, P - 2, 2**256 % P, 2**255 % P, (P - 1) // 2, 2**32 - 1, 2**224]
    xs = edge + [rng.randrange(P) for _ in range(100000)]
    ys = list(reversed(edge)) + [rng.randrange(P) for _ in range(100000)]
    a, b = b"".join(map(le, xs)), b"".join(map(le, ys))
    want = oracle.field_op("FQ_MUL", a, b)
    assert _fq_op(lib, 0, a, b) == want
    assert _fq_op(lib, 1, a, b) == want
    assert _fq_op(lib, 2, a, b) == oracle.field_op("FQ_ADD", a, b)
    assert _fq_op(lib, 3, a, b) == oracle.field_op("FQ_SUB", a, b)
    assert _fq_op(lib, 4, a[:32 * 500]) == oracle.field_op("FQ_INV", a[:32 * 500])
    assert _fq_op(lib, 5, a) == b"".join(le((-x) % P) for x in xs)


def test_fq12_ops(lib, oracle, golden):
    rng = random.Random(2)
    n = 64
    a = b"".join(le(rng.randrange(P)) for _ in range(12 * n))
    b = b"".join(le(rng.randrange(P)) for _ in range(12 * n))
    assert _fq12_op(lib, 0, a, b) == oracle.field_op("FQ12_MUL", a, b)
    assert _fq12_op(lib, 1, a) == oracle.field_op("FQ12_SQR", a)
    assert _fq12_op(lib, 2, a) == oracle.field_op("FQ12_INV", a)
    assert _fq12_op(lib, 3, a) == oracle.field_op("FQ12_FROB1", a)
    assert _fq12_op(lib, 4, a) == oracle.field_op("FQ12_FROB2", a)
    assert _fq12_op(lib, 5, a) == oracle.field_op("FQ12_FROB3", a)
    assert _fq12_op(lib, 6, a) == oracle.field_op("FQ12_CONJ", a)
    g = golden["fq12"]
    assert _fq12_op(lib, 0, H(g["a"]), H(g["b"])).hex() == g["mul"]
    assert _fq12_op(lib, 2, H(g["a"])).hex() == g["inv"]
    e = H(golden["pairing_gen"]["exact"])
    assert _fq12_op(lib, 7, e) == oracle.field_op("FQ12_SQR", e)  # cyclotomic squaring on a GT element


def test_coop_fq12_ops(lib, oracle, golden):
    """6-lane cooperative Fq12 arithmetic (lazy-reduction inner products + warp shuffles) vs the oracle"""
    rng = random.Random(12)
    n = 23  # not a multiple of the 5 groups per warp
    a = b"".join(le(rng.randrange(P)) for _ in range(12 * n))
    b = b"".join(le(rng.randrange(P)) for _ in range(12 * n))
    assert _fq12_op(lib, 20, a, b) == oracle.field_op("FQ12_MUL", a, b)
    assert _fq12_op(lib, 21, a) == oracle.field_op("FQ12_SQR", a)
    assert _fq12_op(lib, 22, a) == oracle.field_op("FQ12_INV", a)
    assert _fq12_op(lib, 23, a) == oracle.field_op("FQ12_FROB1", a)
    assert _fq12_op(lib, 24, a) == oracle.field_op("FQ12_FROB2", a)
    assert _fq12_op(lib, 25, a) == oracle.field_op("FQ12_FROB3", a)
    assert _fq12_op(lib, 26, a) == oracle.field_op("FQ12_CONJ", a)
    A, B = oracle.seeded_inputs(91, 7, threads=4)
    gt = b"".join(oracle.pairing(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128]) for i in range(7))
    assert _fq12_op(lib, 27, gt) == oracle.field_op("FQ12_SQR", gt)          # Granger-Scott on GT elements
    assert _fq12_op(lib, 28, gt) == _fq12_op(lib, 8, gt)                      # cooperative ^x == per-thread ^x
    ml = b"".join(oracle.miller_loop(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128]) for i in range(7))
    assert _fq12_op(lib, 29, ml) == gt                                        # cooperative final exponentiation


@pytest.fixture(params=[(1, 8192, 1), (1, 0, 0), (0, 0, 0)], ids=["split-engines", "split-thread-lines-coop-fe", "per-thread"])
def pipeline(request, sipp):
    """all Miller pipelines must be bit-exact: split = line kernel + cooperative accumulation (default) with the lines made by
    16 lanes per pair (k_lines_wide, the latency-bound rounds) or by one thread per pair (k_lines), and the final
    exponentiation on the 32-lane Fq12 machine or the 6-lane cooperative code; per-thread = one whole loop per thread
    (first-round baseline)"""
    from sipp_b200 import _lib
    sipp.set_option(_lib.OPT_PIPELINE, request.param[0])
    sipp.set_option(_lib.OPT_WIDE_LINES_MAX, request.param[1])
    sipp.set_option(_lib.OPT_FE_ENGINE, request.param[2])
    sipp.set_option(_lib.OPT_WIDE_FOLD_MAX, 1 << 20 if request.param[2] else 0)  # every fold on the engine / none
    sipp.set_option(_lib.OPT_WIDE_ACCUM_MAX, 8192 if request.param[2] else 0)
    yield request.param
    sipp.set_option(_lib.OPT_PIPELINE, 1)
    sipp.set_option(_lib.OPT_WIDE_LINES_MAX, 8192)
    sipp.set_option(_lib.OPT_FE_ENGINE, 1)
    sipp.set_option(_lib.OPT_WIDE_FOLD_MAX, 256)
    sipp.set_option(_lib.OPT_WIDE_ACCUM_MAX, 1536)


def test_pairing_golden_and_oracle(sipp, oracle, golden, pipeline):
    pg = golden["pairing_gen"]
    got = sipp.pairing(H(pg["a"]), H(pg["b"]))
    assert got.hex() == pg["exact"]
    assert hashlib.sha256(got).hexdigest() == "107999c8a16c357ce5236fdb7d765ed2904d57b2c31e6deca0c93063f8463ea2"
    for c in golden["pairing_random"]:
        assert sipp.pairing(H(c["a"]), H(c["b"])).hex() == c["exact"]
    A, B = oracle.seeded_inputs(17, 5)
    for i in range(5):
        assert sipp.pairing(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128]) == oracle.pairing(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128])
    # identity inputs contribute 1
    assert sipp.pairing(bytes(64), B[:128]) == ONE12 and sipp.pairing(A[:64], bytes(128)) == ONE12


def test_final_exp_normalisation_switch(sipp, golden):
    from sipp_b200 import _lib
    pg = golden["pairing_gen"]
    sipp.set_option(_lib.OPT_FE_NORMALISATION, 1)
    try:
        assert sipp.pairing(H(pg["a"]), H(pg["b"])).hex() == pg["ark"]
    finally:
        sipp.set_option(_lib.OPT_FE_NORMALISATION, 0)


def test_seeded_inputs(sipp, oracle):
    for seed, n in ((7, 2), (123, 33)):
        assert sipp.seeded_inputs(seed, n) == oracle.seeded_inputs(seed, n, threads=4)


@pytest.mark.parametrize("n", [1, 2, 3, 19, 20, 21, 63, 64, 65, 200, 1500])
def test_inner_product(sipp, oracle, n, pipeline):
    """prover_native.rs:15-23, incl. ragged (non power of two, partial block) sizes"""
    A, B = oracle.seeded_inputs(1000 + n, n, threads=4)
    assert sipp.inner_product(A, B) == oracle.inner_product(A, B, threads=4)


def test_inner_product_empty_and_mismatch(sipp):
    assert sipp.inner_product(b"", b"") == ONE12  # fold over an empty iterator starts at Fq12::one()
    with pytest.raises(AssertionError):
        sipp.inner_product(bytes(64), b"")  # assert_eq!(A.len(), B.len())


def test_inner_product_with_identities(sipp, oracle, pipeline):
    A, B = oracle.seeded_inputs(5, 8, threads=4)
    A = bytearray(A); B = bytearray(B)
    A[64:128] = bytes(64); B[128 * 5:128 * 6] = bytes(128)
    assert sipp.inner_product(bytes(A), bytes(B)) == oracle.inner_product(bytes(A), bytes(B))


@pytest.mark.parametrize("wide_fold", [1 << 20, 0], ids=["lane-engine", "per-thread-components"])
def test_fold_round(sipp, oracle, wide_fold):
    """prover_native.rs:60-74 via the round-granular context API, incl. exceptional points; both fold kernels"""
    from sipp_b200 import _lib
    sipp.set_option(_lib.OPT_WIDE_FOLD_MAX, wide_fold)
    rng = random.Random(3)
    A, B = oracle.seeded_inputs(31, 16, threads=4)
    A = bytearray(A); B = bytearray(B)
    x = le(rng.randrange(1, R)); xinv = oracle.fr_inverse(x)
    # exceptional cases: identity partner, identity base, a1 == x*a2 (doubling), a1 == -(x*a2) (result identity)
    A[64 * 8:64 * 9] = bytes(64)                       # A2[0] = O
    A[64 * 1:64 * 2] = bytes(64)                       # A1[1] = O
    A[64 * 2:64 * 3] = oracle.g1_mul(bytes(A[64 * 10:64 * 11]), x)
    A[64 * 3:64 * 4] = oracle.g1_mul(bytes(A[64 * 11:64 * 12]), le(R - int.from_bytes(x, "little")))
    B[128 * 8:128 * 9] = bytes(128)
    B[128 * 2:128 * 3] = oracle.g2_mul(bytes(B[128 * 10:128 * 11]), xinv)
    B[128 * 3:128 * 4] = oracle.g2_mul(bytes(B[128 * 11:128 * 12]), le(R - int.from_bytes(xinv, "little")))
    A, B = bytes(A), bytes(B)
    ctx = sipp.ProverContext(A, B)
    zl, zr = ctx.cross_products()
    assert zl == oracle.inner_product(A[64 * 8:], B[:128 * 8]) and zr == oracle.inner_product(A[:64 * 8], B[128 * 8:])
    ctx.fold(x, xinv)
    assert len(ctx) == 8
    a2, b2 = ctx.read()
    wa, wb = oracle.fold_g1(A, x), oracle.fold_g2(B, xinv)
    assert a2 == wa and b2 == wb
    assert wa[64 * 3:64 * 4] == bytes(64) and wb[128 * 3:128 * 4] == bytes(128)
    ctx.close()
    # ragged sizes (partial blocks of the 2- and 4-element layouts) and edge scalars
    A, B = oracle.seeded_inputs(32, 14, threads=4)
    for k in (1, 2, R - 1, rng.randrange(1, R)):
        x = le(k); xinv = oracle.fr_inverse(x)
        ctx = sipp.ProverContext(A, B)
        ctx.fold(x, xinv)
        assert ctx.read() == (oracle.fold_g1(A, x), oracle.fold_g2(B, xinv))
        ctx.close()
    sipp.set_option(_lib.OPT_WIDE_FOLD_MAX, 256)


def test_prove_golden(sipp, golden, pipeline):
    """whole proofs against the committed golden vectors (pure-Python model), incl. identity / doubling inputs"""
    for c in golden["prove"]:
        proof = sipp.sipp_prove_native(H(c["A"]), H(c["B"]))
        assert b"".join(proof).hex() == c["proof"], c["name"]
        assert len(proof) == 2 * (c["n"].bit_length() - 1) + 1
        st = sipp.sipp_verify_native(H(c["A"]), H(c["B"]), proof)
        assert st.final_A.hex() == c["final_A"] and st.final_B.hex() == c["final_B"] and st.final_Z.hex() == c["final_Z"]
        assert st.Z == proof[-1]


def test_round_api_matches_reference_structure(sipp, oracle, golden):
    """A host that keeps its own Transcript (as the Rust host would) and drives sipp_ctx_* per round reproduces
    prover_native.rs:26-80; every folded A/B and challenge is compared with the golden trace."""
    c = [x for x in golden["prove"] if x["name"] == "seed11_n16"][0]
    A, B = H(c["A"]), H(c["B"])
    n = c["n"]
    ctx = sipp.ProverContext(A, B)
    tr = sipp.Transcript()
    proof = []
    Z = ctx.inner_product()
    for i in range(n):
        tr.append_g1(A[64 * i:64 * i + 64]); tr.append_g2(B[128 * i:128 * i + 128])
    proof.append(Z); tr.append_fq12(Z)
    foldedA, foldedB, rnd = b"", b"", 0
    while n > 1:
        zl, zr = ctx.cross_products()
        proof.append(zl); tr.append_fq12(zl)
        proof.append(zr); tr.append_fq12(zr)
        x = tr.get_challenge()
        assert x.hex() == c["challenges"][rnd]
        ctx.fold(x, sipp.fr_inverse(x))
        a, b = ctx.read()
        foldedA += a; foldedB += b
        n //= 2; rnd += 1
    proof.reverse()
    assert b"".join(proof).hex() == c["proof"]
    assert foldedA.hex() == c["foldedA"] and foldedB.hex() == c["foldedB"]


def test_round_api_with_stages(sipp, oracle, golden):
    """the same host loop with sipp_ctx_set_stages on: a host that keeps its own transcript gets the pairing-matrix stages too
    (first stage from the inputs, look-ahead stages, tail) -- every Z, Z_L, Z_R and challenge as on the point-fold route; the
    folded points stay readable during look-ahead stages and are refused once the tail has begun"""
    for n, seed in ((16, None), (128, 1), (1024, 12)):
        if seed is None:
            c = [x for x in golden["prove"] if x["name"] == "seed11_n16"][0]
            A, B, want = H(c["A"]), H(c["B"]), bytes.fromhex(c["proof"])
        else:
            A, B = oracle.seeded_inputs(seed, n, threads=4)
            want = oracle.sipp_prove(A, B, threads=8)
        ctx = sipp.ProverContext(A, B)
        ctx.set_stages(True)
        tr = sipp.Transcript()
        proof = [ctx.inner_product()]
        for i in range(n):
            tr.append_g1(A[64 * i:64 * i + 64]); tr.append_g2(B[128 * i:128 * i + 128])
        tr.append_fq12(proof[0])
        m, refused = n, False
        while m > 1:
            zl, zr = ctx.cross_products()
            proof += [zl, zr]
            tr.append_fq12(zl); tr.append_fq12(zr)
            x = tr.get_challenge()
            ctx.fold(x, sipp.fr_inverse(x))
            m //= 2
            assert len(ctx) == m
            try:
                a, b = ctx.read()
                assert not refused and len(a) == 64 * m
            except sipp.SippError:
                refused = True
        assert refused                                   # the tail stage ran
        proof.reverse()
        assert b"".join(proof) == want, n
        ctx.close()


def test_sipp_native_n64(sipp, oracle):
    """mirror of the reference's own test_sipp_native (verifier_native.rs:96-106) + bit-exactness vs the oracle"""
    A, B = oracle.seeded_inputs(64, 64, threads=4)
    proof = sipp.sipp_prove_native(A, B)
    st = sipp.sipp_verify_native(A, B, proof)          # assert!(sipp_verify_native(&A, &B, &proof).is_ok())
    assert sipp.inner_product(A, B) == proof[-1]        # assert!(&inner_product(&A, &B) == proof.last().unwrap())
    want, tr = oracle.sipp_prove(A, B, threads=4, trace=True)
    assert b"".join(proof) == want
    ok, ost = oracle.sipp_verify(A, B, want, threads=4)
    assert ok and st.final_A == ost["final_A"] and st.final_B == ost["final_B"] and st.final_Z == ost["final_Z"]


def test_matrix_tail_thresholds(sipp, oracle):
    """the pairing-matrix tail (k_mat.cu: the last rounds fold E[i][j] = e(A_i, B_j) in GT instead of the points) must give the
    proof of the point-fold route whatever the round it takes over at -- incl. from the very first round (n <= threshold), with
    identity points on both sides, and under the arkworks final-exponentiation multiple"""
    from sipp_b200 import _lib
    A, B = oracle.seeded_inputs(77, 64, threads=4)
    A = bytearray(A); B = bytearray(B)
    A[64 * 5:64 * 6] = bytes(64)          # identity in G1
    B[128 * 41:128 * 42] = bytes(128)     # identity in G2
    A[64 * 9:64 * 10] = A[64 * 41:64 * 42]  # equal points in the two halves
    A, B = bytes(A), bytes(B)
    want = oracle.sipp_prove(A, B, threads=4)
    try:
        for thr in (0, 2, 4, 8, 16, 32, 64):
            sipp.set_option(_lib.OPT_MATRIX_TAIL, thr)
            assert b"".join(sipp.sipp_prove_native(A, B)) == want, thr
        # look-ahead stages: blocks of points as "virtual points" for log2(R) rounds, the points folded on a side stream
        for thr, bn, br in ((32, 64, 4), (8, 64, 8), (4, 64, 4), (2, 64, 32), (16, 32, 8), (0, 64, 8), (4, 16, 4)):
            sipp.set_option(_lib.OPT_MATRIX_TAIL, thr)
            sipp.set_option(_lib.OPT_MATRIX_BLOCK_N, bn)
            sipp.set_option(_lib.OPT_MATRIX_BLOCK_R, br)
            assert b"".join(sipp.sipp_prove_native(A, B)) == want, (thr, bn, br)
        sipp.set_option(_lib.OPT_MATRIX_BLOCK_N, 256)
        sipp.set_option(_lib.OPT_MATRIX_BLOCK_R, 8)
        # first stage (matrix over the inputs, Z from its diagonal) on / off, for sizes on both sides of every branch of its rule
        for first in (0, 1):
            sipp.set_option(_lib.OPT_MATRIX_FIRST, first)
            for thr in (32, 8, 2):
                sipp.set_option(_lib.OPT_MATRIX_TAIL, thr)
                for m in (64, 32, 16, 8, 4):
                    got = b"".join(sipp.sipp_prove_native(A[:64 * m], B[:128 * m]))
                    assert got == (want if m == 64 else oracle.sipp_prove(A[:64 * m], B[:128 * m], threads=4)), (first, thr, m)
        sipp.set_option(_lib.OPT_MATRIX_FIRST, 1)
        sipp.set_option(_lib.OPT_MATRIX_TAIL, 32)
        # stages with the 16-lane line engine switched off (found by tools/fuzz_single.py: the matrix build must not take its
        # throughput path for single-point blocks)
        sipp.set_option(_lib.OPT_WIDE_LINES_MAX, 0)
        for m in (64, 8):
            assert b"".join(sipp.sipp_prove_native(A[:64 * m], B[:128 * m])) == (want if m == 64 else oracle.sipp_prove(A[:64 * m], B[:128 * m], threads=4)), m
        sipp.set_option(_lib.OPT_WIDE_LINES_MAX, 8192)
        sipp.set_option(_lib.OPT_FE_NORMALISATION, 1)
        sipp.set_option(_lib.OPT_MATRIX_TAIL, 0)
        ref = b"".join(sipp.sipp_prove_native(A[:64 * 16], B[:128 * 16]))
        sipp.set_option(_lib.OPT_MATRIX_TAIL, 8)
        assert b"".join(sipp.sipp_prove_native(A[:64 * 16], B[:128 * 16])) == ref
    finally:
        sipp.set_option(_lib.OPT_FE_NORMALISATION, 0)
        sipp.set_option(_lib.OPT_WIDE_LINES_MAX, 8192)
        sipp.set_option(_lib.OPT_MATRIX_TAIL, 32)
        sipp.set_option(_lib.OPT_MATRIX_BLOCK_N, 256)
        sipp.set_option(_lib.OPT_MATRIX_BLOCK_R, 8)
        sipp.set_option(_lib.OPT_MATRIX_FIRST, 1)


def test_prove_n128_config0(sipp, oracle):
    """BASELINE configs[0]: n = 128, prove + verify, byte-equal to the CPU reference restatement"""
    A, B = oracle.seeded_inputs(1, 128, threads=4)
    proof = sipp.sipp_prove_native(A, B)
    assert b"".join(proof) == oracle.sipp_prove(A, B, threads=4)
    sipp.sipp_verify_native(A, B, proof)


def test_prove_n4096_config1(sipp, oracle):
    """BASELINE configs[1]: n = 2^12 on one GPU, bit-exact vs the oracle transcript"""
    A, B = sipp.seeded_inputs(2, 4096)
    proof = sipp.sipp_prove_native(A, B)
    assert b"".join(proof) == oracle.sipp_prove(A, B, threads=8)
    assert len(proof) == 25


def test_verify_rejects_tampering(sipp, oracle):
    A, B = oracle.seeded_inputs(9, 8, threads=4)
    proof = sipp.sipp_prove_native(A, B)
    bad = [bytearray(p) for p in proof]
    bad[3][40] ^= 1
    with pytest.raises(sipp.VerificationError):
        sipp.sipp_verify_native(A, B, [bytes(p) for p in bad])
    # wrong statement
    A2 = oracle.seeded_inputs(10, 8, threads=4)[0]
    with pytest.raises(sipp.VerificationError):
        sipp.sipp_verify_native(A2, B, proof)
    # short proof: proof.pop().unwrap() panics in the reference
    with pytest.raises(sipp.SippError) as ei:
        sipp.sipp_verify_native(A, B, proof[:-2])
    assert ei.value.code == -5


def test_error_behaviour(sipp, oracle):
    A, B = oracle.seeded_inputs(9, 4, threads=2)
    with pytest.raises(AssertionError):
        sipp.sipp_prove_native(A, B[:128 * 3])           # assert_eq!(A.len(), B.len())
    with pytest.raises(sipp.SippError):
        sipp.sipp_prove_native(A[:64 * 3], B[:128 * 3])  # n not a power of two
    bad = bytearray(A); bad[0:32] = le(P)                 # x coordinate == p: not canonical
    with pytest.raises(sipp.SippError) as ei:
        sipp.sipp_prove_native(bytes(bad), B)
    assert ei.value.code == -6


def test_partials_and_combine(sipp, oracle):
    """the multi-GPU building blocks on one GPU (what a host with its own exchange calls): un-exponentiated partial products in
    device memory, then sipp_combine_partials over two 'ranks' built from a strided split equals the full products"""
    import torch
    from sipp_b200.sharded import shard_points
    A, B = oracle.seeded_inputs(77, 32, threads=4)
    stream = torch.cuda.current_stream().cuda_stream
    parts0, parts1 = [], []
    for r in range(2):
        Al, Bl = shard_points(A, B, r, 2)
        ctx = sipp.ProverContext(Al, Bl)
        for which, dst in ((0, parts0), (1, parts1)):
            out = torch.empty((1 + which) * 384, dtype=torch.uint8, device="cuda")
            ctx.partial_products(which, out.data_ptr(), stream)
            dst.append(out)
    torch.cuda.synchronize()
    g0, g1 = torch.cat(parts0), torch.cat(parts1)
    z = sipp.combine_partials(g0.data_ptr(), 2, 1, stream)
    assert z[0] == oracle.inner_product(A, B, threads=4)
    zs = sipp.combine_partials(g1.data_ptr(), 2, 2, stream)
    assert zs[0] == oracle.inner_product(A[64 * 16:], B[:128 * 16], threads=4)
    assert zs[1] == oracle.inner_product(A[:64 * 16], B[128 * 16:], threads=4)


def test_inner_product_many_pairs_per_group(sipp, oracle):
    """enough pairs that every 6-lane group folds several pairs into one accumulator (shared squaring, kpg > 1)"""
    n = 20000
    A, B = sipp.seeded_inputs(6, 2048)
    A, B = (A * 10)[:64 * n], (B * 10)[:128 * n]
    assert sipp.inner_product(A, B) == oracle.inner_product(A, B, threads=16)


def test_roundtrip_property_large(sipp):
    """size-independent property at a size the oracle does not need to replay: prove -> verify succeeds and
    Z equals the stand-alone inner product (n = 2^13)"""
    n = 1 << 13
    A, B = sipp.seeded_inputs(4, n)
    proof = sipp.sipp_prove_native(A, B)
    assert len(proof) == 27
    assert sipp.inner_product(A, B) == proof[-1]
    sipp.sipp_verify_native(A, B, proof)


def test_bls_producers_against_oracle(sipp, oracle, golden):
    """the demo's input producers (bin/bls_aggregation.rs:95-117) one by one against the oracle's double-and-add: keygen over the
    generator's window table, variable-base signing, fixed-base G2 over a table built on the fly, the sum as a tree; edge scalars
    (0 -> identity, 1, r - 1, single bytes set), identity points, empty and ragged sizes"""
    from sipp_b200 import _lib
    rng = random.Random(23)
    g1, g2 = H(golden["pairing_gen"]["a"]), H(golden["pairing_gen"]["b"])
    ks = [0, 1, 2, R - 1, 255, 256, 1 << 248, (1 << 253) + 5] + [rng.randrange(R) for _ in range(200)]
    kb = b"".join(le(k) for k in ks)
    pks = sipp.g1_generator_mul_batch(kb)
    assert pks[0] == bytes(64) and pks[1] == g1
    assert pks == [oracle.g1_mul(g1, le(k)) for k in ks]
    # signing: one message point per key (variable base), messages = multiples of the generator with a few identities
    ms = [oracle.g2_mul(g2, le(rng.randrange(1, R))) for _ in range(37)]
    ms[5] = bytes(128)
    sk = [le(k) for k in ks[:37]]
    sigs = sipp.g2_mul_batch(ms, sk)
    assert sigs == [oracle.g2_mul(m, k) for m, k in zip(ms, sk)]
    # one base for every scalar: small counts walk the base, larger ones build its window table
    for cnt in (3, 208):
        got = sipp.g2_mul_batch(ms[7], kb[:32 * cnt])
        assert got == [oracle.g2_mul(ms[7], le(k)) for k in ks[:cnt]]
    # sum: sizes around the block / thread shape of the tree, identities inside, P + (-P)
    pts = sipp.g2_mul_batch(g2, b"".join(le(rng.randrange(1, R)) for _ in range(700)))

    def neg2(q):
        x, y0, y1 = q[:64], int.from_bytes(q[64:96], "little"), int.from_bytes(q[96:], "little")
        return x + le((P - y0) % P) + le((P - y1) % P)
    for cnt in (0, 1, 2, 127, 128, 129, 513, 700):
        sel = list(pts[:cnt])
        if cnt >= 129:
            sel[3] = bytes(128)
            sel[100] = neg2(sel[99])
        acc = None
        want = bytes(128)
        if sel:
            # oracle sum through fold_g2 with scalar 1: (B1 + 1 * B2) pairwise tree
            cur = list(sel)
            while len(cur) > 1:
                if len(cur) % 2:
                    cur.append(bytes(128))
                hlen = len(cur) // 2
                folded = oracle.fold_g2(b"".join(cur), le(1))
                cur = [folded[128 * i:128 * i + 128] for i in range(hlen)]
            want = cur[0]
        assert sipp.g2_sum(sel) == want, cnt
    assert sipp.g1_neg_generator() == le(1) + le(P - 2)
    # error behaviour: a scalar >= r, a coordinate >= p
    with pytest.raises(sipp.SippError) as ei:
        sipp.g1_generator_mul_batch(le(R))
    assert ei.value.code == _lib.ERR_ENCODING
    with pytest.raises(sipp.SippError) as ei:
        sipp.g2_sum(le(P) + bytes(96))
    assert ei.value.code == _lib.ERR_ENCODING


def test_bls_aggregation_demo(sipp, oracle, golden):
    """bin/bls_aggregation.rs:93-122 with the library's producers: 127 key pairs, 127 messages in G2, signatures, the aggregate,
    the 128-pair instance (pk_i, m_i) + (-G1, sigma): inner product == 1, prove, verify -- every step against the oracle.
    (The demo's messages come from hash-to-G2, which lives in an un-vendored crate; here m_i = [h_i] G2 for seeded h_i.)"""
    n = 128
    rng = random.Random(24)
    g2 = H(golden["pairing_gen"]["b"])
    sks = [le(rng.randrange(1, R)) for _ in range(n - 1)]                         # private_keys  :95
    pks = sipp.g1_generator_mul_batch(sks)                                        # public_keys   :96-99
    ms = sipp.g2_mul_batch(g2, [le(rng.randrange(1, R)) for _ in range(n - 1)])   # messages in G2 (stand-in for :100-104)
    sigs = sipp.g2_mul_batch(ms, sks)                                             # signatures    :105-109
    sigma = sipp.g2_sum(sigs)                                                     # aggregated    :110-113
    assert all(oracle.g2_on_curve(s) for s in sigs[:5]) and sigs[0] == oracle.g2_mul(ms[0], sks[0])
    A2 = b"".join(pks) + sipp.g1_neg_generator()                                  # a.push(-G1)   :116
    B2 = b"".join(ms) + sigma                                                     # b.push(sigma) :117
    assert sipp.inner_product(A2, B2) == ONE12                                    # assert_eq!(inner_product(&a, &b), Fq12::one())  :119
    proof = sipp.sipp_prove_native(A2, B2)                                        # :120
    st = sipp.sipp_verify_native(A2, B2, proof)                                   # :121
    assert sipp.pairing(st.final_A, st.final_B) == st.final_Z                     # :122
    assert proof[-1] == ONE12 and st.Z == ONE12
    assert b"".join(proof) == oracle.sipp_prove(A2, B2, threads=8)
    # a forged aggregate (one signature dropped) does not give the identity
    assert sipp.inner_product(A2, b"".join(ms) + sipp.g2_sum(sigs[1:])) != ONE12


def test_seeded_inputs_match_oracle_generator(sipp, oracle):
    """the synthetic inputs of every benchmark are keygen over the two generators' window tables: same bytes as the oracle's"""
    for seed, n in ((1, 1), (2, 77), (9, 1500)):
        assert sipp.seeded_inputs(seed, n) == oracle.seeded_inputs(seed, n, threads=8)


def test_fold_large_round_straus_vs_split(sipp, oracle):
    """a round large enough for the throughput fold (k_fold_straus, h >= 16384): same bytes as the lane-split kernel, a sample of
    elements against the oracle's plain double-and-add, exceptional points included"""
    n = 1 << 15
    rng = random.Random(8)
    A, B = sipp.seeded_inputs(12, n)
    x = le(rng.randrange(1, R)); xinv = oracle.fr_inverse(x)
    A = bytearray(A); B = bytearray(B)
    h = n // 2
    A[64 * 5:64 * 6] = bytes(64)                                   # A1[5] = O
    B[128 * (h + 7):128 * (h + 8)] = bytes(128)                    # B2[7] = O
    A[64 * 9:64 * 10] = oracle.g1_mul(bytes(A[64 * (h + 9):64 * (h + 10)]), x)                                   # doubling case
    B[128 * 11:128 * 12] = oracle.g2_mul(bytes(B[128 * (h + 11):128 * (h + 12)]), le(R - int.from_bytes(xinv, "little")))  # sum = O
    A, B = bytes(A), bytes(B)
    outs = []
    for straus in (1, 0):
        sipp.set_option(10, straus)
        try:
            ctx = sipp.ProverContext(A, B)
            ctx.fold(x, xinv)
            outs.append(ctx.read())
            ctx.close()
        finally:
            sipp.set_option(10, 1)
    assert outs[0] == outs[1]
    fa, fb = outs[0]
    assert fb[128 * 11:128 * 12] == bytes(128)
    for i in (0, 5, 7, 9, 11, 123, h - 1):
        assert fa[64 * i:64 * i + 64] == oracle.fold_g1(A[64 * i:64 * i + 64] + A[64 * (h + i):64 * (h + i) + 64], x), i
        assert fb[128 * i:128 * i + 128] == oracle.fold_g2(B[128 * i:128 * i + 128] + B[128 * (h + i):128 * (h + i) + 128], xinv), i


# twist points (on y^2 = x^3 + 3/xi) that are NOT in the order-r subgroup, made with the pure-Python model (a random twist point, and
# r times it: a point whose order divides the cofactor 2p - r)
TWIST_NOT_G2 = H("fbee9e35ab3c0ac619c62937bfe3caf537f37ffb5991752a801d56df83aea70a5de70f4aba744b50f26b23f62869ea32968c0a8a8e5449e0444c86ada8f60b28"
                 "b220740d239210aa74480d7e4ebe546a6b9f5fa2f8cf8c0a0a4580f18c2e612775f82b46ee8dc7599bf9c136ec784e777e061d68a3c8d110c06b2a0e0d572002")
TWIST_COFACTOR_ORDER = H("2b66479ac463b5b5f53b9cfd46efe028fe7bea566ff042edbbc48c198673c32a531c699076c6076499d44db1348fa2825f4f0673f09890f44761c3415a3db52e"
                         "e03c1427e3e064c6723f718785f986642fb9f61a77a8b4fe5d9c49913e5c95273c79ed1afc24de83175a47fc30a1e3aa56e552ca537aba0f361143c88cb49a23")


def test_point_validation(sipp, oracle):
    """what G1Affine::new / G2Affine::new assert when the reference's inputs are built: points off the curve and twist points outside
    the order-r subgroup are refused by every entry point that takes points (SIPP_ERR_ENCODING); valid points, the identity
    included, pass; SIPP_OPT_VALIDATE_POINTS = 0 restores trusted-input behaviour"""
    from sipp_b200 import _lib
    assert oracle.g2_on_curve(TWIST_NOT_G2) and oracle.g2_on_curve(TWIST_COFACTOR_ORDER)
    n = 8
    A, B = oracle.seeded_inputs(91, n)
    ident = bytearray(A), bytearray(B)
    ident[0][64:128] = bytes(64)
    ident[1][128 * 5:128 * 6] = bytes(128)
    assert b"".join(sipp.sipp_prove_native(bytes(ident[0]), bytes(ident[1]))) == oracle.sipp_prove(bytes(ident[0]), bytes(ident[1]))
    proof = sipp.sipp_prove_native(A, B)

    def expect_refused(a, b, what):
        for call in (lambda: sipp.sipp_prove_native(a, b), lambda: sipp.inner_product(a, b), lambda: sipp.sipp_verify_native(a, b, proof),
                     lambda: sipp.sipp_prove_native_batch(a, b, 4), lambda: sipp.sipp_verify_native_batch(a, b, 8, [proof])):
            with pytest.raises(sipp.SippError) as ei:
                call()
            assert ei.value.code == _lib.ERR_ENCODING and what in str(ei.value), str(ei.value)
    bad = bytearray(A)
    bad[64 * 3 + 32] ^= 1                                      # y of A_3: off y^2 = x^3 + 3
    expect_refused(bytes(bad), B, "not on the curve")
    bad = bytearray(B)
    bad[128 * 2 + 64] ^= 1                                     # y.c0 of B_2: off the twist
    expect_refused(A, bytes(bad), "not on the curve")
    for pt in (TWIST_NOT_G2, TWIST_COFACTOR_ORDER):
        bad = bytearray(B)
        bad[128 * 6:128 * 7] = pt
        expect_refused(A, bytes(bad), "subgroup")
    sipp.set_option(_lib.OPT_VALIDATE_POINTS, 0)
    try:
        bad = bytearray(B)
        bad[128 * 6:128 * 7] = TWIST_NOT_G2
        sipp.inner_product(A, bytes(bad))                      # trusted inputs: no check, no error (value unspecified)
    finally:
        sipp.set_option(_lib.OPT_VALIDATE_POINTS, 1)
    # 3,000 valid points in one launch (every block / warp shape of the kernel), none refused
    A3, B3 = sipp.seeded_inputs(92, 3000)
    assert sipp.inner_product(A3, B3) == oracle.inner_product(A3, B3, threads=16)


def test_verifier_refuses_non_canonical_proof(sipp, oracle):
    """a proof element with a coordinate c + p (< 2^256) reduces to the same field element but hashes differently; ark's
    deserialisation cannot produce it, so the verifier refuses it instead of accepting a malleated proof"""
    from sipp_b200 import _lib
    n = 4
    A, B = oracle.seeded_inputs(93, n)
    proof = sipp.sipp_prove_native(A, B)
    z = bytearray(proof[-1])
    c = int.from_bytes(z[:32], "little")
    assert c + P < 2**256
    z[:32] = (c + P).to_bytes(32, "little")
    bad = proof[:-1] + [bytes(z)]
    with pytest.raises(sipp.SippError) as ei:
        sipp.sipp_verify_native(A, B, bad)
    assert ei.value.code == _lib.ERR_ENCODING
    res = (ctypes.c_int * 2)()
    lib = _lib.load()
    assert lib.sipp_verify_native_batch(A + A, B + B, n, 2, b"".join(proof) + b"".join(bad), len(proof), res, None, None, None) == 0
    assert list(res) == [0, _lib.ERR_ENCODING]


def test_gt_fold_machine_matches_oracle(sipp, lib, oracle):
    """Z_L^x * Z * Z_R^(x^-1) (verifier_native.rs:59-61) on the two 32-lane machines (k_gt_fold_eng) == the one-thread-per-power
    kernel == the oracle's generic powers, on arbitrary Fq12 elements (not in the cyclotomic subgroup) and edge exponents"""
    from sipp_b200 import _lib
    rng = random.Random(17)
    for x in (1, 2, 3, R - 1, rng.randrange(1, R), rng.randrange(1, R), 1 << 253):
        xinv = pow(x, -1, R)
        zl, z, zr = (b"".join(le(rng.randrange(P)) for _ in range(12)) for _ in range(3))
        outs = []
        for eng in (1, 0):
            sipp.set_option(_lib.OPT_FE_ENGINE, eng)
            out = ctypes.create_string_buffer(384)
            try:
                assert lib.sipp_gt_fold(zl, z, zr, le(x), le(xinv), out) == 0
            finally:
                sipp.set_option(_lib.OPT_FE_ENGINE, 1)
            outs.append(out.raw)
        assert outs[0] == outs[1]

        def power(base, e):
            acc = ONE12
            for bit in bin(e)[2:]:
                acc = oracle.field_op("FQ12_SQR", acc)
                if bit == "1":
                    acc = oracle.field_op("FQ12_MUL", acc, base)
            return acc
        want = oracle.field_op("FQ12_MUL", oracle.field_op("FQ12_MUL", power(zl, x), z), power(zr, xinv))
        assert outs[0] == want
