"""Batched instances (BASELINE config "4096 independent n=128 SIPP instances"): lock-step proving of independent instances
with one Fiat-Shamir chain per instance on the device.  Bit-exact against the single-instance prover, the CPU oracle and
the golden proofs; the device transcript pieces are checked on their own against the host transcript."""
import ctypes
import random
import struct

import pytest

pytestmark = pytest.mark.gpu

GL_P = 0xFFFFFFFF00000001
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


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


def test_device_poseidon_matches_host(lib):
    """k_transcript.cu's textbook permutation == the host's sparse-matrix / AVX-512 permutation, incl. the upstream KAT inputs"""
    rng = random.Random(11)
    states = [[0] * 12, list(range(12)), [GL_P - 1] * 12] + [[rng.randrange(GL_P) for _ in range(12)] for _ in range(500)]
    states += [[rng.randrange(2**64) for _ in range(12)] for _ in range(100)]  # non-canonical representatives
    flat = [v for s in states for v in s]
    dev = (ctypes.c_uint64 * len(flat))(*flat)
    assert lib.sipp_test_poseidon_device(dev, len(states)) == 0, lib.sipp_last_error()
    for i, s in enumerate(states):
        h = (ctypes.c_uint64 * 12)(*s)
        lib.sipp_poseidon_permute(h)
        assert list(dev[12 * i:12 * i + 12]) == list(h), i


def test_device_transcript_round_matches_host(sipp, lib):
    """absorb (Z,) Z_L, Z_R -> challenge -> inverse -> fold plan on the device == host Transcript + sipp_fr_inverse + host recoding"""
    rng = random.Random(12)
    count = 70
    for nf in (2, 3):
        for order in (0, 1):
            sipp.set_option(2, order)
            try:
                states = [[rng.randrange(GL_P) for _ in range(4)] for _ in range(count)]
                fqs = [[b"".join(rng.randrange(P).to_bytes(32, "little") for _ in range(12)) for _ in range(nf)] for _ in range(count)]
                st = (ctypes.c_uint64 * (4 * count))(*[v for s in states for v in s])
                xs = ctypes.create_string_buffer(64 * count)
                nwords = 6 * 11 + 2
                plans = (ctypes.c_uint32 * (nwords * count))()
                raw = b"".join(b"".join(f) for f in fqs)
                assert lib.sipp_test_transcript_round_device(st, raw, nf, count, xs, plans) == 0, lib.sipp_last_error()
                for j in range(count):
                    t = sipp.Transcript()
                    for k in range(4):
                        t._t.state[k] = states[j][k]
                    for f in fqs[j]:
                        t.append_fq12(f)
                    x = t.get_challenge()
                    xi = sipp.fr_inverse(x)
                    assert xs.raw[64 * j:64 * j + 32] == x and xs.raw[64 * j + 32:64 * j + 64] == xi, (nf, order, j)
                    assert list(st[4 * j:4 * j + 4]) == t.state
                    buf = (ctypes.c_uint32 * 128)()
                    nw = lib.sipp_test_fold_plan(x, xi, buf, 128)
                    assert nw == nwords
                    assert list(plans[nwords * j:nwords * (j + 1)]) == list(buf[:nw]), (nf, order, j)
            finally:
                sipp.set_option(2, 0)


@pytest.mark.parametrize("n,count", [(2, 5), (8, 3), (16, 33), (128, 4)])
def test_batch_equals_single_and_oracle(sipp, oracle, n, count):
    """every instance of a batch == the single-instance GPU proof == the oracle's proof (config 5 at test size)"""
    A, B = oracle.seeded_inputs(500 + n, n * count, threads=8)
    proofs = sipp.sipp_prove_native_batch(A, B, n)
    assert len(proofs) == count
    for j in range(count):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        assert b"".join(proofs[j]) == b"".join(sipp.sipp_prove_native(a, b)), (n, j)
        if j < 2:
            assert b"".join(proofs[j]) == oracle.sipp_prove(a, b, threads=8), (n, j)
            sipp.sipp_verify_native(a, b, proofs[j])


def test_batch_golden(sipp, golden):
    """the committed golden proofs (independent pure-Python model) come out of the batched path too"""
    H = bytes.fromhex
    for c in golden["prove"]:
        A, B = H(c["A"]), H(c["B"])
        proofs = sipp.sipp_prove_native_batch(A * 3, B * 3, c["n"])
        assert len(proofs) == 3
        for p in proofs:
            assert b"".join(p).hex() == c["proof"], c["name"]


def test_batch_kpg_variants_and_identity(sipp, oracle):
    """accumulator-group sizes (pairs of one product sharing the squarings) do not change a bit; identity points contribute 1"""
    n, count = 32, 6
    A, B = oracle.seeded_inputs(77, n * count, threads=8)
    A = bytearray(A); B = bytearray(B)
    A[64 * 5:64 * 6] = bytes(64)            # identity in G1 (instance 0)
    B[128 * 40:128 * 41] = bytes(128)       # identity in G2 (instance 1)
    A, B = bytes(A), bytes(B)
    want = None
    for kpg in (1, 4, 32):
        sipp.set_option(9, kpg)
        try:
            got = sipp.sipp_prove_native_batch(A, B, n)
        finally:
            sipp.set_option(9, 32)
        if want is None:
            want = got
            for j in (0, 1, 5):
                a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
                assert b"".join(got[j]) == oracle.sipp_prove(a, b, threads=8)
        assert got == want, kpg
    sipp.set_option(12, 0)  # line coefficients of B recomputed for Z and the first Z_L / Z_R instead of shared
    try:
        assert sipp.sipp_prove_native_batch(A, B, n) == want
    finally:
        sipp.set_option(12, 1)
    sipp.set_option(10, 0)  # lane-split component fold instead of the shared-doubling (Straus) fold
    try:
        assert sipp.sipp_prove_native_batch(A, B, n) == want
    finally:
        sipp.set_option(10, 1)


def test_batch_errors(sipp, oracle):
    A, B = oracle.seeded_inputs(9, 8, threads=2)
    with pytest.raises(AssertionError):
        sipp.sipp_prove_native_batch(A, B[:128 * 4], 4)
    with pytest.raises(sipp.SippError):
        sipp.sipp_prove_native_batch(A[:64 * 6], B[:128 * 6], 3)   # n not a power of two
    bad = bytearray(A); bad[64 * 5:64 * 5 + 32] = P.to_bytes(32, "little")
    with pytest.raises(sipp.SippError) as ei:
        sipp.sipp_prove_native_batch(bytes(bad), B, 4)
    assert ei.value.code == -6


def test_batch_config5_sample(sipp, oracle):
    """BASELINE config 5 shape at a tenth of its size: 512 x n = 128; a sample of instances against the oracle, all Z against
    the stand-alone inner product of a sample, and the round trip through the verifier"""
    n, count = 128, 512
    A, B = sipp.seeded_inputs(5, n * count)
    proofs = sipp.sipp_prove_native_batch(A, B, n)
    assert len(proofs) == count and all(len(p) == 15 for p in proofs)
    for j in (0, 1, 255, 511):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        assert b"".join(proofs[j]) == oracle.sipp_prove(a, b, threads=16), j
    for j in (7, 300):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        assert proofs[j][-1] == sipp.inner_product(a, b)
        sipp.sipp_verify_native(a, b, proofs[j])


def test_verify_batch(sipp, oracle):
    """batched verifier (verifier_native.rs:14-85 per instance): statements equal the single-instance verifier's and the oracle's;
    a tampered proof / a wrong statement fails alone"""
    n, count = 16, 7
    A, B = oracle.seeded_inputs(91, n * count, threads=8)
    proofs = sipp.sipp_prove_native_batch(A, B, n)
    sts = sipp.sipp_verify_native_batch(A, B, n, proofs)
    for j in range(count):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        one = sipp.sipp_verify_native(a, b, proofs[j])
        assert not isinstance(sts[j], Exception), j
        assert (sts[j].final_A, sts[j].final_B, sts[j].final_Z, sts[j].Z) == (one.final_A, one.final_B, one.final_Z, one.Z), j
        if j < 2:
            ok, ost = oracle.sipp_verify(a, b, b"".join(proofs[j]), threads=4)
            assert ok and sts[j].final_A == ost["final_A"] and sts[j].final_B == ost["final_B"] and sts[j].final_Z == ost["final_Z"]
    bad = [list(p) for p in proofs]
    t = bytearray(bad[3][2]); t[100] ^= 4; bad[3][2] = bytes(t)          # a Z_L / Z_R of instance 3
    t = bytearray(bad[5][-1]); t[0] ^= 1; bad[5][-1] = bytes(t)          # Z of instance 5
    A2 = bytearray(A); A2[64 * n * 1:64 * n * 1 + 64] = A[64 * n * 2:64 * n * 2 + 64]   # instance 1 proves another statement
    sts = sipp.sipp_verify_native_batch(bytes(A2), B, n, bad)
    assert [isinstance(s, sipp.VerificationError) for s in sts] == [False, True, False, True, False, True, False]
    with pytest.raises(sipp.SippError) as ei:
        sipp.sipp_verify_native_batch(A, B, n, [p[:-2] for p in proofs])  # proof.pop().unwrap() on a short proof
    assert ei.value.code == -5


def test_verify_batch_n128(sipp):
    """config-5 shape: every proof of a 64 x n = 128 batch verifies on the GPU, in lock-step"""
    n, count = 128, 64
    A, B = sipp.seeded_inputs(6, n * count)
    proofs = sipp.sipp_prove_native_batch(A, B, n)
    sts = sipp.sipp_verify_native_batch(A, B, n, proofs)
    assert all(not isinstance(s, Exception) for s in sts)
    assert all(s.final_Z != s.Z for s in sts)


def test_batch_config5_full_size(sipp, oracle):
    """BASELINE config 5 at its full size: 4096 x n = 128, instance j = pairs [128 j, 128 (j + 1)) of the seed-5 stream;
    64 of the 4096 proofs against the oracle, ALL of them verified on the GPU (batched verifier), one corrupted proof rejected alone"""
    n, count = 128, 4096
    A, B = sipp.seeded_inputs(5, n * count)
    proofs = sipp.sipp_prove_native_batch(A, B, n)
    assert len(proofs) == count
    for j in range(0, count, 64):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        assert b"".join(proofs[j]) == oracle.sipp_prove(a, b, threads=16), j
    bad = 1234
    t = bytearray(proofs[bad][7]); t[3] ^= 0x10
    proofs[bad] = proofs[bad][:7] + [bytes(t)] + proofs[bad][8:]
    sts = sipp.sipp_verify_native_batch(A, B, n, proofs)
    failed = [j for j, s in enumerate(sts) if isinstance(s, Exception)]
    assert failed == [bad]
    assert all(s.Z == proofs[j][-1] for j, s in enumerate(sts) if j != bad)


def test_batch_degenerate_shapes(sipp, oracle):
    """count = 1, n = 1 (proof = [Z]), and an instance made only of identity pairs (every product is 1)"""
    A, B = oracle.seeded_inputs(3, 8, threads=2)
    assert sipp.sipp_prove_native_batch(A, B, 8) == [sipp.sipp_prove_native(A, B)]
    ones = sipp.sipp_prove_native_batch(A, B, 1)
    assert [p[0] for p in ones] == [sipp.pairing(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128]) for i in range(8)]
    Az = bytes(64 * 4) + A[:64 * 4]
    Bz = B[:128 * 4] + bytes(128 * 4)
    got = sipp.sipp_prove_native_batch(Az, Bz, 4)
    one12 = (1).to_bytes(32, "little") + bytes(352)
    assert got[0] == sipp.sipp_prove_native(Az[:64 * 4], Bz[:128 * 4]) and all(z == one12 for z in got[0])
    assert got[1] == sipp.sipp_prove_native(Az[64 * 4:], Bz[128 * 4:])
    sts = sipp.sipp_verify_native_batch(Az, Bz, 4, got)
    assert all(not isinstance(s, Exception) for s in sts)


def test_batch_random_shapes(sipp):
    """random (n, count) shapes, counts that are not powers of two: every instance of the batch == the single-instance prover"""
    rng = random.Random(2026)
    for trial in range(8):
        n = 1 << rng.randrange(0, 6)
        count = rng.randrange(1, 12)
        A, B = sipp.seeded_inputs(1000 + trial, n * count)
        proofs = sipp.sipp_prove_native_batch(A, B, n)
        assert len(proofs) == count
        for j in range(count):
            a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
            assert proofs[j] == sipp.sipp_prove_native(a, b), (n, count, j)
        sts = sipp.sipp_verify_native_batch(A, B, n, proofs)
        assert all(not isinstance(s, Exception) for s in sts), (n, count)
