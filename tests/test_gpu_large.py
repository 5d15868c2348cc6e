"""BASELINE configs[2] (n = 2^16) and configs[3] (n = 2^20) on one B200, bit for bit against the oracle.

The oracle ran ONCE on these seeded inputs (tests/golden/gen_large_digests.py, all host cores, minutes) and left the proof
bytes, every challenge and digests of every folded vector in tests/golden/sipp_large.json; here the CUDA prover reproduces
them.  These are the only tests in which the throughput kernels run INSIDE a checked proof: k_lines (more than 8,192 pairs per
launch), the chunked line table, k_accum with many pairs per accumulator group, and k_fold_straus in the single-proof path
(h >= 16,384)."""
import hashlib
import json
import os

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def large():
    with open(os.path.join(HERE, "golden", "sipp_large.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def sipp():
    import sipp_b200
    from sipp_b200 import _lib
    _lib.require_gpu_once()
    return sipp_b200


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _inputs(sipp, g):
    A, B = sipp.seeded_inputs(g["seed"], g["n"])
    assert sha(A) == g["sha256_A"] and sha(B) == g["sha256_B"], "GPU input generator differs from the oracle's"
    return A, B


@pytest.mark.parametrize("key", ["seed2_n=2^12", "seed2_n=2^13", "n=2^16"])
def test_round_by_round_against_oracle_trace(sipp, large, key):
    """the round-granular C ABI a Rust host would call (sipp_ctx_*), the transcript on the host: every challenge, every folded
    A / B vector (digest over all rounds), final_A / final_B and the assembled proof equal the oracle's"""
    g = large[key]
    n = g["n"]
    A, B = _inputs(sipp, g)
    ctx = sipp.ProverContext(A, B)
    tr = sipp.Transcript()
    from sipp_b200 import _lib
    import ctypes
    _lib.load().sipp_transcript_append_pairs(ctypes.byref(tr._t), A, B, n)
    fwd = [ctx.inner_product()]
    tr.append_fq12(fwd[0])
    hA, hB = hashlib.sha256(), hashlib.sha256()
    challenges = []
    first = True
    while len(ctx) > 1:
        zl, zr = ctx.cross_products()
        fwd += [zl, zr]
        tr.append_fq12(zl)
        tr.append_fq12(zr)
        x = tr.get_challenge()
        challenges.append(x.hex())
        ctx.fold(x, sipp.fr_inverse(x))
        a, b = ctx.read()
        if first:
            assert sha(a) == g["sha256_foldedA_round1"] and sha(b) == g["sha256_foldedB_round1"], "first fold differs"
            first = False
        hA.update(a)
        hB.update(b)
    assert challenges == g["challenges"]
    assert hA.hexdigest() == g["sha256_foldedA_all_rounds"] and hB.hexdigest() == g["sha256_foldedB_all_rounds"]
    assert a.hex() == g["final_A"] and b.hex() == g["final_B"]
    proof = b"".join(reversed(fwd))
    assert proof == open(os.path.join(HERE, "golden", g["proof_file"]), "rb").read()
    ctx.close()


@pytest.mark.parametrize("key", ["seed2_n=2^14", "seed2_n=2^15", "n=2^16", "n=2^20"])
def test_whole_proof_against_oracle_digest(sipp, large, key):
    g = large[key]
    A, B = _inputs(sipp, g)
    proof = b"".join(sipp.sipp_prove_native(A, B))
    assert sha(proof) == g["sha256_proof"]
    assert proof == open(os.path.join(HERE, "golden", g["proof_file"]), "rb").read()
    if g["n"] <= 1 << 16:
        st = sipp.sipp_verify_native(A, B, [proof[i:i + 384] for i in range(0, len(proof), 384)])
        assert st.final_A.hex() == g["final_A"] and st.final_B.hex() == g["final_B"]
        bad = bytearray(proof)
        bad[384 * 3 + 5] ^= 1
        with pytest.raises(sipp.VerificationError):
            sipp.sipp_verify_native(A, B, [bytes(bad[i:i + 384]) for i in range(0, len(bad), 384)])


def test_throughput_kernels_forced_inside_a_checked_proof(sipp, large):
    """n = 2^13 with the latency-regime kernels switched off: every round of the proof runs k_lines, k_accum, k_fold_split /
    k_fold_straus and the 6-lane final exponentiation -- and must still reproduce the oracle's bytes"""
    from sipp_b200 import _lib
    g = large["seed2_n=2^13"]
    A, B = _inputs(sipp, g)
    saved = {o: _lib.load().sipp_get_option(o) for o in (_lib.OPT_WIDE_LINES_MAX, _lib.OPT_WIDE_FOLD_MAX, _lib.OPT_WIDE_ACCUM_MAX, _lib.OPT_FE_ENGINE)}
    try:
        for o in saved:
            sipp.set_option(o, 0)
        proof = b"".join(sipp.sipp_prove_native(A, B))
    finally:
        for o, v in saved.items():
            sipp.set_option(o, v)
    assert sha(proof) == g["sha256_proof"]
