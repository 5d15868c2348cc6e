"""Differential fuzz of the single-instance GPU prover / verifier against the CPU oracle (test infrastructure): random sizes, seeds,
identity points and pipeline options.  python tools/fuzz_single.py [iterations] [seed]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
from oracle import pyoracle as o

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
defaults = ((_lib.OPT_PIPELINE, 1), (_lib.OPT_FE_ENGINE, 1), (_lib.OPT_WIDE_LINES_MAX, 8192), (_lib.OPT_WIDE_FOLD_MAX, 256), (_lib.OPT_WIDE_ACCUM_MAX, 1536),
            (_lib.OPT_MATRIX_TAIL, 32), (_lib.OPT_MATRIX_BLOCK_N, 256), (_lib.OPT_MATRIX_BLOCK_R, 8), (_lib.OPT_MATRIX_FIRST, 1))
for it in range(iters):
    n = 1 << rng.randrange(0, 10 if it % 8 == 0 else 7)
    seed = rng.randrange(1, 1 << 40)
    A, B = sipp_b200.seeded_inputs(seed, n)
    if rng.random() < 0.4:
        A, B = bytearray(A), bytearray(B)
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(n)
            if rng.random() < 0.5: A[64 * i:64 * i + 64] = bytes(64)
            else: B[128 * i:128 * i + 128] = bytes(128)
        A, B = bytes(A), bytes(B)
    opts = {_lib.OPT_PIPELINE: rng.choice([1, 1, 1, 0]), _lib.OPT_FE_ENGINE: rng.choice([0, 1]), _lib.OPT_WIDE_LINES_MAX: rng.choice([0, 8192]),
            _lib.OPT_WIDE_FOLD_MAX: rng.choice([0, 512]), _lib.OPT_WIDE_ACCUM_MAX: rng.choice([0, 1536]),
            # pairing-matrix stages: tail size, look-ahead stage, first stage over the inputs
            _lib.OPT_MATRIX_TAIL: rng.choice([0, 2, 4, 8, 16, 32]), _lib.OPT_MATRIX_BLOCK_N: rng.choice([0, 64, 256, 512]),
            _lib.OPT_MATRIX_BLOCK_R: rng.choice([4, 8, 16, 32]), _lib.OPT_MATRIX_FIRST: rng.choice([0, 1, 1])}
    for k, v in opts.items():
        sipp_b200.set_option(k, v)
    want = o.sipp_prove(A, B, threads=8)
    try:
        proof = sipp_b200.sipp_prove_native(A, B)
        assert b"".join(proof) == want, ("prove", it, n, seed, opts)
        try:
            st = sipp_b200.sipp_verify_native(A, B, proof)
        except Exception as e:
            raise AssertionError(("verify", repr(e), it, n, seed, opts))
    finally:
        for k, v in defaults:
            sipp_b200.set_option(k, v)
    ok, ost = o.sipp_verify(A, B, want, threads=8)
    assert ok and st.final_A == ost["final_A"] and st.final_B == ost["final_B"] and st.final_Z == ost["final_Z"], (it, n, seed)
print("fuzz ok: %d proofs byte-identical to the oracle, %.1f s" % (iters, time.time() - t0))
