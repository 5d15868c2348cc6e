"""Differential fuzz of the batched prover / verifier against the single-instance prover (different fold, line and final-exponentiation
kernels on the two sides): random shapes, seeds and kernel-choice options.  python tools/fuzz_batch.py [iterations] [seed]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
checked = 0
for it in range(iters):
    n = 1 << rng.randrange(0, 8)
    count = rng.randrange(1, 48 if n <= 16 else 12)
    seed = rng.randrange(1, 1 << 40)
    opts = {_lib.OPT_BATCH_KPG_MAX: rng.choice([1, 2, 8, 32]), _lib.OPT_FOLD_STRAUS: rng.choice([0, 1]), _lib.OPT_BATCH_QLINES: rng.choice([0, 1]),
            _lib.OPT_FE_ENGINE: rng.choice([0, 1]), _lib.OPT_WIDE_LINES_MAX: rng.choice([0, 8192]), _lib.OPT_WIDE_FOLD_MAX: rng.choice([0, 512])}
    A, B = sipp_b200.seeded_inputs(seed, n * count)
    if rng.random() < 0.3:   # sprinkle identity points
        A, B = bytearray(A), bytearray(B)
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(n * count)
            if rng.random() < 0.5: A[64 * i:64 * i + 64] = bytes(64)
            else: B[128 * i:128 * i + 128] = bytes(128)
        A, B = bytes(A), bytes(B)
    for k, v in opts.items():
        sipp_b200.set_option(k, v)
    try:
        proofs = sipp_b200.sipp_prove_native_batch(A, B, n)
        sts = sipp_b200.sipp_verify_native_batch(A, B, n, proofs)
    finally:
        for k, v in ((_lib.OPT_BATCH_KPG_MAX, 32), (_lib.OPT_FOLD_STRAUS, 1), (_lib.OPT_BATCH_QLINES, 1), (_lib.OPT_FE_ENGINE, 1),
                     (_lib.OPT_WIDE_LINES_MAX, 8192), (_lib.OPT_WIDE_FOLD_MAX, 256)):
            sipp_b200.set_option(k, v)
    assert all(not isinstance(s, Exception) for s in sts), (it, n, count, seed, opts)
    for j in rng.sample(range(count), min(count, 3)):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        assert proofs[j] == sipp_b200.sipp_prove_native(a, b), (it, n, count, seed, j, opts)
        checked += 1
print("fuzz ok: %d iterations, %d instances cross-checked, %.1f s" % (iters, checked, time.time() - t0))
