"""A/B of the line kernels inside proves: python tools/wide_ab.py  (prints per-prove kernel-class times for both)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
for n in (4096, 128):
    A, B = sipp_b200.seeded_inputs(2, n)
    ref = None
    for wide in (0, 8192):
        sipp_b200.set_option(_lib.OPT_WIDE_LINES_MAX, wide)
        sipp_b200.set_option(_lib.OPT_FE_ENGINE, 1 if wide else 0)
        sipp_b200.set_option(_lib.OPT_WIDE_FOLD_MAX, 256 if wide else 0)
        sipp_b200.set_option(_lib.OPT_WIDE_ACCUM_MAX, 1536 if wide else 0)
        for rep in range(3):
            sipp_b200.set_option(_lib.OPT_PROFILE, 1)
            sipp_b200.stats(reset=True)
            t0 = time.perf_counter(); proof = sipp_b200.sipp_prove_native(A, B); dt = time.perf_counter() - t0
            st = sipp_b200.stats(reset=True)
        ref = ref or proof
        assert proof == ref, "wide and narrow proofs differ"
        print("n=%d wide_max=%5d  prove %.1f ms  miller %.2f  reduce+fe %.2f  fold %.2f  transcript-exposed %.2f" %
              (n, wide, dt * 1e3, st["miller_ms"], st["reduce_fe_ms"], st["fold_ms"], st["transcript_ms"]))
# one product of m pairs: lines + accum time
for m in (256, 1024, 2048, 4096, 8192, 16384):
    A, B = sipp_b200.seeded_inputs(3, min(m, 4096))
    k = max(1, m // 4096)
    ctx = sipp_b200.ProverContext(A * k, B * k)
    for wide in (0, 1 << 20, 2 << 20):
        sipp_b200.set_option(_lib.OPT_WIDE_LINES_MAX, wide)
        sipp_b200.set_option(_lib.OPT_WIDE_ACCUM_MAX, 1 << 20 if wide == 2 << 20 else 0)
        for rep in range(2):
            sipp_b200.set_option(_lib.OPT_PROFILE, 1)
            sipp_b200.stats(reset=True)
            z = ctx.inner_product()
            st = sipp_b200.stats(reset=True)
        print("m=%6d lines %s accum %s  miller %.3f ms  reduce+fe %.3f ms" % (m, "wide" if wide else "thread", "machine" if wide == 2 << 20 else "6-lane", st["miller_ms"], st["reduce_fe_ms"]))
    ctx.close()
sipp_b200.set_option(_lib.OPT_WIDE_LINES_MAX, 8192)
sipp_b200.set_option(_lib.OPT_WIDE_ACCUM_MAX, 1536)

# fold alone: time per fold of h elements for both kernels
import random
rng = random.Random(1)
for h in (64, 256, 512, 1024, 2048):
    A, B = sipp_b200.seeded_inputs(5, 2 * h)
    for wide in (0, 1 << 20):
        sipp_b200.set_option(_lib.OPT_WIDE_FOLD_MAX, wide)
        outs = []
        for rep in range(2):
            ctx = sipp_b200.ProverContext(A, B)
            x = rng.randrange(1, 1 << 250).to_bytes(32, "little") if rep == 0 else x
            sipp_b200.set_option(_lib.OPT_PROFILE, 1)
            sipp_b200.stats(reset=True)
            ctx.fold(x, sipp_b200.fr_inverse(x))
            outs.append(ctx.read())
            st = sipp_b200.stats(reset=True)
            ctx.close()
        print("fold h=%5d wide=%d  %.3f ms" % (h, 1 if wide else 0, st["fold_ms"]))
sipp_b200.set_option(_lib.OPT_WIDE_FOLD_MAX, 256)
