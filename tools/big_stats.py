"""Kernel-class times of one large proof (profile spans of the library), twice.  python tools/big_stats.py [log2 n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
A, B = sipp_b200.seeded_inputs(k - 16 + 3 if k in (16, 20) else 2, n) if False else sipp_b200.seeded_inputs(4 if k == 20 else 3 if k == 16 else 2, n)
for rep in range(2):
    sipp_b200.set_option(_lib.OPT_PROFILE, 1)
    sipp_b200.stats(reset=True)
    t0 = time.perf_counter()
    proof = sipp_b200.sipp_prove_native(A, B)
    dt = time.perf_counter() - t0
    st = sipp_b200.stats(reset=True)
    sipp_b200.set_option(_lib.OPT_PROFILE, 0)
    print("n=2^%d prove %.3f s  exposed transcript %.1f ms  miller %.1f  fe %.1f  fold %.1f  other %.1f ms  loops %d  launches %d" % (
        k, dt, st["transcript_ms"], st["miller_ms"], st["reduce_fe_ms"], st["fold_ms"], st["other_ms"], st["miller_pairs"], st["launches"]), flush=True)
