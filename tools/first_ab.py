"""A/B of the first-stage budget (SIPP_OPT_MATRIX_FIRST = log2 of the Miller loops it may spend): prove time of n pairs.
python tools/first_ab.py [n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]
for n in sizes:
    A, B = sipp_b200.seeded_inputs(2, n)
    ref = None
    for cap in [int(c) for c in os.environ.get('FIRST_AB_CAPS', '0,14,15,16,17,18,16,17').split(',')]:
        sipp_b200.set_option(_lib.OPT_MATRIX_FIRST, cap)
        sipp_b200.sipp_prove_native(A, B)
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            proof = sipp_b200.sipp_prove_native(A, B)
            ts.append((time.perf_counter() - t0) * 1e3)
        p = b"".join(proof)
        ref = ref or p
        print("n=%d first-stage budget 2^%-2d  min %.2f ms  median %.2f ms  same=%s" % (n, cap, min(ts), sorted(ts)[len(ts) // 2], p == ref), flush=True)
