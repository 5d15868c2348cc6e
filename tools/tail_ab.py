"""A/B of the pairing-matrix tail (SIPP_OPT_MATRIX_TAIL): prove time of n pairs for several thresholds.  python tools/tail_ab.py [n] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
A, B = sipp_b200.seeded_inputs(2, n)
ref = None
for thr, bn, br, first in ((0, 0, 8, 0), (32, 0, 8, 0), (32, 256, 8, 0), (32, 256, 8, 1), (16, 256, 8, 1), (32, 512, 16, 1), (32, 256, 8, 0), (32, 256, 8, 1)):
    sipp_b200.set_option(_lib.OPT_MATRIX_FIRST, first)
    sipp_b200.set_option(_lib.OPT_MATRIX_TAIL, thr)
    sipp_b200.set_option(_lib.OPT_MATRIX_BLOCK_N, bn)
    sipp_b200.set_option(_lib.OPT_MATRIX_BLOCK_R, br)
    sipp_b200.sipp_prove_native(A, B)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        proof = sipp_b200.sipp_prove_native(A, B)
        ts.append((time.perf_counter() - t0) * 1e3)
    p = b"".join(proof)
    ref = ref or p
    print("n=%d tail<=%-3d block<=%-4d R=%-2d first=%d  min %.2f ms  median %.2f ms  same=%s" % (n, thr, bn, br, first, min(ts), sorted(ts)[len(ts) // 2], p == ref), flush=True)
