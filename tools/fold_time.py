"""time of one fold round (sipp_ctx_fold) at several sizes, CUDA-event spans.  python tools/fold_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
x = bytes.fromhex("0123456789abcdef" * 3 + "0123456789abcd0f")
xi = sipp_b200.fr_inverse(x)
for n in (8192, 4096, 2048, 1024, 512):
    A, B = sipp_b200.seeded_inputs(4, n)
    ctx = sipp_b200.ProverContext(A, B); ctx.fold(x, xi); ctx.read(); ctx.close()
    ts = []
    for _ in range(3):
        ctx = sipp_b200.ProverContext(A, B)
        sipp_b200.set_option(_lib.OPT_PROFILE, 1); sipp_b200.stats(reset=True)
        ctx.fold(x, xi); ctx.read()
        ts.append(sipp_b200.stats(reset=True)["fold_ms"]); sipp_b200.set_option(_lib.OPT_PROFILE, 0)
        ctx.close()
    print("fold h=%d  %.3f ms" % (n // 2, min(ts)), flush=True)
