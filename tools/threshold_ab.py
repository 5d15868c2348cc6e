"""A/B of the size thresholds between the latency engines and the throughput kernels inside an n = 2^12 prove (min of 7 proves)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
A, B = sipp_b200.seeded_inputs(2, 4096)
base = b"".join(sipp_b200.sipp_prove_native(A, B))
def run(label):
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter(); p = sipp_b200.sipp_prove_native(A, B); best = min(best, time.perf_counter() - t0)
    assert b"".join(p) == base
    print("%-34s %.2f ms" % (label, best * 1e3))
run("defaults")
for v in (2048, 4096):
    sipp_b200.set_option(_lib.OPT_WIDE_ACCUM_MAX, v); run("wide_accum_max=%d" % v)
sipp_b200.set_option(_lib.OPT_WIDE_ACCUM_MAX, 1536)
for v in (1024, 2048):
    sipp_b200.set_option(_lib.OPT_WIDE_FOLD_MAX, v); run("wide_fold_max=%d" % v)
sipp_b200.set_option(_lib.OPT_WIDE_FOLD_MAX, 256)
for v in (4096, 16384):
    sipp_b200.set_option(_lib.OPT_WIDE_LINES_MAX, v); run("wide_lines_max=%d" % v)
