"""saturated Miller-loop leg: one inner product over many pairs (the workload the ncu --set full capture profiles)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
from sipp_b200 import _lib
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sA, sB = sipp_b200.seeded_inputs(3, 4096)
k = m // 4096
ctx = sipp_b200.ProverContext(sA * k, sB * k)
for i in range(reps):
    sipp_b200.set_option(_lib.OPT_PROFILE, 1)
    sipp_b200.stats(reset=True)
    t0 = time.perf_counter(); ctx.inner_product(); dt = time.perf_counter() - t0
    st = sipp_b200.stats(reset=True)
    print("pairs %d  miller %.3f ms (%.2f M loops/s)  reduce+fe %.3f ms  wall %.1f ms" % (m, st["miller_ms"], m / st["miller_ms"] / 1e3, st["reduce_fe_ms"], dt * 1e3))
ctx.close()
