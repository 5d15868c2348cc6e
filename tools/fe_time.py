import os, sys
sys.path.insert(0, "/root/repo")
import sipp_b200
from sipp_b200 import _lib
A, B = sipp_b200.seeded_inputs(2, 128)
ctx = sipp_b200.ProverContext(A, B)
ctx.cross_products()
sipp_b200.set_option(_lib.OPT_PROFILE, 1)
sipp_b200.stats(reset=True)
for _ in range(20):
    ctx.cross_products()
st = sipp_b200.stats(reset=True)
print("reduce+fe %.3f ms per launch, miller %.3f ms" % (st["reduce_fe_ms"] / 20, st["miller_ms"] / 20))
