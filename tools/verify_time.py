"""prove / verify time of one n-pair statement through the public calls.  python tools/verify_time.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A, B = sipp_b200.seeded_inputs(2, n)
proof = sipp_b200.sipp_prove_native(A, B)
for name, fn in (("prove", lambda: sipp_b200.sipp_prove_native(A, B)), ("verify", lambda: sipp_b200.sipp_verify_native(A, B, proof))):
    fn()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    print("n=%d %-6s min %.2f ms median %.2f ms" % (n, name, min(ts), sorted(ts)[2]), flush=True)
