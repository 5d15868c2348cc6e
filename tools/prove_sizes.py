"""prove / verify time at the default options for a range of sizes.  python tools/prove_sizes.py [sizes...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
for n in sizes:
    A, B = sipp_b200.seeded_inputs(2, n)
    proof = sipp_b200.sipp_prove_native(A, B)
    out = []
    for fn in (lambda: sipp_b200.sipp_prove_native(A, B), lambda: sipp_b200.sipp_verify_native(A, B, proof)):
        fn()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
        out.append(min(ts))
    print("n=%-5d prove %.2f ms  verify %.2f ms" % (n, out[0], out[1]), flush=True)
