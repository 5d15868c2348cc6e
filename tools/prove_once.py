"""a few proves of n pairs (profiling target): python tools/prove_once.py [n] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A, B = sipp_b200.seeded_inputs(2, n)
for _ in range(reps):
    proof = sipp_b200.sipp_prove_native(A, B)
print("ok", len(proof))
