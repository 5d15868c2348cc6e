"""small end-to-end runs for compute-sanitizer (memcheck / racecheck): single proof, batched proof, batched verifier"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sipp_b200
n, count = 8, 3
A, B = sipp_b200.seeded_inputs(31, n * count)
single = [sipp_b200.sipp_prove_native(A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]) for j in range(count)]
batch = sipp_b200.sipp_prove_native_batch(A, B, n)
assert batch == single
sts = sipp_b200.sipp_verify_native_batch(A, B, n, batch)
assert all(not isinstance(s, Exception) for s in sts)
sipp_b200.sipp_verify_native(A[:64 * n], B[:128 * n], single[0])
A2, B2 = sipp_b200.seeded_inputs(32, 2048)          # a round big enough for k_lines / k_accum / k_fold_split
p = sipp_b200.sipp_prove_native(A2, B2)
assert sipp_b200.inner_product(A2, B2) == p[-1]
A3, B3 = sipp_b200.seeded_inputs(33, 128)           # first stage over the inputs (blocks), look-ahead stage, matrix tail (k_mat.cu)
p3 = sipp_b200.sipp_prove_native(A3, B3)
assert sipp_b200.inner_product(A3, B3) == p3[-1]
sipp_b200.sipp_verify_native(A3, B3, p3)
print("sanitize run ok")
