#!/usr/bin/env python
"""Generates tests/golden/sipp_large.json (+ the proof bytes) by running the CPU oracle ONCE on BASELINE configs[2] and [3]:

    n = 2^16, seed 3   and   n = 2^20, seed 4     (seeded generator of oracle_seeded_inputs / sipp_seeded_inputs)

The oracle runs its fast variant (product of Miller loops + one final exponentiation per product -- the same values as the
faithful one, tests/test_oracle_golden.py pins that) on all host cores: about 20 s at 2^16 and a few minutes at 2^20.
Recorded per size: sha256 of the inputs, the proof (binary file), sha256(proof), every challenge, sha256 of the concatenated
folded A / folded B vectors of all rounds, and final_A / final_B.  The -m gpu tests and bench.py compare the CUDA prover
(one GPU and sharded) with these.

    python tests/golden/gen_large_digests.py [--keys n=2^16,n=2^20]
"""
import argparse
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as o  # noqa: E402

# (log2 n, seed, key): the headline sizes of bench.py (seed 2: n = 2^12 on one GPU, 2^12 per GPU up to 2^15 on eight) and the
# two large BASELINE configurations
CASES = [(12, 2, "seed2_n=2^12"), (13, 2, "seed2_n=2^13"), (14, 2, "seed2_n=2^14"), (15, 2, "seed2_n=2^15"), (16, 3, "n=2^16"), (20, 4, "n=2^20")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", default=",".join(c[2] for c in CASES))
    args = ap.parse_args()
    path = os.path.join(HERE, "sipp_large.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    threads = os.cpu_count() or 1
    for k, seed, key in [c for c in CASES if c[2] in args.keys.split(",")]:
        n = 1 << k
        t0 = time.time()
        A, B = o.seeded_inputs(seed, n, threads=threads)
        t1 = time.time()
        proof, tr = o.sipp_prove(A, B, 0, threads, trace=True)
        t2 = time.time()
        h = n // 2
        fa, fb = tr["foldedA"], tr["foldedB"]
        binname = "sipp_large_%s.proof.bin" % key.replace("=2^", "2p")
        open(os.path.join(HERE, binname), "wb").write(proof)
        doc[key] = {
            "n": n, "seed": seed, "fe_normalisation": "exact", "fq12_transcript_order": "w-basis",
            "sha256_A": hashlib.sha256(A).hexdigest(), "sha256_B": hashlib.sha256(B).hexdigest(),
            "proof_file": binname, "sha256_proof": hashlib.sha256(proof).hexdigest(),
            "challenges": [tr["challenges"][32 * i:32 * (i + 1)].hex() for i in range(k)],
            "sha256_foldedA_all_rounds": hashlib.sha256(fa).hexdigest(),
            "sha256_foldedB_all_rounds": hashlib.sha256(fb).hexdigest(),
            "sha256_foldedA_round1": hashlib.sha256(fa[:64 * h]).hexdigest(),
            "sha256_foldedB_round1": hashlib.sha256(fb[:128 * h]).hexdigest(),
            "final_A": fa[-64:].hex(), "final_B": fb[-128:].hex(),
            "oracle_seconds": {"inputs": round(t1 - t0, 1), "prove": round(t2 - t1, 1), "threads": threads},
        }
        print("n=2^%d: inputs %.1f s, prove %.1f s, sha256(proof) = %s" % (k, t1 - t0, t2 - t1, doc[key]["sha256_proof"]), flush=True)
        json.dump(doc, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
