"""CPU checks of the large-configuration fixtures (tests/golden/sipp_large.json, written by gen_large_digests.py): the files are
consistent, and the smallest one is regenerated here from scratch -- the oracle on the seeded inputs, fast and faithful variant --
so the committed digests are reproducible without trusting the generator run."""
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def test_large_fixtures_consistent_and_reproducible(oracle):
    with open(os.path.join(HERE, "golden", "sipp_large.json")) as f:
        large = json.load(f)
    assert {"seed2_n=2^12", "seed2_n=2^13", "seed2_n=2^14", "seed2_n=2^15", "n=2^16", "n=2^20"} <= set(large)
    for key, g in large.items():
        proof = open(os.path.join(HERE, "golden", g["proof_file"]), "rb").read()
        k = g["n"].bit_length() - 1
        assert len(proof) == 384 * (2 * k + 1) and hashlib.sha256(proof).hexdigest() == g["sha256_proof"], key
        assert len(g["challenges"]) == k
    g = large["seed2_n=2^12"]
    A, B = oracle.seeded_inputs(g["seed"], g["n"], threads=os.cpu_count() or 1)
    assert hashlib.sha256(A).hexdigest() == g["sha256_A"] and hashlib.sha256(B).hexdigest() == g["sha256_B"]
    proof, tr = oracle.sipp_prove(A, B, 0, os.cpu_count() or 1, trace=True)
    assert hashlib.sha256(proof).hexdigest() == g["sha256_proof"]
    assert [tr["challenges"][32 * i:32 * i + 32].hex() for i in range(12)] == g["challenges"]
    assert tr["foldedA"][-64:].hex() == g["final_A"] and tr["foldedB"][-128:].hex() == g["final_B"]
    ok, st = oracle.sipp_verify(A, B, proof, 0, os.cpu_count() or 1)
    assert ok and st["final_A"].hex() == g["final_A"]
