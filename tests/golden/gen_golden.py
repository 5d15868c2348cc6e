"""Regenerates tests/golden/sipp_golden.json from the pure-Python model (tests/golden/sipp_model.py).

    python tests/golden/gen_golden.py

The reference (Rust, un-vendored dependencies, no toolchain here) cannot generate vectors, so these are
produced by the independent big-integer model and double as the pin for the C oracle and the CUDA path.
SURVEY.md Appendix D digests (n = 2, 4, 8 and e(G1,G2)) are asserted while generating.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sipp_model as m  # noqa: E402

APPENDIX_D = {
    2: "e85f24adb79484590d5b36a9cd53db89a2d8b1cc052077d34ac0390a1fc06b1b",
    4: "cb0289db955cfd85798b913bc57be5b65a8f9325cd7ae1c4a26f2197c09484f9",
    8: "e09bf270cf1a85f6993320f0d4b2bf713287abcd483983dca44254e4146e1631",
}


def hx(b): return b.hex()


def prove_case(name, A, B):
    tr = {}
    proof = m.sipp_prove_native(A, B, trace=tr)
    ok, st = m.sipp_verify_native(A, B, proof)
    assert ok
    return dict(name=name, n=len(A),
                A=hx(b"".join(m.g1_raw(p) for p in A)), B=hx(b"".join(m.g2_raw(q) for q in B)),
                proof=hx(m.proof_bytes(proof)), proof_sha256=m.sha256_hex(m.proof_bytes(proof)),
                challenges=[hx(x.to_bytes(32, "little")) for x in tr.get("x", [])],
                foldedA=hx(b"".join(m.g1_raw(p) for rnd in tr.get("A", []) for p in rnd)),
                foldedB=hx(b"".join(m.g2_raw(q) for rnd in tr.get("B", []) for q in rnd)),
                final_A=hx(m.g1_raw(st["final_A"])), final_B=hx(m.g2_raw(st["final_B"])), final_Z=hx(m.f12_bytes(st["final_Z"])))


def main():
    rng = random.Random(20261017)
    out = {}
    out["poseidon"] = dict(
        rc_first4=["%016x" % v for v in m.RC[:4]], rc_last4=["%016x" % v for v in m.RC[356:]],
        perm_zero=["%016x" % v for v in m.poseidon_perm([0] * 12)],
        perm_iota=["%016x" % v for v in m.poseidon_perm(list(range(12)))],
        perm_pm1=["%016x" % v for v in m.poseidon_perm([m.PG - 1] * 12)],
        hash_1=["%016x" % v for v in m.hash_no_pad([1])],
        hash_iota20=["%016x" % v for v in m.hash_no_pad(list(range(20)))],
    )
    # challenge derivation incl. the zero-limb-stripping quirk
    ch = []
    for digest in ([1, 2, 3, 4], [0, 5, 0, 7], [2**32 - 1, 2**32, 0, 2**63], [m.PG - 1] * 4, [0, 0, 0, 0], [2**40, 3, 2**33 + 1, 9]):
        digits = []
        for d in digest:
            while d:
                digits.append(d & 0xFFFFFFFF)
                d >>= 32
        v = sum(dg << (32 * j) for j, dg in enumerate(digits)) % m.R
        ch.append(dict(digest=["%016x" % d for d in digest], x=hx(v.to_bytes(32, "little"))))
    out["challenge"] = ch
    # transcript walk: append_g1, append_g2, append_fq12, get_challenge
    t = m.Transcript()
    e = m.pairing(m.G1_GEN, m.G2_GEN)
    t.append_g1(m.G1_GEN); s1 = list(t.state)
    t.append_g2(m.G2_GEN); s2 = list(t.state)
    t.append_fq12(e); s3 = list(t.state)
    out["transcript"] = dict(after_g1=["%016x" % v for v in s1], after_g2=["%016x" % v for v in s2], after_fq12=["%016x" % v for v in s3],
                             challenge=hx(t.get_challenge().to_bytes(32, "little")))
    # pairing
    e_ark = m.f12_pow(e, m.LAMBDA_ARK)
    assert m.sha256_hex(m.f12_bytes(e)) == "107999c8a16c357ce5236fdb7d765ed2904d57b2c31e6deca0c93063f8463ea2"
    assert m.sha256_hex(m.f12_bytes(e_ark)) == "e109983de6d3ff0d8d4e1236dd4d91d2a313d7e7a22e3a15062b6759ad70331c"
    out["pairing_gen"] = dict(a=hx(m.g1_raw(m.G1_GEN)), b=hx(m.g2_raw(m.G2_GEN)), exact=hx(m.f12_bytes(e)), ark=hx(m.f12_bytes(e_ark)))
    pr = []
    for _ in range(3):
        a, b = rng.randrange(1, m.R), rng.randrange(1, m.R)
        pa, qb = m.g1_mul(m.G1_GEN, a), m.g2_mul(m.G2_GEN, b)
        val = m.pairing(pa, qb)
        assert val == m.f12_pow(e, a * b % m.R)  # bilinearity
        pr.append(dict(a=hx(m.g1_raw(pa)), b=hx(m.g2_raw(qb)), exact=hx(m.f12_bytes(val))))
    out["pairing_random"] = pr
    # tower ops
    def rf12(): return m.f12_from_ark([rng.randrange(m.P) for _ in range(12)])
    x, y = rf12(), rf12()
    out["fq12"] = dict(a=hx(m.f12_bytes(x)), b=hx(m.f12_bytes(y)), mul=hx(m.f12_bytes(m.f12_mul(x, y))), sqr=hx(m.f12_bytes(m.f12_sqr(x))),
                       inv=hx(m.f12_bytes(m.f12_inv(x))), frob1=hx(m.f12_bytes(m.f12_frob(x, 1))), frob2=hx(m.f12_bytes(m.f12_frob(x, 2))),
                       frob3=hx(m.f12_bytes(m.f12_frob(x, 3))), conj=hx(m.f12_bytes(m.f12_conj(x))))
    # scalar multiplications / folds
    k = rng.randrange(1, m.R)
    pa, qb = m.g1_mul(m.G1_GEN, 12345), m.g2_mul(m.G2_GEN, 67890)
    out["scalar_mul"] = dict(k=hx(k.to_bytes(32, "little")), a=hx(m.g1_raw(pa)), b=hx(m.g2_raw(qb)),
                             ka=hx(m.g1_raw(m.g1_mul(pa, k))), kb=hx(m.g2_raw(m.g2_mul(qb, k))))
    # provers
    cases = []
    for n in (1, 2, 4, 8):
        A = [m.g1_mul(m.G1_GEN, i + 1) for i in range(n)]
        B = [m.g2_mul(m.G2_GEN, i + 1) for i in range(n)]
        c = prove_case("multiples_n%d" % n, A, B)
        if n in APPENDIX_D:
            assert c["proof_sha256"] == APPENDIX_D[n], (n, c["proof_sha256"])
        cases.append(c)
    A, B = m.seeded_inputs(7, 4)
    cases.append(prove_case("seed7_n4", A, B))
    A, B = m.seeded_inputs(11, 16)
    cases.append(prove_case("seed11_n16", A, B))
    # exceptional inputs: identity points, and a fold partner equal / opposite to its mate
    A, B = m.seeded_inputs(13, 4)
    A[1] = None; B[2] = None; A[2] = A[0]; B[3] = m.g2_neg(B[1])
    cases.append(prove_case("exceptional_n4", A, B))
    out["prove"] = cases
    out["seeded_scalars_seed7_n2"] = [hx(s.to_bytes(32, "little")) for s in m.seeded_scalars(7, 2)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sipp_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
