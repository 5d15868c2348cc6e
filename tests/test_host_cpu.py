"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the header declares,
fails loudly without a GPU, and its Poseidon transcript (host code, like the reference's) matches the golden vectors."""
import ctypes
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = bytes.fromhex


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "sipp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sipp_[a-z0-9_]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol():
    from sipp_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "libsipp_b200.so does not export %s" % name


def test_no_gpu_fails_loudly():
    from sipp_b200 import _lib
    import sipp_b200
    lib = _lib.load()
    if lib.sipp_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sipp_b200.SippError) as ei:
        sipp_b200.inner_product(bytes(64), bytes(128))
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)
    out = ctypes.create_string_buffer(384)
    assert lib.sipp_prove_native(bytes(64), 1, bytes(128), 1, out) == _lib.ERR_CUDA


def test_product_does_not_reference_oracle():
    """the product package must not import / link the oracle"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sipp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "sipp_oracle" not in txt and "libsipp_oracle" not in txt, f


def test_poseidon_and_transcript_golden(golden):
    import sipp_b200
    from sipp_b200 import _lib
    lib = _lib.load()
    g = golden["poseidon"]

    def perm(st):
        arr = (ctypes.c_uint64 * 12)(*st)
        lib.sipp_poseidon_permute(arr)
        return ["%016x" % v for v in arr]
    assert perm([0] * 12) == g["perm_zero"]
    assert perm(list(range(12))) == g["perm_iota"]
    assert perm([2**64 - 2**32] * 12) == g["perm_pm1"]
    t = sipp_b200.Transcript()
    t.append([1])
    # hash_no_pad(state || [1]) with zero state == hash of [0,0,0,0,1]
    pg, tg = golden["pairing_gen"], golden["transcript"]
    t = sipp_b200.Transcript()
    t.append_g1(H(pg["a"])); assert ["%016x" % v for v in t.state] == tg["after_g1"]
    t.append_g2(H(pg["b"])); assert ["%016x" % v for v in t.state] == tg["after_g2"]
    t.append_fq12(H(pg["exact"])); assert ["%016x" % v for v in t.state] == tg["after_fq12"]
    st = t.state
    assert t.get_challenge().hex() == tg["challenge"]
    assert t.state == st


def test_transcript_matches_oracle_random(oracle):
    import random
    import sipp_b200
    rng = random.Random(9)
    t, ot = sipp_b200.Transcript(), oracle.Transcript()
    for _ in range(40):
        n = rng.choice([0, 1, 3, 4, 5, 12, 16, 32, 96])
        msg = [rng.choice([0, 1, 2**32 - 1, rng.randrange(2**32)]) for _ in range(n)]
        t.append(msg); ot.append(msg)
        assert t.state == ot.state
        assert t.get_challenge() == ot.get_challenge()


def test_fr_inverse_host(oracle):
    import random
    import sipp_b200
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    rng = random.Random(4)
    for v in [1, 2, R - 1] + [rng.randrange(1, R) for _ in range(20)]:
        inv = int.from_bytes(sipp_b200.fr_inverse(v.to_bytes(32, "little")), "little")
        assert inv * v % R == 1
    with pytest.raises(sipp_b200.SippError) as ei:
        sipp_b200.fr_inverse(bytes(32))
    assert ei.value.code == -4  # zero challenge: the reference panics on unwrap()


def test_shard_indices():
    from sipp_b200.sharded import shard_indices
    n, G = 64, 8
    owned = [shard_indices(n, r, G) for r in range(G)]
    assert sorted(sum(owned, [])) == list(range(n))
    # i and its fold partner i + n/2 live on the same rank while n >= 2G
    m = n
    while m >= 2 * G:
        for i in range(m // 2):
            assert i % G == (i + m // 2) % G
        m //= 2


def test_fold_plan_recoding():
    """host recoding of the round challenge for the fold kernel (glv.cc): sum_j (+-NAF_j) L^j == k (mod r) for the GLV
    (G1, x) and GLS (G2, x^-1) decompositions, sub-scalars short, NAF digits non-adjacent"""
    import random
    from sipp_b200 import _lib
    lib = _lib.load()
    lib.sipp_test_fold_plan.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_size_t]
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    X = 4965661367192848881
    L2 = 6 * X * X
    L1 = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    assert (L1 * L1 + L1 + 1) % R == 0 and (L2**4 - L2**2 + 1) % R == 0
    rng = random.Random(8)

    def naf_value(words):
        plus = sum(words[i] << (32 * i) for i in range(5))
        minus = sum(words[5 + i] << (32 * i) for i in range(5))
        assert plus & minus == 0
        nz = plus | minus
        assert nz & (nz >> 1) == 0, "adjacent non-zero digits"
        v = plus - minus
        return -v if words[10] else v, nz.bit_length()

    cases = [(1, 1), (2, pow(2, -1, R)), (R - 1, R - 1), (L1, L2), (L2 % R, L1)] + [(rng.randrange(1, R),) * 2 for _ in range(300)]
    for kx, ki in cases:
        buf = (ctypes.c_uint32 * 128)()
        nw = lib.sipp_test_fold_plan(kx.to_bytes(32, "little"), ki.to_bytes(32, "little"), buf, 128)
        assert nw == 6 * 11 + 2
        w = list(buf[:nw])
        comps = [naf_value(w[11 * j:11 * j + 11]) for j in range(6)]
        g1_bits, g2_bits = w[66], w[67]
        assert sum(v * pow(L1, j, R) for j, (v, _) in enumerate(comps[:2])) % R == kx
        assert sum(v * pow(L2, j, R) for j, (v, _) in enumerate(comps[2:])) % R == ki
        assert g1_bits == max(b for _, b in comps[:2]) and g1_bits <= 130
        assert g2_bits == max(b for _, b in comps[2:]) and g2_bits <= 68


def test_poseidon_avx512_matches_portable():
    """the AVX-512 permutation (used when the CPU has it) computes exactly the portable one, incl. non-canonical
    and extreme lanes, over a long chain"""
    import random
    from sipp_b200 import _lib
    lib = _lib.load()
    lib.sipp_poseidon_permute_portable.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
    lib.sipp_poseidon_permute_portable.restype = None
    rng = random.Random(3)
    PG = 2**64 - 2**32 + 1
    special = [0, 1, PG - 1, 2**32 - 1, 2**32, 2**63, PG - 2**32]
    st = [rng.randrange(PG) for _ in range(12)]
    for it in range(3000):
        if it % 5 == 0:
            st[rng.randrange(12)] = rng.choice(special)
        a = (ctypes.c_uint64 * 12)(*st)
        b = (ctypes.c_uint64 * 12)(*st)
        lib.sipp_poseidon_permute(a)
        lib.sipp_poseidon_permute_portable(b)
        assert list(a) == list(b), it
        st = list(a)


def test_stage_policy_table():
    """the sizes of the pairing-matrix stages are host logic: the first stage follows the host's absorb time (box A/B in
    profiles/r02_v8_first_stage_ab.txt), later stages are 8 blocks from 256 points and single points from 32"""
    from sipp_b200 import _lib
    lib = _lib.load()
    lib.sipp_test_stage_blocks.argtypes = [ctypes.c_int, ctypes.c_size_t]
    lib.sipp_test_stage_blocks.restype = ctypes.c_long
    first = {n: lib.sipp_test_stage_blocks(0, n) for n in (2, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536, 2**20)}
    assert first == {2: 2, 16: 16, 32: 32, 64: 8, 128: 8, 256: 8, 512: 4, 1024: 8, 2048: 16, 4096: 32, 8192: 32, 16384: 16, 65536: 4, 2**20: 0}, first
    later = {n: lib.sipp_test_stage_blocks(1, n) for n in (16, 32, 64, 128, 256, 512, 4096)}
    assert later == {16: 16, 32: 32, 64: 4, 128: 8, 256: 8, 512: 0, 4096: 0}, later
    for n, nr in first.items():
        assert nr == 0 or (n % nr == 0 and nr * n <= 2**18)


def test_poseidon_golden_on_every_backend(golden):
    """the permutation known answers and a transcript through each implementation this CPU can run (SIPP_POSEIDON forces the
    slower ones at load time, so each runs in its own process)"""
    import subprocess
    import sys
    code = (
        "import ctypes, json, sys\n"
        "sys.path.insert(0, %r)\n"
        "import sipp_b200\n"
        "from sipp_b200 import _lib\n"
        "lib = _lib.load()\n"
        "def perm(st):\n"
        "    a = (ctypes.c_uint64 * 12)(*st); lib.sipp_poseidon_permute(a); return ['%%016x' %% v for v in a]\n"
        "t = sipp_b200.Transcript(); t.append(list(range(1, 30)))\n"
        "print(json.dumps({'backend': lib.sipp_poseidon_backend(), 'zero': perm([0] * 12), 'iota': perm(list(range(12))),\n"
        "                  'pm1': perm([2**64 - 2**32] * 12), 'state': ['%%016x' %% v for v in t.state], 'ch': t.get_challenge().hex()}))\n"
    ) % ROOT
    seen = {}
    for force in ("", "avx512", "portable"):
        env = dict(os.environ)
        env.pop("SIPP_POSEIDON", None)
        if force:
            env["SIPP_POSEIDON"] = force
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        r = json.loads(out.stdout.strip().splitlines()[-1])
        g = golden["poseidon"]
        assert r["zero"] == g["perm_zero"] and r["iota"] == g["perm_iota"] and r["pm1"] == g["perm_pm1"], (force, r["backend"])
        seen[force] = r
    assert seen["portable"]["backend"] == 0
    assert seen["avx512"]["backend"] <= 1 and seen[""]["backend"] >= seen["avx512"]["backend"]
    assert seen[""]["state"] == seen["avx512"]["state"] == seen["portable"]["state"]
    assert seen[""]["ch"] == seen["avx512"]["ch"] == seen["portable"]["ch"]


def test_poseidon_avx512_rare_paths():
    """the carry paths of the vector code show up once in ~10^5 permutations, the borrow fix-ups of the scalar asm blocks once in
    2^32 calls: a 400,000-permutation chain against the portable code inside the library, and crafted operands for the blocks"""
    import random
    from sipp_b200 import _lib
    lib = _lib.load()
    lib.sipp_test_poseidon_chain.argtypes = [ctypes.c_uint64, ctypes.c_long]
    lib.sipp_test_poseidon_chain.restype = ctypes.c_long
    lib.sipp_test_poseidon_scalar.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    if lib.sipp_poseidon_backend() < 1:
        pytest.skip("no AVX-512 on this CPU")
    assert lib.sipp_test_poseidon_chain(12345, 400000) == -1        # AVX-512, IFMA (when the CPU has it) and portable side by side
    PG, M = 2**64 - 2**32 + 1, 2**64 - 1
    rng = random.Random(5)

    def call(which, *ins):
        a = (ctypes.c_uint64 * 5)(*(list(ins) + [0] * (5 - len(ins))))
        o = (ctypes.c_uint64 * 2)()
        assert lib.sipp_test_poseidon_scalar(which, a, o) == 0
        return o[0], o[1]

    edge = [0, 1, 2**32 - 1, 2**32, 2**63, PG - 1, PG, M, M - 2**32, 0xFFFFFFFF00000000, 0x00000000FFFFFFFF]
    pairs = [(lo, hi) for lo in edge for hi in edge] + [(rng.randrange(2**64), rng.randrange(2**64)) for _ in range(2000)]
    pairs += [(rng.randrange(2**20), (0xFFFFFFFF << 32) | rng.randrange(2**32)) for _ in range(200)]      # lo < hi >> 32: the borrow fix-up
    for lo, hi in pairs:
        assert call(0, lo, hi)[0] % PG == ((hi << 64) + lo) % PG, (hex(lo), hex(hi))
    for _ in range(3000):
        lo, hi, top = rng.choice(edge + [rng.randrange(2**64)]), rng.choice(edge + [rng.randrange(2**64)]), rng.randrange(14)
        if rng.random() < 0.2:
            lo, hi = rng.randrange(2**16), 0                                                                # small sum, large top: the second fix-up
        p7, m00 = rng.choice(edge + [rng.randrange(2**64)]), rng.randrange(PG)
        if ((hi << 64) + lo + p7 * m00) >> 128:
            continue                                                                                          # the caller's sum stays below 2^128 + top
        assert call(1, lo, hi, top, p7, m00)[0] % PG == ((top << 128) + (hi << 64) + lo + p7 * m00) % PG
    for u in edge + [rng.randrange(2**64) for _ in range(2000)]:
        post = rng.choice([0, 1, PG - 1, rng.randrange(PG)])
        p7, x = call(2, u, post)
        assert p7 % PG == pow(u, 7, PG) and x % PG == (pow(u, 7, PG) + post) % PG, hex(u)
    if lib.sipp_poseidon_backend() == 2:
        # closing of an IFMA accumulator lane: a0 < 2^59 and a1 < 2^59 (sums of 52-bit halves, a1 >= 4095 from the + p in the row
        # constant), a2 < 2^32; the borrow of the - 2^8 a2 step (low word of a0 + 2^52 a1 below 2^8 a2) is a 2^-25 event in real data
        for it in range(4000):
            a1 = 4095 + rng.randrange(2**58)
            a2 = rng.randrange(2**32)
            a0 = rng.randrange(2**58)
            if it % 4 == 0:                                            # low 64 bits of a0 + 2^52 a1 tiny: force the borrow
                a1 = (a1 >> 12) << 12
                a0 = rng.randrange(1 + min(2**20, (a2 << 8)))
            if it % 4 == 1:                                            # a0 + (a1 << 52) carries out of the low word
                a1 |= 0xFFF
                a0 = 2**52 + rng.randrange(2**58)
            c, x = rng.choice(edge + [rng.randrange(PG)]), rng.choice(edge + [rng.randrange(2**64)])
            base = a0 + (a1 << 52) - (a2 << 8)
            assert base >= 0
            sc, ve = call(3, a0, a1, a2, c, x)
            assert sc % PG == (base + c * x) % PG, (hex(a0), hex(a1), hex(a2), hex(c), hex(x))
            assert ve % PG == base % PG, (hex(a0), hex(a1), hex(a2))
        # the vector product of the full rounds; its borrow fix-up (low word of the 128-bit product below the top 32 bits) is a cold call
        ops = [(x, y) for x in edge for y in edge] + [(rng.randrange(2**64), rng.randrange(2**64)) for _ in range(3000)]
        ops += [(2**63, 2**33), (2**32, 2**64 - 2**32), (2**48 + 3, 2**48 * 65535), (PG - 1, PG - 1), (M, M)]
        for k in range(33, 64):                                       # x y = 2^(64 + k) (1 + small): low word tiny, top half large
            ops += [(2**k, 2**64 - 2**(64 - k + 32) * rng.randrange(1, 2**(k - 33) + 1)) for _ in range(20)]
        borrows = 0
        for x, y in ops:
            y %= 2**64
            prod = x * y
            borrows += (prod & M) < (prod >> 96)
            m_, s_ = call(4, x, y)
            assert m_ % PG == prod % PG, (hex(x), hex(y))
            assert s_ % PG == x * x % PG, hex(x)
        assert borrows > 50                                           # the crafted operands do reach the fix-up


def test_statement_public_input_vector(golden):
    """SIPPStatement <-> the u32 vector of statements.rs:133-170 (what the untouched plonky2 circuit consumes): layout, the
    MyFq12 coefficient order (against the independent pure-Python model) and the round trip"""
    import random
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import sipp_model as m
    from sipp_b200 import api
    rng = random.Random(4)
    f = [(rng.randrange(m.P), rng.randrange(m.P)) for _ in range(6)]          # w-basis coefficients g_0..g_5
    ark = m.f12_bytes(f)
    my = api.fq12_to_myfq12_bytes(ark)
    assert [int.from_bytes(my[32 * i:32 * i + 32], "little") for i in range(12)] == m.f12_myfq12_coeffs(f)
    assert api.myfq12_bytes_to_fq12(my) == ark
    n = 4
    fq = lambda: rng.randrange(m.P).to_bytes(32, "little")  # noqa: E731
    A = [fq() + fq() for _ in range(n)]
    B = [fq() + fq() + fq() + fq() for _ in range(n)]
    st = api.SIPPStatement(A=A, B=B, Z=ark, final_A=A[1], final_B=B[2], final_Z=m.f12_bytes(list(reversed(f))))
    vec = st.to_vec()
    assert len(vec) == 16 * n + 32 * n + 96 + 16 + 32 + 96 and all(0 <= v < 2**32 for v in vec)
    assert vec[:8] == [int.from_bytes(A[0][4 * i:4 * i + 4], "little") for i in range(8)]            # x of A_0, 8 LE limbs
    assert vec[48 * n:48 * n + 8] == [(m.f12_myfq12_coeffs(f)[0] >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    assert api.SIPPStatement.from_vec(n, vec) == st
    with pytest.raises(AssertionError):
        api.SIPPStatement.from_vec(n, vec[:-1])
    # straight through the C ABI (what a Rust host binds): lengths, round trip, a limb group >= p is refused, H2 switch
    from sipp_b200 import _lib
    lib = _lib.load()
    total = lib.sipp_statement_u32_len(n)
    assert total == 48 * n + 240
    out = (ctypes.c_uint32 * total)()
    assert lib.sipp_statement_to_u32(b"".join(A), b"".join(B), n, st.Z, st.final_A, st.final_B, st.final_Z, out, total) == 0
    assert list(out) == vec
    assert lib.sipp_statement_to_u32(b"".join(A), b"".join(B), n, st.Z, st.final_A, st.final_B, st.final_Z, out, total - 1) == _lib.ERR_LENGTH
    bad = list(vec)
    bad[8:16] = [0xFFFFFFFF] * 8
    with pytest.raises(_lib.SippError) as ei:
        api.SIPPStatement.from_vec(n, bad)
    assert ei.value.code == _lib.ERR_ENCODING
    lib.sipp_set_option(_lib.OPT_FQ12_ORDER, 1)
    try:
        nested = st.to_vec()
        assert nested[48 * n:48 * n + 96] == [int.from_bytes(ark[4 * i:4 * i + 4], "little") for i in range(96)]
        assert api.SIPPStatement.from_vec(n, nested) == st
    finally:
        lib.sipp_set_option(_lib.OPT_FQ12_ORDER, 0)


def test_shard_instances_partition():
    """batched instances are sharded by contiguous slices (no collective): the slices tile [0, count) and differ by at most one"""
    from sipp_b200.sharded import shard_instances
    for count in (1, 7, 8, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            parts = [shard_instances(count, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == count
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_fr_inverse_binary_matches_fermat():
    """the binary inversion of the device transcript (glv_core.h, compiled for the host) == sipp_fr_inverse (x^(r-2)) == pow(x, -1, r)"""
    import random
    from sipp_b200 import _lib
    lib = _lib.load()
    lib.sipp_test_fr_inverse_binary.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    rng = random.Random(5)
    xs = [1, 2, 3, R - 1, R - 2, (R + 1) // 2, 2**253, 2**64, 2**64 - 1] + [rng.randrange(1, R) for _ in range(3000)]
    for x in xs:
        out = ctypes.create_string_buffer(32)
        assert lib.sipp_test_fr_inverse_binary(x.to_bytes(32, "little"), out) == 0
        assert int.from_bytes(out.raw, "little") == pow(x, -1, R), hex(x)
    out = ctypes.create_string_buffer(32)
    assert lib.sipp_test_fr_inverse_binary((0).to_bytes(32, "little"), out) == -2
    assert lib.sipp_test_fr_inverse_binary(R.to_bytes(32, "little"), out) == -1
