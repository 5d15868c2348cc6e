"""ctypes binding of the CPU oracle (oracle/libsipp_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package sipp_b200/ never does.  See oracle/sipp_oracle.h.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsipp_oracle.so")

FE_ARK, FQ12_NESTED, FAITHFUL = 1, 2, 4
OPS = dict(FQ_MUL=0, FQ_ADD=1, FQ_SUB=2, FQ_INV=3, FQ_SQR=4, FQ2_MUL=10, FQ2_SQR=11, FQ2_INV=12, FQ2_MUL_XI=13,
           FQ12_MUL=20, FQ12_SQR=21, FQ12_INV=22, FQ12_FROB1=23, FQ12_FROB2=24, FQ12_FROB3=25, FQ12_CONJ=26,
           FQ12_CYC_SQR=27)


def build(force=False):
    src = os.path.join(_HERE, "sipp_oracle.c")
    if force or not os.path.exists(_LIB) or (os.path.exists(src) and os.path.getmtime(_LIB) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsipp_oracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.oracle_seeded_scalars.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_char_p]
        _lib.oracle_seeded_inputs.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    return _lib


def _buf(n):
    return ctypes.create_string_buffer(n)


def _esize(op):
    return 32 if op < 10 else 64 if op < 20 else 384


def field_op(name, a, b=None):
    op = OPS[name]
    es = _esize(op)
    assert len(a) % es == 0 and (b is None or len(b) == len(a))
    cnt = len(a) // es
    out = _buf(len(a))
    rc = lib().oracle_field_op(op, bytes(a), None if b is None else bytes(b), out, ctypes.c_size_t(cnt))
    if rc:
        raise ValueError("oracle_field_op rc=%d" % rc)
    return out.raw


def g1_mul(a, k):
    out = _buf(64)
    assert lib().oracle_g1_mul(bytes(a), bytes(k), out) == 0
    return out.raw


def g2_mul(b, k):
    out = _buf(128)
    assert lib().oracle_g2_mul(bytes(b), bytes(k), out) == 0
    return out.raw


def g1_on_curve(a): return bool(lib().oracle_g1_on_curve(bytes(a)))
def g2_on_curve(b): return bool(lib().oracle_g2_on_curve(bytes(b)))


def fold_g1(A, x):
    n = len(A) // 64
    out = _buf(64 * (n // 2))
    assert lib().oracle_fold_g1(bytes(A), ctypes.c_size_t(n), bytes(x), out) == 0
    return out.raw


def fold_g2(B, xinv):
    n = len(B) // 128
    out = _buf(128 * (n // 2))
    assert lib().oracle_fold_g2(bytes(B), ctypes.c_size_t(n), bytes(xinv), out) == 0
    return out.raw


def fr_inverse(x):
    out = _buf(32)
    rc = lib().oracle_fr_inverse(bytes(x), out)
    if rc:
        raise ZeroDivisionError("challenge is zero")
    return out.raw


def pairing(a, b, opts=0):
    out = _buf(384)
    assert lib().oracle_pairing(bytes(a), bytes(b), out, ctypes.c_uint(opts)) == 0
    return out.raw


def miller_loop(a, b):
    out = _buf(384)
    assert lib().oracle_miller_loop(bytes(a), bytes(b), out) == 0
    return out.raw


def final_exp(f, opts=0):
    out = _buf(384)
    assert lib().oracle_final_exp(bytes(f), out, ctypes.c_uint(opts)) == 0
    return out.raw


def inner_product(A, B, opts=0, threads=1):
    n = len(A) // 64
    assert len(B) == 128 * n
    out = _buf(384)
    assert lib().oracle_inner_product_mt(bytes(A), bytes(B), ctypes.c_size_t(n), out, ctypes.c_uint(opts), threads) == 0
    return out.raw


def poseidon_perm(state):
    arr = (ctypes.c_uint64 * 12)(*state)
    lib().oracle_poseidon_perm(arr)
    return list(arr)


def hash_no_pad(inputs):
    arr = (ctypes.c_uint64 * max(1, len(inputs)))(*inputs)
    out = (ctypes.c_uint64 * 4)()
    lib().oracle_hash_no_pad(arr, ctypes.c_size_t(len(inputs)), out)
    return list(out)


def round_constants():
    out = (ctypes.c_uint64 * 360)()
    lib().oracle_poseidon_round_constants(out)
    return list(out)


def challenge_from_digest(digest):
    arr = (ctypes.c_uint64 * 4)(*digest)
    out = _buf(32)
    lib().oracle_challenge_from_digest(arr, out)
    return out.raw


class Transcript:
    """mirror of /root/reference/src/transcript_native.rs Transcript<F>"""

    class _S(ctypes.Structure):
        _fields_ = [("state", ctypes.c_uint64 * 4), ("perms", ctypes.c_uint64)]

    def __init__(self, opts=0):
        self._s = self._S()
        self.opts = opts
        lib().oracle_transcript_new(ctypes.byref(self._s))

    @property
    def state(self): return list(self._s.state)
    @property
    def perms(self): return int(self._s.perms)

    def append(self, msg):
        arr = (ctypes.c_uint64 * max(1, len(msg)))(*msg)
        lib().oracle_transcript_append(ctypes.byref(self._s), arr, ctypes.c_size_t(len(msg)))

    def append_g1(self, a): lib().oracle_transcript_append_g1(ctypes.byref(self._s), bytes(a))
    def append_g2(self, b): lib().oracle_transcript_append_g2(ctypes.byref(self._s), bytes(b))
    def append_fq12(self, f): lib().oracle_transcript_append_fq12(ctypes.byref(self._s), bytes(f), ctypes.c_uint(self.opts))

    def get_challenge(self):
        out = _buf(32)
        lib().oracle_transcript_get_challenge(ctypes.byref(self._s), out)
        return out.raw


def n_rounds(n):
    r = 0
    while n > 1:
        n >>= 1
        r += 1
    return r


def sipp_prove(A, B, opts=0, threads=1, trace=False):
    n = len(A) // 64
    assert len(B) == 128 * n
    rounds = n_rounds(n)
    proof = _buf(384 * (2 * rounds + 1))
    ch = _buf(max(1, 32 * rounds)) if trace else None
    fa = _buf(max(1, 64 * (n - 1))) if trace else None
    fb = _buf(max(1, 128 * (n - 1))) if trace else None
    rc = lib().oracle_sipp_prove(bytes(A), bytes(B), ctypes.c_size_t(n), proof, ctypes.c_uint(opts), threads, ch, fa, fb)
    if rc:
        raise ValueError("oracle_sipp_prove rc=%d" % rc)
    if trace:
        return proof.raw, dict(challenges=ch.raw[:32 * rounds], foldedA=fa.raw[:64 * (n - 1)], foldedB=fb.raw[:128 * (n - 1)])
    return proof.raw


def sipp_verify(A, B, proof, opts=0, threads=1):
    n = len(A) // 64
    fa, fb, fz = _buf(64), _buf(128), _buf(384)
    rc = lib().oracle_sipp_verify(bytes(A), bytes(B), ctypes.c_size_t(n), bytes(proof), ctypes.c_size_t(len(proof) // 384),
                                  ctypes.c_uint(opts), threads, fa, fb, fz)
    if rc < 0:
        raise ValueError("oracle_sipp_verify rc=%d" % rc)
    return bool(rc), dict(final_A=fa.raw, final_B=fb.raw, final_Z=fz.raw)


def seeded_scalars(seed, n):
    out = _buf(max(1, 64 * n))
    lib().oracle_seeded_scalars(seed, n, out)
    return out.raw[:64 * n]


def seeded_inputs(seed, n, threads=1):
    A, B = _buf(max(1, 64 * n)), _buf(max(1, 128 * n))
    lib().oracle_seeded_inputs(seed, n, A, B, threads)
    return A.raw[:64 * n], B.raw[:128 * n]
