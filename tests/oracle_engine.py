"""Oracle-backed compute + gloo exchange for the CPU tests of the sharded prover.  TEST ONLY.

The protocol loop under test is the product's own (`sipp_prove_native_sharded_backend` in libsipp_b200.so: host code, no GPU
needed); this module supplies the six callbacks of `sipp_shard_backend`: the arithmetic of a rank's shard comes from the
oracle, the collectives from torch.distributed (gloo)."""
import ctypes

import torch
import torch.distributed as dist

from oracle import pyoracle as o
from sipp_b200 import _lib

ONE = (1).to_bytes(32, "little") + bytes(352)


def _miller_product(A, B):
    acc = ONE
    for i in range(len(A) // 64):
        acc = o.field_op("FQ12_MUL", acc, o.miller_loop(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128]))
    return acc


class OracleBackend:
    def __init__(self, A_local, B_local, rank, world):
        self.A, self.B, self.rank, self.world = bytes(A_local), bytes(B_local), rank, world
        self.collapsed = False
        self.gathered = b""
        self.calls = []
        B_ = _lib.ShardBackend
        self._keep = [B_.LEN_FN(self._len), B_.PRODUCTS_FN(self._products), B_.COMBINE_FN(self._combine), B_.BROADCAST_FN(self._broadcast),
                      B_.FOLD_FN(self._fold), B_.COLLAPSE_FN(self._collapse)]
        self.struct = B_(None, rank, world, *self._keep)

    def _gather(self, mine: bytes) -> bytes:
        if self.world == 1 or self.collapsed:
            return mine
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return b"".join(bytes(x.numpy().tobytes()) for x in out)

    def _len(self, _u):
        return len(self.A) // 64

    def _products(self, _u, which):
        self.calls.append(("products", which, len(self.A) // 64))
        n = len(self.A) // 64
        if which == 0:
            mine = _miller_product(self.A, self.B)
        else:
            h = n // 2
            mine = _miller_product(self.A[64 * h:], self.B[:128 * h]) + _miller_product(self.A[:64 * h], self.B[128 * h:])
        self.gathered = self._gather(mine)
        return 0

    def _combine(self, _u, nprod, out):
        assert self.rank == 0
        count = len(self.gathered) // (384 * nprod)
        res = b""
        for p in range(nprod):
            acc = ONE
            for r in range(count):
                off = 384 * (r * nprod + p)
                acc = o.field_op("FQ12_MUL", acc, self.gathered[off:off + 384])
            res += o.final_exp(acc)
        ctypes.memmove(out, res, len(res))
        return 0

    def _broadcast(self, _u, xs):
        if self.world > 1:
            t = torch.frombuffer(bytearray(ctypes.string_at(xs, _lib.SHARD_XS_BYTES)), dtype=torch.uint8)
            dist.broadcast(t, src=0)
            ctypes.memmove(xs, bytes(t.numpy().tobytes()), _lib.SHARD_XS_BYTES)
        return 0

    def _fold(self, _u, x, xinv):
        self.calls.append(("fold", len(self.A) // 64))
        x, xinv = ctypes.string_at(x, 32), ctypes.string_at(xinv, 32)
        self.A, self.B = o.fold_g1(self.A, x), o.fold_g2(self.B, xinv)
        return 0

    def _collapse(self, _u):
        self.calls.append(("collapse", len(self.A) // 64))
        assert len(self.A) == 64
        g = self._gather(self.A + self.B)
        self.collapsed = True
        if self.rank == 0:
            self.A = b"".join(g[192 * r:192 * r + 64] for r in range(self.world))
            self.B = b"".join(g[192 * r + 64:192 * (r + 1)] for r in range(self.world))
        return 0

    def prove(self, n, A_full, B_full):
        lib = _lib.load()
        proof = ctypes.create_string_buffer(384 * lib.sipp_proof_len(n)) if self.rank == 0 else None
        _lib.check(lib.sipp_prove_native_sharded_backend(ctypes.byref(self.struct), n, A_full, B_full, proof))
        return proof.raw if self.rank == 0 else None
