"""Multi-GPU SIPP prover: one process per GPU, pairs sharded with strided ownership (GPU g owns i = g mod G).

Protocol structure follows /root/reference/src/prover_native.rs:26-80; what is distributed is only the data-parallel
part of each round:

  * index i and its fold partner i + n/2 are congruent mod G while n >= 2G, so each rank folds its local array
    of n/G elements exactly like a single-GPU instance (local partner = j + n/(2G)) -- no data-path exchange;
  * per round every rank computes its partial Miller products for Z_L and Z_R (2 x 384 B), they are all-gathered
    (NCCL over NVLink; gloo in the CPU tests), rank 0 multiplies the G partials, runs ONE final exponentiation per
    product, feeds the transcript and broadcasts the challenge (x, x^-1: 64 B);
  * when a rank holds a single pair (n == G) the G remaining points are gathered to rank 0, which finishes the
    tail rounds alone.

The compute engine is injected: the product engine is `CudaEngine` (the C ABI; raises if the library or a GPU is
missing -- there is no CPU fallback).  The CPU `gloo` tests inject an oracle-backed engine that lives under tests/.
"""
from typing import List, Optional

import torch
import torch.distributed as dist

from . import api
from ._lib import SippError

G1_BYTES, G2_BYTES, FQ12_BYTES = 64, 128, 384


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """global indices owned by `rank`: i = rank (mod world)"""
    return list(range(rank, n, world))


def shard_points(A: bytes, B: bytes, rank: int, world: int):
    n = len(A) // G1_BYTES
    idx = shard_indices(n, rank, world)
    return (b"".join(A[G1_BYTES * i:G1_BYTES * (i + 1)] for i in idx),
            b"".join(B[G2_BYTES * i:G2_BYTES * (i + 1)] for i in idx))


def shard_instances(count: int, rank: int, world: int):
    """Batched instances (BASELINE config 5) are independent: rank g proves the contiguous slice [lo, hi) of the instances
    and no collective is needed.  Slices differ by at most one instance."""
    base, extra = divmod(count, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class CudaEngine:
    """Engine over libsipp_b200.so.  Tensors are uint8 CUDA tensors; partials are in the library's device format."""

    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise SippError(-1, "CudaEngine needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)

    class _Ctx:
        def __init__(self, eng, A, B, device_ptrs=None, n=None):
            self.eng = eng
            self.ctx = api.ProverContext(A, B) if device_ptrs is None else api.ProverContext(device_ptrs=device_ptrs, n=n)

        def __len__(self):
            return len(self.ctx)

        def partial_products(self, which: int) -> torch.Tensor:
            nprod = 1 if which == 0 else 2
            out = torch.empty(nprod * FQ12_BYTES, dtype=torch.uint8, device=self.eng.device)
            self.ctx.partial_products(which, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return out

        def fold(self, x: bytes, x_inv: bytes):
            self.ctx.fold(x, x_inv)

        def read(self):
            return self.ctx.read()

        def prove(self, A, B):
            return self.ctx.prove(A, B)

    def create(self, A: bytes, B: bytes):
        return CudaEngine._Ctx(self, A, B)

    def create_from_device(self, device_ptrs, n: int):
        """shard already resident in HBM in boundary format (dA: n x 64 B, dB: n x 128 B)"""
        return CudaEngine._Ctx(self, None, None, device_ptrs, n)

    def combine(self, gathered: torch.Tensor, count: int, nprod: int) -> List[bytes]:
        gathered = gathered.contiguous()
        return api.combine_partials(gathered.data_ptr(), count, nprod, torch.cuda.current_stream().cuda_stream)

    def tensor_device(self):
        return self.device


def _all_gather(t: torch.Tensor, world: int) -> torch.Tensor:
    out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t) if t.is_cuda else dist.all_gather(list(out.view(world, -1).unbind(0)), t)
    return out


def sharded_prove(engine, A_local: bytes, B_local: bytes, n: int, A_full: Optional[bytes] = None, B_full: Optional[bytes] = None,
                  rank: Optional[int] = None, world: Optional[int] = None, device_ptrs=None) -> Optional[List[bytes]]:
    """Runs the SIPP prover over `world` ranks.  Every rank passes its strided shard; rank 0 also passes the full
    A, B (the transcript absorbs every input point, prover_native.rs:36-39).  Returns the proof on rank 0, None elsewhere."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    assert n % world == 0 and n >= world, "n must be a multiple of the number of ranks"
    assert n & (n - 1) == 0 and world & (world - 1) == 0, "n and the rank count must be powers of two"
    dev = engine.tensor_device()
    ctx = engine.create(A_local, B_local) if device_ptrs is None else engine.create_from_device(device_ptrs, n // world)
    tr = None
    absorb = None
    if rank == 0:
        assert A_full is not None and B_full is not None
        tr = api.Transcript()
        import ctypes
        import threading
        from . import _lib
        # register A and B (:36-39): a strictly serial 8n-permutation hash chain that needs nothing from the GPUs, so it
        # runs on a host thread (ctypes releases the GIL) while every rank computes Z and the first Z_L, Z_R
        absorb = threading.Thread(target=_lib.load().sipp_transcript_append_pairs, args=(ctypes.byref(tr._t), A_full, B_full, n))
        absorb.start()
    proof: List[bytes] = []

    def product_round(which: int) -> Optional[List[bytes]]:
        nprod = 1 if which == 0 else 2
        part = ctx.partial_products(which)
        gathered = _all_gather(part, world) if world > 1 else part
        if rank == 0:
            return engine.combine(gathered, world, nprod)
        return None

    z = product_round(0)                                                    # let Z = inner_product(A, B)   :29
    cur = n
    first = True
    while cur > 1 and cur // world >= 2:                                    # folds stay local while n >= 2G
        zs = product_round(1)                                               # Z_L, Z_R   :48-49
        xb = torch.zeros(64, dtype=torch.uint8, device=dev)
        if rank == 0 and first:
            absorb.join()
            proof.append(z[0]); tr.append_fq12(z[0])                        # :42-43
        first = False
        if rank == 0:
            proof.append(zs[0]); tr.append_fq12(zs[0])                      # :52-53
            proof.append(zs[1]); tr.append_fq12(zs[1])                      # :54-55
            x = tr.get_challenge()                                          # :57
            x_inv = api.fr_inverse(x)                                       # :58
            xb = torch.frombuffer(bytearray(x + x_inv), dtype=torch.uint8).to(dev)
        if world > 1:
            dist.broadcast(xb, src=0)                                       # the challenge is broadcast
        xs = bytes(xb.cpu().numpy().tobytes())
        ctx.fold(xs[:32], xs[32:])                                          # :60-74 on the local shard
        cur //= 2
    if rank == 0 and first:                                                 # no local round ran (n == world)
        absorb.join()
        proof.append(z[0]); tr.append_fq12(z[0])
    if cur > 1:
        # every rank holds exactly one pair: collapse the tail onto rank 0
        a1, b1 = ctx.read()
        pt = torch.frombuffer(bytearray(a1 + b1), dtype=torch.uint8).to(dev)
        gathered = _all_gather(pt, world) if world > 1 else pt
        if rank == 0:
            g = bytes(gathered.cpu().numpy().tobytes())
            rec = G1_BYTES + G2_BYTES
            At = b"".join(g[rec * r:rec * r + G1_BYTES] for r in range(world))
            Bt = b"".join(g[rec * r + G1_BYTES:rec * (r + 1)] for r in range(world))
            tail = engine.create(At, Bt)
            while cur > 1:
                part = tail.partial_products(1)
                zs = engine.combine(part, 1, 2)
                proof.append(zs[0]); tr.append_fq12(zs[0])
                proof.append(zs[1]); tr.append_fq12(zs[1])
                x = tr.get_challenge()
                tail.fold(x, api.fr_inverse(x))
                cur //= 2
    if rank == 0:
        proof.reverse()                                                     # :78
        return proof
    return None
