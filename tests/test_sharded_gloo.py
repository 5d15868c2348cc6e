"""world_size 1 / 2 / 4 `gloo` runs, on CPU, of the multi-rank prover's protocol loop -- the product's own loop inside
libsipp_b200.so (`sipp_prove_native_sharded_backend`, host code) -- with the oracle as each rank's arithmetic and gloo as the
exchange (tests/oracle_engine.py).  The proof must be byte-identical to the single-process oracle proof; the call pattern
must show strided ownership at work: folds stay local while n >= 2 world, then ONE collapse onto rank 0."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, seed, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import pyoracle as o
    from oracle_engine import OracleBackend
    from sipp_b200.sharded import shard_points
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, B = o.seeded_inputs(seed, n)
    Al, Bl = shard_points(A, B, rank, world)
    be = OracleBackend(Al, Bl, rank, world)
    proof = be.prove(n, A if rank == 0 else None, B if rank == 0 else None)
    q.put((rank, proof, be.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8), (2, 2), (4, 16), (4, 4), (2, 32)])
def test_sharded_prover_gloo(world, n, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + world * 7 + n
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 21, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(world):
        r, proof, calls = q.get(timeout=240)
        got[r] = (proof, calls)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, B = oracle.seeded_inputs(21, n)
    assert got[0][0] == oracle.sipp_prove(A, B)
    # strided ownership: every rank folds locally from n / world pairs down to one, then the tail collapses once
    local_folds = [m for m in (n // world >> k for k in range(64)) if m >= 2]
    for r in range(world):
        calls = got[r][1]
        assert [c[1] for c in calls if c[0] == "fold"][:len(local_folds)] == local_folds
        assert [c for c in calls if c[0] == "collapse"] == [("collapse", 1)]
    assert [c[1] for c in got[0][1] if c[0] == "fold"][len(local_folds):] == [m for m in (world >> k for k in range(64)) if m >= 2]
    assert all(len([c for c in got[r][1] if c[0] == "fold"]) == len(local_folds) for r in range(1, world))


def test_sharded_backend_world_of_one(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_engine import OracleBackend
    for n in (1, 2, 8):
        A, B = oracle.seeded_inputs(22, n)
        assert OracleBackend(A, B, 0, 1).prove(n, A, B) == oracle.sipp_prove(A, B)


def test_sharded_backend_argument_errors():
    import ctypes
    from sipp_b200 import _lib
    lib = _lib.load()
    assert lib.sipp_prove_native_sharded_backend(None, 8, None, None, None) == _lib.ERR_ARG
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_engine import OracleBackend
    be = OracleBackend(b"", b"", 0, 2)
    buf = ctypes.create_string_buffer(384 * 7)
    assert lib.sipp_prove_native_sharded_backend(ctypes.byref(be.struct), 6, b"", b"", buf) == _lib.ERR_ARG     # n not a power of two
    assert lib.sipp_prove_native_sharded_backend(ctypes.byref(be.struct), 1, b"", b"", buf) == _lib.ERR_ARG     # n < world
    assert lib.sipp_prove_native_sharded_backend(ctypes.byref(be.struct), 8, None, None, buf) == _lib.ERR_ARG   # rank 0 without A, B
    be3 = OracleBackend(b"", b"", 0, 3)
    assert lib.sipp_prove_native_sharded_backend(ctypes.byref(be3.struct), 8, b"", b"", buf) == _lib.ERR_ARG    # world not a power of two
