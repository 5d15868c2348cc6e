"""world_size-2 (and 4) `gloo` runs of the multi-rank prover host logic on CPU, with the oracle as the compute
engine (tests/oracle_engine.py).  The proof must be byte-identical to the single-process oracle proof."""
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
    from oracle_engine import OracleEngine
    from sipp_b200.sharded import shard_points, sharded_prove
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, B = o.seeded_inputs(seed, n)
    Al, Bl = shard_points(A, B, rank, world)
    proof = sharded_prove(OracleEngine(), Al, Bl, n, A if rank == 0 else None, B if rank == 0 else None)
    if rank == 0:
        q.put(b"".join(proof))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8), (2, 2), (4, 16), (4, 4)])
def test_sharded_prover_gloo(world, n, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + world * 7 + n
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 21, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, B = oracle.seeded_inputs(21, n)
    assert got == oracle.sipp_prove(A, B)
