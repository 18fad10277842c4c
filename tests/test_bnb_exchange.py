"""CPU, world_size 2, gloo: the host-side multi-rank logic (incumbent exchange, partitioning,
bench counter reductions).  The LP path itself never uses a collective."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sypha_b200 import bnb_exchange as ex
    out = {}
    # round 1: nobody has an incumbent
    out["r1"] = ex.exchange_incumbent(math.inf, None, 4)[0]
    # round 2: rank 1 has the better incumbent
    obj = 12.0 if rank == 0 else 9.0
    x = torch.full((4,), float(rank + 1), dtype=torch.float64)
    b, bx, owner = ex.exchange_incumbent(obj, x, 4)
    out["r2"] = (b, bx.tolist(), owner)
    # round 3: tie -> lowest rank wins on every rank
    b, bx, owner = ex.exchange_incumbent(7.0, torch.full((4,), 10.0 + rank, dtype=torch.float64), 4)
    out["r3"] = (b, bx.tolist(), owner)
    out["lb"] = ex.global_lower_bound(5.0 + rank)
    out["cnt"] = ex.reduce_counters(1.0 + rank, [10 * (rank + 1), 1])
    out["part"] = ex.partition_round_robin(list(range(7)), rank, world)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_incumbent_exchange_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        o = res[r]
        assert o["r1"] == math.inf
        assert o["r2"] == (9.0, [2.0] * 4, 1)
        assert o["r3"] == (7.0, [10.0] * 4, 0)
        assert o["lb"] == 5.0
        assert o["cnt"] == (2.0, [30.0, 2.0])
    assert res[0]["part"] == [0, 2, 4, 6] and res[1]["part"] == [1, 3, 5]
