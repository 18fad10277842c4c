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
    # node donation: rank 0 has 7 open nodes (one too deep to travel), rank 1 has 1
    if rank == 0:
        nodes = [(tuple((10 * i + d, d % 2) for d in range(i)), 40.0 + i / 8) for i in range(6)]
        nodes.append((tuple((d, 1) for d in range(9)), -math.inf))          # depth 9 > max_depth 8: stays
    else:
        nodes = [(((3, 0),), 41.5)]
    out["rebalanced"] = ex.rebalance_frontier(nodes, max_depth=8)
    out["balanced_again"] = ex.rebalance_frontier(out["rebalanced"][0], max_depth=8)[1:]
    out["empty"] = ex.rebalance_frontier([], max_depth=8)[1:]
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
    # 7 + 1 nodes -> 4 + 4: rank 0 gives its last three donatable nodes (i = 5, 4, 3), the deep one stays
    n0, total0, sent0, recv0 = res[0]["rebalanced"]
    n1, total1, sent1, recv1 = res[1]["rebalanced"]
    assert (total0, sent0, recv0) == (8, 3, 0) and (total1, sent1, recv1) == (8, 0, 3)
    mk = lambda i: (tuple((10 * i + d, d % 2) for d in range(i)), 40.0 + i / 8)
    assert n0 == [mk(0), mk(1), mk(2), (tuple((d, 1) for d in range(9)), -math.inf)]
    assert n1 == [(((3, 0),), 41.5), mk(5), mk(4), mk(3)]
    assert res[0]["balanced_again"] == (8, 0, 0) and res[1]["balanced_again"] == (8, 0, 0)
    assert res[0]["empty"] == (0, 0, 0)


def test_transfer_plan_is_balanced_and_minimal():
    from sypha_b200.bnb_exchange import plan_transfers
    targets, moves = plan_transfers([10, 0, 3, 3])
    assert targets == [4, 4, 4, 4]
    assert moves == [(0, 1, 4), (0, 2, 1), (0, 3, 1)]
    targets, moves = plan_transfers([0, 0, 0, 9, 0, 0, 0, 0])
    assert targets == [2, 1, 1, 1, 1, 1, 1, 1] and sum(k for _, _, k in moves) == 8
    assert all(s == 3 for s, _, _ in moves)
    assert plan_transfers([5, 5])[1] == [] and plan_transfers([6, 5])[1] == []


def _async_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sypha_b200 import bnb_exchange as ex
    ax = ex.AsyncBoundExchange(lag=2)
    seen = []
    for r in range(5):
        # rank 1 finds 30 in round 1, rank 0 finds 28 in round 3
        inc = math.inf
        if rank == 1 and r >= 1:
            inc = 30.0
        if rank == 0 and r >= 3:
            inc = 28.0
        ax.post(inc, open_nodes=10 * rank + r, processed=r + 1)
        for rows in ax.collect():
            seen.append(rows.tolist())
    tail = [rows.tolist() for rows in ax.collect(drain=True)]
    x = torch.full((3,), 7.0 + rank, dtype=torch.float64) if rank == 0 else None
    final = ax.final_incumbent(28.0 if rank == 0 else math.inf, x, 3)
    q.put((rank, {"seen": seen, "tail": tail, "final": (final[0], final[1].tolist(), final[2]),
                  "posted": ax.posted, "bytes": ax.bytes_posted}))
    dist.barrier()
    dist.destroy_process_group()


def test_async_bound_exchange_world2():
    """Every rank reads the same lagged rows: round r's gather is collected in round r + lag on both ranks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_async_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    inf = math.inf
    expect = [[[inf, r, r + 1], [30.0 if r >= 1 else inf, 10 + r, r + 1]] for r in range(5)]
    expect[3][0][0] = expect[4][0][0] = 28.0
    for r in range(world):
        assert res[r]["seen"] == expect[:3]            # rounds 0..2 are due by round 4 with lag 2
        assert res[r]["tail"] == expect[3:]
        assert res[r]["final"] == (28.0, [7.0] * 3, 0)
        assert res[r]["posted"] == 5 and res[r]["bytes"] == 5 * 24
