"""World-size-2 ``gloo`` checks of the multi-rank plumbing (runs on CPU): rendezvous on 127.0.0.1, disjoint
burst shards, max-over-ranks timing, output gather.  The data path itself has no collective."""
import os
import socket

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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from fbanet_b200.dist import gather_rows, init_from_env, reduce_max, shard_range
    r, _, w = init_from_env("gloo")
    n = 7
    b, e = shard_range(n, r, w)
    mine = torch.arange(n, dtype=torch.float32)[b:e].view(-1, 1) * 10 + r  # rank-tagged rows
    counts = [shard_range(n, i, w)[1] - shard_range(n, i, w)[0] for i in range(w)]
    full = gather_rows(mine, counts)
    slow = reduce_max(100.0 + 50.0 * r)  # rank 1 is "slower": the reported time is the max
    dist.barrier()
    q.put((r, (b, e), full.flatten().tolist(), slow))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, full0, t0), (r1, s1, full1, t1) = res
    assert s0 == (0, 4) and s1 == (4, 7)
    assert full0 == full1 == [0.0, 10.0, 20.0, 30.0, 41.0, 51.0, 61.0]
    assert t0 == t1 == 150.0
