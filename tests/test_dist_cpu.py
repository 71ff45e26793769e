"""CPU, world_size 2, gloo: the multi-rank plumbing bench.py uses -- static sharding with no
data-path collective, and max-over-ranks timing / sum-over-ranks units via all_reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from beatheritage_b200 import segment as seg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    songs = list(seg.shard_range(4096, rank, world))
    elapsed = 1.0 + rank                       # rank 1 is slower
    units = float(len(songs))
    t_max, total = bench.reduce_over_ranks(elapsed, units)
    q.put((rank, len(songs), songs[:3], t_max, total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
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
    assert [r[1] for r in res] == [2048, 2048]
    assert res[0][2] == [0, 2, 4] and res[1][2] == [1, 3, 5]
    for r in res:
        assert r[3] == 2.0 and r[4] == 4096.0    # max over ranks, sum over ranks
