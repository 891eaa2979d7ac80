"""World-size-2 gloo test of the multi-GPU host logic: contiguous shards cover the batch
exactly once, ragged levels included, and the one-off key broadcast delivers identical bytes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tfhe_rs_string_b200 import multigpu as M


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # key arena stand-in: rank 0 holds the bytes, everyone ends up with them
        arena = torch.arange(1 << 16, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(1 << 16, dtype=torch.uint8)
        M.broadcast_bytes(arena, dist, src=0)
        b, e = M.shard_bounds(total, world, rank)
        # each rank "processes" its shard (here: squares), results gathered for the check only
        mine = torch.zeros(total, dtype=torch.int64)
        mine[b:e] = torch.arange(b, e) ** 2
        dist.all_reduce(mine)
        t = torch.tensor([float(e - b)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # the bench's max-over-ranks reduction
        out.put((rank, int(arena.sum()), mine.tolist(), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 4097])
def test_two_rank_shards_and_key_broadcast(total):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_sum = int((np.arange(1 << 16) % 256).sum())
    for rank, s, vals, mx in res:
        assert s == expect_sum
        assert vals == [i * i for i in range(total)]
        assert mx == float(-(-total // 2))


def test_shard_bounds_properties():
    for total in (0, 1, 5, 4096, 4097, 83456):
        for world in (1, 2, 3, 4, 8):
            spans = [M.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
            assert list(M.gather_counts(total, world)) == sizes
    assert M.level_shards([5120, 1024, 3], 2, 1) == [(2560, 5120), (512, 1024), (2, 3)]
    with pytest.raises(ValueError):
        M.shard_bounds(4, 2, 2)
