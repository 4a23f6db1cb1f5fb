"""Multi-GPU host logic on CPU: the point-range rule of the sharded MSM under a world_size-2 gloo group
(every rank derives its own range; together they tile [0, n) in rank order), and its edge cases."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aes_zero_knowledge_proof_circuit_b200 as zk


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = torch.tensor([list(zk.shard_range(n, rank, world)) for n in sizes], dtype=torch.int64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        if rank == 0:
            for i, n in enumerate(sizes):
                pos = 0
                for r in range(world):
                    start, count = int(gathered[r][i][0]), int(gathered[r][i][1])
                    assert start == pos, (n, r, start, pos)
                    pos += count
                assert pos == n
                counts = [int(gathered[r][i][1]) for r in range(world)]
                assert max(counts) - min(counts) <= 1
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_point_ranges_tile_under_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    sizes = [0, 1, 2, 3, 1572862, 25165822, 402653182]  # incl. the SRS sizes of the 16 B / 256 B / 4 KiB keys
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_point_ranges_single_process(world):
    for n in (0, 5, 8, 1000003):
        pos = 0
        for r in range(world):
            s, c = zk.shard_range(n, r, world)
            assert s == pos
            pos += c
        assert pos == n
    with pytest.raises(zk.ZkAesError):
        zk.shard_range(10, world, world)
