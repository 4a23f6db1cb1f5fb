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


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 5, 8, 16])
@pytest.mark.parametrize("ncoset,ntask", [(3, 7), (4, 5), (1, 7), (2, 5)])
def test_coset_plan_is_acyclic_and_balanced(nranks, ncoset, ntask):
    """zkaes_coset_plan (the split of the prover's round-2 / round-3 coset transforms over the ranks; the prover calls it per wave of at most
    `nranks` cosets): owners are j mod N, every transform has exactly one executor, a rank that hands work out never takes work in (so the
    stream-ordered NCCL sends and receives cannot deadlock), and handing work out never makes the busiest rank busier."""
    nw = min(ncoset, nranks)  # one wave
    owner, ex = zk.coset_plan(nranks, nw, ntask, 1.5)
    assert owner == [j % nranks for j in range(nw)] and len(ex) == nw and all(len(r) == ntask for r in ex)
    assert all(0 <= e < nranks for row in ex for e in row)
    gives = {owner[j] for j in range(nw) for p in range(ntask) if ex[j][p] != owner[j]}
    takes = {ex[j][p] for j in range(nw) for p in range(ntask) if ex[j][p] != owner[j]}
    assert not (gives & takes)
    load = [0.0] * nranks
    for j in range(nw):
        load[owner[j]] += 1.5
        for p in range(ntask):
            load[ex[j][p]] += 1.0 + (0.15 if ex[j][p] != owner[j] else 0.0)
    assert max(load) <= ntask + 1.5 + 1e-9            # never worse than everything at the owner
    if nranks >= 2 * nw:
        assert max(load) < 0.7 * (ntask + 1.5)        # with idle ranks around, the critical path really shrinks
    # handed-out transforms are the last-needed ones of their coset (the owner starts with what it needs first)
    for j in range(nw):
        local = [p for p in range(ntask) if ex[j][p] == owner[j]]
        assert local == list(range(len(local)))
    assert (owner, ex) == zk.coset_plan(nranks, nw, ntask, 1.5)
    with pytest.raises(zk.ZkAesError):
        zk.coset_plan(0, 3, 7, 1.5)
