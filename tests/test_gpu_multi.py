"""Sharded-MSM prover on 2 / 3 / 4 / 8 GPUs (skipped where the box has fewer): each rank keeps its share of the SRS, the library all-gathers
the per-rank window sums over NCCL; both ranks must emit the oracle's golden proof bytes."""
import json
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "marlin_proof_16B.json")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import aes_zero_knowledge_proof_circuit_b200 as zk

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with open(GOLD) as f:
            gold = json.load(f)
        ctx = zk.Context(rank)
        uid = [zk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
        pk = ctx.synthesize_keys(16, bytes.fromhex(gold["tau_seed"]), bytes.fromhex(gold["gamma_seed"]))
        ct, proof = ctx.encrypt(pk, bytes.fromhex(gold["message"]), bytes.fromhex(gold["key"]), bytes.fromhex(gold["zk_seed"]))
        q.put((rank, ct.hex() == gold["ciphertext"], proof.hex() == gold["proof"]))
        pk.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_multi_rank_proof_matches_golden(world):
    """every rank of a `world`-GPU proof must emit the golden bytes (3 ranks: the SRS share and the coset ownership do not
    divide evenly; 4: one round-2 coset per rank; 8: more ranks than cosets)"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True, True) for r in range(world)]
