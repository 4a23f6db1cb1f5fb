"""Sharded-MSM prover on 2 / 3 / 4 / 8 GPUs (skipped where the box has fewer): each rank keeps its share of the SRS, the library all-gathers
the per-rank window sums over NCCL; every rank must emit the oracle's golden proof bytes.  Two ways of driving the ranks: one process per
GPU (zkaes_ctx_comm_init) and ONE process for all GPUs (zkaes_ctx_create_multi: worker threads + in-process NCCL), the form a drop-in
`encrypt()` call site needs.  A 256-byte message (16 blocks, |H| = 2^22) checks a size whose MSMs really spread over the ranks."""
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


def _worker256(rank, world, port, q):
    """256-byte message: all ranks must emit identical bytes, and the (host, pairing) verifier must accept them"""
    import hashlib

    import torch.distributed as dist

    import aes_zero_knowledge_proof_circuit_b200 as zk

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = zk.Context(rank)
        uid = [zk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
        pk = ctx.synthesize_keys(256, bytes(range(32)), bytes(range(1, 33)))
        msg = bytes((i * 131 + 7) & 0xFF for i in range(256))
        ct, proof = ctx.encrypt(pk, msg, bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), bytes([7] * 32))
        ok = zk.verify_encryption(pk.verifying_key(), proof, ct) if rank == 0 else True
        q.put((rank, hashlib.sha256(ct + proof + pk.verifying_key()).hexdigest(), ok))
        pk.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_256_byte_proof_agrees_with_one_gpu_and_verifies(world, ctx):
    """the 256-byte proof from `world` ranks equals, byte for byte, the one a single GPU computes (sharding must not change results)"""
    import hashlib

    import aes_zero_knowledge_proof_circuit_b200 as zk

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pk = ctx.synthesize_keys(256, bytes(range(32)), bytes(range(1, 33)))
    msg = bytes((i * 131 + 7) & 0xFF for i in range(256))
    ct, proof = ctx.encrypt(pk, msg, bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), bytes([7] * 32))
    single = hashlib.sha256(ct + proof + pk.verifying_key()).hexdigest()
    assert zk.verify_encryption(pk.verifying_key(), proof, ct)
    pk.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker256, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, single, True) for r in range(world)]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_process_multi_gpu_context(world, tmp_path):
    """zkaes_ctx_create_multi: ONE process, `world` GPUs, no torch.distributed -- encrypt() emits the golden bytes, key files written per
    rank load again, and a second proof (another statement) verifies"""
    import aes_zero_knowledge_proof_circuit_b200 as zk

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with open(GOLD) as f:
        gold = json.load(f)
    ctx = zk.Context(list(range(world)))
    try:
        assert ctx.n_devices == world
        pk = ctx.synthesize_keys(16, bytes.fromhex(gold["tau_seed"]), bytes.fromhex(gold["gamma_seed"]))
        ct, proof = ctx.encrypt(pk, bytes.fromhex(gold["message"]), bytes.fromhex(gold["key"]), bytes.fromhex(gold["zk_seed"]))
        assert ct.hex() == gold["ciphertext"] and proof.hex() == gold["proof"] and pk.verifying_key().hex() == gold["verifying_key"]
        path = str(tmp_path / "key16")
        pk.save(path)
        assert all(os.path.exists(f"{path}.r{r}") for r in range(world))
        pk.close()
        pk2 = ctx.load_keys(path)
        ct2, proof2 = ctx.encrypt(pk2, bytes(range(16)), bytes(range(16, 32)), bytes([9] * 32))
        assert zk.verify_encryption(pk2.verifying_key(), proof2, ct2) and not zk.verify_encryption(pk2.verifying_key(), proof2, ct)
        ct3, proof3 = ctx.encrypt(pk2, bytes.fromhex(gold["message"]), bytes.fromhex(gold["key"]), bytes.fromhex(gold["zk_seed"]))
        assert proof3.hex() == gold["proof"]
        pk2.close()
    finally:
        ctx.close()
