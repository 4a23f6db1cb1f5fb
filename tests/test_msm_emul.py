"""CPU tier: the MSM kernels' per-thread bodies (csrc/msm_core.cuh), run sequentially by tests/msm_emul.cpp, against the
oracle's MSM.  Covers what the GPU tier cannot enumerate cheaply: every combination of slice length, chunk size, window
size and reduction segment on inputs that force buckets to straddle slices and chunks."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.cpu import FR, ints_to_limbs, limbs_to_ints, rand_fr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "msm_emul.cpp")
    out = os.path.join(ROOT, "tests", "_msm_emul.so")
    hdr = os.path.join(ROOT, "aes_zero_knowledge_proof_circuit_b200", "csrc", "msm_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.msm_emul.restype = ctypes.c_int
    lib.msm_emul.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t,
                             ctypes.c_uint32, ctypes.c_void_p]
    lib.msm_plan.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def run(curve, bases, scalars, c=0, L=32, chunk=1 << 27, seg=64):
        bases = np.ascontiguousarray(bases, dtype=np.uint64)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros(12, dtype=np.uint64)
        rc = lib.msm_emul(curve, bases.ctypes.data, scalars.ctypes.data, len(scalars), c, L, chunk, seg, out.ctypes.data)
        assert rc > 0
        return out

    lib.msm_emul_paired.restype = ctypes.c_int
    lib.msm_emul_paired.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t,
                                    ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]

    def run_paired(curve, bases, scalars, c=0, L=32, chunk=1 << 27, G=64, G2=128, R=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint64)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros(12, dtype=np.uint64)
        rc = lib.msm_emul_paired(curve, bases.ctypes.data, scalars.ctypes.data, len(scalars), c, L, chunk, G, G2, R, out.ctypes.data)
        assert rc > 0
        return out

    lib.msm_emul_small.restype = ctypes.c_int
    lib.msm_emul_small.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]

    def run_small(curve, bases, vals, n, start=0, stride=1, c=2, S=32, L=32):
        bases = np.ascontiguousarray(bases, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.int32)
        out = np.zeros(12, dtype=np.uint64)
        rc = lib.msm_emul_small(curve, bases.ctypes.data, vals.ctypes.data, n, start, stride, c, S, L, out.ctypes.data)
        assert rc > 0, rc
        return out

    run.lib = lib
    run.paired = run_paired
    run.small = run_small
    return run


def _inputs(oracle, curve, n, seed):
    rng = np.random.default_rng(seed)
    bases = oracle.g1_walk(curve, 1000 + seed, 3, n)
    sc = rand_fr(rng, curve, n)
    if n > 16:
        sc[1] = 0
        sc[2] = ints_to_limbs([1], 4)[0]
        sc[3] = ints_to_limbs([FR[curve] - 1], 4)[0]
        sc[4] = ints_to_limbs([(FR[curve] - 1) // 2], 4)[0]        # largest scalar that is not folded
        sc[5] = ints_to_limbs([(FR[curve] - 1) // 2 + 1], 4)[0]    # smallest scalar that is
        sc[6] = sc[7]
        bases[8] = bases[9]          # equal points with equal scalars: a doubling inside a bucket
        sc[8] = sc[9]
        bases[10] = 0                # point at infinity
        sc[11:16] = sc[16]           # a heavier bucket
    return bases, sc


@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("c,L,chunk,seg", [(4, 1, 1000, 1), (4, 3, 37, 2), (5, 7, 64, 3), (7, 32, 50, 64), (3, 64, 1000, 4), (9, 5, 200, 16),
                                           (11, 33, 128, 64), (2, 4, 16, 1)])
def test_emulated_pipeline_matches_oracle(emul, oracle, curve, c, L, chunk, seg):
    n = 150
    bases, sc = _inputs(oracle, curve, n, c * 100 + L)
    exp = oracle.g1_msm(curve, bases, sc).reshape(-1)
    assert (emul(curve, bases, sc, c=c, L=L, chunk=chunk, seg=seg) == exp).all()


@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("R", [1, 2, 3])
@pytest.mark.parametrize("c,L,chunk,G,G2", [(4, 3, 1000, 1, 1), (4, 5, 37, 3, 2), (5, 7, 64, 64, 128), (3, 64, 1000, 7, 3), (9, 5, 200, 16, 4), (2, 4, 16, 2, 5)])
def test_emulated_pair_round_matches_oracle(emul, oracle, curve, c, L, chunk, G, G2, R):
    """batched-affine pair round: equal points with equal scalars (doubling inside a pair), P + (-P), infinity, odd buckets"""
    n = 150
    bases, sc = _inputs(oracle, curve, n, c * 100 + L)
    bases[20] = bases[21]
    sc[21] = ints_to_limbs([FR[curve] - limbs_to_ints(sc[20:21])[0]], 4)[0]   # s P + (r - s) P: opposite points meet in every bucket
    exp = oracle.g1_msm(curve, bases, sc).reshape(-1)
    bases[30] = 0                                 # the point at infinity as a pair operand
    exp = oracle.g1_msm(curve, bases, sc).reshape(-1)
    assert (emul.paired(curve, bases, sc, c=c, L=L, chunk=chunk, G=G, G2=G2, R=R) == exp).all()


def test_emulated_pair_round_degenerate_inputs(emul, oracle):
    curve = 377
    bases = oracle.g1_walk(curve, 5, 2, 64)
    same = np.tile(ints_to_limbs([FR[curve] - 12345], 4), (64, 1))
    for R in (1, 2, 4):
        assert (emul.paired(curve, bases, same, c=6, L=5, chunk=40, G=4, G2=3, R=R) == oracle.g1_msm(curve, bases, same).reshape(-1)).all()
        eq = np.tile(bases[:1], (64, 1))                                   # the same point 64 times: every pair is a doubling, in every round
        assert (emul.paired(curve, eq, same, c=6, L=5, G=4, G2=3, R=R) == oracle.g1_msm(curve, eq, same).reshape(-1)).all()
    # (the only points with x = 0, (0, +-1), have order 3: they are outside G1, and the MSM's scalar folding s -> r - s
    #  presupposes points of order r, as every arkworks G1Affine is; so x = 0 can only be the encoded point at infinity)
    zero = np.zeros((64, 4), dtype=np.uint64)
    assert (emul.paired(curve, bases, zero, c=6, L=8) == 0).all()
    one = zero.copy()
    one[:, 0] = 1
    assert (emul.paired(curve, bases, one, c=6, L=8) == oracle.g1_msm(curve, bases, one, algo=1).reshape(-1)).all()
    assert (emul.paired(curve, bases[:1], one[:1]) == bases[0].reshape(-1)).all()


def test_emulated_heavy_buckets_and_small_inputs(emul, oracle):
    curve = 377
    bases = oracle.g1_walk(curve, 5, 2, 64)
    same = np.tile(ints_to_limbs([FR[curve] - 12345], 4), (64, 1))   # every term in the same bucket of every window
    for L in (1, 5, 64, 100):
        assert (emul(curve, bases, same, c=6, L=L, chunk=40, seg=8) == oracle.g1_msm(curve, bases, same).reshape(-1)).all()
    zero = np.zeros((64, 4), dtype=np.uint64)
    assert (emul(curve, bases, zero, c=6, L=8) == 0).all()
    one = zero.copy()
    one[:, 0] = 1
    assert (emul(curve, bases, one, c=6, L=8) == oracle.g1_msm(curve, bases, one, algo=1).reshape(-1)).all()
    assert (emul(curve, bases[:1], one[:1], c=0, L=32) == bases[0].reshape(-1)).all()
    assert (emul(curve, bases[:0], one[:0], c=0, L=32) == 0).all()


def test_automatic_plan(emul):
    def plan(n, bits=253, forced=0, nranks=1, cmax=22):
        out = (ctypes.c_int * 3)()
        emul.lib.msm_plan(n, bits, forced, nranks, cmax, out)
        return tuple(out)

    for n in (1, 100, 1 << 16, 1 << 22, 1 << 27, 3 << 27):
        c, W, nbw = plan(n)
        assert 3 <= c <= 22 and W == -(-253 // c) and nbw == 1 << (c - 1)
    assert plan(1 << 27)[0] == 22 and plan(1 << 27)[1] == 12
    assert plan(1 << 27, nranks=8) == plan(1 << 24)          # the plan follows the per-rank share
    assert plan(1 << 27, cmax=23)[:2] == (23, 11)
    assert plan(1 << 20, bits=255, forced=16)[:2] == (16, 16)


@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("c,S,L,n", [(1, 32, 32, 150), (1, 3, 5, 150), (2, 32, 7, 150), (2, 4, 1, 33), (3, 5, 64, 150), (13, 2, 3, 40), (1, 7, 4, 5), (2, 32, 32, 1)])
def test_emulated_small_scalar_pipeline_matches_oracle(emul, oracle, curve, c, S, L, n):
    """the Lagrange-basis commitments' MSM (msm_small_window_sums): one signed digit per term, S pseudo-windows summed without doublings --
    bits, {-2..2}, the widest digit, fewer terms than pseudo-windows, bucket runs cut by many slices, equal / opposite points, infinity"""
    rng = np.random.default_rng(c * 1000 + S * 10 + L)
    lim = 1 << (c - 1)
    vals = (rng.integers(0, 2, size=n) if c == 1 else rng.integers(-lim, lim + 1, size=n)).astype(np.int32)
    bases = oracle.g1_walk(curve, 77 + n, 3, n)
    if n > 16:
        vals[0], vals[1], vals[2] = lim, -lim if c > 1 else 0, 0
        bases[4] = bases[3]
        vals[3] = vals[4] = 1
        if c > 1:
            bases[6] = bases[5]
            vals[5], vals[6] = 1, -1
        bases[7] = 0
        vals[7] = 1
    sc = ints_to_limbs([int(v) % FR[curve] for v in vals], 4)
    exp = oracle.g1_msm(curve, bases, sc).reshape(-1)
    assert (emul.small(curve, bases, vals, n, c=c, S=S, L=L) == exp).all()


def test_emulated_small_scalar_cyclic_share(emul, oracle):
    """rank r of N reads every N-th value starting at r against its own every-N-th bases (prover.cu lagrange_commit): the shares add up to the whole"""
    curve, n, N = 377, 101, 4
    rng = np.random.default_rng(9)
    vals = rng.integers(-2, 3, size=n).astype(np.int32)
    bases = oracle.g1_walk(curve, 5, 7, n)
    parts = []
    for r in range(N):
        local = np.ascontiguousarray(bases[r::N])
        parts.append(emul.small(curve, local, vals, len(local), start=r, stride=N, c=2, S=8, L=3))
    total = parts[0]
    for p in parts[1:]:
        total = oracle.g1_add(curve, total, p)
    sc = ints_to_limbs([int(v) % FR[curve] for v in vals], 4)
    assert (total == oracle.g1_msm(curve, bases, sc).reshape(-1)).all()
