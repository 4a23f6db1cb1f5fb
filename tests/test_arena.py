"""CPU tier: the host-side scratch arena the prover allocates from (DevArena, csrc/common.cuh) under random traffic."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_arena_random_traffic(seed):
    src = os.path.join(ROOT, "tests", "arena_emul.cpp")
    out = os.path.join(ROOT, "tests", "_arena_emul.so")
    hdr = os.path.join(ROOT, "aes_zero_knowledge_proof_circuit_b200", "csrc", "common.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.arena_selftest.argtypes = [ctypes.c_uint64, ctypes.c_int]
    assert lib.arena_selftest(seed, 20000) == 0
