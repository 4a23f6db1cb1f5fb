"""csrc/fq52.cuh -- the FP64-limb (8 x 52 bit, DFMA hi/lo) Montgomery product of the round-2 experiment.
CPU tier: the same header compiled for the host, with std::fma under round-toward-zero standing in for __fma_rz, against Python
integers.  GPU tier: the device kernel through zkaes_selftest_field variant 3."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 0x1AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
R416_INV = pow(1 << 416, -1, Q)

HOST_SRC = r"""
#include <cfenv>
#include "fq52.cuh"
using namespace zk;
extern "C" void fq52_mul_words(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    for (size_t i = 0; i < n; ++i) {
        Fq52 r = fq52_mont_mul<Fq377P52>(fq52_from_words(a + 12 * i), fq52_from_words(b + 12 * i));
        fq52_to_words(r, out + 12 * i);
    }
    fesetround(old);
}
"""


def _cases(n, seed):
    rnd = random.Random(seed)
    vals = [(0, 0), (1, 1), (Q - 1, Q - 1), (Q - 1, 1), (1, Q - 1), (Q - 2, Q - 3), ((1 << 376) % Q, (1 << 370) % Q), ((1 << 52) - 1, (1 << 52) - 1)]
    vals += [(((1 << 52) - 1) << (52 * k), Q - 1 - k) for k in range(7)]
    vals += [(rnd.randrange(Q), rnd.randrange(Q)) for _ in range(n)]
    words = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    a = np.array([words(x) for x, _ in vals], dtype=np.uint32)
    b = np.array([words(y) for _, y in vals], dtype=np.uint32)
    return vals, a, b


def _check(vals, out):
    for (x, y), row in zip(vals, out):
        got = sum(int(w) << (32 * i) for i, w in enumerate(row))
        assert got == x * y * R416_INV % Q, (hex(x), hex(y))


def test_fq52_product_host_emulation(tmp_path):
    src = tmp_path / "fq52_host.cpp"
    src.write_text(HOST_SRC)
    so = tmp_path / "fq52_host.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-frounding-math", "-shared", "-fPIC", "-I", os.path.join(ROOT, "aes_zero_knowledge_proof_circuit_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = ctypes.CDLL(str(so))
    vals, a, b = _cases(5000, 52)
    out = np.zeros_like(a)
    lib.fq52_mul_words(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(vals)))
    _check(vals, out)


@pytest.mark.gpu
def test_fq52_product_on_the_device(ctx):
    vals, a, b = _cases(1 << 14, 53)
    out = ctx.selftest_field(377, 1, 2, 3, a.view(np.uint64).reshape(-1, 6), b.view(np.uint64).reshape(-1, 6))
    _check(vals, np.ascontiguousarray(out).view(np.uint32).reshape(-1, 12))
