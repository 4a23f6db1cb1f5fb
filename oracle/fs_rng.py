"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Randomness of the Marlin prover, restated.

Follows (pinned third-party crates, not under /root/reference -- SURVEY.md 8(c)):
    rand_chacha 0.3.1   ChaCha20Rng: key = 32-byte seed, 64-bit block counter in words 12-13, stream 0, 20 rounds;
                        rand_core BlockRng: 64-word buffer, next_u64 = two consecutive words (low first)
    ark-marlin 0.3.0    rng.rs FiatShamirRng<Blake2s>: seed = Blake2s(bytes); absorb: seed' = Blake2s(bytes || seed)
    ark-ff 0.3.0        UniformRand for Fp256: draw [u64; 4], clear the top REPR_SHAVE_BITS, accept if < modulus;
                        the accepted words ARE the (Montgomery) representation
Only u64-granular draws are ever made on this path, so the generator is modelled as a stream of u64.
"""
from __future__ import annotations

import hashlib

import numpy as np


def _rotl(x, n):
    return (x << np.uint32(n)) | (x >> np.uint32(32 - n))


def chacha20_blocks(key: bytes, counter0: int, nblocks: int) -> np.ndarray:
    """nblocks keystream blocks starting at 64-bit block counter `counter0` -> (nblocks, 16) uint32."""
    const = np.frombuffer(b"expand 32-byte k", dtype="<u4")
    k = np.frombuffer(key, dtype="<u4")
    ctr = np.arange(counter0, counter0 + nblocks, dtype=np.uint64)
    st = np.zeros((16, nblocks), dtype=np.uint32)
    for i in range(4):
        st[i] = const[i]
    for i in range(8):
        st[4 + i] = k[i]
    st[12] = (ctr & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    st[13] = (ctr >> np.uint64(32)).astype(np.uint32)
    x = st.copy()

    def qr(a, b, c, d):
        x[a] += x[b]; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] += x[d]; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] += x[b]; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] += x[d]; x[b] = _rotl(x[b] ^ x[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        x += st
    return np.ascontiguousarray(x.T)


class ChaCha20Rng:
    """u64 stream of rand_chacha's ChaCha20Rng::from_seed(seed)."""

    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = bytes(seed)
        self.pos = 0  # in u64 units (8 per block)

    def take_u64(self, n: int) -> np.ndarray:
        b0, b1 = self.pos // 8, (self.pos + n + 7) // 8
        words = chacha20_blocks(self.key, b0, max(b1 - b0, 1)).reshape(-1)
        u = words[0::2].astype(np.uint64) | (words[1::2].astype(np.uint64) << np.uint64(32))
        off = self.pos - b0 * 8
        self.pos += n
        return u[off:off + n].copy()

    def next_u64(self) -> int:
        return int(self.take_u64(1)[0])

    def next_u128(self) -> int:  # rand 0.8 Standard for u128: low = first u64, high = second
        lo, hi = self.take_u64(2)
        return (int(hi) << 64) | int(lo)


def fr_rand_many(rng: ChaCha20Rng, modulus: int, n: int) -> np.ndarray:
    """n draws of ark-ff `Fr::rand(rng)` -> (n, 4) uint64 raw (Montgomery) limbs."""
    if n == 0:
        return np.zeros((0, 4), dtype=np.uint64)
    shave = 256 - modulus.bit_length()
    mask = np.uint64((1 << (64 - shave)) - 1)
    mod_limbs = [np.uint64((modulus >> (64 * k)) & ((1 << 64) - 1)) for k in range(4)]
    out = []
    have = 0
    while have < n:
        need = n - have
        m = int(need * (2 ** modulus.bit_length()) / modulus * 1.05) + 16
        start = rng.pos
        c = rng.take_u64(4 * m).reshape(m, 4)
        c[:, 3] &= mask
        lt = np.zeros(m, dtype=bool)
        eq = np.ones(m, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (c[:, k] < mod_limbs[k])
            eq &= c[:, k] == mod_limbs[k]
        idx = np.nonzero(lt)[0]
        if len(idx) >= need:
            last = idx[need - 1]
            rng.pos = start + 4 * (int(last) + 1)  # un-consume the candidates after the last accepted one
            out.append(c[idx[:need]])
            have = n
        else:
            out.append(c[idx])
            have += len(idx)
    return np.ascontiguousarray(np.concatenate(out, axis=0))


def fr_rand(rng: ChaCha20Rng, modulus: int) -> np.ndarray:
    return fr_rand_many(rng, modulus, 1)[0]


class FiatShamirRng:
    """ark-marlin 0.3.0 FiatShamirRng<Blake2s>."""

    def __init__(self, seed_bytes: bytes):
        self.seed = hashlib.blake2s(seed_bytes).digest()
        self.rng = ChaCha20Rng(self.seed)

    def absorb(self, data: bytes):
        self.seed = hashlib.blake2s(data + self.seed).digest()
        self.rng = ChaCha20Rng(self.seed)
