"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess
from ctypes import c_int, c_size_t, c_uint64, c_void_p

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE_DIR)
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")

FR = {377: 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
      381: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001}
FQ = {377: 0x1AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
      381: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB}


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "zk_oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


def ints_to_limbs(vals, nlimbs):
    """list of python ints -> (len, nlimbs) uint64 little-endian limbs"""
    out = np.zeros((len(vals), nlimbs), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        for k in range(nlimbs):
            out[i, k] = (v >> (64 * k)) & mask
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, arr.shape[-1])
    return [sum(int(x) << (64 * k) for k, x in enumerate(row)) for row in arr]


def rand_fr(rng: np.random.Generator, curve, n):
    """n uniform-ish canonical scalars < r as (n,4) uint64 (top limb reduced by rejection on the top bits)."""
    r = FR[curve]
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    top = np.uint64(r >> 192)
    a[:, 3] %= top  # strictly below the modulus' top limb => value < r
    return a


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        vp = c_void_p
        L.orc_field_op.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_size_t]
        L.orc_constant.argtypes = [c_int, c_int, vp]
        L.orc_g1_on_curve.argtypes = [c_int, vp, c_size_t]
        L.orc_g1_mul_gen.argtypes = [c_int, vp, c_size_t, vp]
        L.orc_g1_walk.argtypes = [c_int, c_uint64, c_uint64, c_size_t, vp]
        L.orc_g1_msm.argtypes = [c_int, vp, vp, c_size_t, c_int, vp]
        L.orc_g1_add.argtypes = [c_int, vp, vp, vp]
        L.orc_ntt.argtypes = [c_int, vp, c_int, c_int, c_int]
        L.orc_dft_naive.argtypes = [c_int, vp, vp, c_int, c_int]
        L.orc_aes128_ecb.argtypes = [vp, c_size_t, vp, vp, vp]
        L.orc_set_threads.argtypes = [c_int]
        L.orc_fr_vec.argtypes = [c_int, c_int, vp, c_size_t, vp, c_size_t, vp]
        L.orc_fr_batch_inv.argtypes = [c_int, vp, vp, c_size_t]
        L.orc_fr_poly_eval.argtypes = [c_int, vp, c_size_t, vp, vp]
        L.orc_fr_powers.argtypes = [c_int, vp, vp, c_size_t, vp]
        L.orc_fr_div_linear.argtypes = [c_int, vp, c_size_t, vp, vp]
        L.orc_g1_mul.argtypes = [c_int, vp, vp, c_size_t, vp]
        L.orc_g1_srs.argtypes = [c_int, vp, c_size_t, vp]

    @staticmethod
    def _p(a):
        return c_void_p(a.ctypes.data) if a is not None else None

    def threads(self):
        return self.lib.orc_threads()

    def field_op(self, curve, field, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.zeros_like(a)
        bb = np.ascontiguousarray(b, dtype=np.uint64) if b is not None else None
        rc = self.lib.orc_field_op(curve, field, op, self._p(a), self._p(bb), self._p(out), a.shape[0])
        assert rc == 0
        return out

    def to_mont(self, curve, field, a):
        return self.field_op(curve, field, 4, a)

    def from_mont(self, curve, field, a):
        return self.field_op(curve, field, 5, a)

    def constant(self, curve, what):
        out = np.zeros(12, dtype=np.uint64)
        n = self.lib.orc_constant(curve, what, self._p(out))
        assert n > 0
        return out[:n].copy()

    def g1_on_curve(self, curve, pts):
        pts = np.ascontiguousarray(pts, dtype=np.uint64)
        return self.lib.orc_g1_on_curve(curve, self._p(pts), pts.shape[0]) == 1

    def g1_mul_gen(self, curve, scalars):
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros((scalars.shape[0], 12), dtype=np.uint64)
        assert self.lib.orc_g1_mul_gen(curve, self._p(scalars), scalars.shape[0], self._p(out)) == 0
        return out

    def g1_walk(self, curve, k0, step, n):
        out = np.zeros((n, 12), dtype=np.uint64)
        assert self.lib.orc_g1_walk(curve, k0, step, n, self._p(out)) == 0
        return out

    def g1_msm(self, curve, bases, scalars, algo=0):
        bases = np.ascontiguousarray(bases, dtype=np.uint64)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros(12, dtype=np.uint64)
        assert self.lib.orc_g1_msm(curve, self._p(bases), self._p(scalars), scalars.shape[0], algo, self._p(out)) == 0
        return out

    def g1_add(self, curve, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.zeros(12, dtype=np.uint64)
        assert self.lib.orc_g1_add(curve, self._p(a), self._p(b), self._p(out)) == 0
        return out

    def ntt(self, curve, data, inverse=False, coset=False):
        out = np.ascontiguousarray(data.copy(), dtype=np.uint64)
        n = out.shape[0]
        assert self.lib.orc_ntt(curve, self._p(out), n.bit_length() - 1, int(inverse), int(coset)) == 0
        return out

    def dft_naive(self, curve, data, coset=False):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        out = np.zeros_like(data)
        assert self.lib.orc_dft_naive(curve, self._p(data), self._p(out), data.shape[0].bit_length() - 1, int(coset)) == 0
        return out

    def aes128_ecb(self, msg: bytes, key: bytes, trace=False):
        ct = ctypes.create_string_buffer(len(msg))
        tr = ctypes.create_string_buffer(len(msg) // 16 * 40 * 16) if trace else None
        rc = self.lib.orc_aes128_ecb(msg, len(msg), key, ct, tr)
        if rc != 0:
            raise ValueError("message length must be a multiple of 16")
        return (ct.raw, tr.raw) if trace else ct.raw

    def aes_step(self, name, inp: bytes, key: bytes = None):
        out = ctypes.create_string_buffer(176 if name == "derive_keys" else 16)
        fn = getattr(self.lib, "orc_aes_" + name)
        if name == "add_round_key":
            fn(inp, key, out)
        else:
            fn(inp, out)
        return out.raw

    # ---- Fr vector primitives (Montgomery (n,4) uint64 arrays) used by oracle/marlin_oracle.py ----
    def fr_vec(self, curve, op, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(a)
        assert self.lib.orc_fr_vec(curve, op, self._p(a), a.shape[0], self._p(b), b.shape[0], self._p(out)) == 0
        return out

    def fr_batch_inv(self, curve, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(a)
        assert self.lib.orc_fr_batch_inv(curve, self._p(a), self._p(out), a.shape[0]) == 0
        return out

    def fr_poly_eval(self, curve, coeffs, x):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(4)
        out = np.zeros(4, dtype=np.uint64)
        assert self.lib.orc_fr_poly_eval(curve, self._p(coeffs), coeffs.shape[0], self._p(x), self._p(out)) == 0
        return out

    def fr_powers(self, curve, base, n, c=None):
        base = np.ascontiguousarray(base, dtype=np.uint64).reshape(4)
        c = self.constant(curve, 2) if c is None else np.ascontiguousarray(c, dtype=np.uint64).reshape(4)
        out = np.zeros((n, 4), dtype=np.uint64)
        assert self.lib.orc_fr_powers(curve, self._p(base), self._p(c), n, self._p(out)) == 0
        return out

    def fr_div_linear(self, curve, coeffs, z):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        z = np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
        n = coeffs.shape[0]
        out = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
        assert self.lib.orc_fr_div_linear(curve, self._p(coeffs), n, self._p(z), self._p(out)) == 0
        return out

    def g1_mul(self, curve, pts, scalars):
        pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 12)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(pts)
        assert self.lib.orc_g1_mul(curve, self._p(pts), self._p(scalars), pts.shape[0], self._p(out)) == 0
        return out

    def g1_srs(self, curve, tau_canonical, n):
        tau = np.ascontiguousarray(tau_canonical, dtype=np.uint64).reshape(4)
        out = np.zeros((n, 12), dtype=np.uint64)
        assert self.lib.orc_g1_srs(curve, self._p(tau), n, self._p(out)) == 0
        return out
