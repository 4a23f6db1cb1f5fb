"""ctypes binding of libzkaes_b200.so (the C ABI in include/zkaes_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the symbol
table, and if no CUDA device is usable `Context()` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkaes_b200.so")

CURVE_BLS12_377 = 377
CURVE_BLS12_381 = 381


class ZkAesError(RuntimeError):
    """Non-zero status from the C ABI; mirrors the reference's `anyhow::Result` error style (src/helpers/traits.rs:4-20)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ZkAesError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the prover path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp = c_void_p
    sigs = {
        "zkaes_ctx_create": (c_int, [c_int, POINTER(vp)]),
        "zkaes_ctx_destroy": (None, [vp]),
        "zkaes_coset_plan": (c_int, [c_int, c_int, c_int, ctypes.c_double, vp, vp]),
        "zkaes_ctx_create_multi": (c_int, [vp, c_int, POINTER(vp)]),
        "zkaes_ctx_devices": (c_int, [vp]),
        "zkaes_last_error": (c_char_p, [vp]),
        "zkaes_ctx_stream": (vp, [vp]),
        "zkaes_ctx_launches": (c_uint64, [vp]),
        "zkaes_ctx_sync": (c_int, [vp]),
        "zkaes_ctx_set_msm_window": (c_int, [vp, c_int]),
        "zkaes_ctx_set_tuning": (c_int, [vp, ctypes.c_char_p, c_int]),
        "zkaes_comm_unique_id": (c_int, [vp]),
        "zkaes_ctx_comm_init": (c_int, [vp, c_int, c_int, vp]),
        "zkaes_shard_range": (c_int, [c_size_t, c_int, c_int, POINTER(c_size_t), POINTER(c_size_t)]),
        "zkaes_ctx_profile": (c_int, [vp, c_int]),
        "zkaes_ctx_profile_read": (c_int, [vp, vp]),
        "zkaes_dev_alloc": (c_int, [vp, c_size_t, POINTER(vp)]),
        "zkaes_dev_free": (c_int, [vp, vp]),
        "zkaes_dev_upload": (c_int, [vp, vp, vp, c_size_t]),
        "zkaes_dev_download": (c_int, [vp, vp, vp, c_size_t]),
        "zkaes_msm_g1": (c_int, [vp, c_int, vp, vp, c_size_t, vp]),
        "zkaes_msm_g1_device": (c_int, [vp, c_int, vp, vp, c_size_t, c_int, vp]),
        "zkaes_msm_g1_small": (c_int, [vp, c_int, vp, vp, c_size_t, c_int, vp]),
        "zkaes_msm_g1_prepare_bases": (c_int, [vp, c_int, vp, c_size_t]),
        "zkaes_msm_g1_windows_bytes": (c_size_t, [vp, c_int, c_size_t]),
        "zkaes_msm_g1_windows": (c_int, [vp, c_int, vp, vp, c_size_t, c_size_t, c_int, vp]),
        "zkaes_msm_g1_fold": (c_int, [vp, c_int, vp, c_int, c_size_t, vp]),
        "zkaes_ntt_fr": (c_int, [vp, c_int, vp, c_uint32, c_int, c_int]),
        "zkaes_ntt_fr_device": (c_int, [vp, c_int, vp, c_uint32, c_int, c_int]),
        "zkaes_srs_powers_device": (c_int, [vp, c_int, vp, c_size_t, vp]),
        "zkaes_selftest_field": (c_int, [vp, c_int, c_int, c_int, c_int, vp, vp, vp, c_size_t]),
        "zkaes_selftest_g1": (c_int, [vp, c_int, c_int, vp, vp, vp, c_size_t]),
        "zkaes_selftest_host_field": (c_int, [c_int, c_int, c_int, vp, vp, vp, c_size_t]),
        "zkaes_selftest_host_g1": (c_int, [c_int, c_int, vp, vp, vp, c_size_t]),
        "zkaes_circuit_build": (c_int, [c_size_t, POINTER(vp)]),
        "zkaes_circuit_free": (None, [vp]),
        "zkaes_circuit_info": (c_int, [vp, vp]),
        "zkaes_circuit_matrix": (c_int, [vp, c_int, vp, vp, vp]),
        "zkaes_witness_aes128_ecb": (c_int, [vp, vp, vp, c_size_t, vp, vp, vp]),
        "zkaes_synthesize_keys": (c_int, [vp, c_size_t, vp, vp, POINTER(vp)]),
        "zkaes_pk_free": (None, [vp]),
        "zkaes_pk_save": (c_int, [vp, vp, ctypes.c_char_p, c_int]),
        "zkaes_pk_load": (c_int, [vp, ctypes.c_char_p, POINTER(vp)]),
        "zkaes_pk_info": (c_int, [vp, vp]),
        "zkaes_pk_vk_bytes": (c_int, [vp, vp, POINTER(c_size_t)]),
        "zkaes_encrypt": (c_int, [vp, vp, vp, c_size_t, vp, vp, vp, vp, POINTER(c_size_t)]),
        "zkaes_pk_verifying_key": (c_int, [vp, vp, POINTER(c_size_t)]),
        "zkaes_verify_encryption": (c_int, [vp, c_size_t, vp, c_size_t, vp, c_size_t, POINTER(c_int)]),
        "zkaes_selftest_pairing": (c_int, [vp, vp, vp]),
        "zkaes_proof_deserialize": (c_int, [vp, c_size_t, vp]),
        "zkaes_proof_serialize": (c_int, [vp, vp, POINTER(c_size_t)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    lib._zk_symbols = tuple(sigs)
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _ptr(a):
    """Pointer to a numpy array / bytes-like / int address."""
    if a is None:
        return None
    if isinstance(a, int):
        return c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return c_void_p(a.ctypes.data)
    if isinstance(a, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(a)), c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor
        return c_void_p(a.data_ptr())
    raise TypeError(type(a))


def comm_unique_id() -> bytes:
    buf = np.zeros(128, dtype=np.uint8)
    rc = lib().zkaes_comm_unique_id(_ptr(buf))
    if rc != 0:
        raise ZkAesError(f"zkaes_comm_unique_id failed with {rc} (is libnccl.so.2 loadable?)")
    return buf.tobytes()


def shard_range(n: int, rank: int, nranks: int):
    """point range [start, start + count) of `rank` in an n-point MSM sharded over nranks GPUs"""
    s, c = c_size_t(0), c_size_t(0)
    if lib().zkaes_shard_range(n, rank, nranks, ctypes.byref(s), ctypes.byref(c)) != 0:
        raise ZkAesError("zkaes_shard_range: bad arguments")
    return s.value, c.value


def coset_plan(nranks: int, ncoset: int, ntask: int, own_extra: float):
    """-> (owner[ncoset], exec[ncoset][ntask]): zkaes_coset_plan, the work split of the prover's coset evaluations over the ranks"""
    owner = np.zeros(ncoset, dtype=np.int32)
    ex = np.zeros(ncoset * ntask, dtype=np.int32)
    if lib().zkaes_coset_plan(nranks, ncoset, ntask, float(own_extra), _ptr(owner), _ptr(ex)) != 0:
        raise ZkAesError("zkaes_coset_plan: bad arguments")
    return owner.tolist(), ex.reshape(ncoset, ntask).tolist()


class Context:
    """One prover context: one CUDA device (`Context(0)`, one per process / rank), or several devices driven by this one process
    (`Context([0, 1, 2, 3])`, zkaes_ctx_create_multi: worker threads + an in-process NCCL communicator)."""

    def __init__(self, device=0):
        self._h = c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (c_int * len(device))(*device)
            rc = lib().zkaes_ctx_create_multi(ctypes.cast(ids, c_void_p), len(device), ctypes.byref(self._h))
            what = f"zkaes_ctx_create_multi(devices={list(device)})"
            self.device = device[0] if device else None
        else:
            rc = lib().zkaes_ctx_create(device, ctypes.byref(self._h))
            what = f"zkaes_ctx_create(device={device})"
            self.device = device
        if rc != 0:
            raise ZkAesError(f"{what} failed with {rc}: B200 (sm_100) GPUs are required; there is no CPU path")

    @property
    def n_devices(self) -> int:
        return int(lib().zkaes_ctx_devices(self._h))

    def close(self):
        if self._h:
            lib().zkaes_ctx_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = lib().zkaes_last_error(self._h)
            raise ZkAesError(f"libzkaes_b200 error {rc}: {msg.decode() if msg else ''}")

    @property
    def stream(self) -> int:
        return lib().zkaes_ctx_stream(self._h) or 0

    @property
    def launches(self) -> int:
        return lib().zkaes_ctx_launches(self._h)

    def sync(self):
        self._check(lib().zkaes_ctx_sync(self._h))

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        """join the NCCL communicator of the sharded MSM (call before synthesize_keys); unique_id from comm_unique_id() on rank 0"""
        buf = np.frombuffer(bytes(unique_id), dtype=np.uint8) if unique_id is not None else None
        self._check(lib().zkaes_ctx_comm_init(self._h, rank, nranks, _ptr(buf)))

    def profile(self, enable: bool):
        self._check(lib().zkaes_ctx_profile(self._h, int(enable)))

    def profile_read(self):
        """-> dict(launches, ms, terms, madds) of the MSM bucket-accumulation kernel since the last read"""
        out = np.zeros(4, dtype=np.float64)
        self._check(lib().zkaes_ctx_profile_read(self._h, _ptr(out)))
        return {"launches": int(out[0]), "ms": float(out[1]), "terms": float(out[2]), "madds": float(out[3])}

    def set_msm_window(self, bits: int):
        self._check(lib().zkaes_ctx_set_msm_window(self._h, bits))

    def set_tuning(self, key: str, value: int):
        self._check(lib().zkaes_ctx_set_tuning(self._h, key.encode(), value))

    # ---- device memory ----
    def alloc(self, nbytes: int) -> int:
        p = c_void_p()
        self._check(lib().zkaes_dev_alloc(self._h, nbytes, ctypes.byref(p)))
        return p.value

    def free(self, dev: int):
        self._check(lib().zkaes_dev_free(self._h, c_void_p(dev)))

    def upload(self, dev: int, host: np.ndarray):
        self._check(lib().zkaes_dev_upload(self._h, c_void_p(dev), _ptr(host), host.nbytes))

    def download(self, host: np.ndarray, dev: int):
        self._check(lib().zkaes_dev_download(self._h, _ptr(host), c_void_p(dev), host.nbytes))

    # ---- MSM ----
    def msm_g1(self, curve: int, bases: np.ndarray, scalars: np.ndarray) -> np.ndarray:
        """bases: (n, 12) uint64 affine Montgomery; scalars: (n, 4) uint64 canonical. Returns (12,) uint64 affine."""
        n = scalars.shape[0]
        assert bases.shape == (n, 12) and scalars.shape == (n, 4)
        out = np.zeros(12, dtype=np.uint64)
        self._check(lib().zkaes_msm_g1(self._h, curve, _ptr(bases), _ptr(scalars), n, _ptr(out)))
        return out

    def msm_g1_small(self, curve: int, bases: np.ndarray, values: np.ndarray, value_bits: int) -> np.ndarray:
        """sum values[i] * bases[i] for int32 values with |v| <= 2^(value_bits - 1): the single-pass path of the Lagrange-basis commitments"""
        n = values.shape[0]
        assert bases.shape == (n, 12) and values.dtype == np.int32
        out = np.zeros(12, dtype=np.uint64)
        self._check(lib().zkaes_msm_g1_small(self._h, curve, _ptr(bases), _ptr(values), n, value_bits, _ptr(out)))
        return out

    def msm_g1_device(self, curve: int, bases_dev, scalars_dev, n: int, scalars_montgomery: bool = False, bases_prepared: bool = False) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        flags = int(scalars_montgomery) | (2 if bases_prepared else 0)
        self._check(lib().zkaes_msm_g1_device(self._h, curve, _ptr(bases_dev), _ptr(scalars_dev), n, flags, _ptr(out)))
        return out

    def msm_g1_prepare_bases(self, curve: int, bases_dev, n: int):
        """rewrite device-resident bases in place into the MSM kernels' internal form (do once per SRS)"""
        self._check(lib().zkaes_msm_g1_prepare_bases(self._h, curve, _ptr(bases_dev), n))

    def msm_g1_windows_bytes(self, curve: int, n_total: int) -> int:
        return lib().zkaes_msm_g1_windows_bytes(self._h, curve, n_total)

    def msm_g1_windows(self, curve: int, bases_dev, scalars_dev, n_local: int, n_total: int, windows_dev, scalars_montgomery: bool = False,
                       bases_prepared: bool = False):
        flags = int(scalars_montgomery) | (2 if bases_prepared else 0)
        self._check(lib().zkaes_msm_g1_windows(self._h, curve, _ptr(bases_dev), _ptr(scalars_dev), n_local, n_total, flags, _ptr(windows_dev)))

    def msm_g1_fold(self, curve: int, gathered_dev, n_ranks: int, n_total: int) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        self._check(lib().zkaes_msm_g1_fold(self._h, curve, _ptr(gathered_dev), n_ranks, n_total, _ptr(out)))
        return out

    # ---- NTT ----
    def ntt_fr(self, curve: int, data: np.ndarray, inverse: bool = False, coset: bool = False) -> np.ndarray:
        """data: (n, 4) uint64 Montgomery, n a power of two; transformed copy is returned."""
        n = data.shape[0]
        log_n = n.bit_length() - 1
        assert n == 1 << log_n and data.shape == (n, 4)
        out = np.ascontiguousarray(data.copy())
        self._check(lib().zkaes_ntt_fr(self._h, curve, _ptr(out), log_n, int(inverse), int(coset)))
        return out

    def ntt_fr_device(self, curve: int, data_dev, log_n: int, inverse: bool = False, coset: bool = False):
        self._check(lib().zkaes_ntt_fr_device(self._h, curve, _ptr(data_dev), log_n, int(inverse), int(coset)))

    # ---- SRS ----
    def srs_powers_device(self, curve: int, seed32: bytes, n: int, out_dev):
        assert len(seed32) == 32
        buf = (c_uint8 * 32).from_buffer_copy(seed32)
        self._check(lib().zkaes_srs_powers_device(self._h, curve, ctypes.cast(buf, c_void_p), n, _ptr(out_dev)))

    # ---- K1: witness generation ----
    def witness_aes128_ecb(self, circuit: "Circuit", msg: bytes, key: bytes, want_assignment: bool = True):
        """-> (ciphertext bytes, assignment uint8 array [instance | witness] or None)"""
        assert len(key) == 16
        ct = np.zeros(len(msg), dtype=np.uint8)
        nvar = circuit.info["num_instance"] + circuit.info["num_witness"]
        z = np.zeros(nvar, dtype=np.uint8) if want_assignment else None
        m = np.frombuffer(bytes(msg), dtype=np.uint8)
        k = np.frombuffer(bytes(key), dtype=np.uint8)
        self._check(lib().zkaes_witness_aes128_ecb(self._h, circuit._h, _ptr(m), len(msg), _ptr(k), _ptr(ct), _ptr(z)))
        return ct.tobytes(), z

    # ---- self tests ----
    def selftest_field(self, curve: int, field: int, op: int, variant: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        out = np.zeros_like(a)
        self._check(lib().zkaes_selftest_field(self._h, curve, field, op, variant, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out

    def selftest_g1(self, curve: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        out = np.zeros_like(a)
        self._check(lib().zkaes_selftest_g1(self._h, curve, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out


CIRCUIT_INFO_FIELDS = ("msg_len", "n_blocks", "num_instance", "num_instance_used", "num_witness", "num_witness_real", "num_constraints",
                       "nnz_a", "nnz_b", "nnz_c", "wit_key0", "wit_fixed0", "wit_block0", "wit_block_stride", "fixed_instrs", "block_instrs",
                       "fixed_levels", "block_levels")


class Circuit:
    """Shape of the AES-128-ECB R1CS for a message length (host only; no GPU needed)."""

    def __init__(self, msg_len: int):
        self._h = c_void_p()
        rc = lib().zkaes_circuit_build(msg_len, ctypes.byref(self._h))
        if rc != 0:
            raise ZkAesError(f"zkaes_circuit_build({msg_len}) failed with {rc}: message length must be a non-zero multiple of 16")
        info = np.zeros(len(CIRCUIT_INFO_FIELDS), dtype=np.uint64)
        _ok(lib().zkaes_circuit_info(self._h, _ptr(info)), "zkaes_circuit_info")
        self.info = {k: int(v) for k, v in zip(CIRCUIT_INFO_FIELDS, info)}

    def matrix(self, which: int):
        """-> (row_ptr, col, coeff) CSR arrays of A (0), B (1) or C (2)"""
        nnz = self.info[("nnz_a", "nnz_b", "nnz_c")[which]]
        row_ptr = np.zeros(self.info["num_constraints"] + 1, dtype=np.uint32)
        col = np.zeros(nnz, dtype=np.uint32)
        coeff = np.zeros(nnz, dtype=np.int8)
        _ok(lib().zkaes_circuit_matrix(self._h, which, _ptr(row_ptr), _ptr(col), _ptr(coeff)), "zkaes_circuit_matrix")
        return row_ptr, col, coeff

    def close(self):
        if self._h:
            lib().zkaes_circuit_free(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PK_INFO_FIELDS = ("msg_len", "num_constraints", "num_variables", "nnz_a", "nnz_b", "nnz_c", "h", "k", "x", "max_degree", "num_instance_used",
                  "lagrange_points")


def _ok(rc, what):
    """ABI calls without a context to ask for the message: a non-zero return code is an error, never an `assert`"""
    if rc != 0:
        raise ZkAesError(f"libzkaes_b200: {what} returned {rc}")


class ProvingKey:
    """Device-resident proving key (SRS powers, matrices, index polynomials) for one plaintext length."""

    def __init__(self, ctx: Context, handle):
        self._ctx = ctx
        self._h = handle
        info = np.zeros(len(PK_INFO_FIELDS), dtype=np.uint64)
        _ok(lib().zkaes_pk_info(self._h, _ptr(info)), "zkaes_pk_info")
        self.info = {k: int(v) for k, v in zip(PK_INFO_FIELDS, info)}

    def _bytes_of(self, fn, what) -> bytes:
        n = c_size_t(0)
        _ok(fn(self._h, None, ctypes.byref(n)), what)
        buf = np.zeros(n.value, dtype=np.uint8)
        _ok(fn(self._h, _ptr(buf), ctypes.byref(n)), what)
        return buf.tobytes()

    def vk_bytes(self) -> bytes:
        return self._bytes_of(lib().zkaes_pk_vk_bytes, "zkaes_pk_vk_bytes")

    def verifying_key(self) -> bytes:
        """The VerifyingKey half of synthesize_keys' result (src/lib.rs:138): what verify_encryption takes (ark-serialize bytes)."""
        return self._bytes_of(lib().zkaes_pk_verifying_key, "zkaes_pk_verifying_key")

    def save(self, path: str, srs: bool = True, index_polys: bool = True):
        """zkaes_pk_save: this rank's key file; without the two bulk sections it is ~3 KB and load recomputes them."""
        self._ctx._check(lib().zkaes_pk_save(self._ctx._h, self._h, os.fsencode(path), (1 if srs else 0) | (2 if index_polys else 0)))

    def close(self):
        if self._h:
            lib().zkaes_pk_free(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _seed(b):
    assert len(b) == 32
    return np.frombuffer(bytes(b), dtype=np.uint8)


def _synthesize_keys(self, plaintext_length: int, tau_seed: bytes, gamma_seed: bytes) -> ProvingKey:
    h = c_void_p()
    self._check(lib().zkaes_synthesize_keys(self._h, plaintext_length, _ptr(_seed(tau_seed)), _ptr(_seed(gamma_seed)), ctypes.byref(h)))
    return ProvingKey(self, h)


def _load_keys(self, path: str) -> ProvingKey:
    """zkaes_pk_load: a key saved by ProvingKey.save(), rebuilt on this context's device without the commitment MSMs"""
    h = c_void_p()
    self._check(lib().zkaes_pk_load(self._h, os.fsencode(path), ctypes.byref(h)))
    return ProvingKey(self, h)


def _encrypt(self, pk: ProvingKey, message: bytes, secret_key: bytes, zk_seed: bytes):
    """-> (ciphertext bytes, proof bytes)"""
    assert len(secret_key) == 16
    n = c_size_t(0)
    m = np.frombuffer(bytes(message), dtype=np.uint8)
    k = np.frombuffer(bytes(secret_key), dtype=np.uint8)
    ct = np.zeros(max(len(message), 1), dtype=np.uint8)
    self._check(lib().zkaes_encrypt(self._h, pk._h, _ptr(m), len(message), _ptr(k), _ptr(_seed(zk_seed)), _ptr(ct), None, ctypes.byref(n)))
    proof = np.zeros(n.value, dtype=np.uint8)
    self._check(lib().zkaes_encrypt(self._h, pk._h, _ptr(m), len(message), _ptr(k), _ptr(_seed(zk_seed)), _ptr(ct), _ptr(proof), ctypes.byref(n)))
    return ct[: len(message)].tobytes(), proof[: n.value].tobytes()


def verify_encryption(verifying_key: bytes, proof: bytes, ciphertext: bytes) -> bool:
    """`verify_encryption(verifying_key, proof, ciphertext) -> Result<bool>` (reference src/lib.rs:116-136).  Host only: no
    Context / GPU needed.  Raises ZkAesError when the key or the proof cannot be parsed (the reference's Err)."""
    vk = np.frombuffer(bytes(verifying_key), dtype=np.uint8)
    pf = np.frombuffer(bytes(proof), dtype=np.uint8)
    ct = np.frombuffer(bytes(ciphertext) or b"\0", dtype=np.uint8)
    ok = c_int(0)
    rc = lib().zkaes_verify_encryption(_ptr(vk), len(verifying_key), _ptr(pf), len(proof), _ptr(ct), len(ciphertext), ctypes.byref(ok))
    if rc != 0:
        msg = lib().zkaes_last_error(None)
        raise ZkAesError(f"libzkaes_b200 error {rc}: {msg.decode() if msg else ''}")
    return bool(ok.value)


class _Commitment(ctypes.Structure):
    _fields_ = [("comm", c_uint8 * 96), ("has_shifted", c_uint8), ("shifted", c_uint8 * 96)]


class _Opening(ctypes.Structure):
    _fields_ = [("w", c_uint8 * 96), ("has_random_v", c_uint8), ("random_v", c_uint8 * 32)]


class ProofFields(ctypes.Structure):
    """include/zkaes_b200.h: zkaes_proof_fields"""
    _fields_ = [("n_rounds", c_uint32), ("round_sizes", c_uint32 * 3), ("commitments", _Commitment * 9), ("n_evaluations", c_uint32),
                ("evaluations", (c_uint8 * 32) * 7), ("n_openings", c_uint32), ("openings", _Opening * 2)]


def _host_check(rc):
    if rc != 0:
        msg = lib().zkaes_last_error(None)
        raise ZkAesError(f"libzkaes_b200 error {rc}: {msg.decode() if msg else ''}")


def deserialize_proof(proof: bytes) -> ProofFields:
    """`deserialize_proof(bytes) -> MarlinProof` (re-exported at reference src/lib.rs:52): the proof's commitments (uncompressed),
    evaluations and opening proofs as plain fields.  Host only."""
    out = ProofFields()
    pf = np.frombuffer(bytes(proof), dtype=np.uint8)
    _host_check(lib().zkaes_proof_deserialize(_ptr(pf), len(proof), ctypes.byref(out)))
    return out


def serialize_proof(fields: ProofFields) -> bytes:
    """`serialize_proof(proof) -> bytes`: the ark-serialize 0.3.0 compressed form zkaes_encrypt emits"""
    n = c_size_t(0)
    _host_check(lib().zkaes_proof_serialize(ctypes.byref(fields), None, ctypes.byref(n)))
    buf = np.zeros(n.value, dtype=np.uint8)
    _host_check(lib().zkaes_proof_serialize(ctypes.byref(fields), _ptr(buf), ctypes.byref(n)))
    return buf[: n.value].tobytes()


def pairing_selftest(a: int, b: int) -> bytes:
    """e(a G1, b G2) as 576 canonical bytes (test hook, compared with tools/pairing_model.py)."""
    out = np.zeros(576, dtype=np.uint8)
    ab = np.frombuffer(int(a).to_bytes(32, "little"), dtype=np.uint8)
    bb = np.frombuffer(int(b).to_bytes(32, "little"), dtype=np.uint8)
    _ok(lib().zkaes_selftest_pairing(_ptr(ab), _ptr(bb), _ptr(out)), "zkaes_selftest_pairing")
    return out.tobytes()


Context.synthesize_keys = _synthesize_keys
Context.load_keys = _load_keys
Context.encrypt = _encrypt
