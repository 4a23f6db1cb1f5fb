//! `zk_aes` on NVIDIA B200: the public API of lambdaclass/AES_zero_knowledge_proof_circuit (reference `src/lib.rs:60-174`)
//! over `libzkaes_b200.so`.
//!
//! The three entry points keep the reference's signatures, so a caller switches crates without touching call sites:
//!
//! ```text
//! pub fn synthesize_keys(plaintext_length: usize) -> Result<(ProvingKey, VerifyingKey)>          // src/lib.rs:138
//! pub fn encrypt(message: &[u8], secret_key: &[u8; 16], proving_key: ProvingKey) -> Result<MarlinProof>   // src/lib.rs:60-64
//! pub fn verify_encryption(verifying_key: VerifyingKey, proof: &MarlinProof, ciphertext: &[u8]) -> Result<bool>   // src/lib.rs:116-120
//! ```
//!
//! Differences a caller can observe, all additive:
//! * `ProvingKey` is a handle to a key that stays resident in GPU memory; `clone()` is a reference count, not the multi-GB copy
//!   the reference's benchmark pays per iteration (`benches/benchmark_encrypt.rs:46`).
//! * `MarlinProof` carries the ark-serialize 0.3.0 bytes of `ark_marlin::Proof` (`serialize_proof` / `deserialize_proof`,
//!   reference `src/lib.rs:52`) and, next to them, the ciphertext the circuit computed (`MarlinProof::ciphertext`).
//! * GPUs are chosen with `ZKAES_DEVICES` (comma-separated CUDA ordinals, default `0`); with several, one `encrypt()` call
//!   shards the prover's multi-scalar multiplications over all of them (`zkaes_ctx_create_multi`).
//! * The zero-knowledge randomness comes from `rand::thread_rng()` per call (the reference: `simpleworks::marlin::generate_rand()`);
//!   `encrypt_with_seed` makes a proof reproducible.  The SRS is the same kind of INSECURE test SRS as the reference's
//!   (`README.md:26`), derived from the fixed seeds below.
//!
//! There is no CPU fallback: without a B200 every prover call returns an error.
#![deny(unsafe_op_in_unsafe_fn)]

use anyhow::{anyhow, ensure, Result};
use rand::RngCore;
use std::ffi::{CStr, CString};
use std::os::raw::c_int;
use std::ptr;
use std::sync::{Arc, Mutex, OnceLock};
use zk_aes_b200_sys as sys;

/// Seeds of the test SRS trapdoors (tau, gamma).  INSECURE by construction, like the reference's `generate_universal_srs` on a
/// fresh rng: anyone who knows them can forge proofs.  A production deployment loads keys made from a ceremony SRS instead.
pub const TEST_SRS_TAU_SEED: [u8; 32] = [
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
];
pub const TEST_SRS_GAMMA_SEED: [u8; 32] = [
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
];

// ---------------------------------------------------------------------------------------------------------------------
// context: one per process, created on first use
// ---------------------------------------------------------------------------------------------------------------------
struct Context {
    raw: *mut sys::zkaes_ctx,
}
// The library's context is not re-entrant; every use goes through the Mutex below.
unsafe impl Send for Context {}

impl Drop for Context {
    fn drop(&mut self) {
        // SAFETY: `raw` came from zkaes_ctx_create(_multi) and is destroyed exactly once.
        unsafe { sys::zkaes_ctx_destroy(self.raw) }
    }
}

fn last_error(ctx: *const sys::zkaes_ctx) -> String {
    // SAFETY: zkaes_last_error returns a NUL-terminated string owned by the library (or by the context), valid until the next call.
    unsafe {
        let p = sys::zkaes_last_error(ctx);
        if p.is_null() {
            String::from("unknown error")
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

fn check(ctx: *const sys::zkaes_ctx, rc: c_int, what: &str) -> Result<()> {
    if rc == sys::ZKAES_OK {
        Ok(())
    } else {
        Err(anyhow!("{what}: libzkaes_b200 error {rc}: {}", last_error(ctx)))
    }
}

fn devices_from_env() -> Result<Vec<c_int>> {
    let spec = std::env::var("ZKAES_DEVICES").unwrap_or_else(|_| String::from("0"));
    let mut out = Vec::new();
    for part in spec.split(',') {
        let part = part.trim();
        if part.is_empty() {
            continue;
        }
        out.push(part.parse::<c_int>().map_err(|e| anyhow!("ZKAES_DEVICES: {part:?}: {e}"))?);
    }
    ensure!(!out.is_empty(), "ZKAES_DEVICES names no device");
    Ok(out)
}

fn context() -> Result<&'static Mutex<Context>> {
    static CTX: OnceLock<std::result::Result<Mutex<Context>, String>> = OnceLock::new();
    let slot = CTX.get_or_init(|| {
        let devices = devices_from_env().map_err(|e| e.to_string())?;
        let mut raw: *mut sys::zkaes_ctx = ptr::null_mut();
        // SAFETY: `devices` outlives the call; `raw` receives an owned context on success.
        let rc = unsafe { sys::zkaes_ctx_create_multi(devices.as_ptr(), devices.len() as c_int, &mut raw) };
        if rc != sys::ZKAES_OK || raw.is_null() {
            return Err(format!(
                "zkaes_ctx_create_multi({devices:?}) failed with {rc}: B200 (sm_100) GPUs are required, there is no CPU path"
            ));
        }
        Ok(Mutex::new(Context { raw }))
    });
    slot.as_ref().map_err(|e| anyhow!("{e}"))
}

// ---------------------------------------------------------------------------------------------------------------------
// key and proof types (the reference's aliases, src/lib.rs:53-56)
// ---------------------------------------------------------------------------------------------------------------------
struct PkInner {
    raw: *mut sys::zkaes_pk,
    msg_len: usize,
}
// The handle is only dereferenced while the context Mutex is held.
unsafe impl Send for PkInner {}
unsafe impl Sync for PkInner {}

impl Drop for PkInner {
    fn drop(&mut self) {
        // SAFETY: `raw` came from zkaes_synthesize_keys / zkaes_pk_load and is freed exactly once.
        unsafe { sys::zkaes_pk_free(self.raw) }
    }
}

/// The proving key: SRS share, circuit matrices and index polynomials, resident in GPU memory.  Cloning shares the key.
#[derive(Clone)]
pub struct ProvingKey(Arc<PkInner>);

impl ProvingKey {
    /// Length in bytes of the messages this key proves (`synthesize_keys`' argument).
    pub fn plaintext_length(&self) -> usize {
        self.0.msg_len
    }

    /// Writes the key to `path` (`zkaes_pk_save`; with several GPUs one file per rank, `path.r<rank>`).  `bulk` includes the SRS
    /// share and the index polynomials; without them the file is ~3 KB and `load` recomputes them.
    pub fn save(&self, path: &str, bulk: bool) -> Result<()> {
        let ctx = context()?.lock().map_err(|_| anyhow!("context mutex poisoned"))?;
        let cpath = CString::new(path)?;
        let flags = if bulk { sys::ZKAES_PK_FILE_SRS | sys::ZKAES_PK_FILE_INDEX_POLYS } else { 0 };
        // SAFETY: context and key handles are live; `cpath` is NUL-terminated.
        let rc = unsafe { sys::zkaes_pk_save(ctx.raw, self.0.raw, cpath.as_ptr(), flags) };
        check(ctx.raw, rc, "zkaes_pk_save")
    }

    /// Rebuilds a key from a file written by `save` -- without the twelve commitment MSMs `synthesize_keys` pays.
    pub fn load(path: &str) -> Result<(ProvingKey, VerifyingKey)> {
        let ctx = context()?.lock().map_err(|_| anyhow!("context mutex poisoned"))?;
        let cpath = CString::new(path)?;
        let mut raw: *mut sys::zkaes_pk = ptr::null_mut();
        // SAFETY: as above; `raw` receives an owned key on success.
        let rc = unsafe { sys::zkaes_pk_load(ctx.raw, cpath.as_ptr(), &mut raw) };
        check(ctx.raw, rc, "zkaes_pk_load")?;
        finish_key(raw)
    }
}

/// The verifying key: ark-serialize 0.3.0 bytes of `ark_marlin::IndexVerifierKey<Fr, MarlinKZG10<Bls12_377, _>>`.
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct VerifyingKey(pub Vec<u8>);

/// A Marlin proof of one encryption: ark-serialize 0.3.0 bytes of `ark_marlin::Proof`, plus the ciphertext the circuit computed.
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct MarlinProof {
    bytes: Vec<u8>,
    ciphertext: Vec<u8>,
}

impl MarlinProof {
    /// The AES-128-ECB ciphertext of the message (the statement the proof is about).  Empty for a proof obtained from bytes.
    pub fn ciphertext(&self) -> &[u8] {
        &self.ciphertext
    }
    pub fn as_bytes(&self) -> &[u8] {
        &self.bytes
    }
}

/// `simpleworks::marlin::serialization::serialize_proof` (re-exported by the reference, `src/lib.rs:52`).
pub fn serialize_proof(proof: &MarlinProof) -> Result<Vec<u8>> {
    Ok(proof.bytes.clone())
}

/// `simpleworks::marlin::serialization::deserialize_proof` (`src/lib.rs:52`): validates the encoding strictly (curve points,
/// subgroup, canonical field elements, this protocol's shape) and keeps the bytes.
pub fn deserialize_proof(bytes: Vec<u8>) -> Result<MarlinProof> {
    let mut fields = std::mem::MaybeUninit::<sys::zkaes_proof_fields>::zeroed();
    // SAFETY: `bytes` is a live slice; `fields` is writable storage of the right size (plain-old-data, zero is a valid value).
    let rc = unsafe { sys::zkaes_proof_deserialize(bytes.as_ptr(), bytes.len(), fields.as_mut_ptr()) };
    check(ptr::null(), rc, "deserialize_proof")?;
    Ok(MarlinProof { bytes, ciphertext: Vec::new() })
}

fn query_bytes(f: impl Fn(*mut u8, *mut usize) -> c_int, what: &str) -> Result<Vec<u8>> {
    let mut len: usize = 0;
    ensure!(f(ptr::null_mut(), &mut len) == sys::ZKAES_OK, "{what}: size query failed");
    let mut out = vec![0u8; len];
    ensure!(f(out.as_mut_ptr(), &mut len) == sys::ZKAES_OK, "{what}: read failed");
    out.truncate(len);
    Ok(out)
}

fn finish_key(raw: *mut sys::zkaes_pk) -> Result<(ProvingKey, VerifyingKey)> {
    let mut info = [0u64; sys::ZKAES_PK_INFO_WORDS];
    // SAFETY: `raw` is a live key handle; `info` has ZKAES_PK_INFO_WORDS entries.
    let rc = unsafe { sys::zkaes_pk_info(raw, info.as_mut_ptr()) };
    let inner = PkInner { raw, msg_len: info[0] as usize };
    ensure!(rc == sys::ZKAES_OK, "zkaes_pk_info failed with {rc}");
    // SAFETY: the closure passes the library a buffer of the length it reported.
    let vk = query_bytes(|p, n| unsafe { sys::zkaes_pk_verifying_key(inner.raw, p, n) }, "zkaes_pk_verifying_key")?;
    Ok((ProvingKey(Arc::new(inner)), VerifyingKey(vk)))
}

// ---------------------------------------------------------------------------------------------------------------------
// the reference's three entry points
// ---------------------------------------------------------------------------------------------------------------------

/// Reference `src/lib.rs:138-174`: test SRS, circuit shape, Marlin index; the proving key stays on the GPU(s).
pub fn synthesize_keys(plaintext_length: usize) -> Result<(ProvingKey, VerifyingKey)> {
    ensure!(plaintext_length > 0 && plaintext_length % 16 == 0, "plaintext length must be a positive multiple of 16 bytes (AES-128-ECB blocks)");
    let ctx = context()?.lock().map_err(|_| anyhow!("context mutex poisoned"))?;
    let mut raw: *mut sys::zkaes_pk = ptr::null_mut();
    // SAFETY: the seeds are 32-byte arrays; `raw` receives an owned key on success.
    let rc = unsafe {
        sys::zkaes_synthesize_keys(ctx.raw, plaintext_length, TEST_SRS_TAU_SEED.as_ptr(), TEST_SRS_GAMMA_SEED.as_ptr(), &mut raw)
    };
    check(ctx.raw, rc, "synthesize_keys")?;
    finish_key(raw)
}

/// Reference `src/lib.rs:60-114`: AES-128-ECB witness generation and the Marlin proof, on the GPU(s).  Takes the key by value, as
/// the reference does; clone the (reference-counted) key first to keep using it.
pub fn encrypt(message: &[u8], secret_key: &[u8; 16], proving_key: ProvingKey) -> Result<MarlinProof> {
    let mut seed = [0u8; 32];
    rand::thread_rng().fill_bytes(&mut seed);
    encrypt_with_seed(message, secret_key, &proving_key, &seed)
}

/// `encrypt` with caller-chosen zero-knowledge randomness: the same inputs give the same proof bytes.
pub fn encrypt_with_seed(message: &[u8], secret_key: &[u8; 16], proving_key: &ProvingKey, zk_seed: &[u8; 32]) -> Result<MarlinProof> {
    ensure!(
        message.len() == proving_key.0.msg_len,
        "message is {} bytes, the proving key was synthesised for {}",
        message.len(),
        proving_key.0.msg_len
    );
    let ctx = context()?.lock().map_err(|_| anyhow!("context mutex poisoned"))?;
    let mut ciphertext = vec![0u8; message.len()];
    let mut len: usize = 0;
    // SAFETY: a null proof buffer asks for the size only.
    let rc = unsafe {
        sys::zkaes_encrypt(ctx.raw, proving_key.0.raw, message.as_ptr(), message.len(), secret_key.as_ptr(), zk_seed.as_ptr(), ciphertext.as_mut_ptr(), ptr::null_mut(), &mut len)
    };
    check(ctx.raw, rc, "encrypt (size query)")?;
    let mut bytes = vec![0u8; len];
    // SAFETY: all buffers are live and of the lengths passed; `ciphertext` holds message.len() bytes.
    let rc = unsafe {
        sys::zkaes_encrypt(ctx.raw, proving_key.0.raw, message.as_ptr(), message.len(), secret_key.as_ptr(), zk_seed.as_ptr(), ciphertext.as_mut_ptr(), bytes.as_mut_ptr(), &mut len)
    };
    check(ctx.raw, rc, "encrypt")?;
    bytes.truncate(len);
    Ok(MarlinProof { bytes, ciphertext })
}

/// Reference `src/lib.rs:116-136`: the ciphertext becomes 8 public-input bits per byte (`src/helpers/mod.rs:84-93`) and the
/// Marlin verifier runs on the host CPU (no GPU, no context).  `Ok(false)` = proof rejected; `Err` = key or proof unparsable.
pub fn verify_encryption(verifying_key: VerifyingKey, proof: &MarlinProof, ciphertext: &[u8]) -> Result<bool> {
    let mut accepted: c_int = 0;
    // SAFETY: the three slices are live for the duration of the call.
    let rc = unsafe {
        sys::zkaes_verify_encryption(
            verifying_key.0.as_ptr(),
            verifying_key.0.len(),
            proof.bytes.as_ptr(),
            proof.bytes.len(),
            ciphertext.as_ptr(),
            ciphertext.len(),
            &mut accepted,
        )
    };
    check(ptr::null(), rc, "verify_encryption")?;
    Ok(accepted == 1)
}
