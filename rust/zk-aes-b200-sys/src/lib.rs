//! Raw bindings to `libzkaes_b200.so`, one declaration per export of `include/zkaes_b200.h` (same order).
//!
//! Every function returns `ZKAES_OK` (0) or a negative error code and never unwinds across the boundary;
//! `zkaes_last_error` gives the message.  The safe wrapper that keeps the reference's public API
//! (`synthesize_keys` / `encrypt` / `verify_encryption`, reference `src/lib.rs:60-174`) is the `zk-aes-b200` crate.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const ZKAES_OK: c_int = 0;
pub const ZKAES_ERR_ARG: c_int = -1;
pub const ZKAES_ERR_CUDA: c_int = -2;
pub const ZKAES_ERR_STATE: c_int = -3;
pub const ZKAES_ERR_UNSUPPORTED: c_int = -4;

pub const ZKAES_CURVE_BLS12_377: c_int = 377;
pub const ZKAES_CURVE_BLS12_381: c_int = 381;

pub const ZKAES_MSM_SCALARS_MONTGOMERY: c_int = 1;
pub const ZKAES_MSM_BASES_PREPARED: c_int = 2;
pub const ZKAES_PK_FILE_SRS: c_int = 1;
pub const ZKAES_PK_FILE_INDEX_POLYS: c_int = 2;
pub const ZKAES_CIRCUIT_INFO_WORDS: usize = 18;
pub const ZKAES_PK_INFO_WORDS: usize = 12;

#[repr(C)]
pub struct zkaes_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zkaes_pk {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zkaes_circuit {
    _private: [u8; 0],
}

/// `zkaes_proof_fields.commitments[i]`
#[repr(C)]
#[derive(Clone, Copy)]
pub struct zkaes_commitment {
    pub comm: [u8; 96],
    pub has_shifted: u8,
    pub shifted: [u8; 96],
}
/// `zkaes_proof_fields.openings[i]`
#[repr(C)]
#[derive(Clone, Copy)]
pub struct zkaes_opening {
    pub w: [u8; 96],
    pub has_random_v: u8,
    pub random_v: [u8; 32],
}
/// Plain fields of `ark_marlin::Proof` (what `deserialize_proof` gives a Rust caller of the reference).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct zkaes_proof_fields {
    pub n_rounds: u32,
    pub round_sizes: [u32; 3],
    pub commitments: [zkaes_commitment; 9],
    pub n_evaluations: u32,
    pub evaluations: [[u8; 32]; 7],
    pub n_openings: u32,
    pub openings: [zkaes_opening; 2],
}

extern "C" {
    // ---- context
    pub fn zkaes_ctx_create(device_id: c_int, out: *mut *mut zkaes_ctx) -> c_int;
    pub fn zkaes_ctx_create_multi(device_ids: *const c_int, n_devices: c_int, out: *mut *mut zkaes_ctx) -> c_int;
    pub fn zkaes_ctx_devices(ctx: *const zkaes_ctx) -> c_int;
    pub fn zkaes_ctx_destroy(ctx: *mut zkaes_ctx);
    pub fn zkaes_last_error(ctx: *const zkaes_ctx) -> *const c_char;
    pub fn zkaes_ctx_stream(ctx: *mut zkaes_ctx) -> *mut c_void;
    pub fn zkaes_ctx_launches(ctx: *const zkaes_ctx) -> u64;
    pub fn zkaes_ctx_sync(ctx: *mut zkaes_ctx) -> c_int;
    pub fn zkaes_comm_unique_id(out128: *mut u8) -> c_int;
    pub fn zkaes_ctx_comm_init(ctx: *mut zkaes_ctx, rank: c_int, nranks: c_int, unique_id128: *const u8) -> c_int;
    pub fn zkaes_shard_range(n: usize, rank: c_int, nranks: c_int, start: *mut usize, count: *mut usize) -> c_int;
    pub fn zkaes_coset_plan(nranks: c_int, ncoset: c_int, ntask: c_int, own_extra: f64, owner_out: *mut c_int, exec_out: *mut c_int) -> c_int;
    pub fn zkaes_ctx_profile(ctx: *mut zkaes_ctx, enable: c_int) -> c_int;
    pub fn zkaes_ctx_profile_read(ctx: *mut zkaes_ctx, out4: *mut f64) -> c_int;
    pub fn zkaes_ctx_set_msm_window(ctx: *mut zkaes_ctx, window_bits: c_int) -> c_int;
    pub fn zkaes_ctx_set_tuning(ctx: *mut zkaes_ctx, key: *const c_char, value: c_int) -> c_int;
    // ---- device memory
    pub fn zkaes_dev_alloc(ctx: *mut zkaes_ctx, bytes: usize, out_dev: *mut *mut c_void) -> c_int;
    pub fn zkaes_dev_free(ctx: *mut zkaes_ctx, dev: *mut c_void) -> c_int;
    pub fn zkaes_dev_upload(ctx: *mut zkaes_ctx, dev: *mut c_void, host: *const c_void, bytes: usize) -> c_int;
    pub fn zkaes_dev_download(ctx: *mut zkaes_ctx, host: *mut c_void, dev: *const c_void, bytes: usize) -> c_int;
    // ---- MSM (ark-ec VariableBaseMSM seam)
    pub fn zkaes_msm_g1(ctx: *mut zkaes_ctx, curve_id: c_int, bases_host: *const c_void, scalars_host: *const c_void, n: usize, out_affine96: *mut c_void) -> c_int;
    pub fn zkaes_msm_g1_small(ctx: *mut zkaes_ctx, curve_id: c_int, bases_host: *const c_void, values_host: *const i32, n: usize, value_bits: c_int, out_affine96: *mut c_void) -> c_int;
    pub fn zkaes_msm_g1_device(ctx: *mut zkaes_ctx, curve_id: c_int, bases_dev: *const c_void, scalars_dev: *const c_void, n: usize, flags: c_int, out_affine96_host: *mut c_void) -> c_int;
    pub fn zkaes_msm_g1_prepare_bases(ctx: *mut zkaes_ctx, curve_id: c_int, bases_dev: *mut c_void, n: usize) -> c_int;
    pub fn zkaes_msm_g1_windows_bytes(ctx: *mut zkaes_ctx, curve_id: c_int, n_total: usize) -> usize;
    pub fn zkaes_msm_g1_windows(ctx: *mut zkaes_ctx, curve_id: c_int, bases_dev: *const c_void, scalars_dev: *const c_void, n_local: usize, n_total: usize, flags: c_int, windows_dev: *mut c_void) -> c_int;
    pub fn zkaes_msm_g1_fold(ctx: *mut zkaes_ctx, curve_id: c_int, gathered_windows_dev: *const c_void, n_ranks: c_int, n_total: usize, out_affine96_host: *mut c_void) -> c_int;
    // ---- NTT (ark-poly Radix2EvaluationDomain seam)
    pub fn zkaes_ntt_fr(ctx: *mut zkaes_ctx, curve_id: c_int, data_host: *mut c_void, log_n: u32, inverse: c_int, coset: c_int) -> c_int;
    pub fn zkaes_ntt_fr_device(ctx: *mut zkaes_ctx, curve_id: c_int, data_dev: *mut c_void, log_n: u32, inverse: c_int, coset: c_int) -> c_int;
    // ---- test SRS
    pub fn zkaes_srs_powers_device(ctx: *mut zkaes_ctx, curve_id: c_int, seed32: *const u8, n: usize, out_bases_dev: *mut c_void) -> c_int;
    // ---- self tests
    pub fn zkaes_selftest_field(ctx: *mut zkaes_ctx, curve_id: c_int, field: c_int, op: c_int, variant: c_int, a_host: *const c_void, b_host: *const c_void, out_host: *mut c_void, count: usize) -> c_int;
    pub fn zkaes_selftest_g1(ctx: *mut zkaes_ctx, curve_id: c_int, op: c_int, a_host: *const c_void, b_host: *const c_void, out_host: *mut c_void, count: usize) -> c_int;
    pub fn zkaes_selftest_host_field(curve_id: c_int, field: c_int, op: c_int, a: *const c_void, b: *const c_void, out: *mut c_void, count: usize) -> c_int;
    pub fn zkaes_selftest_host_g1(curve_id: c_int, op: c_int, a: *const c_void, b: *const c_void, out: *mut c_void, count: usize) -> c_int;
    pub fn zkaes_selftest_pairing(a32: *const u8, b32: *const u8, out576: *mut u8) -> c_int;
    // ---- circuit shape (host only)
    pub fn zkaes_circuit_build(msg_len: usize, out: *mut *mut zkaes_circuit) -> c_int;
    pub fn zkaes_circuit_free(c: *mut zkaes_circuit);
    pub fn zkaes_circuit_info(c: *const zkaes_circuit, info: *mut u64) -> c_int;
    pub fn zkaes_circuit_matrix(c: *const zkaes_circuit, which: c_int, row_ptr: *mut u32, col: *mut u32, coeff: *mut i8) -> c_int;
    // ---- K1: witness generation
    pub fn zkaes_witness_aes128_ecb(ctx: *mut zkaes_ctx, c: *const zkaes_circuit, msg: *const u8, msg_len: usize, key16: *const u8, ct_out: *mut u8, assignment_out: *mut u8) -> c_int;
    // ---- keys and encrypt() (reference src/lib.rs:138-174, 60-114)
    pub fn zkaes_synthesize_keys(ctx: *mut zkaes_ctx, plaintext_len: usize, tau_seed32: *const u8, gamma_seed32: *const u8, out: *mut *mut zkaes_pk) -> c_int;
    pub fn zkaes_pk_free(pk: *mut zkaes_pk);
    pub fn zkaes_pk_save(ctx: *mut zkaes_ctx, pk: *const zkaes_pk, path: *const c_char, flags: c_int) -> c_int;
    pub fn zkaes_pk_load(ctx: *mut zkaes_ctx, path: *const c_char, out: *mut *mut zkaes_pk) -> c_int;
    pub fn zkaes_pk_info(pk: *const zkaes_pk, info: *mut u64) -> c_int;
    pub fn zkaes_pk_vk_bytes(pk: *const zkaes_pk, out: *mut u8, len: *mut usize) -> c_int;
    pub fn zkaes_encrypt(ctx: *mut zkaes_ctx, pk: *const zkaes_pk, msg: *const u8, msg_len: usize, key16: *const u8, zk_seed32: *const u8, ct_out: *mut u8, proof_out: *mut u8, proof_len: *mut usize) -> c_int;
    // ---- verify_encryption (reference src/lib.rs:116-136) and the proof wire format (src/lib.rs:52)
    pub fn zkaes_pk_verifying_key(pk: *const zkaes_pk, out: *mut u8, len: *mut usize) -> c_int;
    pub fn zkaes_verify_encryption(vk: *const u8, vk_len: usize, proof: *const u8, proof_len: usize, ciphertext: *const u8, ct_len: usize, accepted: *mut c_int) -> c_int;
    pub fn zkaes_proof_deserialize(proof: *const u8, proof_len: usize, out: *mut zkaes_proof_fields) -> c_int;
    pub fn zkaes_proof_serialize(fields: *const zkaes_proof_fields, out: *mut u8, len: *mut usize) -> c_int;
}
