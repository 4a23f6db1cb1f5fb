/* libzkaes_b200 -- C ABI of the B200-native prover path for lambdaclass/AES_zero_knowledge_proof_circuit.
 *
 * This is the drop-in boundary.  The reference is a pure-Rust crate with no FFI of its own
 * (`#![forbid(unsafe_code)]`, src/lib.rs:2); each entry point below names the reference interface it stands
 * in for.  A Rust `-sys` shim binds exactly these symbols (INTEGRATION.md shows it); tests and bench.py bind
 * them through ctypes.
 *
 * Conventions: return 0 on success, negative on error (never aborts, never throws across the boundary);
 * zkaes_last_error(ctx) gives the message.  All pointers are caller-owned except opaque handles, which are
 * released by their matching *_free / *_destroy.  "host" pointers may be pageable.  Calls are blocking
 * (internally asynchronous on the context's stream).  A context is not re-entrant.  Multi-GPU: either one process drives
 * all GPUs through zkaes_ctx_create_multi, or one process (one context) per GPU joins a communicator with
 * zkaes_ctx_comm_init; the stand-alone MSM entry points exchange their small partials with an all-gather between
 * zkaes_msm_g1_windows and zkaes_msm_g1_fold.
 *
 * Wire formats (identical to arkworks 0.3.0 in-memory layouts, little-endian):
 *   Fr element : 32 B, 4 x u64 limbs, Montgomery form (R = 2^256)             [ark-ff Fp256]
 *   scalar     : 32 B, 4 x u64 limbs, canonical integer (`into_repr()`)       [ark-ff BigInteger256]
 *   G1 affine  : 96 B, x || y, each 6 x u64 limbs Montgomery (R = 2^384); x = y = 0 encodes infinity
 *                (ark-ec's separate `infinity: bool` is mapped by the shim)   [ark-ec GroupAffine]
 * curve_id: 377 = BLS12-377 (the reference's proving curve, src/lib.rs:47), 381 = BLS12-381.
 */
#ifndef ZKAES_B200_H
#define ZKAES_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZKAES_OK 0
#define ZKAES_ERR_ARG (-1)
#define ZKAES_ERR_CUDA (-2)
#define ZKAES_ERR_STATE (-3)
#define ZKAES_ERR_UNSUPPORTED (-4)

#define ZKAES_CURVE_BLS12_377 377
#define ZKAES_CURVE_BLS12_381 381

typedef struct zkaes_ctx zkaes_ctx;

/* ---- context ------------------------------------------------------------------------------------------ */
/* Creates a context on CUDA device `device_id` (fails loudly if there is no usable GPU: there is no CPU path). */
int zkaes_ctx_create(int device_id, zkaes_ctx** out);
/* One context, several GPUs, ONE process (`zkaes_ctx_create(const int* device_ids, int n_devices, ..)`): the
 * returned context is rank 0 on device_ids[0] and owns one peer context per further device, each driven by its own host thread,
 * plus an in-process NCCL communicator.  zkaes_synthesize_keys / zkaes_encrypt / zkaes_pk_save / zkaes_pk_load on such a context
 * run on every rank at once (MSMs sharded by point, as described below) and return rank 0's result after checking that every
 * rank produced the same bytes -- so the reference's single-call `encrypt()` (src/lib.rs:60-64) drives all GPUs of the box.
 * Key files of a multi-GPU context are written per rank as `path.r<rank>`.  n_devices = 1 is zkaes_ctx_create(device_ids[0]).
 * The MSM / NTT / witness entry points act on rank 0's device only. */
int zkaes_ctx_create_multi(const int* device_ids, int n_devices, zkaes_ctx** out);
/* number of GPUs (ranks) behind the context */
int zkaes_ctx_devices(const zkaes_ctx* ctx);
void zkaes_ctx_destroy(zkaes_ctx* ctx);
const char* zkaes_last_error(const zkaes_ctx* ctx);
/* The CUDA stream the context launches on (a cudaStream_t), so callers can record events on it. */
void* zkaes_ctx_stream(zkaes_ctx* ctx);
/* Number of kernels this library has launched on the context so far (bench.py's gpu_launches). */
uint64_t zkaes_ctx_launches(const zkaes_ctx* ctx);
/* Blocks until all work queued on the context's stream has finished. */
int zkaes_ctx_sync(zkaes_ctx* ctx);
/* ---- multi-GPU (one process and one context per GPU) ---------------------------------------------------------------
 * The prover's MSMs shard by point: rank r keeps the SRS points i = r (mod N) (cyclic, so polynomials of every length
 * spread evenly), runs the bucket method on them and the per-rank window sums (W x 192 B) are all-gathered over NCCL and
 * folded on every rank; witness generation and NTTs stay
 * per GPU (every rank computes them redundantly, so transcripts stay identical).  Rank 0 obtains an id with
 * zkaes_comm_unique_id and distributes it out of band (bench.py: torch.distributed broadcast); every rank then calls
 * zkaes_ctx_comm_init BEFORE zkaes_synthesize_keys.  All ranks must issue the same sequence of key / encrypt calls.
 * zkaes_shard_range is the contiguous point-range rule for callers that shard their own zkaes_msm_g1_windows inputs
 * (host only). */
int zkaes_comm_unique_id(uint8_t out128[128]);
int zkaes_ctx_comm_init(zkaes_ctx* ctx, int rank, int nranks, const uint8_t unique_id128[128]);
int zkaes_shard_range(size_t n, int rank, int nranks, size_t* start, size_t* count);
/* How the prover splits the coset evaluations of rounds 2 and 3 over the ranks (host only; exposed for tests and for reading the
 * phase traces): ncoset cosets of ntask forward transforms each plus own_extra transforms' worth of owner-only work.
 * owner_out[ncoset]: the rank that assembles coset j; exec_out[ncoset * ntask]: the rank that computes transform p of coset j
 * (it sends the result to the owner when it is not the owner).  Ranks that hand work out never take work in. */
int zkaes_coset_plan(int nranks, int ncoset, int ntask, double own_extra, int* owner_out, int* exec_out);

/* Per-kernel timing of the dominant kernel (the MSM bucket accumulation) for bench.py's roofline: when enabled, every
 * launch is bracketed by CUDA events on the context's stream.  profile_read synchronises and returns
 * out = {launches, total ms, total MSM terms, total mixed additions (upper bound)} since the last read, then resets. */
int zkaes_ctx_profile(zkaes_ctx* ctx, int enable);
int zkaes_ctx_profile_read(zkaes_ctx* ctx, double out[4]);
/* Tuning: force the MSM window width (0 = automatic). */
int zkaes_ctx_set_msm_window(zkaes_ctx* ctx, int window_bits);
/* Tuning knobs that never change results: "msm_window_max" (cap of the automatic window choice, 3..24; the bucket array
 * is 2^(c-1) * ceil(253/c) points of 192 B), "msm_acc_blocks" (resident blocks per SM of the bucket accumulation: 3 or 4),
 * "msm_pair_round" (R = 0..4: add the entries of every bucket in pairs as affine points with a shared inversion, R times
 * over, before the XYZZ accumulation; 0 = plain accumulation). */
/* Further keys: "msm_madd_call" (0 / 1: ten inlined products per mixed addition / ten calls of one out-of-line multiplier, the default),
 * "msm_prefetch" (0 / 1 / 2: stage the next entry's point in shared memory with cp.async / cp.async.bulk + mbarrier; measured no gain,
 * default 0), "msm_plan_ranks" (the number of ranks that share a zkaes_msm_g1_windows / zkaes_msm_g1_fold MSM, so that the window plan
 * fits the per-rank share; every rank must set the same value; default 1), "r1_lagrange" (1 / 0: zkaes_encrypt commits to w, z_A, z_B
 * through the key's Lagrange-basis points with one small digit per term, or through the SRS powers as ark-marlin does; same commitments,
 * same proof bytes; default 1, without effect on a key that holds no such points -- zkaes_pk_info word 11). */
int zkaes_ctx_set_tuning(zkaes_ctx* ctx, const char* key, int value);

/* ---- device memory (thin wrappers so non-CUDA hosts can keep inputs resident in HBM) -------------------- */
int zkaes_dev_alloc(zkaes_ctx* ctx, size_t bytes, void** out_dev);
int zkaes_dev_free(zkaes_ctx* ctx, void* dev);
int zkaes_dev_upload(zkaes_ctx* ctx, void* dev, const void* host, size_t bytes);
int zkaes_dev_download(zkaes_ctx* ctx, void* host, const void* dev, size_t bytes);

/* ---- S4 seam: MSM --------------------------------------------------------------------------------------
 * Stands in for ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul(&[G1Affine], &[BigInteger256]) -> G1Projective`
 * (reference Cargo.lock:118-120, reached from src/lib.rs:111).  Result is returned in affine form.
 */
/* host buffers in, 96-byte affine result out (host).
 * Bases must lie in the prime-order subgroup G1 (as every arkworks G1Affine does) and scalars must be canonical (< r):
 * the kernels fold s to min(s, r - s) with the point negated, which presupposes r P = O. */
int zkaes_msm_g1(zkaes_ctx* ctx, int curve_id, const void* bases_host, const void* scalars_host, size_t n, void* out_affine96);
/* device-resident inputs.  flags: bit 0 (ZKAES_MSM_SCALARS_MONTGOMERY) = the scalars are Fr elements in Montgomery form
 * (as the prover's coefficient vectors are) and are converted on the fly; bit 1 (ZKAES_MSM_BASES_PREPARED) = the bases
 * were rewritten in place by zkaes_msm_g1_prepare_bases. */
#define ZKAES_MSM_SCALARS_MONTGOMERY 1
#define ZKAES_MSM_BASES_PREPARED 2
int zkaes_msm_g1_device(zkaes_ctx* ctx, int curve_id, const void* bases_dev, const void* scalars_dev, size_t n,
                        int flags, void* out_affine96_host);
/* Small signed scalars: sum_i values[i] * bases[i] for |values[i]| <= 2^(value_bits - 1), value_bits in 1..13 (host buffers).
 * One signed digit per term is the whole scalar, so the bucket method runs a single pass of at most n mixed additions (the
 * general entry points spend W = 11 windows on a 253-bit scalar).  The kernel path of the prover's Lagrange-basis
 * commitments to w, z_A, z_B (ark-marlin 0.3.0 ahp/prover.rs first round, whose evaluations over H are bits and sums of a
 * few bits), exported so that it can be checked against zkaes_msm_g1 term by term. */
int zkaes_msm_g1_small(zkaes_ctx* ctx, int curve_id, const void* bases_host, const int32_t* values_host, size_t n, int value_bits,
                       void* out_affine96);
/* Kept for ABI stability: the kernels read the arkworks form directly (12 x u32 Montgomery limbs per coordinate), so
 * "preparing" bases is the identity and ZKAES_MSM_BASES_PREPARED changes nothing. */
int zkaes_msm_g1_prepare_bases(zkaes_ctx* ctx, int curve_id, void* bases_dev, size_t n);
/* multi-GPU split: (1) per-rank window sums of this rank's point range, written to a device buffer of
 * zkaes_msm_g1_windows_bytes(n_total) bytes; the window plan is derived from n_total so all ranks agree.
 * (2) after an all-gather of those buffers, fold n_ranks sets into the affine result (host). */
size_t zkaes_msm_g1_windows_bytes(zkaes_ctx* ctx, int curve_id, size_t n_total);
int zkaes_msm_g1_windows(zkaes_ctx* ctx, int curve_id, const void* bases_dev, const void* scalars_dev, size_t n_local,
                         size_t n_total, int flags, void* windows_dev);
int zkaes_msm_g1_fold(zkaes_ctx* ctx, int curve_id, const void* gathered_windows_dev, int n_ranks, size_t n_total,
                      void* out_affine96_host);

/* ---- S4 seam: NTT --------------------------------------------------------------------------------------
 * Stands in for ark-poly 0.3.0 `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place(&mut Vec<Fr>)`
 * (reference Cargo.lock:234-236).  In place, natural order in and out, n = 2^log_n Montgomery Fr elements.
 */
int zkaes_ntt_fr(zkaes_ctx* ctx, int curve_id, void* data_host, uint32_t log_n, int inverse, int coset);
int zkaes_ntt_fr_device(zkaes_ctx* ctx, int curve_id, void* data_dev, uint32_t log_n, int inverse, int coset);

/* ---- synthetic G1 bases (test SRS) ---------------------------------------------------------------------
 * Stands in for the KZG10 setup inside simpleworks::marlin::generate_universal_srs (src/lib.rs:141):
 * out[i] = tau^i * G for i < n, tau derived from `seed32` -- an INSECURE test SRS exactly like the
 * reference's (README.md:26).  Output stays on the device (n x 96 B). */
int zkaes_srs_powers_device(zkaes_ctx* ctx, int curve_id, const uint8_t seed32[32], size_t n, void* out_bases_dev);

/* ---- on-device self test of the field / curve arithmetic (used by tests/, not by the product path) -------
 * field: 0 = Fr, 1 = Fq.  op: 0 add, 1 sub, 2 mul, 3 square of a (Fq on the device: the generated dedicated squaring).  variant: 0 = generated PTX multiplier (inlined), 1 = portable CIOS,
 * 2 = the out-of-line copy of the PTX multiplier that the MSM inner loop and the curve formulas call,
 * 3 = the FP64-limb product of csrc/fq52.cuh (an experiment: curve 377, field 1, op 2 only; out = a b 2^-416 mod q, i.e. its own
 *     Montgomery radix -- tests compare against big integers).
 * a, b, out are host arrays of `count` elements (32 B or 48 B each). */
int zkaes_selftest_field(zkaes_ctx* ctx, int curve_id, int field, int op, int variant, const void* a_host, const void* b_host,
                         void* out_host, size_t count);
/* out[i] = a[i] + b[i] on G1 (affine 96 B each) through the device XYZZ formulas: op 0 = mixed add, 1 = full add,
 * 2 = double a[i] */
int zkaes_selftest_g1(zkaes_ctx* ctx, int curve_id, int op, const void* a_host, const void* b_host, void* out_host, size_t count);

/* The same templates executed on the host CPU (the host uses them for the O(W*c) window fold and the affine
 * normalisation).  No context / GPU needed.  field op: 0 add, 1 sub, 2 mul, 3 inverse(a), 4 neg(a). */
int zkaes_selftest_host_field(int curve_id, int field, int op, const void* a, const void* b, void* out, size_t count);
int zkaes_selftest_host_g1(int curve_id, int op, const void* a, const void* b, void* out, size_t count);

/* ---- circuit shape (host only, no GPU needed) ------------------------------------------------------------------
 * The R1CS that the reference obtains by running `encrypt_and_generate_constraints` (src/lib.rs:176-293) against
 * arkworks' constraint system, after Marlin's padding.  Used by zkaes_synthesize_keys; exposed so that tests can
 * compare the shape with the oracle's gadget model (variable order, rows of A/B/C, witness program).
 * info[]: 0 msg_len, 1 n_blocks, 2 num_instance (padded), 3 num_instance_used, 4 num_witness, 5 num_witness_real,
 *         6 num_constraints, 7 nnz(A), 8 nnz(B), 9 nnz(C), 10 first key-bit witness, 11 first key-schedule witness,
 *         12 first block witness, 13 witnesses per block, 14 fixed-program instrs, 15 block-program instrs,
 *         16 fixed-program levels, 17 block-program levels */
typedef struct zkaes_circuit zkaes_circuit;
#define ZKAES_CIRCUIT_INFO_WORDS 18
int zkaes_circuit_build(size_t msg_len, zkaes_circuit** out);
void zkaes_circuit_free(zkaes_circuit* c);
int zkaes_circuit_info(const zkaes_circuit* c, uint64_t info[ZKAES_CIRCUIT_INFO_WORDS]);
/* which: 0 = A, 1 = B, 2 = C.  row_ptr: num_constraints + 1 entries; col / coeff: nnz entries. */
int zkaes_circuit_matrix(const zkaes_circuit* c, int which, uint32_t* row_ptr, uint32_t* col, int8_t* coeff);

/* ---- K1: AES-128-ECB witness generation -----------------------------------------------------------------------
 * Stands in for the value side of the reference's circuit synthesis in encrypt() (src/lib.rs:66-98, 176-293): runs
 * the block program of `c` for every 16-byte block on the device.  msg_len must equal the circuit's length.
 * ct_out (msg_len bytes, host) receives the ciphertext; assignment_out (nullable, host, num_instance + num_witness
 * bytes) receives the full variable assignment, one byte per variable (all variables of this circuit are bits), in
 * R1CS column order [instance | witness]. */
int zkaes_witness_aes128_ecb(zkaes_ctx* ctx, const zkaes_circuit* c, const uint8_t* msg, size_t msg_len, const uint8_t key[16],
                             uint8_t* ct_out, uint8_t* assignment_out);

/* ---- S1/S3 seam: keys and encrypt() ------------------------------------------------------------------------------
 * zkaes_synthesize_keys stands in for `synthesize_keys(plaintext_length)` (src/lib.rs:138-174): test SRS from the two
 * seeds (tau, gamma: INSECURE, like the reference's -- README.md:26), circuit shape, Marlin index, all left resident in
 * HBM behind the opaque handle.  Unlike the reference's hard-coded bounds (src/lib.rs:141) the SRS is sized for the
 * requested length.  Limits of this build: |K| = next_pow2(max nnz) <= 2^28 (messages up to 8 KiB; larger lengths return
 * ZKAES_ERR_UNSUPPORTED), and the circuit must have nnz(A) < nnz(B) (ark-marlin's balance_matrices is then the identity; true for
 * every length of this circuit, checked at key synthesis).
 * zkaes_encrypt stands in for `encrypt(message, secret_key, proving_key)` (src/lib.rs:60-114): witness generation and
 * the Marlin proof.  `zk_seed32` seeds the prover's zero-knowledge randomness (the reference draws it from
 * simpleworks::marlin::generate_rand(); an explicit seed makes runs reproducible).  ct_out receives msg_len bytes.
 * proof_out may be NULL to query the size; *proof_len is in/out (capacity in, bytes written out).  The bytes are the
 * ark-serialize 0.3.0 CanonicalSerialize form of ark_marlin::Proof (what `deserialize_proof`, src/lib.rs:52, reads).
 * info[]: 0 msg_len, 1 num_constraints, 2 num_variables, 3-5 nnz(A,B,C), 6 |H|, 7 |K|, 8 |X|, 9 SRS max degree,
 *         10 instance variables used, 11 Lagrange-basis points held per basis on this rank (0 = round 1 commits through the
 *         SRS powers; see the tuning key "r1_lagrange"). */
typedef struct zkaes_pk zkaes_pk;
#define ZKAES_PK_INFO_WORDS 12
int zkaes_synthesize_keys(zkaes_ctx* ctx, size_t plaintext_len, const uint8_t tau_seed32[32], const uint8_t gamma_seed32[32], zkaes_pk** out);
void zkaes_pk_free(zkaes_pk* pk);
/* Key files (the reference regenerates SRS and keys in every process, src/lib.rs:138-174, and its
 * ProvingKey never leaves the process).  zkaes_pk_save writes the key of THIS rank to `path`; flags: bit 0 (ZKAES_PK_FILE_SRS) =
 * include the rank's SRS share, bit 1 (ZKAES_PK_FILE_INDEX_POLYS) = include the 12 index polynomials.  With neither the file is
 * ~3 KB (seeds of the test SRS, sizes, the 12 index commitments, the verifying key).  zkaes_pk_load rebuilds the key on the
 * context's device: the circuit shape and matrices from the message length, whatever the file lacks by recomputation (SRS
 * from its seeds, index polynomials from the matrices) -- the twelve |K|-term commitment MSMs of synthesize_keys are never
 * repeated -- and checks the file's verifying key against the one its contents give.  Multi-GPU: one file per rank; a file
 * only loads on a context with the rank / world size it was saved from. */
#define ZKAES_PK_FILE_SRS 1
#define ZKAES_PK_FILE_INDEX_POLYS 2
int zkaes_pk_save(zkaes_ctx* ctx, const zkaes_pk* pk, const char* path, int flags);
int zkaes_pk_load(zkaes_ctx* ctx, const char* path, zkaes_pk** out);
int zkaes_pk_info(const zkaes_pk* pk, uint64_t info[ZKAES_PK_INFO_WORDS]);
/* Verifying-key bytes as they enter the Fiat-Shamir transcript: index info (3 x u64 LE) || 12 index commitments
 * (ark-ff ToBytes of marlin_pc::Commitment, 195 bytes each).  out may be NULL to query the size. */
int zkaes_pk_vk_bytes(const zkaes_pk* pk, uint8_t* out, size_t* len);
int zkaes_encrypt(zkaes_ctx* ctx, const zkaes_pk* pk, const uint8_t* msg, size_t msg_len, const uint8_t key[16], const uint8_t zk_seed32[32],
                  uint8_t* ct_out, uint8_t* proof_out, size_t* proof_len);

/* ---- S1 seam: verify_encryption ------------------------------------------------------------------------------------
 * zkaes_pk_verifying_key exports the VerifyingKey half of `synthesize_keys`' result (src/lib.rs:138,173) as the ark-serialize
 * 0.3.0 CanonicalSerialize bytes of ark_marlin::IndexVerifierKey<Fr, MarlinKZG10<Bls12_377, DensePolynomial<Fr>>>: index_info
 * (4 x u64), the 12 index commitments (compressed G1, no shifted part), marlin_pc::VerifierKey {kzg10 vk: g, gamma_g (compressed
 * G1), h, beta_h (compressed G2); degree bounds and shift powers; max_degree; supported_degree} -- layout in csrc/verifier.cpp.
 * out may be NULL to query the size.
 * Statement length: like ark-marlin 0.3.0 (which derives the input domain from public_input.len() + 1 and zero-pads the input
 * itself), a ciphertext is a statement of the key iff next_pow2(8 * ct_len + 1) equals the key's input-domain size; callers that
 * know the key's message length should compare lengths themselves, as the reference's callers must.
 * zkaes_verify_encryption stands in for `verify_encryption(verifying_key, proof, ciphertext)` (src/lib.rs:116-136): the
 * ciphertext becomes 8 public-input bits per byte (src/helpers/mod.rs:84-93), the Marlin verifier replays the
 * transcript and checks the two KZG openings with a BLS12-377 pairing.  HOST ONLY -- no context, no device (the
 * reference verifies on the CPU too).  Returns 0 with *accepted = 1 / 0 (Ok(true) / Ok(false)); a key or proof that
 * cannot be parsed returns ZK_ERR_ARG (Err(..)) and leaves the reason in zkaes_last_error(NULL). */
int zkaes_pk_verifying_key(const zkaes_pk* pk, uint8_t* out, size_t* len);
int zkaes_verify_encryption(const uint8_t* vk, size_t vk_len, const uint8_t* proof, size_t proof_len, const uint8_t* ciphertext, size_t ct_len,
                            int* accepted);
/* ---- proof wire format (src/lib.rs:52 re-exports simpleworks' serialize_proof / deserialize_proof) -----------------------
 * The proof bytes zkaes_encrypt emits are the ark-serialize 0.3.0 CanonicalSerialize form of ark_marlin::Proof (compressed
 * G1 points).  zkaes_proof_deserialize unpacks them into plain fields -- what `deserialize_proof(bytes) -> MarlinProof` gives
 * a Rust caller -- and zkaes_proof_serialize packs them again (byte-identical round trip).  HOST ONLY.
 * Points: x || y, 48 + 48 bytes canonical little-endian (not Montgomery); the point at infinity is all zero.  Field elements:
 * 32 bytes canonical little-endian.  Commitments are in protocol order: w, z_a, z_b, mask | t, g_1, h_1 | g_2, h_2. */
typedef struct zkaes_proof_fields {
    uint32_t n_rounds;               /* 3 */
    uint32_t round_sizes[3];         /* 4, 3, 2 */
    struct {
        uint8_t comm[96];
        uint8_t has_shifted;         /* degree-bounded polynomials (g_1, g_2) carry a second commitment */
        uint8_t shifted[96];
    } commitments[9];
    uint32_t n_evaluations;          /* 7: a_denom b_denom c_denom g_1 g_2 t z_b (sorted by label) */
    uint8_t evaluations[7][32];
    uint32_t n_openings;             /* 2: at beta, at gamma */
    struct {
        uint8_t w[96];
        uint8_t has_random_v;
        uint8_t random_v[32];
    } openings[2];
} zkaes_proof_fields;
int zkaes_proof_deserialize(const uint8_t* proof, size_t proof_len, zkaes_proof_fields* out);
int zkaes_proof_serialize(const zkaes_proof_fields* in, uint8_t* out, size_t* len);  /* out may be NULL to query the size */

/* Test hook: e(a G1, b G2) for canonical 32-byte LE scalars, as 12 x 48 canonical LE bytes (tools/pairing_model.py layout). */
int zkaes_selftest_pairing(const uint8_t a32[32], const uint8_t b32[32], uint8_t out576[576]);

#ifdef __cplusplus
}
#endif
#endif /* ZKAES_B200_H */
