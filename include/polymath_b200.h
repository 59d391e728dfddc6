/* polymath_b200 — C ABI of the B200 (sm_100a) prover backend for sigma0-dev/polymath.
 *
 * The reference is pure Rust with no FFI of its own; these entry points
 * are what a `polymath-b200-sys` crate would bind to replace the
 * bodies of the reference's private helpers.  Each declaration cites the reference
 * interface it replaces (file:line under sigma0-dev/polymath).
 *
 * Conventions
 *  - every function returns 0 (PM_OK) on success, a PM_ERR_* code otherwise; the
 *    message of the last failure on the calling thread is available from pm_last_error();
 *    no C++ exception crosses the boundary;
 *  - the caller owns all host buffers; the library owns device memory behind opaque handles;
 *  - Fr elements are 32 bytes, Fq elements 48 bytes: little-endian limbs of the
 *    MONTGOMERY form (R = 2^256 / 2^384) — byte-identical to arkworks' in-memory
 *    `Fp<MontBackend<..>, 4|6>`;
 *  - G1 affine points are read with a caller-given stride: x (48 B) then y (48 B); the
 *    point at infinity is (0,0) for stride 96, or flagged by a non-zero byte at offset
 *    96 for stride >= 104 (arkworks `Affine { x, y, infinity }`).  Points are written
 *    packed (96 B, infinity = (0,0)) and are the canonical affine image;
 *  - calls are thread-safe: the contexts of a process share one CUDA stream and one set of kernel
 *    workspaces, so the library serialises its entry points internally (one proof at a time per GPU —
 *    a single proof already fills the device); a context's phases must still be called in order;
 *  - all work runs on the CUDA device current to the calling thread (one process per GPU).
 */
#ifndef POLYMATH_B200_H
#define POLYMATH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PM_OK = 0,
    PM_ERR_CUDA = 1,        /* CUDA runtime failure */
    PM_ERR_ARG = 2,         /* invalid argument */
    PM_ERR_UNSATISFIED = 3, /* witness does not satisfy the SAP: reference panics at src/prover.rs:108 */
    PM_ERR_DEGENERATE = 4,  /* h == 0 or deg h > n-2: reference panics at src/prover.rs:107 */
    PM_ERR_REMAINDER = 5,   /* opening remainder non-zero: reference panics at src/prover.rs:221 */
    PM_ERR_STATE = 6        /* prove phases called out of order */
};

#define PM_FR_BYTES 32
#define PM_FQ_BYTES 48
#define PM_G1_BYTES 96
#define PM_G1_COMPRESSED_BYTES 48 /* zcash / ark-bls12-381 compressed G1 */
#define PM_XYZZ_BYTES 192 /* (X, Y, ZZ, ZZZ) partial sum of a sharded MSM; x = X/ZZ, y = Y/ZZZ */

const char* pm_last_error(void);
/* ABI version of this header; bumped on any incompatible change. */
int pm_abi_version(void);
/* Optional, before the first CUDA call of the process on `device` (-1 = current): sets cudaDeviceLmemResizeToMax so
 * that the driver keeps the local-memory pool of the context at its high-water mark instead of shrinking and regrowing
 * it between kernels with different stack frames (a device-wide synchronisation inside a phase).  Returns 1 when the
 * flag is in place, 0 when the context already existed without it, < 0 on error. */
int pm_runtime_configure(int device);
/* Number of CUDA devices visible; <= 0 means the library cannot run (no CPU fallback exists). */
int pm_device_count(void);
/* Kernels launched by this library on the current device since load (bench accounting). */
uint64_t pm_kernel_launches(void);

/* ---- standalone kernels (host buffers; used by the kernel sweep and the parity tests) ---- */

/* out[i] = a[i] * b[i] in Fr.  Replaces ark-ff `Fp::mul` (every `F` product, e.g. src/prover.rs:266-277). */
int pm_fr_mul_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* out[i] = a[i] * b[i] in Fq (base-field product used inside all G1 arithmetic). */
int pm_fq_mul_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* out[i] = a[i] + b[i]  /  a[i] - b[i] in Fr. */
/* out[i] = a[i]^-1 (0 -> 0): ark-ff `Field::inverse` (src/common.rs:41,46; every projective -> affine conversion). */
int pm_fr_inv_batch(const uint8_t* a, uint8_t* out, size_t n);
int pm_fq_inv_batch(const uint8_t* a, uint8_t* out, size_t n);
int pm_fr_add_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
int pm_fr_sub_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);

/* In-place size-2^log_n NTT over Fr, natural order in and out.
 * inverse == 0: `Radix2EvaluationDomain::fft` (src/prover.rs:319);
 * inverse != 0: `ifft_in_place` incl. the n^-1 scaling (src/prover.rs:241,325).
 * coset_gen (nullable, 32 B): forward evaluates on coset_gen*H; inverse interpolates from it. */
int pm_ntt_fr(uint8_t* data, unsigned log_n, int inverse, const uint8_t* coset_gen);

/* Sharded NTT of size 2^log_n over G = 2^log_g ranks (G <= 8), one process per GPU; same transform as pm_ntt_fr.
 * Rank g holds the interleaved subsequence x[j*G + g], j < 2^log_n / G, in DEVICE memory (32-byte Montgomery Fr).
 *   pm_ntt_dist_local   : local (N/G)-point transform in place in `data_dev`, twiddled and packed destination-major
 *                         into `send_dev` (G equal blocks: block h goes to rank h);
 *   <one all-to-all of the blocks: send block h -> rank h, received block g <- rank g>  (caller: NCCL / NVLink)
 *   pm_ntt_dist_combine : G-point transform across the received blocks `recv_dev` -> `out_dev`:
 *                         X[k] for k = rank (mod G) at local index (k - rank)/G, i.e. interleaved like the input.
 * All pointers are device pointers of 2^log_n / G elements; work is enqueued on `cuda_stream` (a cudaStream_t;
 * NULL = the default stream, as in the CUDA runtime) and not synchronised.  Replaces `Radix2EvaluationDomain::{fft, ifft_in_place}` (src/prover.rs:241,319,325)
 * for domains sharded across GPUs. */
int pm_ntt_dist_local(void* data_dev, void* send_dev, unsigned log_n, unsigned log_g, unsigned rank, int inverse, void* cuda_stream);
int pm_ntt_dist_combine(const void* recv_dev, void* out_dev, unsigned log_n, unsigned log_g, int inverse, void* cuda_stream);

/* out = sum_i scalars[i] * bases[i].  Replaces `VariableBaseMSM::msm_unchecked` behind
 * `msm()` (src/prover.rs:380-384). */
int pm_msm_g1(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, uint8_t out[PM_G1_BYTES]);
/* Same, with an explicit window width (0 = automatic); test hook for the bucket pipeline. */
int pm_msm_g1_window(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                     int heavy_threshold, uint8_t out[PM_G1_BYTES]);

/* Same through `levels` precomputed multiples per base (built on the fly); test hook for the fixed-base tables. */
int pm_msm_g1_levels(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                     int levels, uint8_t out[PM_G1_BYTES]);

/* Process-wide tuning of the bucket accumulation of every later MSM (standalone and inside the prover):
 * rounds = number of batched-affine pair rounds before the XYZZ walk (-1 = automatic, 0 = none).
 * Results never depend on it; test and sweep hook. */
int pm_msm_set_tuning(int rounds);

/* G1 point (de)compression on the device: the 48-byte zcash encoding ark-bls12-381 uses for `CanonicalSerialize`
 * (big-endian x; byte 0: 0x80 compressed, 0x40 infinity, 0x20 y is the larger root).  Replaces the per-point work of
 * `ProvingKey::{deserialize_compressed, serialize_compressed}` (src/data_structures.rs:55-73): one Fq square root per
 * decoded point.  in/out: n x 48 bytes <-> n x 96 bytes (Montgomery affine, (0,0) = infinity).
 * validate != 0 also checks the prime-order subgroup (`deserialize_compressed`); 0 stops at the on-curve check
 * (a superset of `deserialize_compressed_unchecked`).  A bad encoding fails with PM_ERR_ARG naming the first index. */
int pm_g1_decompress_batch(const uint8_t* in, size_t n, int validate, uint8_t* out);
int pm_g1_compress_batch(const uint8_t* in, size_t n, uint8_t* out);

/* out[i] = scalars[i] * G (G = the BLS12-381 G1 generator), canonical affine.
 * Replaces `generate()` (src/generator.rs:169-177). */
int pm_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out);

/* ---- prover context: proving key resident on the device, one call per protocol phase ---- */

/* R1CS matrices as produced by `cs.to_matrices()` (src/generator.rs:46-54), flattened to CSR:
 * row r of matrix M holds entries [row_ptr[r], row_ptr[r+1]) of (col, val); columns follow
 * arkworks (0 = the constant one, 1..m0-1 = instance, m0.. = witness).  Duplicate columns in a
 * row keep first-match semantics like `m_at` (src/common.rs:100-105). */
typedef struct {
    uint64_t num_instance_variables;      /* m0, includes the leading 1 (SAPMatrices, src/common.rs:113-127) */
    uint64_t num_r1cs_witness_variables;  /* mw */
    uint64_t num_r1cs_constraints;        /* nr */
    const uint64_t* a_row_ptr; const uint32_t* a_col; const uint8_t* a_val;
    const uint64_t* b_row_ptr; const uint32_t* b_col; const uint8_t* b_val;
    const uint64_t* c_row_ptr; const uint32_t* c_col; const uint8_t* c_val;
} pm_r1cs_view;

/* Borrowed view of a `ProvingKey` (src/data_structures.rs:56-73).  All point arrays share `point_stride`:
 * 96 (packed Montgomery affine), >= 104 (arkworks' in-memory `Affine{x,y,infinity}`) or 48 = the compressed
 * encoding of `serialize_compressed` (decoded on the device while uploading; on-curve checked). */
typedef struct {
    pm_r1cs_view r1cs;                    /* pk.sap_matrices */
    uint64_t n;                           /* pk.vk.n  (domain size) */
    uint64_t sigma;                       /* pk.vk.sigma = n + 3 */
    size_t point_stride;
    const uint8_t* x_powers_g1;               uint64_t x_powers_g1_len;               /* n + 1 */
    const uint8_t* x_powers_y_alpha_g1;       uint64_t x_powers_y_alpha_g1_len;       /* 3 */
    const uint8_t* x_powers_zh_by_y_alpha_g1; uint64_t x_powers_zh_by_y_alpha_g1_len; /* n - 1 */
    const uint8_t* x_powers_y_gamma_g1;       uint64_t x_powers_y_gamma_g1_len;       /* 2 */
    const uint8_t* x_powers_y_gamma_z_g1;     uint64_t x_powers_y_gamma_z_g1_len;     /* 10n + 23 */
    const uint8_t* uj_wj_lcs_by_y_alpha_g1;   uint64_t uj_wj_lcs_by_y_alpha_g1_len;   /* cols - m0 */
} pm_pk_view;

typedef struct pm_ctx pm_ctx;

/* Upload (and repack) a proving key once; later `prove` calls reuse the device copy
 * (`prove(&pk, ..)` borrows the key on every call, src/lib.rs:72-78).  Compressed vectors (point_stride 48) are
 * validated like `ProvingKey::deserialize_compressed`: canonical x, flags, on the curve AND in the prime-order
 * subgroup; pm_ctx_create_unchecked (below) skips the subgroup check only, like `deserialize_compressed_unchecked`. */
int pm_ctx_create(const pm_pk_view* pk, pm_ctx** out);
void pm_ctx_destroy(pm_ctx* ctx);

/* Setup on the device: the six G1 vectors of `generate_proving_key` (src/generator.rs:81-137)
 * from the trapdoors x, z (32 B each, Montgomery), kept resident in a new context.  Also
 * writes [x]_2 and [z]_2 (src/generator.rs:144-145) as 2 x 192 bytes (x.c0, x.c1, y.c0, y.c1; Montgomery). */
int pm_setup(const pm_r1cs_view* r1cs, const uint8_t x[PM_FR_BYTES], const uint8_t z[PM_FR_BYTES], pm_ctx** out,
             uint8_t x_g2[192], uint8_t z_g2[192]);
/* Sizes (n, sigma, columns) of the context's key. */
int pm_ctx_dims(const pm_ctx* ctx, uint64_t* n, uint64_t* sigma, uint64_t* num_columns);
/* Copy one of the key's vectors back to the host (to build/serialise a ProvingKey).
 * which: 0 x_powers_g1, 1 x_powers_y_alpha_g1, 2 x_powers_zh_by_y_alpha_g1, 3 x_powers_y_gamma_g1,
 *        4 x_powers_y_gamma_z_g1, 5 uj_wj_lcs_by_y_alpha_g1.  `out` receives len*stride bytes
 * (stride 96: packed; >= 104: arkworks layout with the infinity flag byte at offset 96). */
int pm_ctx_key_len(const pm_ctx* ctx, int which, uint64_t* len);
int pm_ctx_export_key(const pm_ctx* ctx, int which, uint8_t* out, size_t stride);

/* Phase 1 (src/prover.rs:73-123): witness -> y vector -> U.z, W.z -> u, w, h -> [a]_1, [c]_1.
 * x: m0 instance values (x[0] = 1), w: mw witness values, r_a: the two blinding coefficients
 * drawn by the caller's RNG (src/prover.rs:110, coefficient 0 first). */
int pm_prove_phase1(pm_ctx* ctx, const uint8_t* x, const uint8_t* w, const uint8_t r_a[2 * PM_FR_BYTES],
                    uint8_t a_out[PM_G1_BYTES], uint8_t c_out[PM_G1_BYTES]);
/* Same, split so the assignment upload can be kept out of a timed region. */
int pm_ctx_set_assignment(pm_ctx* ctx, const uint8_t* x, const uint8_t* w);
int pm_prove_phase1_resident(pm_ctx* ctx, const uint8_t r_a[2 * PM_FR_BYTES], uint8_t a_out[PM_G1_BYTES],
                             uint8_t c_out[PM_G1_BYTES]);
/* Phase 2 (src/prover.rs:128-132): a(x1) = u(x1) + r_a(x1) * y1^alpha. */
int pm_prove_phase2(pm_ctx* ctx, const uint8_t x1[PM_FR_BYTES], const uint8_t y1_alpha[PM_FR_BYTES],
                    uint8_t a_at_x1_out[PM_FR_BYTES]);
/* Phase 3 (src/prover.rs:142-229): opening quotient by (X - x1) and [d]_1. */
int pm_prove_phase3(pm_ctx* ctx, const uint8_t x2[PM_FR_BYTES], const uint8_t c_at_x1[PM_FR_BYTES],
                    uint8_t d_out[PM_G1_BYTES]);
/* ---- multi-GPU: one process per GPU, MSMs split by point range ----
 * A sharded context holds the points g = k*world + rank of every key vector (interleaved split, so
 * the a-, c- and d-side MSMs are all balanced).  The polynomial work is replicated on every rank.
 * Each MSM phase is split in two: *_partial runs the device work and returns this rank's XYZZ sum(s);
 * the caller all-gathers them (e.g. ncclAllGather of world * 384 / 192 bytes) and every rank calls
 * *_finish on the gathered buffer, which adds the partials on the host and returns the same affine
 * point everywhere.  Group addition is not an NCCL reduction, hence all-gather + local add. */
int pm_setup_sharded(const pm_r1cs_view* r1cs, const uint8_t x[PM_FR_BYTES], const uint8_t z[PM_FR_BYTES], int rank,
                     int world, pm_ctx** out, uint8_t x_g2[192], uint8_t z_g2[192]);
int pm_ctx_create_sharded(const pm_pk_view* pk, int rank, int world, pm_ctx** out);
/* Opt-in `deserialize_compressed_unchecked`: no prime-order-subgroup check of compressed key vectors (rank 0, world 1
 * for an unsharded context).  Encoding and on-curve checks still run. */
int pm_ctx_create_unchecked(const pm_pk_view* pk, int rank, int world, pm_ctx** out);
int pm_ctx_shard(const pm_ctx* ctx, int* rank, int* world);
int pm_prove_phase1_partial(pm_ctx* ctx, const uint8_t r_a[2 * PM_FR_BYTES], uint8_t partials_out[2 * PM_XYZZ_BYTES]);
int pm_prove_phase1_finish(pm_ctx* ctx, const uint8_t* gathered, int count, uint8_t a_out[PM_G1_BYTES],
                           uint8_t c_out[PM_G1_BYTES]);
int pm_prove_phase3_partial(pm_ctx* ctx, const uint8_t x2[PM_FR_BYTES], const uint8_t c_at_x1[PM_FR_BYTES],
                            uint8_t partial_out[PM_XYZZ_BYTES]);
int pm_prove_phase3_finish(pm_ctx* ctx, const uint8_t* gathered, int count, uint8_t d_out[PM_G1_BYTES]);
/* The same exchange INSIDE the phase: with an NCCL communicator attached to the sharded context, the per-rank sums of
 * each MSM are all-gathered on the device (ncclAllGather over NVLink, enqueued on the library's stream right behind
 * the MSM kernels: no host round trip between compute and collective) and every rank finishes on its host.
 * NCCL is bound at run time: `libnccl_path` names the libnccl.so.2 the process already uses (NULL = "libnccl.so.2"
 * from the loader path).  Rank 0 calls pm_nccl_unique_id and distributes the 128 bytes (any transport); every rank
 * then calls pm_ctx_attach_nccl (collective: returns when all ranks of the context's world have joined).
 * pm_prove_phase{1,3}_collective replace the *_partial / all-gather / *_finish triple. */
int pm_nccl_unique_id(const char* libnccl_path, uint8_t id_out[128]);
int pm_ctx_attach_nccl(pm_ctx* ctx, const char* libnccl_path, const uint8_t id[128]);
int pm_ctx_has_collective(const pm_ctx* ctx);
int pm_prove_phase1_collective(pm_ctx* ctx, const uint8_t r_a[2 * PM_FR_BYTES], uint8_t a_out[PM_G1_BYTES],
                               uint8_t c_out[PM_G1_BYTES]);
int pm_prove_phase3_collective(pm_ctx* ctx, const uint8_t x2[PM_FR_BYTES], const uint8_t c_at_x1[PM_FR_BYTES],
                               uint8_t d_out[PM_G1_BYTES]);
/* Host-only helper (no device needed): canonical affine sum of `count` XYZZ records spaced `stride` bytes. */
int pm_host_sum_partials(const uint8_t* parts, int count, size_t stride, uint8_t out[PM_G1_BYTES]);

/* Test hook: copy an intermediate of the last proof to the host.
 * which: 0 u coeffs (n), 1 w coeffs (n), 2 witness-u coeffs (n), 3 u^2 coeffs (2n), 4 [x|w|y] (cols - m0),
 *        5 phase-1 c-side scalars, 6 opening quotient D (10n + 22).  Returns the element count in *len. */
int pm_ctx_debug_read(pm_ctx* ctx, int which, uint8_t* out, uint64_t capacity_elems, uint64_t* len);
/* Test hook for the multi-GPU kernels on ONE GPU: on an unsharded context that has just completed a proof, replays the
 * polynomial work the way `virtual_world` (2, 4 or 8) ranks of the sharded-resident flow would — SAP rows of each residue
 * class, sharded transforms with the all-to-all emulated by device copies, ring exchange, local scalar assembly, the
 * (X - x1) division by chunk ranges with its carry exchange — and returns in *mismatches how many elements differ from
 * what the unsharded path computed (0 = identical).  Needs n >= 8 * virtual_world^2. */
int pm_ctx_selftest_resident(pm_ctx* ctx, int virtual_world, uint64_t* mismatches);
/* Milliseconds spent on the device by the last phase-1 / phase-2 / phase-3 call (CUDA events). */
int pm_ctx_phase_ms(const pm_ctx* ctx, double ms[3]);

/* ---- host mirror of Polymath::setup / prove (C++; the reference's Rust host flow restated) ---- */

/* rand 0.8 `StdRng` (ChaCha12) as used by the reference's tests and bench (benches/bench.rs:65). */
typedef struct pm_rng pm_rng;
pm_rng* pm_rng_seed_from_u64(uint64_t seed);
pm_rng* pm_rng_from_seed(const uint8_t seed[32]);
void pm_rng_free(pm_rng* rng);
uint64_t pm_rng_next_u64(pm_rng* rng);
/* ark-ff `Fr::rand(rng)`; writes the Montgomery form. */
void pm_rng_fr_rand(pm_rng* rng, uint8_t out[PM_FR_BYTES]);
/* merlin known-answer vector ("test protocol" / "some label" / "some data" / "challenge"). */
int pm_merlin_test_vector(uint8_t out[32]);

/* `Polymath::setup` / `generate_proving_key` (src/generator.rs:24-167) for already-synthesised
 * R1CS matrices: samples the trapdoors x, z from `rng` exactly like the reference, builds the key
 * on the device (kept resident in *ctx_out) and writes the compressed VerifyingKey (392 bytes,
 * src/data_structures.rs:25-50). */
int pm_polymath_setup(const pm_r1cs_view* r1cs, pm_rng* rng, pm_ctx** ctx_out, uint8_t vk_out[392]);
/* `create_proof_with_assignment` (src/prover.rs:66-237): instance = m0 values (leading 1 first),
 * witness = mw values, both Montgomery; draws r_a from `rng`; runs the Merlin transcript
 * (src/common.rs:21-37) on the host between the device phases; writes the compressed Proof
 * (176 bytes, src/data_structures.rs:10-19).  Instance values must be reduced (limbs < r), else PM_ERR_ARG.
 * RNG: r_a is drawn before the device work starts; when phase 1 fails (PM_ERR_UNSATISFIED / PM_ERR_DEGENERATE: the
 * reference panics at src/prover.rs:107-108, BEFORE `F::rand` at :110) the generator is restored, so after a failed
 * proof the caller's stream has not advanced, exactly like the reference's. */
int pm_polymath_prove(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, pm_rng* rng, uint8_t proof_out[176]);

/* `Polymath::verify` / `verify_proof` (src/verifier.rs:19-62) on the HOST — no device work (BASELINE north_star:
 * pairing-based verification stays on the host): deserialises the compressed VerifyingKey (392 bytes) and Proof
 * (176 bytes) with ark-serialize's validation (on curve, prime-order subgroup, canonical scalars), recomputes
 * x1, c(x1), x2 through the Merlin transcript and checks
 *   e([a]_1 + x2 [c]_1 - (a(x1) + x2 c(x1)) [1]_1, [z]_2) * e(-[d]_1, [x]_2 - x1 [1]_2) == 1.
 * public_inputs: num_public x 32 bytes (Montgomery), WITHOUT the leading one (the verifier prepends it,
 * src/verifier.rs:26).  *accepted = 1 / 0 = the reference's Ok(true) / Ok(false); a key or proof that does not
 * deserialise fails with PM_ERR_ARG (the reference's SerializationError).  A public input whose limbs are not
 * below r fails with PM_ERR_ARG as well: raw limbs are never reduced, so x and x + r cannot alias (the reference
 * takes typed field elements). */
int pm_polymath_verify(const uint8_t vk[392], const uint8_t* public_inputs, size_t num_public, const uint8_t proof[176],
                       int* accepted);
/* The reference's other two `Transcript` implementations (src/transcript/keccak256.rs, blake3.rs; exercised by
 * tests/dummy.rs:76-80): same flows with the Fiat-Shamir challenges drawn from H(transcript || label) mod r.
 * transcript: PM_TRANSCRIPT_MERLIN (what pm_polymath_prove / pm_polymath_verify use), _KECCAK256, _BLAKE3. */
enum { PM_TRANSCRIPT_MERLIN = 0, PM_TRANSCRIPT_KECCAK256 = 1, PM_TRANSCRIPT_BLAKE3 = 2 };
int pm_polymath_prove_transcript(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, pm_rng* rng, int transcript,
                                 uint8_t proof_out[176]);
int pm_polymath_verify_transcript(const uint8_t vk[392], const uint8_t* public_inputs, size_t num_public,
                                  const uint8_t proof[176], int transcript, int* accepted);
/* Host-only test hook: out = Keccak-256 (kind 1, the `sha3` crate's Keccak256: original padding) or BLAKE3 (kind 2). */
int pm_host_hash(int kind, const uint8_t* data, size_t len, uint8_t out[32]);
/* Batch form: `count` proofs under one key, proof i with the public inputs
 * public_inputs[i * num_public ..].  The pairing equations are combined with 128-bit coefficients (r_0 = 1) drawn
 * from rand `StdRng::from_seed(BLAKE3(seed | vk | proofs | public inputs))` — bound to the statements, so knowing
 * the caller's seed in advance does not help to craft proofs whose errors cancel — into ONE product of three pairings:
 *   e(sum r_i L_i, [z]_2) * e(-sum r_i [d_i]_1, [x]_2) * e(sum r_i x1_i [d_i]_1, [1]_2) == 1.
 * *accepted = 1 iff the combined check holds (all proofs valid, up to 2^-128 soundness error); count == 0 accepts
 * (seed may then be NULL).  The seed should still be fresh per call. */
int pm_polymath_verify_batch(const uint8_t vk[392], size_t count, const uint8_t* public_inputs, size_t num_public,
                             const uint8_t* proofs, const uint8_t seed[32], int* accepted);
/* Host-only test hook: prod_i e(P_i, Q_i) == 1 for `count` pairs of affine points (G1: 96 bytes, G2: 192 bytes
 * x.c0, x.c1, y.c0, y.c1; Montgomery; all-zero = infinity).  `E::multi_pairing(..).0.is_one()` (src/verifier.rs:50-61). */
int pm_host_pairing_product_is_one(const uint8_t* g1_points, const uint8_t* g2_points, int count, int* is_one);

/* Collective supplied by the caller for the sharded flow: gather `bytes` from every rank into
 * recv (world * bytes, rank order).  Returns 0 on success. */
typedef int (*pm_allgather_fn)(void* user, const uint8_t* send, size_t bytes, uint8_t* recv);
/* Sharded variants: every rank calls them with the same arguments and its own sharded context; all
 * ranks obtain identical vk / proof bytes.  `upload` = 0 reuses the resident assignment.  `allgather` may be NULL
 * when the context has an NCCL communicator (pm_ctx_attach_nccl): the phases then run their own collective. */
int pm_polymath_setup_sharded(const pm_r1cs_view* r1cs, pm_rng* rng, int rank, int world, pm_ctx** ctx_out,
                              uint8_t vk_out[392]);
int pm_polymath_prove_sharded(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, int upload, pm_rng* rng,
                              pm_allgather_fn allgather, void* user, uint8_t proof_out[176]);
/* Calls `allgather` with a rank-tagged payload and verifies the result (plumbing self-test). */
int pm_allgather_selftest(pm_allgather_fn allgather, void* user, int rank, int world);

/* Same as pm_polymath_prove but uses the assignment already resident from pm_ctx_set_assignment
 * (bench.py's device-resident timing leg); `instance` is still needed for the transcript. */
int pm_polymath_prove_resident(pm_ctx* ctx, const uint8_t* instance, pm_rng* rng, uint8_t proof_out[176]);

/* ---- measurement hooks (bench.py; synthetic device-resident inputs, CUDA-event timing) ---- */

/* Issue rate of dependency-free IMAD.WIDE.U32 (32x32+64 multiply-adds per second, whole GPU):
 * the INT32 IMAD-pipe roofline denominator for the MSM / field kernels. */
int pm_bench_imad_peak(double* mads_per_s);
/* Register-resident Montgomery products per second; field: 0 = Fr, 1 = Fq; variants of the same dependent chain:
 * 2 = Fq squaring, 3 = Fq Karatsuba product, 4 = Fr squaring, 5 = Fr Karatsuba product. */
int pm_bench_field_mul(int field, double* muls_per_s);
/* Average milliseconds of `iters` size-2^log_n transforms on resident data (after one warm-up).
 * inverse: bit 0 = inverse transform, bit 1 = the coset variant of pm_ntt_fr (the g^i / g^-i scaling pass included). */
int pm_bench_ntt(unsigned log_n, int inverse, int iters, double* ms_avg);
/* Milliseconds of one device-resident decompression / compression of n synthetic G1 points. */
int pm_bench_g1_codec(size_t n, double* ms_decompress, double* ms_compress);
/* Average milliseconds of `iters` fixed-base batches of n scalar multiplications [s_i]G with the canonical-affine
 * normalisation, resident scalars (after one warm-up): the generator's `generate()` (src/generator.rs:169-177). */
int pm_bench_fixed_base(size_t n, int iters, double* ms_avg);
/* Average milliseconds of `iters` n-point MSMs on resident synthetic bases/scalars (after one warm-up);
 * ms_accumulate (nullable) receives the average time of the bucket-accumulation kernel alone. */
int pm_bench_msm(size_t n, int window_bits, int iters, double* ms_avg, double* ms_accumulate);
/* Same with `levels` precomputed multiples 2^(c*l) P per base (fixed-base tables; 0/1 = none). */
int pm_bench_msm_levels(size_t n, int window_bits, int levels, int iters, double* ms_avg, double* ms_accumulate);
/* Input distribution of the later pm_bench_msm* calls: 0 = uniform scalars (default), 1 = the skewed case of
 * SURVEY.md 8d — 89 % of the scalars one repeated value, 10 % zero, 1 % uniform, and 1 % of the bases at infinity
 * (S-dummy keys, benches/bench.rs:38-61: one hot bucket per window, the chunked heavy-run path). */
int pm_bench_set_msm_skew(int mode);

/* CUDA-event stopwatch on the library's stream: start records an event, stop records another,
 * synchronises it and returns the elapsed device-timeline milliseconds. */
int pm_timer_start(void);
int pm_timer_stop(double* ms);
/* Enable/disable per-kernel event timing inside the MSM and NTT engines, and read the last values:
 * ms[0] = bucket-accumulation kernel of the last MSM, ms[1] = passes of the last NTT. */
int pm_bench_set_kernel_timing(int enable);
int pm_bench_last_kernel_ms(double ms[2]);
/* Last MSM of the main engine (kernel timing enabled): out[0] = ms of the whole bucket-accumulation stage,
 * out[1] = ms of its heaviest kernel (first-round k_pairs_backward; 0 when no pair rounds ran),
 * out[2] = pair rounds used, out[3] = sorted-list entries (points x windows, upper bound). */
int pm_bench_last_msm(double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* POLYMATH_B200_H */
