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
 *  - one context may be used by one thread at a time; distinct contexts are independent;
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

const char* pm_last_error(void);
/* ABI version of this header; bumped on any incompatible change. */
int pm_abi_version(void);
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
int pm_fr_add_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
int pm_fr_sub_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);

/* In-place size-2^log_n NTT over Fr, natural order in and out.
 * inverse == 0: `Radix2EvaluationDomain::fft` (src/prover.rs:319);
 * inverse != 0: `ifft_in_place` incl. the n^-1 scaling (src/prover.rs:241,325).
 * coset_gen (nullable, 32 B): forward evaluates on coset_gen*H; inverse interpolates from it. */
int pm_ntt_fr(uint8_t* data, unsigned log_n, int inverse, const uint8_t* coset_gen);

/* out = sum_i scalars[i] * bases[i].  Replaces `VariableBaseMSM::msm_unchecked` behind
 * `msm()` (src/prover.rs:380-384). */
int pm_msm_g1(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, uint8_t out[PM_G1_BYTES]);
/* Same, with an explicit window width (0 = automatic); test hook for the bucket pipeline. */
int pm_msm_g1_window(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                     int heavy_threshold, uint8_t out[PM_G1_BYTES]);

/* out[i] = scalars[i] * G (G = the BLS12-381 G1 generator), canonical affine.
 * Replaces `generate()` (src/generator.rs:169-177). */
int pm_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out);

/* ---- measurement hooks (bench.py; synthetic device-resident inputs, CUDA-event timing) ---- */

/* Issue rate of dependency-free IMAD.WIDE.U32 (32x32+64 multiply-adds per second, whole GPU):
 * the INT32 IMAD-pipe roofline denominator for the MSM / field kernels. */
int pm_bench_imad_peak(double* mads_per_s);
/* Register-resident Montgomery products per second; field: 0 = Fr, 1 = Fq. */
int pm_bench_field_mul(int field, double* muls_per_s);
/* Average milliseconds of `iters` size-2^log_n transforms on resident data (after one warm-up). */
int pm_bench_ntt(unsigned log_n, int inverse, int iters, double* ms_avg);
/* Average milliseconds of `iters` n-point MSMs on resident synthetic bases/scalars (after one warm-up);
 * ms_accumulate (nullable) receives the average time of the bucket-accumulation kernel alone. */
int pm_bench_msm(size_t n, int window_bits, int iters, double* ms_avg, double* ms_accumulate);

#ifdef __cplusplus
}
#endif
#endif /* POLYMATH_B200_H */
