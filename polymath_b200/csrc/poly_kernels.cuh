// K2 / K6 — SAP evaluation (CSR SpMV + closed-form rows), Hadamard/quotient helpers, chunked
// Horner evaluation and the (X - x1) linear-recurrence division (interface).  See poly_kernels.cu.
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace pm {

// R1CS matrix in CSR form on the device (duplicates removed on the host with first-match
// semantics, mirroring `m_at`, /root/reference/src/common.rs:100-105).
struct DevCsr {
    const uint32_t* row_ptr = nullptr;  // [rows + 1]
    const uint32_t* col = nullptr;      // [nnz]
    const Fr* val = nullptr;            // [nnz], Montgomery
};

struct SapDims {
    uint32_t m0, mw, nr;   // instance vars (incl. the leading 1), witness vars, R1CS constraints
    uint64_t n;            // domain size
};

// status bits written by the device checks (read back once per phase)
enum : uint32_t {
    ST_REMAINDER_NONZERO = 1u,   // (u^2 - w) mod Z_H != 0          -> PM_ERR_UNSATISFIED
    ST_H_ZERO = 2u,              // h == 0                            -> PM_ERR_DEGENERATE
    ST_H_DEGREE = 4u,            // deg h > n - 2                     -> PM_ERR_DEGENERATE
    ST_OPENING_REMAINDER = 8u,   // numerator(x1) != 0                -> PM_ERR_REMAINDER
};

// ztail = [x | w | y]; on entry x and w are filled, on exit y (m0 + nr entries) too.
// u_ev / w_ev / wu_ev (n entries each) receive U.z, W.z and the witness-column part of U.z.
void launch_sap_evals(const SapDims& d, const DevCsr& A, const DevCsr& B, const DevCsr& C, Fr* ztail,
                      Fr* u_ev, Fr* w_ev, Fr* wu_ev, cudaStream_t stream);

// data[i] = data[i]^2
void launch_square(Fr* data, size_t n, cudaStream_t stream);

// Checks on u2 (2n coeffs) against w (n coeffs): remainder zero, h = u2[n..2n) non-zero, u2[2n-1] == 0.
void launch_quotient_checks(const Fr* u2, const Fr* w, uint64_t n, uint32_t* status, cudaStream_t stream);

// ra_ext[0..2) = r_a, ra_ext[2..5) = r_a^2 coefficients (single thread)
void launch_ra_square(Fr* ra_ext, cudaStream_t stream);

// Layout of the concatenated c-side base / scalar array.  Every key vector starts at a multiple of 8 (the largest
// world size), so that under the interleaved split (global index g -> rank g % G) a rank's share of EACH vector is
// the residue class of its own rank: the slice of u / h / [x|w|y] a rank holds after its sharded transforms is exactly
// what its MSM shard needs, with no exchange.  The gaps hold infinity bases and zero scalars (skipped by the MSM).
struct CLayout {
    uint64_t n = 0, tail = 0;         // domain size, cols - m0
    uint64_t off_ya = 0, off_yg = 0, off_zh = 0, off_lcs = 0, len_c = 0, len_a = 0;
    static CLayout make(uint64_t n, uint64_t tail) {
        CLayout L;
        auto up = [](uint64_t v) { return (v + 7) & ~(uint64_t)7; };
        L.n = n; L.tail = tail;
        L.off_ya = up(n + 1);                 // x_powers (n + 1) at 0
        L.off_yg = L.off_ya + 8;              // x_powers_y_alpha (3)
        L.off_zh = L.off_yg + 8;              // x_powers_y_gamma (2)
        L.off_lcs = up(L.off_zh + (n ? n - 1 : 0));   // x_powers_zh_by_y_alpha (n - 1)
        L.len_c = L.off_lcs + tail;           // uj_wj_lcs (tail)
        L.len_a = L.off_ya + 3;               // the a-side MSM reads [u | r_a] over the first len_a bases
        return L;
    }
};

// scal_a = [u (n) | 0.. | r0 r1 0]; scal_c = [2 r_a u (n+1) | r_a^2 (3) | r_a (2) | h (n-1) | ztail (tail)] in CLayout
void launch_assemble_phase1_scalars(const Fr* u, const Fr* u2, const Fr* ztail, const CLayout& L, const Fr* ra_ext,
                                    Fr* scal_a, Fr* scal_c, cudaStream_t stream);

// ---- sharded-resident phase 1: rank `rank` of `world` (a power of two <= 8) holds the residue class
// k = rank (mod world) of every polynomial -------------------------------------------------------------------------
// Rows k = rank + t * world of U.z, W.z and the witness part of U.z (t < n / world): "row-range SpMV" over the
// interleaved row split, from the replicated assignment ztail = [x | w | ..].
void launch_sap_evals_strided(const SapDims& d, const DevCsr& A, const DevCsr& B, const DevCsr& C, const Fr* ztail,
                              uint32_t rank, uint32_t world, Fr* u_loc, Fr* w_loc, Fr* wu_loc, cudaStream_t stream);
// zt_loc[t] = [x | w | y][rank + t * world] for t < count: this rank's scalars of the lcs MSM (y computed on the fly)
void launch_ztail_strided(const SapDims& d, const DevCsr& A, const DevCsr& B, const Fr* ztail, uint32_t rank, uint32_t world,
                          uint64_t count, Fr* zt_loc, cudaStream_t stream);
// quotient checks on the slices: u2_loc (2n / world), w_loc (n / world)
void launch_quotient_checks_strided(const Fr* u2_loc, const Fr* w_loc, uint64_t n, uint32_t rank, uint32_t world,
                                    uint32_t* status, cudaStream_t stream);
// local c-side / a-side scalars: scal_c_loc[t] = scal_c[rank + t * world] etc.  u_prev = the slice of rank - 1 (mod world).
void launch_assemble_phase1_strided(const Fr* u_loc, const Fr* u_prev, const Fr* u2_loc, const Fr* zt_loc, const CLayout& L,
                                    const Fr* ra_ext, uint32_t rank, uint32_t world, uint64_t cnt_a, uint64_t cnt_c,
                                    Fr* scal_a_loc, Fr* scal_c_loc, cudaStream_t stream);

// ---- chunked polynomial machinery ------------------------------------------------------
// A polynomial is described by a "source": either a plain coefficient array or the virtual
// opening numerator of prover.rs:142-209 assembled on the fly from its five blocks.
struct NumeratorSrc {
    const Fr* u;        // n
    const Fr* wu;       // n
    const Fr* u2;       // 2n  (= witness_w + h_numerator)
    const Fr* ra_ext;   // r_a (2) | r_a^2 (3)
    const Fr* consts;   // [0] = x2, [1] = a(x1) + x2*c(x1)
    uint64_t n, sigma, len;  // len = 8*sigma + 2n - 1
};

constexpr int kChunk = 4096;  // coefficients per CTA in the chunked kernels

// chunk_vals[c] = sum_{k in chunk c} p_k * x^(k - c*kChunk)
void launch_chunk_eval_plain(const Fr* coeffs, uint64_t len, const Fr* x, Fr* chunk_vals, cudaStream_t stream);
void launch_chunk_eval_numerator(const NumeratorSrc& src, const Fr* x, Fr* chunk_vals, cudaStream_t stream);
// out[0] = sum_c chunk_vals[c] * x^(c*kChunk)   (+ addend[0] * addend[1]... see launch_a_at_x1)
void launch_combine_chunks(const Fr* chunk_vals, uint64_t nchunks, const Fr* x, Fr* out, cudaStream_t stream);
// a_at_x1 = u(x1) + (r0 + r1*x1) * y1_alpha ; inputs: u_at_x1 (device), ra_ext, consts2 = [x1, y1_alpha]
void launch_a_at_x1(const Fr* u_at_x1, const Fr* ra_ext, const Fr* x1_y1a, Fr* out, cudaStream_t stream);
// q[k-1] = p_k + x * q_k for the virtual numerator; q has len-1 entries; sets ST_OPENING_REMAINDER if
// p(x) != 0.  `work` needs 2*(nchunks+1) + 2*(nchunks/kChunk+2) + 2 elements.  Returns the launch count.
int launch_divide_numerator(const NumeratorSrc& src, const Fr* x, Fr* q, Fr* work, uint32_t* status, cudaStream_t stream);
// The same division cut into chunk ranges (one per rank; SURVEY.md 8e "contiguous coefficient ranges, one exchange"):
//   launch_numerator_range_eval : range_val[0] = sum_{k in chunks [c_lo, c_lo + cnt)} p_k x^(k - c_lo * kChunk)
//   <all-gather of the G range values>
//   launch_numerator_range_carry: from all range values, carry_in[0] = the quotient coefficient entering this rank's
//       range from above, and ST_OPENING_REMAINDER if p(x) != 0 (every rank sees the same total)
//   launch_numerator_range_divide: q[k-1] for the numerator indices k of the range.
// `work` as above (sized for all chunks).  range_lo[s] = first chunk of rank s (world + 1 entries, device).
int launch_numerator_range_eval(const NumeratorSrc& src, const Fr* x, uint64_t c_lo, uint64_t cnt, Fr* work, Fr* range_val,
                                cudaStream_t stream);
void launch_numerator_range_carry(const Fr* range_vals, const uint64_t* range_lo, uint32_t rank, uint32_t world, const Fr* x,
                                  Fr* carry_in, uint32_t* status, cudaStream_t stream);
int launch_numerator_range_divide(const NumeratorSrc& src, const Fr* x, uint64_t c_lo, uint64_t cnt, const Fr* carry_in, Fr* q,
                                  Fr* work, cudaStream_t stream);
// materialise the virtual numerator (tests / debugging)
void launch_materialize_numerator(const NumeratorSrc& src, Fr* out, cudaStream_t stream);

}  // namespace pm
