// K2 / K6 — SAP evaluation, quotient checks and the opening-quotient scan for sm_100a.
//
// K2 replaces the reference's dense SAP evaluation (/root/reference/src/prover.rs:75-96,
// 156-166, 245-313 and the accessors of src/common.rs:138-207) with three CSR SpMVs over the
// R1CS matrices and the closed form of U.z / W.z:
//   rows i < m0           : (Uz)_i = 1 + x_i          (Wz)_i = 4 x_i + y_i
//   rows m0 + k           : 1 - x_k                   y_k
//   rows 2m0 + r          : ((A+B) z')_r              4 (C z')_r + y_{m0+r}
//   rows 2m0 + nr + r     : ((A-B) z')_r              y_{m0+r}
// with y = [0] | (1 - x_j)^2 | ((A-B) z')_r^2 (compute_y_vec, prover.rs:279-302).
//
// K6 replaces `u_poly.evaluate(&x1)` (prover.rs:132), the sparse shift/add assembly of
// A*Y^-g + x2*C*Y^-g (prover.rs:142-209) and `divide_with_q_and_r` by (X - x1)
// (prover.rs:211-225).  The numerator is never materialised: its five blocks are read in
// place ("virtual numerator") by a chunked Horner pass, a short serial carry pass over the
// chunk values, and a chunked division pass (q_{k-1} = p_k + x1 q_k).
#include "poly_kernels.cuh"

namespace pm {

namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = kChunk / kThreads;  // 16 coefficients per thread

__device__ __forceinline__ Fr fr_small(uint32_t k) {
    // k * R mod r for tiny k via repeated addition of one (k <= 4)
    Fr one = Fr::one(), acc = Fr::zero();
    for (uint32_t i = 0; i < k; i++) acc = acc + one;
    return acc;
}

__device__ __forceinline__ Fr csr_row_dot(const DevCsr& m, uint32_t r, const Fr* __restrict__ z) {
    Fr acc = Fr::zero();
    uint32_t beg = m.row_ptr[r], end = m.row_ptr[r + 1];
    for (uint32_t k = beg; k < end; k++) acc = acc + m.val[k] * z[m.col[k]];
    return acc;
}

__global__ void __launch_bounds__(128) k_sap_constraint_rows(SapDims d, DevCsr A, DevCsr B, DevCsr C, Fr* __restrict__ ztail,
                                                             Fr* __restrict__ u_ev, Fr* __restrict__ w_ev, Fr* __restrict__ wu_ev) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.nr) return;
    Fr az = csr_row_dot(A, r, ztail);
    Fr bz = csr_row_dot(B, r, ztail);
    Fr cz = csr_row_dot(C, r, ztail);
    Fr diff = az - bz;
    Fr y = diff.sqr();
    Fr sum = az + bz;
    Fr c4 = cz.dbl().dbl();
    size_t i1 = (size_t)2 * d.m0 + r, i2 = i1 + d.nr;
    ztail[(size_t)d.m0 + d.mw + d.m0 + r] = y;
    u_ev[i1] = sum;   wu_ev[i1] = sum;
    u_ev[i2] = diff;  wu_ev[i2] = diff;
    w_ev[i1] = c4 + y;
    w_ev[i2] = y;
}

__global__ void k_sap_public_rows(SapDims d, Fr* __restrict__ ztail, Fr* __restrict__ u_ev, Fr* __restrict__ w_ev,
                                  Fr* __restrict__ wu_ev) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.m0) return;
    Fr one = Fr::one();
    Fr x = ztail[i];
    Fr omx = one - x;
    Fr y = (i == 0) ? Fr::zero() : omx.sqr();
    ztail[(size_t)d.m0 + d.mw + i] = y;
    u_ev[i] = one + x;
    w_ev[i] = x.dbl().dbl() + y;
    u_ev[d.m0 + i] = omx;
    w_ev[d.m0 + i] = y;
    wu_ev[i] = Fr::zero();
    wu_ev[d.m0 + i] = Fr::zero();
}

__global__ void k_zero_tail(Fr* a, Fr* b, Fr* c, uint64_t from, uint64_t n) {
    uint64_t i = from + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr z = Fr::zero();
    a[i] = z; b[i] = z; c[i] = z;
}

__global__ void __launch_bounds__(256) k_square(Fr* data, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] = data[i].sqr();
}

__global__ void __launch_bounds__(256) k_quotient_checks(const Fr* __restrict__ u2, const Fr* __restrict__ w, uint64_t n,
                                                         uint32_t* __restrict__ status, uint32_t* __restrict__ h_nonzero) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Fr h = u2[n + k];
    Fr rem = u2[k] - w[k] + h;
    if (!rem.is_zero()) atomicOr(status, ST_REMAINDER_NONZERO);
    if (!h.is_zero()) {
        if (k == n - 1) atomicOr(status, ST_H_DEGREE);
        else atomicOr(h_nonzero, 1u);
    }
}
__global__ void k_quotient_finish(uint32_t* status, const uint32_t* h_nonzero) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && h_nonzero[0] == 0) atomicOr(status, ST_H_ZERO);
}

__global__ void k_ra_square(Fr* ra) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr r0 = ra[0], r1 = ra[1];
    ra[2] = r0.sqr();
    ra[3] = (r0 * r1).dbl();
    ra[4] = r1.sqr();
}

// one scalar of the c-side / a-side arrays at global index k of the CLayout, from accessors for u, h = u2[n + .], ztail
template <class U, class H, class Z>
__device__ __forceinline__ void phase1_scalar(const CLayout& L, uint64_t k, const Fr* __restrict__ ra, U u_at, H h_at, Z z_at,
                                              Fr& sc, Fr& sa) {
    const uint64_t n = L.n;
    sc = Fr::zero();
    sa = Fr::zero();
    if (k <= n) {
        // 2 r_a(X) u(X): coefficient k = 2 (r0 u_k + r1 u_{k-1})   (compute_r_g1, prover.rs:340-347)
        Fr acc = Fr::zero();
        if (k < n) { sa = u_at(k); acc = ra[0] * sa; }
        if (k >= 1) acc = acc + ra[1] * u_at(k - 1);
        sc = acc.dbl();
    } else if (k < L.off_ya) {
    } else if (k < L.off_ya + 3) {
        const uint64_t j = k - L.off_ya;
        sc = ra[2 + j];                       // r_a^2 against x_powers_y_alpha
        if (j < 2) sa = ra[j];                // r_a against x_powers_y_alpha (compute_a_g1, prover.rs:336)
    } else if (k < L.off_yg) {
    } else if (k < L.off_yg + 2) {
        sc = ra[k - L.off_yg];                // r_a against x_powers_y_gamma
    } else if (k < L.off_zh) {
    } else if (k < L.off_zh + (n - 1)) {
        sc = h_at(k - L.off_zh);              // h = u2[n .. 2n-1) against x_powers_zh_by_y_alpha
    } else if (k < L.off_lcs) {
    } else if (k < L.len_c) {
        sc = z_at(k - L.off_lcs);             // [x | w | y] against uj_wj_lcs
    }
}

__global__ void __launch_bounds__(256) k_assemble_phase1(const Fr* __restrict__ u, const Fr* __restrict__ u2,
                                                         const Fr* __restrict__ ztail, CLayout L, const Fr* __restrict__ ra,
                                                         Fr* __restrict__ scal_a, Fr* __restrict__ scal_c) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.len_c) return;
    Fr sc, sa;
    phase1_scalar(L, k, ra, [&](uint64_t i) { return u[i]; }, [&](uint64_t j) { return u2[L.n + j]; },
                  [&](uint64_t j) { return ztail[j]; }, sc, sa);
    scal_c[k] = sc;
    if (k < L.len_a) scal_a[k] = sa;
}

// ---- sharded-resident variants: thread t of rank r works on the global index r + t * G ---------------------------------
__device__ __forceinline__ void public_row(const SapDims& d, const Fr* __restrict__ z, uint32_t i, Fr& one_plus, Fr& one_minus, Fr& y) {
    const Fr one = Fr::one();
    const Fr x = z[i];
    one_plus = one + x;
    one_minus = one - x;
    y = (i == 0) ? Fr::zero() : one_minus.sqr();      // compute_y_vec: y_0 = 0, y_j = (1 - x_j)^2
}

__global__ void __launch_bounds__(128) k_sap_rows_strided(SapDims d, DevCsr A, DevCsr B, DevCsr C, const Fr* __restrict__ z,
                                                          uint32_t rank, uint32_t world, Fr* __restrict__ u_loc,
                                                          Fr* __restrict__ w_loc, Fr* __restrict__ wu_loc) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t k = (uint64_t)rank + t * world;
    if (k >= d.n) return;
    Fr u = Fr::zero(), w = Fr::zero(), wu = Fr::zero();
    const uint64_t m0 = d.m0, nr = d.nr;
    if (k < 2 * m0) {
        Fr op, om, y;
        public_row(d, z, (uint32_t)(k < m0 ? k : k - m0), op, om, y);
        if (k < m0) { u = op; w = z[k].dbl().dbl() + y; }
        else { u = om; w = y; }
    } else if (k < 2 * m0 + 2 * nr) {
        const bool second = k >= 2 * m0 + nr;
        const uint32_t r = (uint32_t)(k - 2 * m0 - (second ? nr : 0));
        const Fr az = csr_row_dot(A, r, z), bz = csr_row_dot(B, r, z);
        const Fr diff = az - bz;
        const Fr y = diff.sqr();
        if (second) { u = diff; w = y; }
        else { u = az + bz; w = csr_row_dot(C, r, z).dbl().dbl() + y; }
        wu = u;
    }
    u_loc[t] = u;
    w_loc[t] = w;
    wu_loc[t] = wu;
}

__global__ void __launch_bounds__(128) k_ztail_strided(SapDims d, DevCsr A, DevCsr B, const Fr* __restrict__ z, uint32_t rank,
                                                       uint32_t world, uint64_t count, Fr* __restrict__ zt_loc) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint64_t j = (uint64_t)rank + t * world;
    const uint64_t zw = (uint64_t)d.m0 + d.mw;
    Fr v;
    if (j < zw) {
        v = z[j];
    } else if (j - zw < d.m0) {
        Fr op, om;
        public_row(d, z, (uint32_t)(j - zw), op, om, v);
    } else {
        const uint32_t r = (uint32_t)(j - zw - d.m0);
        v = (csr_row_dot(A, r, z) - csr_row_dot(B, r, z)).sqr();
    }
    zt_loc[t] = v;
}

__global__ void __launch_bounds__(256) k_quotient_checks_strided(const Fr* __restrict__ u2_loc, const Fr* __restrict__ w_loc,
                                                                 uint64_t n, uint32_t rank, uint32_t world,
                                                                 uint32_t* __restrict__ status, uint32_t* __restrict__ h_nonzero) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = n / world;
    if (t >= per) return;
    const uint64_t k = (uint64_t)rank + t * world;
    Fr h = u2_loc[per + t];
    Fr rem = u2_loc[t] - w_loc[t] + h;
    if (!rem.is_zero()) atomicOr(status, ST_REMAINDER_NONZERO);
    if (!h.is_zero()) {
        if (k == n - 1) atomicOr(status, ST_H_DEGREE);
        else atomicOr(h_nonzero, 1u);
    }
}

__global__ void __launch_bounds__(256) k_assemble_phase1_strided(const Fr* __restrict__ u_loc, const Fr* __restrict__ u_prev,
                                                                 const Fr* __restrict__ u2_loc, const Fr* __restrict__ zt_loc,
                                                                 CLayout L, const Fr* __restrict__ ra, uint32_t rank, uint32_t world,
                                                                 uint64_t cnt_a, uint64_t cnt_c, Fr* __restrict__ scal_a_loc,
                                                                 Fr* __restrict__ scal_c_loc) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt_c) return;
    const uint64_t k = (uint64_t)rank + t * world;
    const uint64_t per = L.n / world;
    Fr sc, sa;
    // u_i for i = k (own residue class) or i = k - 1 (the class of rank - 1: u_prev; for rank 0 one local index lower)
    auto u_at = [&](uint64_t i) {
        if (i % world == rank) return u_loc[i / world];
        return u_prev[i / world];
    };
    phase1_scalar(L, k, ra, u_at, [&](uint64_t j) { return u2_loc[per + j / world]; }, [&](uint64_t j) { return zt_loc[j / world]; },
                  sc, sa);
    scal_c_loc[t] = sc;
    if (t < cnt_a) scal_a_loc[t] = sa;
}

// ---- virtual polynomial sources ---------------------------------------------------------
struct PlainSrc {
    const Fr* c;
    uint64_t len;
    __device__ __forceinline__ Fr at(uint64_t k) const { return c[k]; }
};

struct NumSrc {
    NumeratorSrc s;
    // prover.rs:142-209 with sigma = n + 3 (blocks never overlap):
    //   [0, 2)                 x2 * r_a
    //   [2s, 2s+3)             r_a + x2 * r_a^2
    //   [3s, 3s+n)             x2 * wu
    //   [5s, 5s+n+1)           u + x2 * 2 r_a u  - (a(x1) + x2 c(x1)) at 5s
    //   [8s, 8s+2n-1)          x2 * (ww + h_num) = x2 * u^2
    __device__ __forceinline__ Fr at(uint64_t k) const {
        const uint64_t sg = s.sigma, n = s.n;
        const Fr x2 = s.consts[0];
        if (k >= 8 * sg) {
            uint64_t j = k - 8 * sg;
            return (j < 2 * n - 1) ? x2 * s.u2[j] : Fr::zero();
        }
        if (k >= 5 * sg) {
            uint64_t j = k - 5 * sg;
            if (j > n) return Fr::zero();
            Fr t = Fr::zero();
            if (j < n) t = s.ra_ext[0] * s.u[j];
            if (j >= 1) t = t + s.ra_ext[1] * s.u[j - 1];
            Fr v = x2 * t.dbl();
            if (j < n) v = v + s.u[j];
            if (j == 0) v = v - s.consts[1];
            return v;
        }
        if (k >= 3 * sg) {
            uint64_t j = k - 3 * sg;
            return (j < n) ? x2 * s.wu[j] : Fr::zero();
        }
        if (k >= 2 * sg) {
            uint64_t j = k - 2 * sg;
            if (j >= 3) return Fr::zero();
            Fr v = x2 * s.ra_ext[2 + j];
            if (j < 2) v = v + s.ra_ext[j];
            return v;
        }
        if (k < 2) return x2 * s.ra_ext[k];
        return Fr::zero();
    }
};

// chunk value: sum_{k in chunk} p_k x^(k - lo).  Thread t owns kPerThread consecutive
// coefficients; per-thread Horner, then a weighted block reduction with powers of x^kPerThread.
template <class Src>
__global__ void __launch_bounds__(kThreads) k_chunk_eval(Src src, uint64_t len, const Fr* __restrict__ xp, Fr* __restrict__ chunk_vals,
                                                         uint64_t chunk0 = 0) {
    __shared__ Fr sh[kThreads];
    const Fr x = xp[0];
    const uint64_t chunk = chunk0 + blockIdx.x;
    const uint64_t lo = chunk * kChunk + (uint64_t)threadIdx.x * kPerThread;
    Fr acc = Fr::zero();
#pragma unroll 1
    for (int j = kPerThread - 1; j >= 0; j--) {
        uint64_t k = lo + j;
        acc = acc * x;
        if (k < len) acc = acc + src.at(k);
    }
    // weight by x^(kPerThread * tid)
    Fr xe = x;
#pragma unroll 1
    for (int b = 1; b < kPerThread; b <<= 1) xe = xe.sqr();
    Fr wgt = xe.pow_u64((uint64_t)threadIdx.x);
    sh[threadIdx.x] = acc * wgt;
    __syncthreads();
    for (int stride = kThreads / 2; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + stride];
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_vals[chunk] = sh[0];
}

__global__ void k_combine_chunks(const Fr* __restrict__ chunk_vals, uint64_t nchunks, const Fr* __restrict__ xp, Fr* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr xc = xp[0].pow_u64((uint64_t)kChunk);
    Fr acc = Fr::zero();
    for (uint64_t c = nchunks; c-- > 0;) acc = acc * xc + chunk_vals[c];
    out[0] = acc;
}

__global__ void k_a_at_x1(const Fr* u_at_x1, const Fr* ra, const Fr* x1_y1a, Fr* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr x1 = x1_y1a[0], y1a = x1_y1a[1];
    out[0] = u_at_x1[0] + (ra[0] + ra[1] * x1) * y1a;
}

// out[0] = x^kChunk, out[1] = x^(kChunk^2)
__global__ void k_chunk_powers(const Fr* __restrict__ xp, Fr* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr xc = xp[0].pow_u64((uint64_t)kChunk);
    out[0] = xc;
    out[1] = xc.pow_u64((uint64_t)kChunk);
}

// carries[c] = q_{(c+1)*kChunk - 1}: the quotient coefficient entering chunk c from above
// (xcp[0] = the evaluation point raised to the chunk length).  Serial: used on <= a few hundred values.
// carry_in (nullable): the quotient coefficient entering the topmost chunk (a rank's range of a sharded division).
__global__ void k_chunk_carries(const Fr* __restrict__ chunk_vals, uint64_t nchunks, const Fr* __restrict__ xcp,
                                Fr* __restrict__ carries, uint32_t* __restrict__ status, const Fr* __restrict__ carry_in = nullptr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr xc = xcp[0];
    Fr q = carry_in ? carry_in[0] : Fr::zero();
    for (uint64_t c = nchunks; c-- > 0;) {
        carries[c] = q;
        q = chunk_vals[c] + xc * q;
    }
    if (status != nullptr && !q.is_zero()) atomicOr(status, ST_OPENING_REMAINDER);  // q == p(x1): the remainder
}

template <class Src>
__global__ void __launch_bounds__(kThreads) k_chunk_divide(Src src, uint64_t len, const Fr* __restrict__ xp,
                                                           const Fr* __restrict__ carries, Fr* __restrict__ q, uint64_t chunk0 = 0) {
    __shared__ Fr sh[kThreads];
    const Fr x = xp[0];
    const int tid = threadIdx.x;
    const uint64_t chunk = chunk0 + blockIdx.x;
    const uint64_t lo = chunk * kChunk + (uint64_t)tid * kPerThread;
    Fr coef[kPerThread];
    Fr acc = Fr::zero();
#pragma unroll 1
    for (int j = kPerThread - 1; j >= 0; j--) {
        uint64_t k = lo + j;
        coef[j] = (k < len) ? src.at(k) : Fr::zero();
        acc = acc * x + coef[j];
    }
    // suffix scan over threads: in_t = sum_{t' > t} H_{t'} x^(E (t'-t-1)) + x^(E (T-1-t)) * carry_chunk
    // Kogge-Stone on (value) with the uniform multiplier x^(E * 2^s) at step s.
    Fr xe = x;
#pragma unroll 1
    for (int b = 1; b < kPerThread; b <<= 1) xe = xe.sqr();
    // fold the chunk carry into the last thread's value: H'_{T-1} = H_{T-1} + x^E * carry
    if (tid == kThreads - 1) acc = acc + xe * carries[chunk];
    sh[tid] = acc;
    __syncthreads();
    Fr mult = xe;
    for (int d = 1; d < kThreads; d <<= 1) {
        Fr add = (tid + d < kThreads) ? sh[tid + d] * mult : Fr::zero();
        __syncthreads();
        if (tid + d < kThreads) sh[tid] = sh[tid] + add;
        __syncthreads();
        mult = mult.sqr();
    }
    // sh[t] = sum_{t' >= t} H'_{t'} x^(E (t'-t)); carry into thread t is sh[t+1] (or the chunk carry)
    Fr carry = (tid + 1 < kThreads) ? sh[tid + 1] : carries[chunk];
#pragma unroll 1
    for (int j = kPerThread - 1; j >= 0; j--) {
        uint64_t k = lo + j;
        carry = coef[j] + x * carry;
        if (k >= 1 && k < len) q[k - 1] = carry;
    }
}

// Exchange step of the sharded division: E_s = value of rank s's chunk range at x (relative to its first chunk lo_s).
// carry into rank r's range = sum_{s > r} E_s X^(lo_s - hi_r), X = x^kChunk; p(x) = sum_s E_s X^lo_s must vanish.
__global__ void k_range_carry(const Fr* __restrict__ range_vals, const uint64_t* __restrict__ range_lo, uint32_t rank, uint32_t world,
                              const Fr* __restrict__ xp, Fr* __restrict__ carry_in, uint32_t* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Fr X = xp[0].pow_u64((uint64_t)kChunk);
    Fr total = Fr::zero(), carry = Fr::zero();
    const uint64_t hi = range_lo[rank + 1];
    for (uint32_t s = 0; s < world; s++) {
        if (range_lo[s + 1] == range_lo[s]) continue;                // empty range
        total = total + range_vals[s] * X.pow_u64(range_lo[s]);
        if (s > rank) carry = carry + range_vals[s] * X.pow_u64(range_lo[s] - hi);
    }
    carry_in[0] = carry;
    if (!total.is_zero()) atomicOr(status, ST_OPENING_REMAINDER);
}

// v[0] += x^kChunk * carry: folds the carry entering a range from above into its top chunk value (see
// launch_numerator_range_divide)
__global__ void k_fold_carry(Fr* __restrict__ v, const Fr* __restrict__ xcp, const Fr* __restrict__ carry) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    v[0] = v[0] + xcp[0] * carry[0];
}

template <class Src>
__global__ void k_materialize(Src src, uint64_t len, Fr* out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < len) out[k] = src.at(k);
}

}  // namespace

void launch_sap_evals(const SapDims& d, const DevCsr& A, const DevCsr& B, const DevCsr& C, Fr* ztail, Fr* u_ev, Fr* w_ev,
                      Fr* wu_ev, cudaStream_t stream) {
    k_sap_public_rows<<<ceil_div(d.m0, 128), 128, 0, stream>>>(d, ztail, u_ev, w_ev, wu_ev);
    PM_LAUNCH_CHECK();
    if (d.nr) {
        k_sap_constraint_rows<<<ceil_div(d.nr, 128), 128, 0, stream>>>(d, A, B, C, ztail, u_ev, w_ev, wu_ev);
        PM_LAUNCH_CHECK();
    }
    uint64_t rows = 2 * ((uint64_t)d.m0 + d.nr);
    if (rows < d.n) {
        k_zero_tail<<<ceil_div(d.n - rows, 256), 256, 0, stream>>>(u_ev, w_ev, wu_ev, rows, d.n);
        PM_LAUNCH_CHECK();
    }
}

void launch_square(Fr* data, size_t n, cudaStream_t stream) {
    k_square<<<ceil_div(n, 256), 256, 0, stream>>>(data, n);
    PM_LAUNCH_CHECK();
}

void launch_quotient_checks(const Fr* u2, const Fr* w, uint64_t n, uint32_t* status, cudaStream_t stream) {
    // status[1] is scratch for "h has a non-zero coefficient"
    PM_CUDA(cudaMemsetAsync(status + 1, 0, sizeof(uint32_t), stream));
    k_quotient_checks<<<ceil_div(n, 256), 256, 0, stream>>>(u2, w, n, status, status + 1);
    PM_LAUNCH_CHECK();
    k_quotient_finish<<<1, 32, 0, stream>>>(status, status + 1);
    PM_LAUNCH_CHECK();
}

void launch_ra_square(Fr* ra_ext, cudaStream_t stream) {
    k_ra_square<<<1, 32, 0, stream>>>(ra_ext);
    PM_LAUNCH_CHECK();
}

void launch_assemble_phase1_scalars(const Fr* u, const Fr* u2, const Fr* ztail, const CLayout& L, const Fr* ra_ext,
                                    Fr* scal_a, Fr* scal_c, cudaStream_t stream) {
    k_assemble_phase1<<<ceil_div(L.len_c, 256), 256, 0, stream>>>(u, u2, ztail, L, ra_ext, scal_a, scal_c);
    PM_LAUNCH_CHECK();
}

void launch_sap_evals_strided(const SapDims& d, const DevCsr& A, const DevCsr& B, const DevCsr& C, const Fr* ztail,
                              uint32_t rank, uint32_t world, Fr* u_loc, Fr* w_loc, Fr* wu_loc, cudaStream_t stream) {
    const uint64_t per = d.n / world;
    k_sap_rows_strided<<<ceil_div(per, 128), 128, 0, stream>>>(d, A, B, C, ztail, rank, world, u_loc, w_loc, wu_loc);
    PM_LAUNCH_CHECK();
}
void launch_ztail_strided(const SapDims& d, const DevCsr& A, const DevCsr& B, const Fr* ztail, uint32_t rank, uint32_t world,
                          uint64_t count, Fr* zt_loc, cudaStream_t stream) {
    if (!count) return;
    k_ztail_strided<<<ceil_div(count, 128), 128, 0, stream>>>(d, A, B, ztail, rank, world, count, zt_loc);
    PM_LAUNCH_CHECK();
}
void launch_quotient_checks_strided(const Fr* u2_loc, const Fr* w_loc, uint64_t n, uint32_t rank, uint32_t world,
                                    uint32_t* status, cudaStream_t stream) {
    PM_CUDA(cudaMemsetAsync(status + 1, 0, sizeof(uint32_t), stream));
    k_quotient_checks_strided<<<ceil_div(n / world, 256), 256, 0, stream>>>(u2_loc, w_loc, n, rank, world, status, status + 1);
    PM_LAUNCH_CHECK();
    // "h == 0" needs every rank's slice: status[1] travels with the partial sums and is combined on the host
}
void launch_assemble_phase1_strided(const Fr* u_loc, const Fr* u_prev, const Fr* u2_loc, const Fr* zt_loc, const CLayout& L,
                                    const Fr* ra_ext, uint32_t rank, uint32_t world, uint64_t cnt_a, uint64_t cnt_c,
                                    Fr* scal_a_loc, Fr* scal_c_loc, cudaStream_t stream) {
    if (!cnt_c) return;
    k_assemble_phase1_strided<<<ceil_div(cnt_c, 256), 256, 0, stream>>>(u_loc, u_prev, u2_loc, zt_loc, L, ra_ext, rank, world, cnt_a,
                                                                      cnt_c, scal_a_loc, scal_c_loc);
    PM_LAUNCH_CHECK();
}

void launch_chunk_eval_plain(const Fr* coeffs, uint64_t len, const Fr* x, Fr* chunk_vals, cudaStream_t stream) {
    PlainSrc s{coeffs, len};
    k_chunk_eval<PlainSrc><<<ceil_div(len, kChunk), kThreads, 0, stream>>>(s, len, x, chunk_vals);
    PM_LAUNCH_CHECK();
}
void launch_chunk_eval_numerator(const NumeratorSrc& src, const Fr* x, Fr* chunk_vals, cudaStream_t stream) {
    NumSrc s{src};
    k_chunk_eval<NumSrc><<<ceil_div(src.len, kChunk), kThreads, 0, stream>>>(s, src.len, x, chunk_vals);
    PM_LAUNCH_CHECK();
}
void launch_combine_chunks(const Fr* chunk_vals, uint64_t nchunks, const Fr* x, Fr* out, cudaStream_t stream) {
    k_combine_chunks<<<1, 32, 0, stream>>>(chunk_vals, nchunks, x, out);
    PM_LAUNCH_CHECK();
}
void launch_a_at_x1(const Fr* u_at_x1, const Fr* ra_ext, const Fr* x1_y1a, Fr* out, cudaStream_t stream) {
    k_a_at_x1<<<1, 32, 0, stream>>>(u_at_x1, ra_ext, x1_y1a, out);
    PM_LAUNCH_CHECK();
}
// q[k-1] = p_k + x q_k for the virtual numerator (q: len-1 entries); sets ST_OPENING_REMAINDER when
// p(x) != 0.  Two-level carry propagation: chunk values -> (chunks of chunk values) -> serial over
// the few level-2 values -> parallel division of the chunk polynomial -> parallel division proper.
// work: >= 2*(nchunks + 1) + 2*(nchunks/kChunk + 2) + 2 elements.
int launch_divide_numerator(const NumeratorSrc& src, const Fr* x, Fr* q, Fr* work, uint32_t* status, cudaStream_t stream) {
    NumSrc s{src};
    const uint64_t c1 = (src.len + kChunk - 1) / kChunk;
    const uint64_t c2 = (c1 + kChunk - 1) / kChunk;
    Fr* vals1 = work;
    Fr* carr1 = vals1 + c1 + 1;
    Fr* vals2 = carr1 + c1 + 1;
    Fr* carr2 = vals2 + c2 + 1;
    Fr* xpow = carr2 + c2 + 1;   // [x^kChunk, x^(kChunk^2)]
    int launches = 0;
    k_chunk_powers<<<1, 32, 0, stream>>>(x, xpow);
    k_chunk_eval<NumSrc><<<(unsigned)c1, kThreads, 0, stream>>>(s, src.len, x, vals1);
    PM_LAUNCH_CHECK();
    launches += 2;
    if (c1 <= 128) {
        k_chunk_carries<<<1, 32, 0, stream>>>(vals1, c1, xpow, carr1, status);
        launches += 1;
    } else {
        PlainSrc ps{vals1, c1};
        k_chunk_eval<PlainSrc><<<(unsigned)c2, kThreads, 0, stream>>>(ps, c1, xpow, vals2);
        k_chunk_carries<<<1, 32, 0, stream>>>(vals2, c2, xpow + 1, carr2, status);
        // carries of level 1 = quotient of the chunk polynomial by (Y - x^kChunk), shifted by one
        PM_CUDA(cudaMemsetAsync(carr1 + (c1 - 1), 0, sizeof(Fr), stream));
        k_chunk_divide<PlainSrc><<<(unsigned)c2, kThreads, 0, stream>>>(ps, c1, xpow, carr2, carr1);
        launches += 3;
    }
    PM_LAUNCH_CHECK();
    k_chunk_divide<NumSrc><<<(unsigned)c1, kThreads, 0, stream>>>(s, src.len, x, carr1, q);
    PM_LAUNCH_CHECK();
    return launches + 1;
}
namespace {
struct RangeWork {      // the layout launch_divide_numerator uses for `work`, shared by the range functions
    Fr *vals1, *carr1, *vals2, *carr2, *xpow;
    RangeWork(Fr* work, uint64_t len) {
        const uint64_t c1 = (len + kChunk - 1) / kChunk, c2 = (c1 + kChunk - 1) / kChunk;
        vals1 = work;
        carr1 = vals1 + c1 + 1;
        vals2 = carr1 + c1 + 1;
        carr2 = vals2 + c2 + 1;
        xpow = carr2 + c2 + 1;
    }
};
}  // namespace

int launch_numerator_range_eval(const NumeratorSrc& src, const Fr* x, uint64_t c_lo, uint64_t cnt, Fr* work, Fr* range_val,
                                cudaStream_t stream) {
    NumSrc s{src};
    RangeWork w(work, src.len);
    k_chunk_powers<<<1, 32, 0, stream>>>(x, w.xpow);
    PM_LAUNCH_CHECK();
    if (cnt == 0) {
        PM_CUDA(cudaMemsetAsync(range_val, 0, sizeof(Fr), stream));
        return 1;
    }
    k_chunk_eval<NumSrc><<<(unsigned)cnt, kThreads, 0, stream>>>(s, src.len, x, w.vals1, c_lo);
    PM_LAUNCH_CHECK();
    if (cnt <= 128) {
        k_combine_chunks<<<1, 32, 0, stream>>>(w.vals1 + c_lo, cnt, x, range_val);
        PM_LAUNCH_CHECK();
        return 3;
    }
    PlainSrc ps{w.vals1 + c_lo, cnt};
    const uint64_t c2 = (cnt + kChunk - 1) / kChunk;
    k_chunk_eval<PlainSrc><<<(unsigned)c2, kThreads, 0, stream>>>(ps, cnt, w.xpow, w.vals2);
    k_combine_chunks<<<1, 32, 0, stream>>>(w.vals2, c2, w.xpow, range_val);     // point x^kChunk: raised to kChunk inside
    PM_LAUNCH_CHECK();
    return 4;
}

void launch_numerator_range_carry(const Fr* range_vals, const uint64_t* range_lo, uint32_t rank, uint32_t world, const Fr* x,
                                  Fr* carry_in, uint32_t* status, cudaStream_t stream) {
    k_range_carry<<<1, 32, 0, stream>>>(range_vals, range_lo, rank, world, x, carry_in, status);
    PM_LAUNCH_CHECK();
}

int launch_numerator_range_divide(const NumeratorSrc& src, const Fr* x, uint64_t c_lo, uint64_t cnt, const Fr* carry_in, Fr* q,
                                  Fr* work, cudaStream_t stream) {
    if (cnt == 0) return 0;
    NumSrc s{src};
    RangeWork w(work, src.len);
    int launches = 0;
    if (cnt <= 128) {
        k_chunk_carries<<<1, 32, 0, stream>>>(w.vals1 + c_lo, cnt, w.xpow, w.carr1 + c_lo, nullptr, carry_in);
        launches += 1;
    } else {
        // The carry K entering the range from above is the quotient coefficient above its top chunk: q_{top-1} =
        // p_top + x K, i.e. a zero incoming carry with the top coefficient raised by x K — for the chunk polynomial: the top
        // chunk value raised by x^kChunk K.  (Feeding K itself into the second-level recurrence would be wrong: its top
        // chunk is zero-padded, and a carry does not commute with the padding.)  Then the unsharded two-level scheme:
        // carries of level 1 = quotient of the chunk polynomial by (Y - x^kChunk), shifted by one; the top one is K.
        PlainSrc ps{w.vals1 + c_lo, cnt};
        const uint64_t c2 = (cnt + kChunk - 1) / kChunk;
        k_fold_carry<<<1, 32, 0, stream>>>(w.vals1 + c_lo + (cnt - 1), w.xpow, carry_in);
        k_chunk_eval<PlainSrc><<<(unsigned)c2, kThreads, 0, stream>>>(ps, cnt, w.xpow, w.vals2);
        k_chunk_carries<<<1, 32, 0, stream>>>(w.vals2, c2, w.xpow + 1, w.carr2, nullptr, nullptr);
        PM_CUDA(cudaMemcpyAsync(w.carr1 + c_lo + (cnt - 1), carry_in, sizeof(Fr), cudaMemcpyDeviceToDevice, stream));
        k_chunk_divide<PlainSrc><<<(unsigned)c2, kThreads, 0, stream>>>(ps, cnt, w.xpow, w.carr2, w.carr1 + c_lo);
        launches += 4;
    }
    PM_LAUNCH_CHECK();
    k_chunk_divide<NumSrc><<<(unsigned)cnt, kThreads, 0, stream>>>(s, src.len, x, w.carr1, q, c_lo);
    PM_LAUNCH_CHECK();
    return launches + 1;
}

void launch_materialize_numerator(const NumeratorSrc& src, Fr* out, cudaStream_t stream) {
    NumSrc s{src};
    k_materialize<NumSrc><<<ceil_div(src.len, 256), 256, 0, stream>>>(s, src.len, out);
    PM_LAUNCH_CHECK();
}

}  // namespace pm
