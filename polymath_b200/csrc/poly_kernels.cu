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

__global__ void __launch_bounds__(256) k_assemble_phase1(const Fr* __restrict__ u, const Fr* __restrict__ u2,
                                                         const Fr* __restrict__ ztail, uint64_t tail,
                                                         const Fr* __restrict__ ra, uint64_t n, Fr* __restrict__ scal_a,
                                                         Fr* __restrict__ scal_c) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t len_c = (n + 1) + 3 + 2 + (n - 1) + tail;
    if (k >= len_c) return;
    if (k <= n) {
        // 2 r_a(X) u(X): coefficient k = 2 (r0 u_k + r1 u_{k-1})
        Fr acc = Fr::zero();
        if (k < n) acc = ra[0] * u[k];
        if (k >= 1) acc = acc + ra[1] * u[k - 1];
        scal_c[k] = acc.dbl();
        // a-side scalars share the first n + 4 bases
        scal_a[k] = (k < n) ? u[k] : Fr::zero();
    } else if (k < n + 4) {
        uint64_t j = k - (n + 1);
        scal_c[k] = ra[2 + j];
        scal_a[k] = (j < 2) ? ra[j] : Fr::zero();
    } else if (k < n + 6) {
        scal_c[k] = ra[k - (n + 4)];
    } else if (k < n + 6 + (n - 1)) {
        scal_c[k] = u2[n + (k - (n + 6))];   // h = u2[n .. 2n-1)
    } else {
        scal_c[k] = ztail[k - (n + 6 + (n - 1))];
    }
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
__global__ void __launch_bounds__(kThreads) k_chunk_eval(Src src, uint64_t len, const Fr* __restrict__ xp, Fr* __restrict__ chunk_vals) {
    __shared__ Fr sh[kThreads];
    const Fr x = xp[0];
    const uint64_t lo = (uint64_t)blockIdx.x * kChunk + (uint64_t)threadIdx.x * kPerThread;
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
    if (threadIdx.x == 0) chunk_vals[blockIdx.x] = sh[0];
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
__global__ void k_chunk_carries(const Fr* __restrict__ chunk_vals, uint64_t nchunks, const Fr* __restrict__ xcp,
                                Fr* __restrict__ carries, uint32_t* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr xc = xcp[0];
    Fr q = Fr::zero();
    for (uint64_t c = nchunks; c-- > 0;) {
        carries[c] = q;
        q = chunk_vals[c] + xc * q;
    }
    if (status != nullptr && !q.is_zero()) atomicOr(status, ST_OPENING_REMAINDER);  // q == p(x1): the remainder
}

template <class Src>
__global__ void __launch_bounds__(kThreads) k_chunk_divide(Src src, uint64_t len, const Fr* __restrict__ xp,
                                                           const Fr* __restrict__ carries, Fr* __restrict__ q) {
    __shared__ Fr sh[kThreads];
    const Fr x = xp[0];
    const int tid = threadIdx.x;
    const uint64_t lo = (uint64_t)blockIdx.x * kChunk + (uint64_t)tid * kPerThread;
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
    if (tid == kThreads - 1) acc = acc + xe * carries[blockIdx.x];
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
    Fr carry = (tid + 1 < kThreads) ? sh[tid + 1] : carries[blockIdx.x];
#pragma unroll 1
    for (int j = kPerThread - 1; j >= 0; j--) {
        uint64_t k = lo + j;
        carry = coef[j] + x * carry;
        if (k >= 1 && k < len) q[k - 1] = carry;
    }
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

void launch_assemble_phase1_scalars(const Fr* u, const Fr* u2, const Fr* ztail, uint64_t tail, const Fr* ra_ext, uint64_t n,
                                    Fr* scal_a, Fr* scal_c, cudaStream_t stream) {
    const uint64_t len_c = (n + 1) + 3 + 2 + (n - 1) + tail;
    k_assemble_phase1<<<ceil_div(len_c, 256), 256, 0, stream>>>(u, u2, ztail, tail, ra_ext, n, scal_a, scal_c);
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
void launch_materialize_numerator(const NumeratorSrc& src, Fr* out, cudaStream_t stream) {
    NumSrc s{src};
    k_materialize<NumSrc><<<ceil_div(src.len, 256), 256, 0, stream>>>(s, src.len, out);
    PM_LAUNCH_CHECK();
}

}  // namespace pm
