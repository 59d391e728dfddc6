// Generator (setup) on the device.
//
// Follows `generate_proving_key` (/root/reference/src/generator.rs:60-157):
//   y = x^sigma, y^alpha = (1/y)^3, y^-alpha = y^3, y^gamma = (1/y)^5           generator.rs:72-77
//   x_powers_g1                = [x^j]                 j <= n                     :82
//   x_powers_y_alpha_g1        = [x^j y^alpha]         j <= 2                     :86
//   x_powers_y_gamma_g1        = [x^j y^gamma]         j <= 1                     :90
//   x_powers_y_gamma_z_g1      = [x^j y^gamma z]       j <= 2(n-1) + 8 sigma      :94-100
//   x_powers_zh_by_y_alpha_g1  = [x^j Z_H(x) y^-alpha] j <= n-2                   :105-108
//   uj_wj_lcs_by_y_alpha_g1    = [(u_j(x) y^gamma + w_j(x)) y^-alpha]  j >= m0    :112-135
// The reference evaluates u_j(x), w_j(x) by dense column dots over the virtual SAP matrices
// (O(n*m)); here they are transposed SpMVs of the Lagrange vector over the CSC forms of
// A, B, C with the closed form of the SAP view.  Scalars are produced by
// parallel prefix of x, the points by the fixed-base engine (K5).  [x]_2 and [z]_2
// (generator.rs:144-145) are two G2 scalar multiplications done by two device threads.
#include "fixed_base.cuh"
#include "prover.cuh"
#include "roots.cuh"

namespace pm {

namespace {

constexpr int kSeq = 16;  // consecutive outputs per thread in the prefix kernels

enum ConstSlot { C_X = 0, C_Z, C_ONE, C_Y_ALPHA, C_Y_GAMMA, C_Y_GAMMA_Z, C_ZH_Y3, C_Y3, C_LAG_PREF, C_COUNT };

__global__ void k_setup_consts(Fr* c, uint64_t n, uint64_t sigma) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr x = c[C_X], z = c[C_Z];
    Fr y = x.pow_u64(sigma);
    Fr yinv = y.inv();
    Fr y3 = y.sqr() * y;                          // y^-alpha = y^3       generator.rs:75
    Fr ya = yinv.sqr() * yinv;                    // y^alpha  = (1/y)^3   generator.rs:74
    Fr yg = ya * yinv.sqr();                      // y^gamma  = (1/y)^5   generator.rs:76
    Fr zh = x.pow_u64(n) - Fr::one();             // Z_H(x)               generator.rs:106
    Fr nf = Fr::zero();
    nf.v[0] = (uint32_t)n; nf.v[1] = (uint32_t)(n >> 32);
    nf = nf.to_mont();
    c[C_ONE] = Fr::one();
    c[C_Y_ALPHA] = ya;
    c[C_Y_GAMMA] = yg;
    c[C_Y_GAMMA_Z] = yg * z;
    c[C_ZH_Y3] = zh * y3;
    c[C_Y3] = y3;
    c[C_LAG_PREF] = zh * nf.inv();                // Z_H(x) / n
}

// out[j] = mult * x^j for j < count
__global__ void __launch_bounds__(128) k_powers(Fr* __restrict__ out, uint64_t count, const Fr* __restrict__ xp,
                                                const Fr* __restrict__ multp) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = t * kSeq;
    if (lo >= count) return;
    Fr x = xp[0];
    Fr cur = multp[0] * x.pow_u64(lo);
    uint64_t hi = lo + kSeq < count ? lo + kSeq : count;
    for (uint64_t j = lo; j < hi; j++) {
        out[j] = cur;
        cur = cur * x;
    }
}

// L_i(x) = (Z_H(x)/n) * w^i / (x - w^i)   (evaluate_all_lagrange_coefficients, generator.rs:113)
__global__ void __launch_bounds__(128) k_lagrange(Fr* __restrict__ L, uint64_t n, int log_n, const Fr* __restrict__ c) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = t * kSeq;
    if (lo >= n) return;
    int cnt = (int)((n - lo) < (uint64_t)kSeq ? (n - lo) : (uint64_t)kSeq);
    Fr x = c[C_X], pref = c[C_LAG_PREF];
    Fr w = root_of_unity(log_n, false);
    Fr wi = w.pow_u64(lo);
    Fr num[kSeq], den[kSeq], prefix[kSeq];
    Fr acc = Fr::one();
    for (int k = 0; k < cnt; k++) {
        num[k] = wi;
        den[k] = x - wi;
        prefix[k] = acc;
        acc = acc * den[k];
        wi = wi * w;
    }
    Fr inv = acc.inv();
    for (int k = cnt - 1; k >= 0; k--) {
        Fr dinv = inv * prefix[k];
        inv = inv * den[k];
        L[lo + k] = pref * num[k] * dinv;
    }
}

struct DevCsc {
    const uint32_t* col_ptr;
    const uint32_t* row;
    const Fr* val;
};

// scalars of uj_wj_lcs_by_y_alpha_g1 (generator.rs:115-135) via the closed form of the SAP view
__global__ void __launch_bounds__(128) k_lcs_scalars(DevCsc A, DevCsc B, DevCsc C, const Fr* __restrict__ L,
                                                     uint32_t m0, uint32_t mw, uint32_t nr, const Fr* __restrict__ c,
                                                     Fr* __restrict__ out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)2 * m0 + mw + nr;
    if (j >= total) return;
    const Fr y3 = c[C_Y3];
    const Fr* L1 = L + 2 * (uint64_t)m0;       // rows 2 m0 + r
    const Fr* L2 = L1 + nr;                    // rows 2 m0 + nr + r
    if (j < (uint64_t)m0 + mw) {
        uint32_t k = (uint32_t)j;              // R1CS column
        Fr ua = Fr::zero(), wa = Fr::zero();
        for (uint32_t e = A.col_ptr[k]; e < A.col_ptr[k + 1]; e++) {
            uint32_t r = A.row[e];
            ua = ua + A.val[e] * (L1[r] + L2[r]);
        }
        for (uint32_t e = B.col_ptr[k]; e < B.col_ptr[k + 1]; e++) {
            uint32_t r = B.row[e];
            ua = ua + B.val[e] * (L1[r] - L2[r]);
        }
        for (uint32_t e = C.col_ptr[k]; e < C.col_ptr[k + 1]; e++) wa = wa + C.val[e] * L1[C.row[e]];
        if (k < m0) wa = wa + L[k];            // W[k][m0 + k] = 4
        wa = wa.dbl().dbl();
        out[j] = (ua * c[C_Y_GAMMA] + wa) * y3;
    } else {
        uint64_t t = j - ((uint64_t)m0 + mw);
        Fr v = (t < m0) ? (L[t] + L[m0 + t]) : (L1[t - m0] + L2[t - m0]);
        out[j] = v * y3;
    }
}

// dst[k] = src[k * stride + offset] for k < count   (this rank's share of a scalar vector)
__global__ void k_gather_strided(const Fr* __restrict__ src, uint64_t stride, uint64_t offset, uint64_t count, Fr* __restrict__ dst) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dst[k] = src[k * stride + offset];
}

// ---- G2 = E'(Fq2): y^2 = x^3 + 4(1 + u), Jacobian coordinates, a = 0 --------------------
struct Fq2 {
    Fq c0, c1;
};
__device__ __forceinline__ Fq2 f2_add(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
__device__ __forceinline__ Fq2 f2_sub(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
__device__ __noinline__ Fq2 f2_mul(const Fq2& a, const Fq2& b) {
    Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    Fq s = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {t0 - t1, s - t0 - t1};
}
__device__ __forceinline__ Fq2 f2_sqr(const Fq2& a) { return f2_mul(a, a); }
__device__ __forceinline__ Fq2 f2_dbl(const Fq2& a) { return f2_add(a, a); }
__device__ __forceinline__ bool f2_is_zero(const Fq2& a) { return a.c0.is_zero() && a.c1.is_zero(); }
__device__ __noinline__ Fq2 f2_inv(const Fq2& a) {
    Fq d = (a.c0.sqr() + a.c1.sqr()).inv();
    return {a.c0 * d, (a.c1 * d).neg()};
}
struct G2Jac {
    Fq2 x, y, z;
};
__device__ __noinline__ void g2_dbl(G2Jac& p) {  // dbl-2009-l
    if (f2_is_zero(p.z)) return;
    Fq2 A = f2_sqr(p.x), B = f2_sqr(p.y), C = f2_sqr(B);
    Fq2 t = f2_add(p.x, B);
    Fq2 D = f2_dbl(f2_sub(f2_sub(f2_sqr(t), A), C));
    Fq2 E = f2_add(f2_dbl(A), A);
    Fq2 F = f2_sqr(E);
    Fq2 x3 = f2_sub(F, f2_dbl(D));
    Fq2 c8 = f2_dbl(f2_dbl(f2_dbl(C)));
    Fq2 z3 = f2_dbl(f2_mul(p.y, p.z));
    p.y = f2_sub(f2_mul(E, f2_sub(D, x3)), c8);
    p.x = x3;
    p.z = z3;
}
__device__ __noinline__ void g2_add_affine(G2Jac& p, const Fq2& qx, const Fq2& qy) {  // madd, distinct points
    Fq2 one{Fq::one(), Fq::zero()};
    if (f2_is_zero(p.z)) { p.x = qx; p.y = qy; p.z = one; return; }
    Fq2 z1z1 = f2_sqr(p.z);
    Fq2 u2 = f2_mul(qx, z1z1);
    Fq2 s2 = f2_mul(f2_mul(qy, p.z), z1z1);
    Fq2 h = f2_sub(u2, p.x), r = f2_sub(s2, p.y);
    if (f2_is_zero(h)) {
        if (f2_is_zero(r)) { g2_dbl(p); return; }
        p.z = Fq2{Fq::zero(), Fq::zero()};
        return;
    }
    Fq2 hh = f2_sqr(h), hhh = f2_mul(h, hh), v = f2_mul(p.x, hh);
    Fq2 x3 = f2_sub(f2_sub(f2_sqr(r), hhh), f2_dbl(v));
    p.y = f2_sub(f2_mul(r, f2_sub(v, x3)), f2_mul(p.y, hhh));
    p.x = x3;
    p.z = f2_mul(p.z, h);
}

static __device__ __constant__ uint32_t G2_X0[12] = {0x02940a10u, 0xf5f28fa2u, 0x87b4961au, 0xb3f5fb26u, 0x3e2ae580u, 0xa1a893b5u, 0x1a3caee9u, 0x9894999du, 0x1863366bu, 0x6f67b763u, 0x4350bcd7u, 0x05819192u};
static __device__ __constant__ uint32_t G2_X1[12] = {0x9e23f606u, 0xa5a9c075u, 0xbccd60c3u, 0xaaa0c59du, 0xe2867806u, 0x3bb17e18u, 0x8541b367u, 0x1b1ab6ccu, 0xf2158547u, 0xc2b6ed0eu, 0x7360edf3u, 0x11922a09u};
static __device__ __constant__ uint32_t G2_Y0[12] = {0x60494c4au, 0x4c730af8u, 0x5e369c5au, 0x597cfa1fu, 0xaa0a635au, 0xe7e6856cu, 0x6e0d495fu, 0xbbefb5e9u, 0xf0ef25a2u, 0x07d3a975u, 0x7e80dae5u, 0x0083fd8eu};
static __device__ __constant__ uint32_t G2_Y1[12] = {0xdf64b05du, 0xadc0fc92u, 0x2b1461dcu, 0x18aa270au, 0x3be4eba0u, 0x86adac6au, 0xc93da33au, 0x79495c4eu, 0xa43ccaedu, 0xe7175850u, 0x63de1bf2u, 0x0b2bc2a1u};

// out[t] = affine(scalars[t] * G2) as (x.c0, x.c1, y.c0, y.c1); infinity = all zero.  One thread per scalar.
__global__ void k_g2_mul(const Fr* __restrict__ scalars, int count, Fq* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    Fq2 gx, gy;
#pragma unroll
    for (int i = 0; i < 12; i++) { gx.c0.v[i] = G2_X0[i]; gx.c1.v[i] = G2_X1[i]; gy.c0.v[i] = G2_Y0[i]; gy.c1.v[i] = G2_Y1[i]; }
    Fr s = scalars[t].from_mont();
    G2Jac acc;
    acc.x = acc.y = acc.z = Fq2{Fq::zero(), Fq::zero()};
    for (int bit = 254; bit >= 0; bit--) {
        g2_dbl(acc);
        if ((s.v[bit >> 5] >> (bit & 31)) & 1) g2_add_affine(acc, gx, gy);
    }
    Fq* o = out + 4 * t;
    if (f2_is_zero(acc.z)) {
        o[0] = o[1] = o[2] = o[3] = Fq::zero();
        return;
    }
    Fq2 zi = f2_inv(acc.z);
    Fq2 zi2 = f2_sqr(zi);
    Fq2 ax = f2_mul(acc.x, zi2);
    Fq2 ay = f2_mul(acc.y, f2_mul(zi2, zi));
    o[0] = ax.c0; o[1] = ax.c1; o[2] = ay.c0; o[3] = ay.c1;
}

}  // namespace

void run_setup(ProverCtx& ctx, const uint8_t* x, const uint8_t* z, uint8_t* x_g2, uint8_t* z_g2) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    const uint64_t n = ctx.n;
    DevBuf consts_b, lag_b, scal_b, g2_b;
    Fr* c = consts_b.as<Fr>(C_COUNT);
    PM_CUDA(cudaMemcpyAsync(c + C_X, x, sizeof(Fr), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemcpyAsync(c + C_Z, z, sizeof(Fr), cudaMemcpyHostToDevice, s));
    k_setup_consts<<<1, 32, 0, s>>>(c, n, ctx.sigma);
    PM_LAUNCH_CHECK();

    auto powers = [&](Fr* out, uint64_t count, int mult_slot) {
        uint64_t threads = (count + kSeq - 1) / kSeq;
        k_powers<<<ceil_div(threads, 128), 128, 0, s>>>(out, count, c + C_X, c + mult_slot);
        PM_LAUNCH_CHECK();
        rt.extra_launches++;
    };

    // c-side scalars, laid out like ProverCtx::bases_c
    const uint64_t len_c = ctx.len_c(), len_d = ctx.len_d();
    const CLayout lay = ctx.lay();
    Fr* sc = scal_b.as<Fr>(len_c > len_d ? len_c : len_d);
    PM_CUDA(cudaMemsetAsync(sc, 0, len_c * sizeof(Fr), s));      // the gaps of the layout: [0]G = the point at infinity
    powers(sc, n + 1, C_ONE);
    powers(sc + lay.off_ya, 3, C_Y_ALPHA);
    powers(sc + lay.off_yg, 2, C_Y_GAMMA);
    powers(sc + lay.off_zh, n - 1, C_ZH_Y3);
    Fr* L = lag_b.as<Fr>(n);
    {
        uint64_t threads = (n + kSeq - 1) / kSeq;
        k_lagrange<<<ceil_div(threads, 128), 128, 0, s>>>(L, n, ctx.log_n, c);
        PM_LAUNCH_CHECK();
        DevCsc A{ctx.A.col_ptr.get<uint32_t>(), ctx.A.row.get<uint32_t>(), ctx.A.cval.get<Fr>()};
        DevCsc B{ctx.B.col_ptr.get<uint32_t>(), ctx.B.row.get<uint32_t>(), ctx.B.cval.get<Fr>()};
        DevCsc C{ctx.C.col_ptr.get<uint32_t>(), ctx.C.row.get<uint32_t>(), ctx.C.cval.get<Fr>()};
        const uint64_t total = ctx.cols - ctx.m0;
        k_lcs_scalars<<<ceil_div(total, 128), 128, 0, s>>>(A, B, C, L, (uint32_t)ctx.m0, (uint32_t)ctx.mw, (uint32_t)ctx.nr, c,
                                                          sc + lay.off_lcs);
        PM_LAUNCH_CHECK();
        rt.extra_launches += 3;
    }
    // the expensive part — one fixed-base multiplication per point — runs only on this rank's share
    DevBuf share_b;
    auto run_share = [&](uint64_t total, G1Affine* out) {
        if (ctx.world == 1) { rt.fixed_base.run(sc, total, out, s); return; }
        const uint64_t cnt = ctx.local_count(total);
        Fr* share = share_b.as<Fr>(cnt + 1);
        if (cnt) {
            k_gather_strided<<<ceil_div(cnt, 256), 256, 0, s>>>(sc, (uint64_t)ctx.world, (uint64_t)ctx.rank, cnt, share);
            PM_LAUNCH_CHECK();
            rt.extra_launches++;
        }
        rt.fixed_base.run(share, cnt, out, s);
    };
    run_share(len_c, ctx.bases_c.get<G1Affine>());
    // d-side: contiguous range of this rank (the chunk range of its share of the opening division)
    powers(sc, len_d, C_Y_GAMMA_Z);
    if (ctx.world == 1) rt.fixed_base.run(sc, len_d, ctx.bases_d.get<G1Affine>(), s);
    else rt.fixed_base.run(sc + ctx.d_lo(ctx.rank), ctx.d_count(), ctx.bases_d.get<G1Affine>(), s);
    // vk: [x]_2, [z]_2
    Fq* g2 = g2_b.as<Fq>(8);
    k_g2_mul<<<1, 32, 0, s>>>(c + C_X, 2, g2);
    PM_LAUNCH_CHECK();
    rt.extra_launches++;
    uint8_t host[8 * PM_FQ_BYTES];
    PM_CUDA(cudaMemcpyAsync(host, g2, sizeof host, cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaStreamSynchronize(s));
    memcpy(x_g2, host, 4 * PM_FQ_BYTES);
    memcpy(z_g2, host + 4 * PM_FQ_BYTES, 4 * PM_FQ_BYTES);
}

}  // namespace pm
