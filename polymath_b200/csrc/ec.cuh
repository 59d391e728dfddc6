// BLS12-381 G1 on the device: affine bases (96 B, Montgomery limbs, (0,0) = infinity) and
// XYZZ accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ == 0 is infinity).
//
// Replaces ark-ec's `Projective<g1::Config>` arithmetic used inside
// `VariableBaseMSM::msm_unchecked` (/root/reference/src/prover.rs:380-384) and `g * scalar`
// (/root/reference/src/generator.rs:169-177).  Only the affine image leaves the device, so
// the coordinate system is free; XYZZ gives the cheapest mixed addition (8M + 2S).
#pragma once
#include "field.cuh"

namespace pm {

struct alignas(16) G1Affine {
    Fq x, y;
    __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
    __device__ __forceinline__ static G1Affine inf() { return {Fq::zero(), Fq::zero()}; }
};

struct alignas(16) G1XYZZ {
    Fq x, y, zz, zzz;
    __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }
    __device__ __forceinline__ static G1XYZZ inf() { return {Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
    __device__ __forceinline__ static G1XYZZ from_affine(const G1Affine& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, Fq::one(), Fq::one()};
    }
};

// Multiplication policies for the hot loop: fully inlined products, or one shared out-of-line copy
// of the Fq product (keeps the loop body inside the instruction caches).
struct MulInline {
    static __device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return a * b; }
    static __device__ __forceinline__ Fq sqr(const Fq& a) { return a.sqr_wide(); }
};
static __device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) { return a * b; }
// dedicated squaring (222 instead of 288 wide IMADs, field.cuh): one more out-of-line copy
static __device__ __noinline__ Fq fq_sqr_call(Fq a) { return a.sqr_wide(); }
struct MulCall {
    static __device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
    static __device__ __forceinline__ Fq sqr(const Fq& a) { return fq_sqr_call(a); }
};

// The general-purpose group operations below (full addition, doubling) go through the ONE out-of-line product too:
// with every product inlined a single xyzz_add is ~100 KB of SASS, and the latency-bound kernels that call it from a
// few warps per SM (bucket reduction, slice sums) stalled on instruction fetch rather than on the multiplier.

// acc = 2 * p (p affine, not infinity)   EFD mdbl-2008-s-1
static __device__ __noinline__ G1XYZZ xyzz_dbl_affine(const G1Affine& p) {
    G1XYZZ r;
    Fq u = p.y.dbl();
    Fq v = fq_sqr_call(u);
    Fq w = fq_mul_call(u, v);
    Fq s = fq_mul_call(p.x, v);
    Fq xx = fq_sqr_call(p.x);
    Fq m = xx.dbl() + xx;
    r.x = fq_sqr_call(m) - s.dbl();
    r.y = fq_mul_call(m, s - r.x) - fq_mul_call(w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// acc = 2 * acc   EFD dbl-2008-s-1 (a = 0)
static __device__ __noinline__ void xyzz_dbl(G1XYZZ& a) {
    if (a.is_inf()) return;
    Fq u = a.y.dbl();
    Fq v = fq_sqr_call(u);
    Fq w = fq_mul_call(u, v);
    Fq s = fq_mul_call(a.x, v);
    Fq xx = fq_sqr_call(a.x);
    Fq m = xx.dbl() + xx;
    Fq x3 = fq_sqr_call(m) - s.dbl();
    a.y = fq_mul_call(m, s - x3) - fq_mul_call(w, a.y);
    a.x = x3;
    a.zz = fq_mul_call(v, a.zz);
    a.zzz = fq_mul_call(w, a.zzz);
}


// acc += p (p affine; `neg` adds -p)   EFD madd-2008-s, products through policy M
template <class M>
__device__ __forceinline__ void xyzz_madd_t(G1XYZZ& a, const G1Affine& p_in, bool neg) {
    if (p_in.is_inf()) return;
    Fq py = neg ? p_in.y.neg() : p_in.y;
    if (a.is_inf()) {
        a.x = p_in.x; a.y = py; a.zz = Fq::one(); a.zzz = Fq::one();
        return;
    }
    Fq u2 = M::mul(p_in.x, a.zz);
    Fq s2 = M::mul(py, a.zzz);
    Fq p = u2 - a.x;
    Fq r = s2 - a.y;
    if (p.is_zero()) {
        if (r.is_zero()) {
            G1Affine q{p_in.x, py};
            a = xyzz_dbl_affine(q);
        } else {
            a = G1XYZZ::inf();
        }
        return;
    }
    Fq pp = M::sqr(p);
    Fq ppp = M::mul(p, pp);
    Fq q = M::mul(a.x, pp);
    Fq x3 = M::sqr(r) - ppp - q.dbl();
    a.y = M::mul(r, q - x3) - M::mul(a.y, ppp);
    a.x = x3;
    a.zz = M::mul(a.zz, pp);
    a.zzz = M::mul(a.zzz, ppp);
}

// acc += p (p affine; `neg` adds -p)   EFD madd-2008-s
__device__ __forceinline__ void xyzz_madd(G1XYZZ& a, const G1Affine& p_in, bool neg = false) {
    if (p_in.is_inf()) return;
    Fq py = neg ? p_in.y.neg() : p_in.y;
    if (a.is_inf()) {
        a.x = p_in.x; a.y = py; a.zz = Fq::one(); a.zzz = Fq::one();
        return;
    }
    Fq u2 = p_in.x * a.zz;
    Fq s2 = py * a.zzz;
    Fq p = u2 - a.x;
    Fq r = s2 - a.y;
    if (p.is_zero()) {
        if (r.is_zero()) {
            G1Affine q{p_in.x, py};
            a = xyzz_dbl_affine(q);
        } else {
            a = G1XYZZ::inf();
        }
        return;
    }
    Fq pp = p.sqr();
    Fq ppp = p * pp;
    Fq q = a.x * pp;
    Fq x3 = r.sqr() - ppp - q.dbl();
    a.y = r * (q - x3) - a.y * ppp;
    a.x = x3;
    a.zz = a.zz * pp;
    a.zzz = a.zzz * ppp;
}

// acc += b   EFD add-2008-s
static __device__ __noinline__ void xyzz_add(G1XYZZ& a, const G1XYZZ& b) {
    if (b.is_inf()) return;
    if (a.is_inf()) { a = b; return; }
    Fq u1 = fq_mul_call(a.x, b.zz);
    Fq u2 = fq_mul_call(b.x, a.zz);
    Fq s1 = fq_mul_call(a.y, b.zzz);
    Fq s2 = fq_mul_call(b.y, a.zzz);
    Fq p = u2 - u1;
    Fq r = s2 - s1;
    if (p.is_zero()) {
        if (r.is_zero()) xyzz_dbl(a);
        else a = G1XYZZ::inf();
        return;
    }
    Fq pp = fq_sqr_call(p);
    Fq ppp = fq_mul_call(p, pp);
    Fq q = fq_mul_call(u1, pp);
    Fq x3 = fq_sqr_call(r) - ppp - q.dbl();
    a.y = fq_mul_call(r, q - x3) - fq_mul_call(s1, ppp);
    a.x = x3;
    a.zz = fq_mul_call(fq_mul_call(a.zz, b.zz), pp);
    a.zzz = fq_mul_call(fq_mul_call(a.zzz, b.zzz), ppp);
}

// canonical affine image: x = X/ZZ, y = Y/ZZZ with 1/ZZ = (ZZ/ZZZ)^2
static __device__ __noinline__ G1Affine xyzz_to_affine(const G1XYZZ& a) {
    if (a.is_inf()) return G1Affine::inf();
    Fq izzz = a.zzz.inv();
    Fq t = fq_mul_call(a.zz, izzz);
    Fq izz = fq_sqr_call(t);
    return {fq_mul_call(a.x, izz), fq_mul_call(a.y, izzz)};
}

// acc = k * acc for a small unsigned k (double-and-add, MSB first)
static __device__ __noinline__ G1XYZZ xyzz_mul_small(const G1XYZZ& p, uint32_t k) {
    G1XYZZ r = G1XYZZ::inf();
    for (int bit = 31 - __clz(k | 1); bit >= 0; bit--) {
        xyzz_dbl(r);
        if ((k >> bit) & 1) xyzz_add(r, p);
    }
    return r;
}

}  // namespace pm
