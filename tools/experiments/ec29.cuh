// G1 mixed addition on radix-2^29 lazy field elements (field29.cuh): the MSM hot loop.
// Same formulas as ec.cuh (EFD madd-2008-s / mdbl-2008-s-1); every subtraction names the bound of
// its subtrahend.  Invariant of an accumulator between additions:
//     x < 8p,  y < 4p,  zz < 2p,  zzz < 2p      (products are < 1.04p, written "< 2p" below)
// and zz == 0 exactly (all limbs) marks the point at infinity.
#pragma once
#include "../../polymath_b200/csrc/ec.cuh"
#include "field29.cuh"

namespace pm {

struct alignas(16) G1Affine29 {   // 112 bytes; (0,0) = infinity; coordinates canonical-range products (< 1.04p)
    Fq29 x, y;
    __device__ __forceinline__ bool is_inf() const { return x.is_zero_repr() && y.is_zero_repr(); }
};

struct G1XYZZ29 {
    Fq29 x, y, zz, zzz;
    __device__ __forceinline__ bool is_inf() const { return zz.is_zero_repr(); }
    __device__ __forceinline__ static G1XYZZ29 inf() { return {Fq29::zero(), Fq29::zero(), Fq29::zero(), Fq29::zero()}; }
};

struct Mul29Inline {
    static __device__ __forceinline__ Fq29 mul(const Fq29& a, const Fq29& b) { return mul29(a, b); }
};
struct Mul29Call {
    static __device__ __forceinline__ Fq29 mul(const Fq29& a, const Fq29& b) { return mul29_call(a, b); }
};

// 2 * (x, y) for an affine point (not infinity)
template <class M>
__device__ __noinline__ G1XYZZ29 xyzz29_dbl_affine(const Fq29& px, const Fq29& py) {
    G1XYZZ29 r;
    Fq29 u = add29(py, py);                          // < 4p
    Fq29 v = M::mul(u, u);                           // < 2p
    Fq29 w = M::mul(u, v);
    Fq29 s = M::mul(px, v);
    Fq29 xx = M::mul(px, px);
    Fq29 m = add29(add29(xx, xx), xx);               // < 6p
    Fq29 mm = M::mul(m, m);
    Fq29 s2 = add29(s, s);                           // < 4p
    r.x = sub29(mm, s2, SPREAD_4P);                  // < 6p  (<= 8p)
    Fq29 t = sub29(s, r.x, SPREAD_8P);               // < 10p
    Fq29 a = M::mul(m, t);
    Fq29 b = M::mul(w, py);
    r.y = sub29(a, b, SPREAD_2P);                    // < 4p
    r.zz = v;
    r.zzz = w;
    return r;
}

// acc += (px, py)  (`neg` adds the negated point); (px, py) affine with coordinates < 2p, limbs normalised
template <class M>
__device__ __forceinline__ void xyzz29_madd(G1XYZZ29& a, const G1Affine29& p_in, bool neg) {
    if (p_in.is_inf()) return;
    Fq29 py = p_in.y;
    if (neg) {
        Fq29 z = Fq29::zero();
        py = sub29(z, p_in.y, SPREAD_2P);            // 2p - y  (< 2p)
    }
    if (a.is_inf()) {
        a.x = p_in.x; a.y = py; a.zz = Fq29::one(); a.zzz = Fq29::one();
        return;
    }
    Fq29 u2 = M::mul(p_in.x, a.zz);
    Fq29 s2 = M::mul(py, a.zzz);
    Fq29 pd = sub29(u2, a.x, SPREAD_8P);             // x1 < 8p  -> < 10p
    Fq29 rd = sub29(s2, a.y, SPREAD_4P);             // y1 < 4p  -> < 6p
    Fq29 pp = M::mul(pd, pd);
    if (is_zero_mod_p_lt2p(pp)) {
        // same x coordinate: doubling if the y's agree, infinity otherwise
        Fq29 rr0 = M::mul(rd, rd);
        if (is_zero_mod_p_lt2p(rr0)) a = xyzz29_dbl_affine<M>(p_in.x, py);
        else a = G1XYZZ29::inf();
        return;
    }
    Fq29 ppp = M::mul(pd, pp);
    Fq29 q = M::mul(a.x, pp);
    Fq29 rr = M::mul(rd, rd);
    // x3 = rr - ppp - 2q  = rr + (2p - ppp) + (4p - 2q)        < 8p
    Fq29 x3;
#pragma unroll
    for (int i = 0; i < 14; i++) x3.v[i] = rr.v[i] + SPREAD_2P[i] - ppp.v[i] + SPREAD_4P[i] - 2u * q.v[i];
    x3.norm();
    Fq29 t = sub29(q, x3, SPREAD_8P);                // < 10p
    Fq29 y3a = M::mul(rd, t);
    Fq29 y3b = M::mul(a.y, ppp);
    a.y = sub29(y3a, y3b, SPREAD_2P);                // < 4p
    a.x = x3;
    a.zz = M::mul(a.zz, pp);
    a.zzz = M::mul(a.zzz, ppp);
}

// wire point -> radix-29 table entry
__device__ __forceinline__ G1Affine29 to_affine29(const G1Affine& p) {
    G1Affine29 r;
    if (p.is_inf()) { r.x = Fq29::zero(); r.y = Fq29::zero(); return r; }
    r.x = to_fq29(p.x);
    r.y = to_fq29(p.y);
    return r;
}
// accumulator -> wire XYZZ (canonical coordinates)
__device__ __forceinline__ G1XYZZ from_xyzz29(const G1XYZZ29& a) {
    if (a.is_inf()) return G1XYZZ::inf();
    G1XYZZ r;
    r.x = from_fq29(a.x);
    r.y = from_fq29(a.y);
    r.zz = from_fq29(a.zz);
    r.zzz = from_fq29(a.zzz);
    return r;
}

}  // namespace pm
