// Host-side BLS12-381 tower, G1/G2 arithmetic, point decoding and the pairing-product check for
// `Polymath::verify` (/root/reference/src/verifier.rs:19-62, which ends in
// `E::multi_pairing(..).0.is_one()`).  BASELINE.json's north_star keeps pairing-based verification on the
// host; this is that host side for the C++ mirror (the Rust integration keeps using ark-ec's pairing).
// Product code: independent of oracle/.
//
// Representation: Fq2 = Fq[u]/(u^2 + 1); Fq12 = Fq2[w]/(w^6 - xi), xi = 1 + u (six Fq2 coefficients).
// G2 is the M-type twist y^2 = x^3 + 4 xi; untwisting (x', y') -> (x'/w^2, y'/w^3) turns a twist line with
// slope lam through T into  l(P) * w^3 = (lam x_T - y_T) - lam x_P w^2 + y_P w^3  (w^3 lies in Fq4 and is
// killed by the final exponentiation).  The final exponentiation computes f^(3 (q^12 - 1)/r) with the
// (u - 1)^2 (u + q)(u^2 + q^2 - 1) + 3 chain; the cube does not change the "== 1" predicate (3 does not divide r).
#pragma once
#include "fp_host.hpp"

namespace pm { namespace host {

struct Fq2H {
    FqH c0, c1;
    static Fq2H zero() { return {FqH::zero(), FqH::zero()}; }
    static Fq2H one() { return {FqH::one(), FqH::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2H& b) const { return c0 == b.c0 && c1 == b.c1; }
    Fq2H operator+(const Fq2H& b) const { return {c0 + b.c0, c1 + b.c1}; }
    Fq2H operator-(const Fq2H& b) const { return {c0 - b.c0, c1 - b.c1}; }
    Fq2H neg() const { return {c0.neg(), c1.neg()}; }
    Fq2H conj() const { return {c0, c1.neg()}; }
    Fq2H operator*(const Fq2H& b) const {   // Karatsuba, u^2 = -1
        FqH aa = c0 * b.c0, bb = c1 * b.c1, s = (c0 + c1) * (b.c0 + b.c1);
        return {aa - bb, s - aa - bb};
    }
    Fq2H sqr() const {
        FqH a = (c0 + c1) * (c0 - c1), b = c0 * c1;
        return {a, b + b};
    }
    Fq2H scale(const FqH& k) const { return {c0 * k, c1 * k}; }
    Fq2H mul_xi() const { return {c0 - c1, c0 + c1}; }   // * (1 + u)
    Fq2H inv() const {
        FqH d = (c0.sqr() + c1.sqr()).inv();
        return {c0 * d, (c1 * d).neg()};
    }
    Fq2H pow(const uint64_t* e, int words) const {
        Fq2H acc = one();
        for (int w = words - 1; w >= 0; w--)
            for (int bit = 63; bit >= 0; bit--) { acc = acc.sqr(); if ((e[w] >> bit) & 1) acc = acc * *this; }
        return acc;
    }
    // zcash "lexicographically largest": c1 decides unless it is zero
    bool lex_largest() const { return c1.is_zero() ? c0.canonical_gt_half() : c1.canonical_gt_half(); }
};

// ---- small big-integer helpers on the modulus (exponents derived from q at first use) ----
inline void fq_exp_q_plus1_div4(uint64_t e[6]) {   // (q + 1) / 4: square roots in Fq (q = 3 mod 4)
    u128 c = 1;
    for (int i = 0; i < 6; i++) { c += FQ_PARAMS.mod[i]; e[i] = (uint64_t)c; c >>= 64; }
    for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i + 1 < 6 ? e[i + 1] << 62 : 0);
}
inline void fq_exp_q_minus1_div6(uint64_t e[6]) {  // (q - 1) / 6: Frobenius constant of w
    uint64_t t[6];
    memcpy(t, FQ_PARAMS.mod, sizeof t);
    t[0] -= 1;   // q is odd: no borrow
    u128 rem = 0;
    for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | t[i]; e[i] = (uint64_t)(cur / 6); rem = cur % 6; }
}

// square root in Fq; false when v is not a square
struct FqExp { uint64_t w[6]; };
inline bool fq_sqrt(const FqH& v, FqH& out) {
    static const FqExp e = [] { FqExp x; fq_exp_q_plus1_div4(x.w); return x; }();   // thread-safe one-time init
    FqH s = v.pow(e.w, 6);
    if (!(s.sqr() == v)) return false;
    out = s;
    return true;
}
// square root in Fq2 (complex method)
inline bool fq2_sqrt(const Fq2H& a, Fq2H& out) {
    if (a.is_zero()) { out = Fq2H::zero(); return true; }
    FqH s;
    if (a.c1.is_zero()) {
        if (fq_sqrt(a.c0, s)) { out = {s, FqH::zero()}; return true; }
        if (fq_sqrt(a.c0.neg(), s)) { out = {FqH::zero(), s}; return true; }   // (s u)^2 = -s^2
        return false;
    }
    FqH norm = a.c0.sqr() + a.c1.sqr(), n;
    if (!fq_sqrt(norm, n)) return false;
    FqH half = (FqH::one() + FqH::one()).inv();
    FqH delta = (a.c0 + n) * half, x0;
    if (!fq_sqrt(delta, x0)) {
        delta = (a.c0 - n) * half;
        if (!fq_sqrt(delta, x0)) return false;
    }
    FqH x1 = a.c1 * (x0 + x0).inv();
    Fq2H r{x0, x1};
    if (!(r.sqr() == a)) return false;
    out = r;
    return true;
}

// ---- short-Weierstrass (a = 0) arithmetic over F = FqH (G1) or Fq2H (G2) ----
template <class F>
struct AffH {
    F x, y;
    bool inf;
    static AffH infinity() { return {F::zero(), F::zero(), true}; }
    AffH neg() const { return {x, y.neg(), inf}; }
};
template <class F>
struct JacH {
    F x, y, z;
    bool is_inf() const { return z.is_zero(); }
    static JacH infinity() { return {F::one(), F::one(), F::zero()}; }
};
template <class F>
inline void jac_dbl(JacH<F>& p) {   // EFD dbl-2009-l
    if (p.is_inf()) return;
    F A = p.x.sqr(), B = p.y.sqr(), C = B.sqr();
    F t = (p.x + B).sqr() - A - C, D = t + t;
    F E = A + A + A, Fv = E.sqr();
    F x3 = Fv - D - D;
    F c8 = C + C; c8 = c8 + c8; c8 = c8 + c8;
    F y3 = E * (D - x3) - c8;
    F z3 = p.y * p.z;
    p = {x3, y3, z3 + z3};
}
template <class F>
inline void jac_add_affine(JacH<F>& p, const AffH<F>& q) {
    if (q.inf) return;
    if (p.is_inf()) { p = {q.x, q.y, F::one()}; return; }
    F z2 = p.z.sqr(), u2 = q.x * z2, s2 = q.y * p.z * z2;
    F h = u2 - p.x, r = s2 - p.y;
    if (h.is_zero()) {
        if (r.is_zero()) jac_dbl(p); else p = JacH<F>::infinity();
        return;
    }
    F hh = h.sqr(), hhh = h * hh, v = p.x * hh;
    F x3 = r.sqr() - hhh - v - v;
    F y3 = r * (v - x3) - p.y * hhh;
    p = {x3, y3, p.z * h};
}
template <class F>
inline AffH<F> jac_to_affine(const JacH<F>& p) {
    if (p.is_inf()) return AffH<F>::infinity();
    F zi = p.z.inv(), zi2 = zi.sqr();
    return {p.x * zi2, p.y * zi2 * zi, false};
}
// k * P, k = canonical little-endian limbs
template <class F>
inline JacH<F> scalar_mul(const AffH<F>& p, const uint64_t* k, int words) {
    JacH<F> acc = JacH<F>::infinity();
    for (int w = words - 1; w >= 0; w--)
        for (int bit = 63; bit >= 0; bit--) {
            jac_dbl(acc);
            if ((k[w] >> bit) & 1) jac_add_affine(acc, p);
        }
    return acc;
}
template <class F>
inline JacH<F> scalar_mul(const AffH<F>& p, const FrH& k) {
    FrH c = k.from_mont();
    return scalar_mul(p, c.v, 4);
}
template <class F>
inline AffH<F> aff_add(const AffH<F>& a, const AffH<F>& b) {
    if (a.inf) return b;
    JacH<F> j{a.x, a.y, F::one()};
    jac_add_affine(j, b);
    return jac_to_affine(j);
}
// prime-order subgroup membership: [r]P = O
template <class F>
inline bool in_subgroup(const AffH<F>& p) {
    if (p.inf) return true;
    return scalar_mul(p, FR_PARAMS.mod, 4).is_inf();
}

using G1H = AffH<FqH>;
using G2H = AffH<Fq2H>;

inline FqH fq_from_u64(uint64_t v) { return FqH::from_u64(v); }
inline bool g1_on_curve(const G1H& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + fq_from_u64(4); }
inline Fq2H g2_b() { FqH four = fq_from_u64(4); return {four, four}; }   // 4 (1 + u)
inline bool g2_on_curve(const G2H& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + g2_b(); }

// ---- zcash / ark-bls12-381 compressed encodings -> affine (validated like `deserialize_compressed`) ----
// 48 big-endian bytes -> canonical Fq (top three bits already masked by the caller); false when >= q
inline bool fq_from_be(const uint8_t* b, uint8_t mask_top, FqH& out) {
    uint8_t le[48];
    for (int i = 0; i < 48; i++) le[i] = b[47 - i];
    le[47] &= mask_top;
    uint64_t limbs[6];
    memcpy(limbs, le, 48);
    if (FqH::geq_mod(limbs)) return false;
    out = FqH::from_canonical_le(le);
    return true;
}
inline bool g1_decompress(const uint8_t b[48], G1H& out, bool check_subgroup = true) {
    if (!(b[0] & 0x80)) return false;                     // compressed form only
    if (b[0] & 0x40) {                                    // infinity: everything else zero
        if (b[0] & 0x3f) return false;
        for (int i = 1; i < 48; i++) if (b[i]) return false;
        out = G1H::infinity();
        return true;
    }
    FqH x, y;
    if (!fq_from_be(b, 0x1f, x)) return false;
    if (!fq_sqrt(x.sqr() * x + fq_from_u64(4), y)) return false;
    if (y.canonical_gt_half() != ((b[0] & 0x20) != 0)) y = y.neg();
    out = {x, y, false};
    return !check_subgroup || in_subgroup(out);
}
inline bool g2_decompress(const uint8_t b[96], G2H& out, bool check_subgroup = true) {
    if (!(b[0] & 0x80)) return false;
    if (b[0] & 0x40) {
        if (b[0] & 0x3f) return false;
        for (int i = 1; i < 96; i++) if (b[i]) return false;
        out = G2H::infinity();
        return true;
    }
    Fq2H x, y;
    if (!fq_from_be(b, 0x1f, x.c1)) return false;         // c1 first, then c0
    if (!fq_from_be(b + 48, 0xff, x.c0)) return false;
    if (!fq2_sqrt(x.sqr() * x + g2_b(), y)) return false;
    if (y.lex_largest() != ((b[0] & 0x20) != 0)) y = y.neg();
    out = {x, y, false};
    return !check_subgroup || in_subgroup(out);
}
// 32 little-endian bytes -> Fr; false when >= r (ark-serialize rejects non-canonical scalars)
inline bool fr_from_canonical(const uint8_t b[32], FrH& out) {
    uint64_t limbs[4];
    memcpy(limbs, b, 32);
    if (FrH::geq_mod(limbs)) return false;
    out = FrH::from_canonical_le(b);
    return true;
}

// ---- Fq12 = Fq2[w]/(w^6 - xi) ----
struct Fq12H {
    Fq2H c[6];
    static Fq12H one() { Fq12H r; for (int i = 0; i < 6; i++) r.c[i] = Fq2H::zero(); r.c[0] = Fq2H::one(); return r; }
    bool is_one() const {
        if (!(c[0] == Fq2H::one())) return false;
        for (int i = 1; i < 6; i++) if (!c[i].is_zero()) return false;
        return true;
    }
    bool operator==(const Fq12H& b) const { for (int i = 0; i < 6; i++) if (!(c[i] == b.c[i])) return false; return true; }
    Fq12H operator*(const Fq12H& b) const {
        Fq2H t[11];
        for (int k = 0; k < 11; k++) t[k] = Fq2H::zero();
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) t[i + j] = t[i + j] + c[i] * b.c[j];
        Fq12H r;
        for (int k = 0; k < 6; k++) r.c[k] = t[k];
        for (int k = 6; k < 11; k++) r.c[k - 6] = r.c[k - 6] + t[k].mul_xi();
        return r;
    }
    Fq12H sqr() const {   // symmetric schoolbook: 15 cross products (doubled) + 6 squares instead of 36 products
        Fq2H t[11];
        for (int k = 0; k < 11; k++) t[k] = Fq2H::zero();
        for (int i = 0; i < 6; i++) {
            t[2 * i] = t[2 * i] + c[i].sqr();
            for (int j = i + 1; j < 6; j++) {
                Fq2H p = c[i] * c[j];
                t[i + j] = t[i + j] + p + p;
            }
        }
        Fq12H r;
        for (int k = 0; k < 6; k++) r.c[k] = t[k];
        for (int k = 6; k < 11; k++) r.c[k - 6] = r.c[k - 6] + t[k].mul_xi();
        return r;
    }
    // * (l0 + l2 w^2 + l3 w^3), l3 in Fq: the Miller-loop line
    Fq12H mul_line(const Fq2H& l0, const Fq2H& l2, const FqH& l3) const {
        Fq2H t[9];
        for (int k = 0; k < 9; k++) t[k] = Fq2H::zero();
        for (int i = 0; i < 6; i++) {
            t[i] = t[i] + c[i] * l0;
            t[i + 2] = t[i + 2] + c[i] * l2;
            t[i + 3] = t[i + 3] + c[i].scale(l3);
        }
        Fq12H r;
        for (int k = 0; k < 6; k++) r.c[k] = t[k];
        for (int k = 6; k < 9; k++) r.c[k - 6] = r.c[k - 6] + t[k].mul_xi();
        return r;
    }
    Fq12H conj6() const {   // the q^6 Frobenius: w -> -w
        Fq12H r = *this;
        for (int i = 1; i < 6; i += 2) r.c[i] = r.c[i].neg();
        return r;
    }
    struct Gammas { Fq2H g[6]; };
    static const Fq2H* frob_gamma() {   // gamma_i = xi^(i (q - 1)/6): w^(i q) = gamma_i w^i
        static const Gammas t = [] {       // thread-safe one-time init
            Gammas x;
            uint64_t e[6];
            fq_exp_q_minus1_div6(e);
            Fq2H xi{FqH::one(), FqH::one()};
            x.g[0] = Fq2H::one();
            x.g[1] = xi.pow(e, 6);
            for (int i = 2; i < 6; i++) x.g[i] = x.g[i - 1] * x.g[1];
            return x;
        }();
        return t.g;
    }
    Fq12H frobenius() const {
        const Fq2H* g = frob_gamma();
        Fq12H r;
        for (int i = 0; i < 6; i++) r.c[i] = c[i].conj() * g[i];
        return r;
    }
    // a^-1 = conj6(a) / (a conj6(a)); the norm a conj6(a) lies in Fq6 = Fq2[v]/(v^3 - xi), v = w^2
    Fq12H inv() const {
        Fq12H cj = conj6();
        Fq12H nrm = *this * cj;
        const Fq2H a0 = nrm.c[0], a1 = nrm.c[2], a2 = nrm.c[4];
        Fq2H t0 = a0.sqr() - (a1 * a2).mul_xi();
        Fq2H t1 = a2.sqr().mul_xi() - a0 * a1;
        Fq2H t2 = a1.sqr() - a0 * a2;
        Fq2H d = (a0 * t0 + (a2 * t1 + a1 * t2).mul_xi()).inv();
        Fq12H ninv;
        for (int i = 0; i < 6; i++) ninv.c[i] = Fq2H::zero();
        ninv.c[0] = t0 * d;
        ninv.c[2] = t1 * d;
        ninv.c[4] = t2 * d;
        return cj * ninv;
    }
    Fq12H pow(const uint64_t* e, int words) const {
        Fq12H acc = one();
        for (int w = words - 1; w >= 0; w--)
            for (int bit = 63; bit >= 0; bit--) { acc = acc.sqr(); if ((e[w] >> bit) & 1) acc = acc * *this; }
        return acc;
    }
};

constexpr uint64_t kBlsXAbs = 0xd201000000010000ull;   // |u|; the BLS12-381 parameter u is negative

struct PairingTerm { G1H p; G2H q; };

// prod_i f_{|u|, Q_i}(P_i), conjugated because u < 0 (up to factors the final exponentiation removes): one shared
// accumulator, so the 63 squarings are paid once for all pairs.  Pairs with a point at infinity contribute 1.
inline Fq12H multi_miller_loop(const PairingTerm* terms, int count) {
    struct State { const PairingTerm* t; Fq2H xt, yt; };
    State st[8];
    int m = 0;
    Fq12H f = Fq12H::one(), rest = Fq12H::one();
    for (int i = 0; i < count; i++) {
        if (terms[i].p.inf || terms[i].q.inf) continue;
        if (m == 8) {   // more than eight live pairs: the remaining ones in a loop of their own
            rest = multi_miller_loop(terms + i, count - i);
            break;
        }
        st[m++] = {&terms[i], terms[i].q.x, terms[i].q.y};
    }
    if (m == 0) return f;
    for (int bit = 62; bit >= 0; bit--) {     // bit 63 is the leading one
        f = f.sqr();
        for (int k = 0; k < m; k++) {         // tangent at T
            State& s = st[k];
            Fq2H xx = s.xt.sqr();
            Fq2H lam = (xx + xx + xx) * (s.yt + s.yt).inv();
            f = f.mul_line(lam * s.xt - s.yt, lam.scale(s.t->p.x).neg(), s.t->p.y);
            Fq2H x3 = lam.sqr() - s.xt - s.xt;
            s.yt = lam * (s.xt - x3) - s.yt;
            s.xt = x3;
        }
        if ((kBlsXAbs >> bit) & 1) {
            for (int k = 0; k < m; k++) {     // chord through T and Q
                State& s = st[k];
                const G2H& q = s.t->q;
                Fq2H lam2 = (q.y - s.yt) * (q.x - s.xt).inv();
                f = f.mul_line(lam2 * s.xt - s.yt, lam2.scale(s.t->p.x).neg(), s.t->p.y);
                Fq2H x4 = lam2.sqr() - s.xt - q.x;
                s.yt = lam2 * (s.xt - x4) - s.yt;
                s.xt = x4;
            }
        }
    }
    return f.conj6() * rest;
}
inline Fq12H miller_loop(const G1H& p, const G2H& q) {
    PairingTerm t{p, q};
    return multi_miller_loop(&t, 1);
}

// g^u for g in the cyclotomic subgroup (inverse = conj6)
inline Fq12H cyclo_exp_u(const Fq12H& g) {
    uint64_t e = kBlsXAbs;
    return g.pow(&e, 1).conj6();
}

// f^(3 (q^12 - 1)/r)
inline Fq12H final_exponentiation_cubed(const Fq12H& f) {
    Fq12H f1 = f.conj6() * f.inv();                         // f^(q^6 - 1)
    Fq12H f2 = f1.frobenius().frobenius() * f1;             // ^(q^2 + 1): now in the cyclotomic subgroup
    Fq12H y0 = cyclo_exp_u(f2) * f2.conj6();                // f2^(u - 1)
    Fq12H y1 = cyclo_exp_u(y0) * y0.conj6();                // ^(u - 1)
    Fq12H y2 = cyclo_exp_u(y1) * y1.frobenius();            // ^(u + q)
    Fq12H y3 = cyclo_exp_u(cyclo_exp_u(y2)) * y2.frobenius().frobenius() * y2.conj6();   // ^(u^2 + q^2 - 1)
    return y3 * f2.sqr() * f2;                              // * f2^3
}

// prod_i e(P_i, Q_i) == 1 ?
inline bool pairing_product_is_one(const PairingTerm* terms, int count) {
    return final_exponentiation_cubed(multi_miller_loop(terms, count)).is_one();
}

}}  // namespace pm::host
