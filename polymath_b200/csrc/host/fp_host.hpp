// Host-side prime-field arithmetic for the protocol glue that stays on the CPU
// (Fiat-Shamir scalar algebra of /root/reference/src/common.rs:40-97 and point compression).
// 64-bit-limb Montgomery form, byte-identical to the device / arkworks in-memory form.
// Product code: independent of oracle/.
#pragma once
#include <cstdint>
#include <cstring>

namespace pm { namespace host {

typedef unsigned __int128 u128;

template <int N>
struct FpParams {
    uint64_t mod[N];
    uint64_t r2[N];    // R^2 mod p
    uint64_t one[N];   // R mod p
    uint64_t inv;      // -p^-1 mod 2^64
};

template <int N, const FpParams<N>& P>
struct FpH {
    uint64_t v[N];

    static FpH zero() { FpH r; memset(r.v, 0, sizeof r.v); return r; }
    static FpH one() { FpH r; memcpy(r.v, P.one, sizeof r.v); return r; }
    static FpH from_wire(const uint8_t* b) { FpH r; memcpy(r.v, b, sizeof r.v); return r; }   // Montgomery LE limbs
    void to_wire(uint8_t* b) const { memcpy(b, v, sizeof v); }
    static FpH from_u64(uint64_t x) { FpH r = zero(); r.v[0] = x; return r.to_mont(); }
    // canonical integer as little-endian bytes (ark-serialize Fp) -> field element; caller guarantees < p
    static FpH from_canonical_le(const uint8_t* b) { FpH r; memcpy(r.v, b, sizeof r.v); return r.to_mont(); }
    void to_canonical_le(uint8_t* b) const { FpH c = from_mont(); memcpy(b, c.v, sizeof c.v); }

    bool is_zero() const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= v[i]; return o == 0; }
    bool operator==(const FpH& b) const { return memcmp(v, b.v, sizeof v) == 0; }

    static bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) { if (a[i] != P.mod[i]) return a[i] > P.mod[i]; }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        u128 borrow = 0;
        for (int i = 0; i < N; i++) { u128 t = (u128)a[i] - P.mod[i] - (uint64_t)borrow; a[i] = (uint64_t)t; borrow = (t >> 64) & 1; }
    }
    FpH operator+(const FpH& b) const {
        FpH r; u128 c = 0;
        for (int i = 0; i < N; i++) { c += (u128)v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    FpH operator-(const FpH& b) const {
        FpH r; u128 borrow = 0;
        for (int i = 0; i < N; i++) { u128 t = (u128)v[i] - b.v[i] - (uint64_t)borrow; r.v[i] = (uint64_t)t; borrow = (t >> 64) & 1; }
        if (borrow) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)r.v[i] + P.mod[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    FpH neg() const { return zero() - *this; }
    // CIOS Montgomery product
    FpH operator*(const FpH& b) const {
        uint64_t t[N + 2];
        memset(t, 0, sizeof t);
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) {
            u128 c = 0;
#pragma GCC unroll 8
            for (int j = 0; j < N; j++) { c += (u128)v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P.inv;
            c = (u128)m * P.mod[0] + t[0]; c >>= 64;
#pragma GCC unroll 8
            for (int j = 1; j < N; j++) { c += (u128)m * P.mod[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        FpH r; memcpy(r.v, t, sizeof r.v);
        if (t[N] || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    FpH sqr() const { return *this * *this; }
    FpH to_mont() const { FpH r2; memcpy(r2.v, P.r2, sizeof r2.v); return *this * r2; }
    FpH from_mont() const { FpH o = zero(); o.v[0] = 1; return *this * o; }
    FpH pow(const uint64_t* e, int words) const {
        FpH acc = one();
        for (int w = words - 1; w >= 0; w--)
            for (int bit = 63; bit >= 0; bit--) { acc = acc.sqr(); if ((e[w] >> bit) & 1) acc = acc * *this; }
        return acc;
    }
    FpH pow_u64(uint64_t e) const { return pow(&e, 1); }
    // binary extended Euclid on the raw limbs (aR)^-1 = a^-1 R^-1, then one Montgomery product by R^3;
    // inverse of zero is zero (callers check).  ~10x cheaper than the Fermat chain: the Miller loop of the host
    // verifier inverts once per step.
    FpH inv() const {
        if (is_zero()) return zero();
        uint64_t u[N], w[N], x1[N], x2[N];
        memcpy(u, v, sizeof u);
        memcpy(w, P.mod, sizeof w);
        memset(x1, 0, sizeof x1); x1[0] = 1;
        memset(x2, 0, sizeof x2);
        auto is_one = [](const uint64_t* a) { uint64_t o = a[0] ^ 1; for (int i = 1; i < N; i++) o |= a[i]; return o == 0; };
        auto shr1 = [](uint64_t* a, uint64_t top) { for (int i = 0; i < N; i++) a[i] = (a[i] >> 1) | ((i + 1 < N ? a[i + 1] : top) << 63); };
        auto halve = [&](uint64_t* a) {          // a / 2 mod p
            uint64_t top = 0;
            if (a[0] & 1) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)a[i] + P.mod[i]; a[i] = (uint64_t)c; c >>= 64; } top = (uint64_t)c; }
            shr1(a, top);
        };
        auto geq = [](const uint64_t* a, const uint64_t* b) { for (int i = N - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i]; return true; };
        auto sub = [](uint64_t* a, const uint64_t* b) { u128 br = 0; for (int i = 0; i < N; i++) { u128 t = (u128)a[i] - b[i] - (uint64_t)br; a[i] = (uint64_t)t; br = (t >> 64) & 1; } return (uint64_t)br; };
        auto sub_modp = [&](uint64_t* a, const uint64_t* b) {   // a - b mod p, a, b < p
            if (sub(a, b)) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)a[i] + P.mod[i]; a[i] = (uint64_t)c; c >>= 64; } }
        };
        while (!is_one(u) && !is_one(w)) {
            while (!(u[0] & 1)) { shr1(u, 0); halve(x1); }
            while (!(w[0] & 1)) { shr1(w, 0); halve(x2); }
            if (geq(u, w)) { sub(u, w); sub_modp(x1, x2); } else { sub(w, u); sub_modp(x2, x1); }
        }
        FpH r, r2;
        memcpy(r.v, is_one(u) ? x1 : x2, sizeof r.v);
        memcpy(r2.v, P.r2, sizeof r2.v);
        return r * (r2 * r2);
    }
    // canonical value > (p-1)/2 ?   (zcash "lexicographically largest" flag)
    bool canonical_gt_half() const {
        FpH c = from_mont();
        // compare 2c with p: c > (p-1)/2  <=>  2c > p - 1  <=>  2c >= p  (p odd => 2c != p)
        uint64_t d[N + 1]; uint64_t carry = 0;
        for (int i = 0; i < N; i++) { d[i] = (c.v[i] << 1) | carry; carry = c.v[i] >> 63; }
        if (carry) return true;
        return geq_mod(d);
    }
};

inline constexpr FpParams<4> FR_PARAMS = {
    {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
    {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull},
    {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
    0xfffffffeffffffffull};
inline constexpr FpParams<6> FQ_PARAMS = {
    {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
    {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull},
    {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
    0x89f3fffcfffcfffdull};

using FrH = FpH<4, FR_PARAMS>;
using FqH = FpH<6, FQ_PARAMS>;

// 2^32-th root of unity of Fr (Montgomery limbs), arkworks TWO_ADIC_ROOT_OF_UNITY
inline FrH fr_two_adic_root() {
    FrH r;
    const uint64_t l[4] = {0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull};
    memcpy(r.v, l, sizeof l);
    return r;
}
// Radix2EvaluationDomain::group_gen for a domain of size 2^log_n
inline FrH fr_group_gen(int log_n) {
    FrH w = fr_two_adic_root();
    for (int i = log_n; i < 32; i++) w = w.sqr();
    return w;
}

}}  // namespace pm::host
