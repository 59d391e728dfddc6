// Fq in radix 2^29 (14 limbs) for the MSM bucket accumulation.
//
// Measured on B200: the carry-propagating IMAD.WIDE.U32.X issues at half the rate of the plain
// IMAD.WIDE.U32 (profiles/microbench_r1_int32_pipe.txt).  With 29-bit limbs every partial product is
// < 2^58, a whole Montgomery column (14 + 14 products + carry) fits one 64-bit accumulator, and the
// product-scanning multiplication below needs NO carry flags: 392 plain IMAD.WIDE + 14 IMAD instead of
// 276 carry-form ones (552 plain-equivalents) for the 12 x 32-bit form.
//
// Representation: value = sum v[i] * 2^(29 i), Montgomery form with R' = 2^406, LAZY: values live in
// [0, k*p) for a small k tracked per call site; limbs 0..12 are kept < 2^29 ("normalised"), the top
// limb holds the rest.  A product of inputs bounded by a*p and b*p with a*b <= 2^20 is < 1.04 p.
// Subtractions add a multiple K*p whose limbs are pre-spread (each >= 2^30 - 2) so that no limb goes
// negative; the K used at each call site is >= the bound of the subtrahend.
#pragma once
#include "../../polymath_b200/csrc/field.cuh"

namespace pm {

constexpr uint32_t M29 = (1u << 29) - 1u;
static __device__ __constant__ uint32_t P29[14] = {0x1fffaaabu, 0x0ff7ffffu, 0x14ffffeeu, 0x17fffd62u, 0x0f6241eau, 0x09507b58u, 0x0afd9cc3u,
                                                  0x109e70a2u, 0x1764774bu, 0x121a5d66u, 0x12c6e9edu, 0x12ffcd34u, 0x00111ea3u, 0x0000000du};
constexpr uint32_t INV29 = 0x1ffcfffdu;   // -p^-1 mod 2^29
// K*p with limbs spread (limb i borrows 2*2^29 from limb i+1): every lower limb is in [2^30 - 2, 2^30 + 2^29)
static __device__ __constant__ uint32_t SPREAD_1P[14] = {0x5fffaaabu, 0x4ff7fffdu, 0x54ffffecu, 0x57fffd60u, 0x4f6241e8u, 0x49507b56u, 0x4afd9cc1u,
                                                        0x509e70a0u, 0x57647749u, 0x521a5d64u, 0x52c6e9ebu, 0x52ffcd32u, 0x40111ea1u, 0x0000000bu};
static __device__ __constant__ uint32_t SPREAD_2P[14] = {0x5fff5556u, 0x5feffffdu, 0x49ffffdau, 0x4ffffac3u, 0x5ec483d3u, 0x52a0f6aeu, 0x55fb3984u,
                                                        0x413ce142u, 0x4ec8ee95u, 0x4434bacbu, 0x458dd3d9u, 0x45ff9a67u, 0x40223d45u, 0x00000018u};
static __device__ __constant__ uint32_t SPREAD_4P[14] = {0x5ffeaaacu, 0x5fdffffdu, 0x53ffffb7u, 0x5ffff588u, 0x5d8907a8u, 0x4541ed5fu, 0x4bf6730bu,
                                                        0x4279c287u, 0x5d91dd2cu, 0x48697598u, 0x4b1ba7b4u, 0x4bff34d0u, 0x40447a8cu, 0x00000032u};
static __device__ __constant__ uint32_t SPREAD_8P[14] = {0x5ffd5558u, 0x5fbffffdu, 0x47ffff71u, 0x5fffeb13u, 0x5b120f53u, 0x4a83dac1u, 0x57ece618u,
                                                        0x44f38510u, 0x5b23ba5au, 0x50d2eb33u, 0x56374f6au, 0x57fe69a2u, 0x4088f51au, 0x00000066u};
// R'^2 / R mod p: mul29(repack(x*R), TO29) = x*R'      (R = 2^384, the wire Montgomery radix)
static __device__ __constant__ uint32_t TO29[14] = {0x1fddebbdu, 0x1a4f5474u, 0x0291f399u, 0x14d03b3cu, 0x0f6cad2cu, 0x1b4cabcau, 0x1592827cu,
                                                   0x021c6ac7u, 0x1ec52a84u, 0x16fd5ec4u, 0x0c960da6u, 0x0fd2af6bu, 0x13263591u, 0x0000000bu};
// R mod p as a plain integer: mul29(x*R', RMODP29) = x*R (lazy)
static __device__ __constant__ uint32_t RMODP29[14] = {0x0002fffdu, 0x10480000u, 0x0300009du, 0x08001788u, 0x158baebfu, 0x0c2ba9e3u, 0x1d157d22u,
                                                      0x0a6e0a4au, 0x0d77ce58u, 0x1d12b763u, 0x1701c6a5u, 0x1501c926u, 0x1f65ec3fu, 0x0000000au};
static __device__ __constant__ uint32_t ONE29[14] = {0x03a9fb84u, 0x0ba00690u, 0x071288f1u, 0x0f59bcc5u, 0x126cb614u, 0x0585bf36u, 0x1b85ac3du,
                                                    0x1cf856fau, 0x1891ecbdu, 0x1a7eec05u, 0x155a88f0u, 0x0741ac6du, 0x1317c30fu, 0x00000009u};

struct alignas(8) Fq29 {
    uint32_t v[14];

    __device__ __forceinline__ static Fq29 zero() {
        Fq29 r;
#pragma unroll
        for (int i = 0; i < 14; i++) r.v[i] = 0;
        return r;
    }
    __device__ __forceinline__ static Fq29 one() {
        Fq29 r;
#pragma unroll
        for (int i = 0; i < 14; i++) r.v[i] = ONE29[i];
        return r;
    }
    // all limbs zero (exact zero representation; lazy multiples of p are NOT detected here)
    __device__ __forceinline__ bool is_zero_repr() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 14; i++) o |= v[i];
        return o == 0;
    }
    // carry propagation: limbs 0..12 < 2^29 afterwards
    __device__ __forceinline__ void norm() {
#pragma unroll
        for (int i = 0; i < 13; i++) {
            v[i + 1] += v[i] >> 29;
            v[i] &= M29;
        }
    }
};

// a + b (bounds add)
__device__ __forceinline__ Fq29 add29(const Fq29& a, const Fq29& b) {
    Fq29 r;
#pragma unroll
    for (int i = 0; i < 14; i++) r.v[i] = a.v[i] + b.v[i];
    r.norm();
    return r;
}
// a - b + K*p, where SPREAD = spread limbs of K*p and b < K*p, b normalised
__device__ __forceinline__ Fq29 sub29(const Fq29& a, const Fq29& b, const uint32_t* spread) {
    Fq29 r;
#pragma unroll
    for (int i = 0; i < 14; i++) r.v[i] = a.v[i] + spread[i] - b.v[i];
    r.norm();
    return r;
}

// Montgomery product a*b/R' mod p (lazy, < 1.04 p); inputs normalised, bounds a_k * b_k <= 2^20
__device__ __forceinline__ Fq29 mul29(const Fq29& a, const Fq29& b) {
    uint64_t t = 0;
    uint32_t m[14];
    Fq29 r;
#pragma unroll
    for (int k = 0; k < 14; k++) {
#pragma unroll
        for (int i = 0; i <= k; i++) t += (uint64_t)a.v[i] * b.v[k - i];
#pragma unroll
        for (int i = 0; i < k; i++) t += (uint64_t)m[i] * P29[k - i];
        m[k] = ((uint32_t)t * INV29) & M29;
        t += (uint64_t)m[k] * P29[0];
        t >>= 29;
    }
#pragma unroll
    for (int k = 14; k < 27; k++) {
#pragma unroll
        for (int i = k - 13; i < 14; i++) t += (uint64_t)a.v[i] * b.v[k - i];
#pragma unroll
        for (int i = k - 13; i < 14; i++) t += (uint64_t)m[i] * P29[k - i];
        r.v[k - 14] = (uint32_t)t & M29;
        t >>= 29;
    }
    r.v[13] = (uint32_t)t;
    return r;
}
static __device__ __noinline__ Fq29 mul29_call(Fq29 a, Fq29 b) { return mul29(a, b); }

// value == 0 mod p for a value known to be < 2p (a product): it is 0 or p
__device__ __forceinline__ bool is_zero_mod_p_lt2p(const Fq29& a) {
    uint32_t z = 0, e = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) { z |= a.v[i]; e |= a.v[i] ^ P29[i]; }
    return z == 0 || e == 0;
}

// ---- conversions with the wire form (12 x 32-bit limbs, R = 2^384, canonical) ----------------
__device__ __forceinline__ Fq29 repack_32_to_29(const Fq& x) {
    Fq29 r;
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const int bit = 29 * i, w = bit >> 5, off = bit & 31;
        uint64_t lo = (w < 12) ? x.v[w] : 0u;
        uint64_t hi = (w + 1 < 12) ? x.v[w + 1] : 0u;
        uint64_t two = lo | (hi << 32);
        r.v[i] = (uint32_t)(two >> off) & (i < 13 ? M29 : 0xffffffffu);
    }
    return r;
}
// limbs normalised and value < 2^384
__device__ __forceinline__ Fq repack_29_to_32(const Fq29& a) {
    Fq r;
#pragma unroll
    for (int w = 0; w < 12; w++) {
        // bits [32w, 32w+32): gather from the 29-bit limbs that overlap
        uint64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 14; i++) {
            const int lo_bit = 29 * i - 32 * w;   // position of limb i relative to word w
            if (lo_bit > -29 - 3 && lo_bit < 32) {
                if (lo_bit >= 0) acc |= (uint64_t)a.v[i] << lo_bit;
                else acc |= (uint64_t)a.v[i] >> (-lo_bit);
            }
        }
        r.v[w] = (uint32_t)acc;
    }
    return r;
}
// wire Fq (x*R, canonical) -> Fq29 (x*R', < 1.04p)
__device__ __forceinline__ Fq29 to_fq29(const Fq& x) {
    Fq29 c;
#pragma unroll
    for (int i = 0; i < 14; i++) c.v[i] = TO29[i];
    return mul29_call(repack_32_to_29(x), c);
}
// Fq29 (x*R', bound <= 2^10 p) -> wire Fq (x*R, canonical)
__device__ __forceinline__ Fq from_fq29(const Fq29& a) {
    Fq29 c;
#pragma unroll
    for (int i = 0; i < 14; i++) c.v[i] = RMODP29[i];
    Fq29 t = mul29_call(a, c);      // x*R, < 1.04p: subtract p at most once
    Fq29 d;
    // d = t - p with signed borrow propagation
    int64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
        int64_t s = (int64_t)t.v[i] - (int64_t)P29[i] + borrow;
        if (i < 13) { d.v[i] = (uint32_t)(s & M29); borrow = s >> 29; }
        else { d.v[i] = (uint32_t)s; borrow = s >> 32; }
    }
    const bool neg = borrow < 0;
    Fq29 sel;
#pragma unroll
    for (int i = 0; i < 14; i++) sel.v[i] = neg ? t.v[i] : d.v[i];
    return repack_29_to_32(sel);
}

}  // namespace pm
