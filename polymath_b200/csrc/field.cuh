// K1 — BLS12-381 Fr (8 x u32) and Fq (12 x u32) Montgomery arithmetic for sm_100a.
//
// Replaces every `F` / `Fq` operation that the reference performs inside arkworks
// (ark-ff `Fp<MontBackend<_,4|6>>`; call sites e.g. /root/reference/src/prover.rs:98-108,
// 260-277, generator.rs:72-77).  In-memory form is identical to arkworks': little-endian
// limbs of a*R mod p with R = 2^256 (Fr) / 2^384 (Fq), so host buffers need no conversion.
//
// Multiplication is a word-serial Montgomery product whose partial products are split
// into an "even" and an "odd" accumulator so that every 32x32->64 product lands on an
// aligned register pair: each mad.lo.cc/madc.hi.cc pair becomes one IMAD.WIDE.U32(.X)
// in SASS and the carry chains stay inside the INT32 IMAD pipe.
#pragma once
#include <cstdint>

namespace pm {

// ---------------------------------------------------------------------------------------
// one PTX instruction per wrapper; `volatile` keeps the implicit carry-flag order
// ---------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint64_t mul_wide(uint32_t a, uint32_t b) { uint64_t r; asm volatile("mul.wide.u32 %0,%1,%2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm volatile("mad.wide.u32 %0,%1,%2,%3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.cc.u64 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t addc_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.cc.u64 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t addc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.u64 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
}  // namespace ptx

// ---------------------------------------------------------------------------------------
// field parameters (constant bank: usable directly as IMAD operands)
// ---------------------------------------------------------------------------------------
// Fr: r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
static __device__ __constant__ uint32_t FR_MOD[8] = {
    0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
static __device__ __constant__ uint32_t FR_ONE[8] = {  // R mod r
    0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
static __device__ __constant__ uint32_t FR_R2[8] = {   // R^2 mod r
    0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
// Fq: q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
static __device__ __constant__ uint32_t FQ_MOD[12] = {
    0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
static __device__ __constant__ uint32_t FQ_ONE[12] = {  // R mod q
    0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
static __device__ __constant__ uint32_t FQ_R2[12] = {   // R^2 mod q
    0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};

struct FrP {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
    __device__ __forceinline__ static const uint32_t* mod() { return FR_MOD; }
    __device__ __forceinline__ static const uint32_t* one() { return FR_ONE; }
    __device__ __forceinline__ static const uint32_t* r2() { return FR_R2; }
};
struct FqP {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -q^-1 mod 2^32
    __device__ __forceinline__ static const uint32_t* mod() { return FQ_MOD; }
    __device__ __forceinline__ static const uint32_t* one() { return FQ_ONE; }
    __device__ __forceinline__ static const uint32_t* r2() { return FQ_R2; }
};

// ---------------------------------------------------------------------------------------
// Fp<P>: value type, Montgomery form, always fully reduced in [0, p)
// ---------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    __device__ __forceinline__ static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    __device__ __forceinline__ static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one()[i];
        return r;
    }
    __device__ __forceinline__ bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= v[i];
        return o == 0;
    }
    __device__ __forceinline__ bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    __device__ __forceinline__ bool operator!=(const Fp& b) const { return !(*this == b); }

    // r = t - p if t >= p else t   (t < 2p, no overflow out of N limbs)
    __device__ __forceinline__ static void final_sub(uint32_t* t) {
        uint32_t s[N];
        s[0] = ptx::sub_cc(t[0], P::mod()[0]);
#pragma unroll
        for (int i = 1; i < N; i++) s[i] = ptx::subc_cc(t[i], P::mod()[i]);
        uint32_t borrow = ptx::subc(0, 0);  // 0 if no borrow, 0xffffffff if borrow
#pragma unroll
        for (int i = 0; i < N; i++) t[i] = borrow ? t[i] : s[i];
    }

    __device__ __forceinline__ friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
        r.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);
        final_sub(r.v);
        return r;
    }
    __device__ __forceinline__ friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = ptx::subc(0, 0);
        // add p back when the subtraction borrowed
        r.v[0] = ptx::add_cc(r.v[0], P::mod()[0] & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod()[i] & borrow);
        r.v[N - 1] = ptx::addc(r.v[N - 1], P::mod()[N - 1] & borrow);
        return r;
    }
    __device__ __forceinline__ Fp neg() const { return zero() - *this; }
    __device__ __forceinline__ Fp dbl() const { return *this + *this; }

    // ---- Montgomery product on 64-bit columns -------------------------------------------
    // The running sum T is held as two arrays of H = N/2 64-bit columns: A is aligned to
    // limb 0 (column k = limbs 2k, 2k+1) and B to limb 1 (column k = limbs 2k+1, 2k+2), so
    // every 32x32 product a[j]*b lands on one whole column of A (j even) or B (j odd) and
    // `mul.wide.u32` + `add(c).cc.u64` compiles to one IMAD.WIDE.U32(.X) with carry in/out.
    static constexpr int H = N / 2;

    // X[k] += x[2k] * b for all columns; fresh carry chain, carry out returned in CC.
    __device__ __forceinline__ static void mad_cols(uint64_t* X, const uint32_t* x, uint32_t b) {
        X[0] = ptx::add_cc64(X[0], ptx::mul_wide(x[0], b));
#pragma unroll
        for (int k = 1; k < H; k++) X[k] = ptx::addc_cc64(X[k], ptx::mul_wide(x[2 * k], b));
    }
    // m = A.limb0 * (-p^-1);  B += m * p_odd, A += m * p_even  =>  A.limb0 == 0
    __device__ __forceinline__ static void reduce(uint64_t* A, uint64_t* B) {
        uint32_t m = (uint32_t)A[0] * P::INV;
        mad_cols(B, P::mod() + 1, m);  // cannot carry out: T < 2^(32(N+1))
        mad_cols(A, P::mod(), m);
        uint32_t c = ptx::addc(0, 0);
        B[H - 1] += (uint64_t)c << 32;
    }
    // One word of b.  A: aligned to limb 0 of T/2^32 (it was the limb-1 array before the
    // shift); B: the previous limb-0 array whose limb 0 is now zero; it is shifted down one
    // column while the odd products are added, its limb 1 is folded into A's first product.
    __device__ __forceinline__ static void step(uint64_t* A, uint64_t* B, const uint32_t* a, uint32_t bi) {
        uint64_t t0 = ptx::mad_wide(a[0], bi, B[0] >> 32);
        B[0] = ptx::add_cc64(B[1], ptx::mul_wide(a[1], bi));
#pragma unroll
        for (int k = 1; k < H - 1; k++) B[k] = ptx::addc_cc64(B[k + 1], ptx::mul_wide(a[2 * k + 1], bi));
        B[H - 1] = ptx::addc64(ptx::mul_wide(a[N - 1], bi), 0);
        A[0] = ptx::add_cc64(A[0], t0);
#pragma unroll
        for (int k = 1; k < H; k++) A[k] = ptx::addc_cc64(A[k], ptx::mul_wide(a[2 * k], bi));
        uint32_t c = ptx::addc(0, 0);
        B[H - 1] += (uint64_t)c << 32;
        reduce(A, B);
    }

    __device__ __forceinline__ friend Fp operator*(const Fp& a, const Fp& b) {
        uint64_t ev[H], od[H];
#pragma unroll
        for (int k = 0; k < H; k++) {
            ev[k] = ptx::mul_wide(a.v[2 * k], b.v[0]);
            od[k] = ptx::mul_wide(a.v[2 * k + 1], b.v[0]);
        }
        reduce(ev, od);
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            step(od, ev, a.v, b.v[i]);
            if (i + 1 < N) step(ev, od, a.v, b.v[i + 1]);
        }
        // N even: the last step ran with A = od (limb 0 now zero), B = ev.  Result = ev + (od >> 32).
        Fp r;
        uint64_t s;
        s = (od[0] >> 32) | (od[1] << 32);
        uint64_t t = ptx::add_cc64(ev[0], s);
        r.v[0] = (uint32_t)t; r.v[1] = (uint32_t)(t >> 32);
#pragma unroll
        for (int k = 1; k < H; k++) {
            s = (k + 1 < H) ? ((od[k] >> 32) | (od[k + 1] << 32)) : (od[k] >> 32);
            t = (k + 1 < H) ? ptx::addc_cc64(ev[k], s) : ptx::addc64(ev[k], s);
            r.v[2 * k] = (uint32_t)t; r.v[2 * k + 1] = (uint32_t)(t >> 32);
        }
        final_sub(r.v);
        return r;
    }
    __device__ __forceinline__ Fp sqr() const { return *this * *this; }

    __device__ __forceinline__ Fp& operator+=(const Fp& b) { *this = *this + b; return *this; }
    __device__ __forceinline__ Fp& operator-=(const Fp& b) { *this = *this - b; return *this; }
    __device__ __forceinline__ Fp& operator*=(const Fp& b) { *this = *this * b; return *this; }

    // Montgomery <-> canonical
    __device__ __forceinline__ Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }
    __device__ __forceinline__ Fp to_mont() const {
        Fp r2;
#pragma unroll
        for (int i = 0; i < N; i++) r2.v[i] = P::r2()[i];
        return *this * r2;
    }

    // this^e for a little-endian multi-word exponent (not constant time; exponents are public)
    __device__ __noinline__ Fp pow(const uint32_t* e, int words) const {
        Fp acc = one();
        bool started = false;
        for (int w = words - 1; w >= 0; w--) {
            for (int bit = 31; bit >= 0; bit--) {
                if (started) acc = acc.sqr();
                if ((e[w] >> bit) & 1) {
                    acc = started ? acc * *this : *this;
                    started = true;
                }
            }
        }
        return acc;
    }
    __device__ Fp pow_u64(uint64_t e) const {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow(w, 2);
    }
    // Fermat inverse a^(p-2); inverse of zero is zero.  ~1.5 * bits dependent products: kept as the
    // cross-check of inv() (field_kernels.cu: k_inv_selftest).
    __device__ Fp inv_fermat() const {
        uint32_t e[N];
        e[0] = ptx::sub_cc(P::mod()[0], 2u);
#pragma unroll
        for (int i = 1; i < N - 1; i++) e[i] = ptx::subc_cc(P::mod()[i], 0u);
        e[N - 1] = ptx::subc(P::mod()[N - 1], 0u);
        return pow(e, N);
    }

    // x / 2^k mod p for 1 <= k <= 31: add the multiple of p that clears the low k bits, then shift
    __device__ __forceinline__ static void div_pow2(uint32_t* x, int k) {
        const uint32_t m = (x[0] * P::INV) & ((1u << k) - 1u);
        uint32_t t[N + 1];
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint64_t sum = (uint64_t)m * P::mod()[i] + x[i] + carry;
            t[i] = (uint32_t)sum;
            carry = sum >> 32;
        }
        t[N] = (uint32_t)carry;
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __funnelshift_r(t[i], t[i + 1], k);
    }
    // Inverse by the binary extended Euclidean algorithm (subtract, then strip all trailing zero bits at
    // once); inverse of zero is zero.  Invariants: x1 * a = u * R^2, x2 * a = v * R^2 (mod p) on the integer
    // behind the Montgomery form, so the result is again in Montgomery form.  About 0.7 * 2 * bits rounds of
    // ~100 ALU/IMAD instructions instead of ~1.5 * bits full products.
    __device__ __noinline__ Fp inv() const {
        if (is_zero()) return zero();
        uint32_t u[N], v[N], x1[N], x2[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            u[i] = this->v[i];
            v[i] = P::mod()[i];
            x1[i] = P::r2()[i];
            x2[i] = 0;
        }
        for (;;) {
            // strip the trailing zeros of u (u != 0), dividing x1 alongside
            while ((u[0] & 1u) == 0) {
                int k = u[0] ? __ffs((int)u[0]) - 1 : 31;
                if (k > 31) k = 31;
#pragma unroll
                for (int i = 0; i < N; i++) u[i] = __funnelshift_r(u[i], i + 1 < N ? u[i + 1] : 0u, k);
                div_pow2(x1, k);
            }
            // d = u - v
            uint32_t d[N];
            d[0] = ptx::sub_cc(u[0], v[0]);
#pragma unroll
            for (int i = 1; i < N; i++) d[i] = ptx::subc_cc(u[i], v[i]);
            const uint32_t borrow = ptx::subc(0, 0);
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) nz |= d[i];
            if (nz == 0) break;                     // u == v == gcd = 1
            if (borrow) {
                // u < v: swap the pairs, d = v - u
#pragma unroll
                for (int i = 0; i < N; i++) {
                    uint32_t t = u[i]; u[i] = v[i]; v[i] = t;
                    t = x1[i]; x1[i] = x2[i]; x2[i] = t;
                }
                d[0] = ptx::sub_cc(u[0], v[0]);
#pragma unroll
                for (int i = 1; i < N - 1; i++) d[i] = ptx::subc_cc(u[i], v[i]);
                d[N - 1] = ptx::subc(u[N - 1], v[N - 1]);
            }
            // u <- u - v (even, non-zero), x1 <- x1 - x2 mod p
#pragma unroll
            for (int i = 0; i < N; i++) u[i] = d[i];
            x1[0] = ptx::sub_cc(x1[0], x2[0]);
#pragma unroll
            for (int i = 1; i < N; i++) x1[i] = ptx::subc_cc(x1[i], x2[i]);
            const uint32_t b2 = ptx::subc(0, 0);
            x1[0] = ptx::add_cc(x1[0], P::mod()[0] & b2);
#pragma unroll
            for (int i = 1; i < N - 1; i++) x1[i] = ptx::addc_cc(x1[i], P::mod()[i] & b2);
            x1[N - 1] = ptx::addc(x1[N - 1], P::mod()[N - 1] & b2);
        }
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = x1[i];
        return r;
    }
};

using Fr = Fp<FrP>;
using Fq = Fp<FqP>;

}  // namespace pm
