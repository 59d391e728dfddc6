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
#ifndef PM_HOST_EMU   // tests/csrc/field_emu_test.cpp supplies the same functions on the host (explicit carry flag)
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
__device__ __forceinline__ uint64_t sub_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("sub.cc.u64 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t subc_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("subc.cc.u64 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
}  // namespace ptx
#endif

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
    // One reduction word WITHOUT a multiplicand word: step() with b_i = 0, the carried limb t0 folded into the first
    // product of the A chain (p_0 m + t0 < 2^64), so a word costs N wide IMADs like reduce().
    __device__ __forceinline__ Fp sqr() const { return *this * *this; }
    __device__ __forceinline__ static void redc_step(uint64_t* A, uint64_t* B) {
        const uint32_t t0 = (uint32_t)(B[0] >> 32);
        const uint32_t m = ((uint32_t)A[0] + t0) * P::INV;
        B[0] = ptx::add_cc64(B[1], ptx::mul_wide(P::mod()[1], m));
#pragma unroll
        for (int k = 1; k < H - 1; k++) B[k] = ptx::addc_cc64(B[k + 1], ptx::mul_wide(P::mod()[2 * k + 1], m));
        B[H - 1] = ptx::addc64(ptx::mul_wide(P::mod()[N - 1], m), 0);
        A[0] = ptx::add_cc64(A[0], ptx::mad_wide(P::mod()[0], m, t0));
#pragma unroll
        for (int k = 1; k < H; k++) A[k] = ptx::addc_cc64(A[k], ptx::mul_wide(P::mod()[2 * k], m));
        uint32_t c = ptx::addc(0, 0);
        B[H - 1] += (uint64_t)c << 32;
    }
    // Montgomery reduction of a 2N-limb value T < p * 2^(32N) given as N 64-bit columns (T[k] = limbs 2k, 2k+1):
    // T_hi + (T_lo + M p) / 2^(32N) < 2p, one conditional subtraction.  N^2 wide IMADs.
    __device__ __forceinline__ static Fp redc_wide(const uint64_t* T) {
        uint64_t ev[H], od[H];
#pragma unroll
        for (int k = 0; k < H; k++) { ev[k] = T[k]; od[k] = 0; }
        reduce(ev, od);
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            redc_step(od, ev);
            if (i + 1 < N) redc_step(ev, od);
        }
        // low part: ev + (od >> 32), then the high half of T on top
        Fp r;
        uint64_t lo[H];
        uint64_t sft = (od[0] >> 32) | (od[1] << 32);
        lo[0] = ptx::add_cc64(ev[0], sft);
#pragma unroll
        for (int k = 1; k < H; k++) {
            sft = (k + 1 < H) ? ((od[k] >> 32) | (od[k + 1] << 32)) : (od[k] >> 32);
            lo[k] = (k + 1 < H) ? ptx::addc_cc64(ev[k], sft) : ptx::addc64(ev[k], sft);
        }
        lo[0] = ptx::add_cc64(lo[0], T[H]);
#pragma unroll
        for (int k = 1; k < H; k++) lo[k] = (k + 1 < H) ? ptx::addc_cc64(lo[k], T[H + k]) : ptx::addc64(lo[k], T[H + k]);
#pragma unroll
        for (int k = 0; k < H; k++) { r.v[2 * k] = (uint32_t)lo[k]; r.v[2 * k + 1] = (uint32_t)(lo[k] >> 32); }
        final_sub(r.v);
        return r;
    }
    // ---- Karatsuba product (one level) + redc_wide ------------------------------------------------------------
    // Z (HL 64-bit columns) = x * y for HL-limb operands: rows in increasing i, two carry chains per row (positions
    // i + j even -> E, odd -> O), the carry-out of a chain absorbed by the column above its top (carries only so far).
    static constexpr int HL = N / 2;      // limbs of a half operand (even: N is a multiple of 4)
    static constexpr int HC = N / 4;      // 64-bit columns of a half operand
    __device__ __forceinline__ static void mul_half(const uint32_t* x, const uint32_t* y, uint64_t* Z) {
        uint64_t E[HL], O[HL];
#pragma unroll
        for (int k = 0; k < HL; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
        for (int j = 0; j < HL; j++) {
            if (j & 1) O[(j - 1) / 2] = ptx::mul_wide(x[j], y[0]);
            else E[j / 2] = ptx::mul_wide(x[j], y[0]);
        }
#pragma unroll
        for (int i = 1; i < HL; i++) {
            {   // even positions: j = i mod 2, +2, ...
                int last = -1;
#pragma unroll
                for (int j = i & 1; j < HL; j += 2) {
                    const int k = (i + j) / 2;
                    E[k] = (j == (i & 1)) ? ptx::add_cc64(E[k], ptx::mul_wide(x[j], y[i])) : ptx::addc_cc64(E[k], ptx::mul_wide(x[j], y[i]));
                    last = k;
                }
                if (last + 1 < HL) E[last + 1] = ptx::addc64(E[last + 1], 0);
            }
            {   // odd positions
                int last = -1;
#pragma unroll
                for (int j = 1 - (i & 1); j < HL; j += 2) {
                    const int k = (i + j - 1) / 2;
                    O[k] = (j == 1 - (i & 1)) ? ptx::add_cc64(O[k], ptx::mul_wide(x[j], y[i])) : ptx::addc_cc64(O[k], ptx::mul_wide(x[j], y[i]));
                    last = k;
                }
                if (last + 1 < HL) O[last + 1] = ptx::addc64(O[last + 1], 0);
            }
        }
        Z[0] = ptx::add_cc64(E[0], O[0] << 32);
#pragma unroll
        for (int k = 1; k < HL; k++) {
            const uint64_t o = (O[k] << 32) | (O[k - 1] >> 32);
            Z[k] = (k + 1 < HL) ? ptx::addc_cc64(E[k], o) : ptx::addc64(E[k], o);
        }
    }
    // a * b = z0 + (M - z0 - z2) 2^(32 HL) + z2 2^(64 HL) with z0 = a_lo b_lo, z2 = a_hi b_hi, M = (a_lo + a_hi)(b_lo + b_hi):
    // 3 (N/2)^2 wide IMADs instead of N^2 (Fq: 108 instead of 144; with the reduction 252 instead of 288); the extra
    // ~110 additions run on the ALU pipe.  MEASURED AND NOT USED (pm_bench_field_mul variants 3 / 5): 28.6 G/s against
    // 29.4 G/s for the word-serial product in Fq, 60.5 against 65.5 in Fr — the saved multiplier slots are eaten by the
    // additions and the longer dependency chain.  Kept as the checked alternative (tests/csrc/field_emu_test.cpp).
    __device__ __forceinline__ static Fp mul_karatsuba(const Fp& a, const Fp& b) {
        uint64_t T[N];
        mul_half(a.v, b.v, T);                    // z0 -> T[0, HL)
        mul_half(a.v + HL, b.v + HL, T + HL);     // z2 -> T[HL, N)
        uint32_t sa[HL], sb[HL];
        sa[0] = ptx::add_cc(a.v[0], a.v[HL]);
#pragma unroll
        for (int i = 1; i < HL; i++) sa[i] = ptx::addc_cc(a.v[i], a.v[HL + i]);
        const uint32_t ca = ptx::addc(0, 0);
        sb[0] = ptx::add_cc(b.v[0], b.v[HL]);
#pragma unroll
        for (int i = 1; i < HL; i++) sb[i] = ptx::addc_cc(b.v[i], b.v[HL + i]);
        const uint32_t cb = ptx::addc(0, 0);
        uint64_t M[HL];
        mul_half(sa, sb, M);
        // the carried bits: + ca * sb * 2^(32 HL) + cb * sa * 2^(32 HL) + ca * cb * 2^(64 HL); top word beyond the HL columns
        const uint64_t ma = 0 - (uint64_t)ca, mb = 0 - (uint64_t)cb;
        uint64_t mtop = ca & cb;
        M[HC] = ptx::add_cc64(M[HC], ma & ((uint64_t)sb[0] | ((uint64_t)sb[1] << 32)));
#pragma unroll
        for (int k = 1; k < HC; k++) M[HC + k] = ptx::addc_cc64(M[HC + k], ma & ((uint64_t)sb[2 * k] | ((uint64_t)sb[2 * k + 1] << 32)));
        mtop = ptx::addc64(mtop, 0);
        M[HC] = ptx::add_cc64(M[HC], mb & ((uint64_t)sa[0] | ((uint64_t)sa[1] << 32)));
#pragma unroll
        for (int k = 1; k < HC; k++) M[HC + k] = ptx::addc_cc64(M[HC + k], mb & ((uint64_t)sa[2 * k] | ((uint64_t)sa[2 * k + 1] << 32)));
        mtop = ptx::addc64(mtop, 0);
        // z1 = M - z0 - z2 (non-negative, < 2^(64 HL + 1)): the borrows come out of mtop
        M[0] = ptx::sub_cc64(M[0], T[0]);
#pragma unroll
        for (int k = 1; k < HL; k++) M[k] = ptx::subc_cc64(M[k], T[k]);
        mtop = ptx::subc_cc64(mtop, 0);
        M[0] = ptx::sub_cc64(M[0], T[HL]);
#pragma unroll
        for (int k = 1; k < HL; k++) M[k] = ptx::subc_cc64(M[k], T[HL + k]);
        mtop = ptx::subc_cc64(mtop, 0);
        // T += z1 * 2^(32 HL): columns HC .. HC + HL, then the carry runs to the top
        T[HC] = ptx::add_cc64(T[HC], M[0]);
#pragma unroll
        for (int k = 1; k < HL; k++) T[HC + k] = ptx::addc_cc64(T[HC + k], M[k]);
#pragma unroll
        for (int k = HC + HL; k < N; k++) T[k] = (k + 1 < N) ? ptx::addc_cc64(T[k], k == HC + HL ? mtop : 0) : ptx::addc64(T[k], k == HC + HL ? mtop : 0);
        return redc_wide(T);
    }
    // Squaring: the N(N-1)/2 products a_i a_j (i < j) once, doubled, plus the N squares a_i^2 — N(N+1)/2 wide IMADs
    // instead of N^2 — then redc_wide: 222 instead of 288 for Fq.  Columns as in the product: E[k] = limbs (2k, 2k+1),
    // O[k] = limbs (2k+1, 2k+2); a_i a_j lands on E[(i+j)/2] or O[(i+j-1)/2].  Rows are added in increasing i, each as
    // two carry chains (even / odd positions); the column above a chain's top has only seen carries so far, so the
    // chain's carry-out is absorbed there without further propagation.
    // Measured on B200 (profiles/r2_summary.md): in a register-resident chain of independent squarings 46 G/s against
    // 29 G/s for the product, but in the kernels the gain is small — the group operations are bound by the latency of
    // their dependent carry chains at 25 % occupancy, not by the IMAD count: standalone MSM 2^22 +1.2 %, prove +0 %;
    // the Fermat / square-root chains of the G1 codec run 6 % SLOWER with it (longer critical path per squaring), so
    // sqr() below stays the product and only the MSM group operations call sqr_wide().
    __device__ __forceinline__ Fp sqr_wide() const {
        const uint32_t* a = v;
        uint64_t E[N], O[N];
#pragma unroll
        for (int k = 0; k < N; k++) { E[k] = 0; O[k] = 0; }
        // row 0 initialises its columns
#pragma unroll
        for (int j = 1; j < N; j++) {
            if (j & 1) O[(j - 1) / 2] = ptx::mul_wide(a[0], a[j]);
            else E[j / 2] = ptx::mul_wide(a[0], a[j]);
        }
#pragma unroll
        for (int i = 1; i < N - 1; i++) {
            // odd positions i + j: j = i+1, i+3, ...
            {
                int last = -1;
#pragma unroll
                for (int j = i + 1; j < N; j += 2) {
                    const int k = (i + j - 1) / 2;
                    O[k] = (j == i + 1) ? ptx::add_cc64(O[k], ptx::mul_wide(a[i], a[j])) : ptx::addc_cc64(O[k], ptx::mul_wide(a[i], a[j]));
                    last = k;
                }
                if (last >= 0) O[last + 1] = ptx::addc64(O[last + 1], 0);
            }
            // even positions: j = i+2, i+4, ...
            {
                int last = -1;
#pragma unroll
                for (int j = i + 2; j < N; j += 2) {
                    const int k = (i + j) / 2;
                    E[k] = (j == i + 2) ? ptx::add_cc64(E[k], ptx::mul_wide(a[i], a[j])) : ptx::addc_cc64(E[k], ptx::mul_wide(a[i], a[j]));
                    last = k;
                }
                if (last >= 0) E[last + 1] = ptx::addc64(E[last + 1], 0);
            }
        }
        // S = E + (O << 32) on the E grid, doubled, plus the squares on the diagonal
        uint64_t S[N];
        S[0] = ptx::add_cc64(E[0], O[0] << 32);
#pragma unroll
        for (int k = 1; k < N; k++) {
            const uint64_t o = (O[k] << 32) | (O[k - 1] >> 32);
            S[k] = (k + 1 < N) ? ptx::addc_cc64(E[k], o) : ptx::addc64(E[k], o);
        }
        uint64_t T[N];
        T[0] = ptx::add_cc64(S[0] << 1, ptx::mul_wide(a[0], a[0]));
#pragma unroll
        for (int k = 1; k < N; k++) {
            const uint64_t d = (S[k] << 1) | (S[k - 1] >> 63);
            T[k] = (k + 1 < N) ? ptx::addc_cc64(d, ptx::mul_wide(a[k], a[k])) : ptx::addc64(d, ptx::mul_wide(a[k], a[k]));
        }
        return redc_wide(T);
    }

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
