// Host emulation of polymath_b200/csrc/field.cuh: the PTX wrappers are replaced by functions with an explicit carry
// flag, so the SAME template code (product, squaring, reduction, add/sub) runs on the CPU and can be checked against
// independent big-integer arithmetic (unsigned __int128 schoolbook below) without a GPU.
// Build + run: tests/test_field_emu_cpu.py
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define PM_HOST_EMU 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __restrict__

static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, int k) { return k ? (uint32_t)((((uint64_t)hi << 32) | lo) >> (k & 31)) : lo; }
static inline int __ffs(int x) { return x ? __builtin_ctz((unsigned)x) + 1 : 0; }

namespace pm { namespace ptx {
static uint32_t CF = 0;   // carry (add chains) or borrow (sub chains), like CC.CF
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r = a - b; CF = a < b; return r; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - CF; CF = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF; }
inline uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
inline uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }
inline uint64_t add_cc64(uint64_t a, uint64_t b) { unsigned __int128 t = (unsigned __int128)a + b; CF = (uint32_t)(t >> 64); return (uint64_t)t; }
inline uint64_t addc_cc64(uint64_t a, uint64_t b) { unsigned __int128 t = (unsigned __int128)a + b + CF; CF = (uint32_t)(t >> 64); return (uint64_t)t; }
inline uint64_t addc64(uint64_t a, uint64_t b) { return a + b + CF; }
inline uint64_t sub_cc64(uint64_t a, uint64_t b) { uint64_t r = a - b; CF = a < b; return r; }
inline uint64_t subc_cc64(uint64_t a, uint64_t b) { unsigned __int128 t = (unsigned __int128)a - b - CF; CF = (uint32_t)((t >> 64) & 1); return (uint64_t)t; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)(((uint64_t)a * b) >> 32), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc((uint32_t)(((uint64_t)a * b) >> 32), c); }
}}

#include "../../polymath_b200/csrc/field.cuh"

using namespace pm;

// ---- independent reference: a * b * R^-1 mod p with plain 32-bit schoolbook + trial subtraction ----
template <int N>
static void ref_montmul(const uint32_t* a, const uint32_t* b, const uint32_t* p, uint32_t inv, uint32_t* out) {
    uint32_t t[2 * N + 2] = {0};
    for (int i = 0; i < N; i++) {
        uint64_t c = 0;
        for (int j = 0; j < N; j++) { uint64_t s = (uint64_t)a[i] * b[j] + t[i + j] + c; t[i + j] = (uint32_t)s; c = s >> 32; }
        for (int k = i + N; c; k++) { uint64_t s = (uint64_t)t[k] + c; t[k] = (uint32_t)s; c = s >> 32; }
    }
    for (int i = 0; i < N; i++) {
        uint32_t m = t[i] * inv;
        uint64_t c = 0;
        for (int j = 0; j < N; j++) { uint64_t s = (uint64_t)m * p[j] + t[i + j] + c; t[i + j] = (uint32_t)s; c = s >> 32; }
        for (int k = i + N; c; k++) { uint64_t s = (uint64_t)t[k] + c; t[k] = (uint32_t)s; c = s >> 32; }
    }
    // result = t[N..2N] (N+1 limbs), subtract p if >= p
    uint32_t r[N + 1];
    memcpy(r, t + N, sizeof r);
    bool ge = r[N] != 0;
    if (!ge) {
        ge = true;
        for (int i = N - 1; i >= 0; i--) { if (r[i] != p[i]) { ge = r[i] > p[i]; break; } }
    }
    if (ge) { uint64_t bw = 0; for (int i = 0; i < N; i++) { uint64_t s = (uint64_t)r[i] - p[i] - bw; r[i] = (uint32_t)s; bw = (s >> 32) & 1; } }
    memcpy(out, r, N * sizeof(uint32_t));
}

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

template <class F, class P>
static int run(const char* name, int iters) {
    constexpr int N = F::N;
    int bad = 0;
    for (int it = 0; it < iters; it++) {
        F a, b;
        for (int i = 0; i < N; i++) { a.v[i] = rnd(); b.v[i] = rnd(); }
        // special operands: zero, one, p - 1, all-ones-below-p patterns, equal operands
        if (it % 16 == 1) a = F::zero();
        if (it % 16 == 2) a = F::one();
        if (it % 16 == 3) for (int i = 0; i < N; i++) a.v[i] = P::mod()[i] - (i == 0);
        if (it % 16 == 4) for (int i = 0; i < N; i++) a.v[i] = 0xffffffffu;
        if (it % 16 == 5) for (int i = 0; i < N; i++) a.v[i] = (i & 1) ? 0xffffffffu : 0;
        if (it % 16 == 6) for (int i = 0; i < N; i++) a.v[i] = (i & 1) ? 0 : 0xffffffffu;
        // operands must be reduced: clear high bits, then subtract p while >= p
        auto reduce_in = [&](F& x) {
            x.v[N - 1] &= (N == 8) ? 0x7fffffffu : 0x1fffffffu;
            F::final_sub(x.v);
            F::final_sub(x.v);
        };
        reduce_in(a);
        reduce_in(b);
        uint32_t want[N];
        F got = a * b;
        ref_montmul<N>(a.v, b.v, P::mod(), P::INV, want);
        if (memcmp(got.v, want, sizeof want) != 0) { bad++; if (bad < 4) printf("%s mul mismatch at %d\n", name, it); }
        F sq = a.sqr_wide();
        ref_montmul<N>(a.v, a.v, P::mod(), P::INV, want);
        if (memcmp(sq.v, want, sizeof want) != 0) { bad++; if (bad < 4) printf("%s sqr mismatch at %d\n", name, it); }
        F mk = F::mul_karatsuba(a, b);
        ref_montmul<N>(a.v, b.v, P::mod(), P::INV, want);
        if (memcmp(mk.v, want, sizeof want) != 0) { bad++; if (bad < 4) printf("%s karatsuba mismatch at %d\n", name, it); }
    }
    printf("%s: %d iterations, %d mismatches\n", name, iters, bad);
    return bad;
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 200000;
    int bad = run<Fr, FrP>("Fr", iters) + run<Fq, FqP>("Fq", iters);
    return bad ? 1 : 0;
}
