// Instruction- and field-level microbenchmarks for the INT32 pipe on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../polymath_b200/csrc/field.cuh"
using namespace pm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_wide_nocarry(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++)
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = ptx::mad_wide(x, y, acc[k]);
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567812345678ull) sink[0] = s;
}

// two independent carry chains of four IMAD.WIDE.U32(.X) each
__global__ void __launch_bounds__(256) k_wide_carry(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
            acc[0] = ptx::add_cc64(acc[0], ptx::mul_wide(x, y));
            acc[1] = ptx::addc_cc64(acc[1], ptx::mul_wide(x, y));
            acc[2] = ptx::addc_cc64(acc[2], ptx::mul_wide(x, y));
            acc[3] = ptx::addc64(acc[3], ptx::mul_wide(x, y));
            acc[4] = ptx::add_cc64(acc[4], ptx::mul_wide(x, y));
            acc[5] = ptx::addc_cc64(acc[5], ptx::mul_wide(x, y));
            acc[6] = ptx::addc_cc64(acc[6], ptx::mul_wide(x, y));
            acc[7] = ptx::addc64(acc[7], ptx::mul_wide(x, y));
        }
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567812345678ull) sink[0] = s;
}

__global__ void __launch_bounds__(256) k_imad32(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(acc[k]) : "r"(x), "r"(y));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x12345678u) sink[0] = s;
}

__global__ void __launch_bounds__(256) k_imadhi32(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.hi.u32 %0,%1,%2,%0;" : "+r"(acc[k]) : "r"(x), "r"(y));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x12345678u) sink[0] = s;
}

// 32-bit lo/hi pairs with carry (what mad.lo.cc/madc.hi.cc compile to)
__global__ void __launch_bounds__(256) k_pair_carry(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
            acc[0] = ptx::mad_lo_cc(x, y, acc[0]);
            acc[1] = ptx::madc_hi_cc(x, y, acc[1]);
            acc[2] = ptx::madc_lo_cc(x, y, acc[2]);
            acc[3] = ptx::madc_hi_cc(x, y, acc[3]);
            acc[4] = ptx::madc_lo_cc(x, y, acc[4]);
            acc[5] = ptx::madc_hi_cc(x, y, acc[5]);
            acc[6] = ptx::madc_lo_cc(x, y, acc[6]);
            acc[7] = ptx::madc_hi(x, y, acc[7]);
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x12345678u) sink[0] = s;
}

__global__ void __launch_bounds__(256) k_iadd3_carry(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
            acc[0] = ptx::add_cc(acc[0], x);
            acc[1] = ptx::addc_cc(acc[1], x);
            acc[2] = ptx::addc_cc(acc[2], x);
            acc[3] = ptx::addc(acc[3], x);
            acc[4] = ptx::add_cc(acc[4], x);
            acc[5] = ptx::addc_cc(acc[5], x);
            acc[6] = ptx::addc_cc(acc[6], x);
            acc[7] = ptx::addc(acc[7], x);
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x12345678u) sink[0] = s;
}

// ---- v1 multiplier: 32-bit lo/hi pairs (ptxas emits IMAD / IMAD.HI / IADD3.X) ----
template <class P>
struct MulV1 {
    static constexpr int N = P::N;
    __device__ __forceinline__ static void mad_pairs(uint32_t* A, const uint32_t* x, uint32_t b) {
        A[0] = ptx::mad_lo_cc(x[0], b, A[0]);
        A[1] = ptx::madc_hi_cc(x[0], b, A[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            A[j] = ptx::madc_lo_cc(x[j], b, A[j]);
            A[j + 1] = ptx::madc_hi_cc(x[j], b, A[j + 1]);
        }
    }
    __device__ __forceinline__ static void reduce(uint32_t* A, uint32_t* B) {
        uint32_t m = A[0] * P::INV;
        mad_pairs(B, P::mod() + 1, m);
        mad_pairs(A, P::mod(), m);
        B[N - 1] = ptx::addc(B[N - 1], 0);
    }
    __device__ __forceinline__ static void step(uint32_t* A, uint32_t* B, const uint32_t* a, uint32_t bi) {
        A[0] = ptx::add_cc(A[0], B[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            B[j] = ptx::madc_lo_cc(a[j + 1], bi, B[j + 2]);
            B[j + 1] = ptx::madc_hi_cc(a[j + 1], bi, B[j + 3]);
        }
        B[N - 2] = ptx::madc_lo_cc(a[N - 1], bi, 0);
        B[N - 1] = ptx::madc_hi(a[N - 1], bi, 0);
        mad_pairs(A, a, bi);
        B[N - 1] = ptx::addc(B[N - 1], 0);
        reduce(A, B);
    }
    __device__ __forceinline__ static Fp<P> mul(const Fp<P>& a, const Fp<P>& b) {
        uint32_t ev[N], od[N];
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            uint64_t pe = (uint64_t)a.v[j] * b.v[0];
            uint64_t po = (uint64_t)a.v[j + 1] * b.v[0];
            ev[j] = (uint32_t)pe; ev[j + 1] = (uint32_t)(pe >> 32);
            od[j] = (uint32_t)po; od[j + 1] = (uint32_t)(po >> 32);
        }
        reduce(ev, od);
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            step(od, ev, a.v, b.v[i]);
            if (i + 1 < N) step(ev, od, a.v, b.v[i + 1]);
        }
        Fp<P> r;
        r.v[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(ev[k], od[k + 1]);
        r.v[N - 1] = ptx::addc(ev[N - 1], 0);
        Fp<P>::final_sub(r.v);
        return r;
    }
};

template <class F, int V, int TPB>
__global__ void __launch_bounds__(TPB) k_mul_rate(F* sink, int depth) {
    F x, y;
    for (int i = 0; i < F::N; i++) { x.v[i] = threadIdx.x * 7u + i; y.v[i] = blockIdx.x * 13u + i + 1; }
    x.v[F::N - 1] &= 0x0fffffffu;
    y.v[F::N - 1] &= 0x0fffffffu;
    for (int it = 0; it < depth; it++) {
        if (V == 1) { x = MulV1<typename std::conditional<F::N == 8, FrP, FqP>::type>::mul(x, y); y = MulV1<typename std::conditional<F::N == 8, FrP, FqP>::type>::mul(y, x); }
        else { x = x * y; y = y * x; }
    }
    if (x.v[0] == 0xdeadbeefu && y.v[1] == 0x12345u) sink[0] = x;
}

template <class K, class... A>
float time_ms(K kernel, int blocks, int tpb, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<blocks, tpb>>>(args...);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kernel<<<blocks, tpb>>>(args...);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("sms=%d clock_khz=%d\n", sms, khz);
    uint64_t* sink; CK(cudaMalloc(&sink, 4096));
    const int iters = 2000, blocks = sms * 8;
    double ops = (double)blocks * 256 * iters * 64.0;
    struct { const char* name; float ms; } r[] = {
        {"imad_wide_nocarry", time_ms(k_wide_nocarry, blocks, 256, sink, 0x9e3779b9u, 0x7f4a7c15u, iters)},
        {"imad_wide_carry4", time_ms(k_wide_carry, blocks, 256, sink, 0x9e3779b9u, 0x7f4a7c15u, iters)},
        {"imad32_lo", time_ms(k_imad32, blocks, 256, sink, 0x9e3779b9u, 0x7f4a7c15u, iters)},
        {"imad32_hi", time_ms(k_imadhi32, blocks, 256, sink, 0x9e3779b9u, 0x7f4a7c15u, iters)},
        {"pair_lo_hi_carry", time_ms(k_pair_carry, blocks, 256, sink, 0x9e3779b9u, 0x7f4a7c15u, iters)},
        {"iadd3_carry", time_ms(k_iadd3_carry, blocks, 256, sink, 0x9e3779b9u, 0x7f4a7c15u, iters)},
    };
    for (auto& x : r) printf("%-20s %8.3f ms  %7.2f Tops/s  %6.1f ops/clk/SM\n", x.name, x.ms, ops / x.ms / 1e9, ops / (x.ms * 1e-3) / sms / (khz * 1e3));
    Fr* fsink = (Fr*)sink; Fq* qsink = (Fq*)sink;
    const int depth = 1000;
    for (int occ = 0; occ < 3; occ++) {
        int tpb = occ == 0 ? 256 : (occ == 1 ? 128 : 64);
        int bl = sms * 8;
        double muls = (double)bl * tpb * 2.0 * depth;
        float a, b, c, d;
        if (tpb == 256) { a = time_ms(k_mul_rate<Fr, 2, 256>, bl, tpb, fsink, depth); b = time_ms(k_mul_rate<Fr, 1, 256>, bl, tpb, fsink, depth); c = time_ms(k_mul_rate<Fq, 2, 256>, bl, tpb, qsink, depth); d = time_ms(k_mul_rate<Fq, 1, 256>, bl, tpb, qsink, depth); }
        else if (tpb == 128) { a = time_ms(k_mul_rate<Fr, 2, 128>, bl, tpb, fsink, depth); b = time_ms(k_mul_rate<Fr, 1, 128>, bl, tpb, fsink, depth); c = time_ms(k_mul_rate<Fq, 2, 128>, bl, tpb, qsink, depth); d = time_ms(k_mul_rate<Fq, 1, 128>, bl, tpb, qsink, depth); }
        else { a = time_ms(k_mul_rate<Fr, 2, 64>, bl, tpb, fsink, depth); b = time_ms(k_mul_rate<Fr, 1, 64>, bl, tpb, fsink, depth); c = time_ms(k_mul_rate<Fq, 2, 64>, bl, tpb, qsink, depth); d = time_ms(k_mul_rate<Fq, 1, 64>, bl, tpb, qsink, depth); }
        printf("threads/SM=%4d  fr_mul wide %6.2f G/s  pairs %6.2f G/s | fq_mul wide %6.2f G/s  pairs %6.2f G/s\n", tpb * 8,
               muls / a / 1e6, muls / b / 1e6, muls / c / 1e6, muls / d / 1e6);
    }
    return 0;
}
