// DFMA throughput with distinct operands vs IMAD.WIDE with distinct operands (next-round candidate:
// FP64-based 52-bit-limb Montgomery product on the separate fp64 pipe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* sink, uint64_t* isink, int iters) {
    // 8 accumulators, 8 distinct a's, 8 distinct b's per thread
    double a[8], b[8], c[8];
    uint32_t ia[8], ib[8];
    uint64_t ic[8];
    for (int i = 0; i < 8; i++) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i); b[i] = 1.0 - 1e-9 * (blockIdx.x + 2 * i); c[i] = i;
        ia[i] = threadIdx.x * 2654435761u + i; ib[i] = blockIdx.x * 40503u + 7 * i + 1; ic[i] = i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) c[i] = fma(a[(i + r) & 7], b[i], c[i]);
                else ic[i] += (uint64_t)ia[(i + r) & 7] * ib[i];
            }
        }
    }
    double s = 0; uint64_t is = 0;
    for (int i = 0; i < 8; i++) { s += c[i]; is ^= ic[i]; }
    if (s == 1234.5678) sink[0] = s;
    if (is == 0x123456789abcdefull) isink[0] = is;
}
template <int MODE>
float run(int blocks, int iters, double* d, uint64_t* u) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, u, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, u, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* d; uint64_t* u; cudaMalloc(&d, 64); cudaMalloc(&u, 64);
    const int iters = 2000, blocks = sms * 8;
    double ops = (double)blocks * 256 * iters * 64.0;
    float f = run<0>(blocks, iters, d, u), i = run<1>(blocks, iters, d, u);
    printf("DFMA distinct operands      : %.2f T/s  %.1f per clk per SM\n", ops / f / 1e9, ops / (f * 1e-3) / sms / (khz * 1e3));
    printf("IMAD.WIDE distinct operands : %.2f T/s  %.1f per clk per SM\n", ops / i / 1e9, ops / (i * 1e-3) / sms / (khz * 1e3));
    return 0;
}
