// K1 drivers: elementwise batch kernels over Fr / Fq, plus the microbenchmarks that give the
// INT32 IMAD-pipe roofline denominator used for the MSM / field-multiplication kernels
// ("to be measured on the box with a dependency-free mad.wide.u32 microkernel").
#include "field_kernels.cuh"
#include "runtime.cuh"

namespace pm {

namespace {

template <class F, int OP>
__global__ void __launch_bounds__(256) k_batch(const F* __restrict__ a, const F* __restrict__ b, F* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        F x = a[i], y = b[i];
        out[i] = OP == 0 ? x * y : (OP == 1 ? x + y : (OP == 2 ? x - y : x.inv()));
    }
}

template <class F>
void launch(FieldOp op, const F* a, const F* b, F* out, size_t n, cudaStream_t stream) {
    if (n == 0) return;
    unsigned grid = ceil_div(n, 256);
    unsigned cap = (unsigned)sm_count() * 16;
    if (grid > cap) grid = cap;
    switch (op) {
        case FieldOp::Mul: k_batch<F, 0><<<grid, 256, 0, stream>>>(a, b, out, n); break;
        case FieldOp::Add: k_batch<F, 1><<<grid, 256, 0, stream>>>(a, b, out, n); break;
        case FieldOp::Sub: k_batch<F, 2><<<grid, 256, 0, stream>>>(a, b, out, n); break;
        case FieldOp::Inv: k_batch<F, 3><<<grid, 256, 0, stream>>>(a, b, out, n); break;
    }
    PM_LAUNCH_CHECK();
}

// 8 independent 64-bit accumulators per thread, each fed by IMAD.WIDE.U32: no dependency
// between consecutive instructions closer than 8 issues.
__global__ void __launch_bounds__(256) k_imad_peak(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = ptx::mad_wide(x, y, acc[k]);
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567812345678ull) sink[0] = s;  // never true in practice; keeps the loop alive
}

template <class F, int MODE>
__global__ void __launch_bounds__(256) k_mul_rate(F* sink, int depth) {
    F x, y;
#pragma unroll
    for (int i = 0; i < F::N; i++) { x.v[i] = threadIdx.x * 7u + i; y.v[i] = blockIdx.x * 13u + i + 1; }
    x.v[F::N - 1] &= 0x0fffffffu;
    y.v[F::N - 1] &= 0x0fffffffu;
    for (int it = 0; it < depth; it++) {
        if (MODE == 1) {
            x = x.sqr_wide();
            y = y.sqr_wide();
        } else if (MODE == 2) {
            x = F::mul_karatsuba(x, y);
            y = F::mul_karatsuba(y, x);
        } else {
            x = x * y;
            y = y * x;
        }
    }
    if (x.v[0] == 0xdeadbeefu && y.v[1] == 0x12345u) sink[0] = x;
}

__global__ void k_fr_inverse(const Fr* in, Fr* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = in[0].inv();
}

template <class K, class... A>
double time_kernel_ms(K kernel, dim3 grid, dim3 block, A... args) {
    cudaEvent_t e0, e1;
    PM_CUDA(cudaEventCreate(&e0));
    PM_CUDA(cudaEventCreate(&e1));
    kernel<<<grid, block>>>(args...);  // warm-up
    PM_CUDA(cudaDeviceSynchronize());
    PM_CUDA(cudaEventRecord(e0));
    kernel<<<grid, block>>>(args...);
    PM_CUDA(cudaEventRecord(e1));
    PM_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

}  // namespace

void launch_fr_batch(FieldOp op, const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t s) { launch<Fr>(op, a, b, out, n, s); }
void launch_fq_batch(FieldOp op, const Fq* a, const Fq* b, Fq* out, size_t n, cudaStream_t s) { launch<Fq>(op, a, b, out, n, s); }

void launch_fr_inverse(const Fr* in, Fr* out, cudaStream_t stream) {
    k_fr_inverse<<<1, 32, 0, stream>>>(in, out);
    PM_LAUNCH_CHECK();
}

double measure_imad_peak(int iters) {
    DevBuf sink;
    uint64_t* d = sink.as<uint64_t>(1);
    const int blocks = sm_count() * 8;
    double ms = time_kernel_ms(k_imad_peak, dim3(blocks), dim3(256), d, 0x9e3779b9u, 0x7f4a7c15u, iters);
    double ops = (double)blocks * 256.0 * (double)iters * 64.0;
    return ops / (ms * 1e-3);
}

double measure_fr_mul_rate(int depth) {
    DevBuf sink;
    Fr* d = sink.as<Fr>(1);
    const int blocks = sm_count() * 8;
    double ms = time_kernel_ms(k_mul_rate<Fr, 0>, dim3(blocks), dim3(256), d, depth);
    return (double)blocks * 256.0 * 2.0 * depth / (ms * 1e-3);
}

double measure_fq_mul_rate(int depth) {
    DevBuf sink;
    Fq* d = sink.as<Fq>(1);
    const int blocks = sm_count() * 8;
    double ms = time_kernel_ms(k_mul_rate<Fq, 0>, dim3(blocks), dim3(256), d, depth);
    return (double)blocks * 256.0 * 2.0 * depth / (ms * 1e-3);
}

// mode 1 = dedicated squaring, 2 = Karatsuba product (Fq; same chain shape as measure_fq_mul_rate)
double measure_fq_variant_rate(int mode, int depth) {
    DevBuf sink;
    Fq* d = sink.as<Fq>(1);
    const int blocks = sm_count() * 8;
    double ms = mode == 1 ? time_kernel_ms(k_mul_rate<Fq, 1>, dim3(blocks), dim3(256), d, depth)
                          : time_kernel_ms(k_mul_rate<Fq, 2>, dim3(blocks), dim3(256), d, depth);
    return (double)blocks * 256.0 * 2.0 * depth / (ms * 1e-3);
}
double measure_fr_variant_rate(int mode, int depth) {
    DevBuf sink;
    Fr* d = sink.as<Fr>(1);
    const int blocks = sm_count() * 8;
    double ms = mode == 1 ? time_kernel_ms(k_mul_rate<Fr, 1>, dim3(blocks), dim3(256), d, depth)
                          : time_kernel_ms(k_mul_rate<Fr, 2>, dim3(blocks), dim3(256), d, depth);
    return (double)blocks * 256.0 * 2.0 * depth / (ms * 1e-3);
}

}  // namespace pm
