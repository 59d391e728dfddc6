// K5 — fixed-base batch scalar multiplication for the generator.
//
// Replaces `generate()` (/root/reference/src/generator.rs:169-177), which computes
// `(g * f(j)).into()` sequentially with a full double-and-add and one inversion per point.
// Here: a windowed table of multiples of G (kWindows x (2^kWBits - 1) affine points, built
// once per process on the device), one thread per scalar doing kWindows mixed additions in
// XYZZ, then a batched normalisation with one Fq inversion per kBatch points.
#include "fixed_base.cuh"

namespace pm {

namespace {

constexpr int kWBits = 12;
constexpr int kWindows = (255 + kWBits - 1) / kWBits;  // 22
constexpr int kEntries = (1 << kWBits) - 1;
constexpr int kBatch = 16;

__device__ __constant__ uint32_t G1_GEN_X[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                                                0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
__device__ __constant__ uint32_t G1_GEN_Y[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                                                0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};

// bases[w] = 2^(kWBits*w) * G as XYZZ (single thread; 21*12 doublings)
__global__ void k_window_bases(G1XYZZ* bases) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1XYZZ p;
#pragma unroll
    for (int i = 0; i < 12; i++) { p.x.v[i] = G1_GEN_X[i]; p.y.v[i] = G1_GEN_Y[i]; }
    p.zz = Fq::one();
    p.zzz = Fq::one();
    for (int w = 0; w < kWindows; w++) {
        bases[w] = p;
        for (int k = 0; k < kWBits; k++) xyzz_dbl(p);
    }
}

// table_xyzz[w][d-1] = d * bases[w]
__global__ void k_table_entries(const G1XYZZ* __restrict__ bases, G1XYZZ* __restrict__ table_xyzz) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kWindows * kEntries) return;
    int w = t / kEntries, d = t % kEntries + 1;
    table_xyzz[t] = xyzz_mul_small(bases[w], (uint32_t)d);
}

__global__ void __launch_bounds__(128) k_fixed_base(const G1Affine* __restrict__ table, const Fr* __restrict__ scalars, size_t n,
                                                    G1XYZZ* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = scalars[i].from_mont();
    uint32_t limbs[8];
#pragma unroll
    for (int k = 0; k < 8; k++) limbs[k] = s.v[k];
    G1XYZZ acc = G1XYZZ::inf();
    for (int w = 0; w < kWindows; w++) {
        int pos = w * kWBits;
        int limb = pos >> 5, off = pos & 31;
        uint64_t v = limbs[limb];
        if (limb + 1 < 8) v |= (uint64_t)limbs[limb + 1] << 32;
        uint32_t d = (uint32_t)(v >> off) & ((1u << kWBits) - 1u);
        if (d == 0) continue;
        G1Affine p = table[(size_t)w * kEntries + (d - 1)];
        xyzz_madd(acc, p, false);
    }
    out[i] = acc;
}

// Montgomery batch inversion of the ZZZ coordinates of kBatch consecutive points per thread.
__global__ void __launch_bounds__(128) k_batch_to_affine(const G1XYZZ* __restrict__ in, size_t n, G1Affine* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * kBatch;
    if (lo >= n) return;
    int cnt = (int)((n - lo) < (size_t)kBatch ? (n - lo) : (size_t)kBatch);
    Fq prefix[kBatch];
    Fq acc = Fq::one();
    for (int k = 0; k < cnt; k++) {
        prefix[k] = acc;
        Fq z = in[lo + k].zzz;
        if (!z.is_zero()) acc = acc * z;
    }
    Fq inv = acc.inv();
    for (int k = cnt - 1; k >= 0; k--) {
        G1XYZZ p = in[lo + k];
        G1Affine a;
        if (p.zzz.is_zero()) {
            a = G1Affine::inf();
        } else {
            Fq izzz = inv * prefix[k];
            inv = inv * p.zzz;
            Fq tq = p.zz * izzz;
            Fq izz = tq.sqr();
            a.x = p.x * izz;
            a.y = p.y * izzz;
        }
        out[lo + k] = a;
    }
}

}  // namespace

void launch_batch_to_affine(const G1XYZZ* in, size_t n, G1Affine* out, cudaStream_t stream) {
    if (n == 0) return;
    size_t threads = (n + kBatch - 1) / kBatch;
    k_batch_to_affine<<<ceil_div(threads, 128), 128, 0, stream>>>(in, n, out);
    PM_LAUNCH_CHECK();
}

void FixedBaseEngine::ensure_table(cudaStream_t stream) {
    if (built_) return;
    const size_t total = (size_t)kWindows * kEntries;
    G1Affine* table = table_.as<G1Affine>(total);
    DevBuf bases, tmp;
    G1XYZZ* b = bases.as<G1XYZZ>(kWindows);
    G1XYZZ* tx = tmp.as<G1XYZZ>(total);
    k_window_bases<<<1, 32, 0, stream>>>(b);
    PM_LAUNCH_CHECK();
    k_table_entries<<<ceil_div(total, 128), 128, 0, stream>>>(b, tx);
    PM_LAUNCH_CHECK();
    launch_batch_to_affine(tx, total, table, stream);
    PM_CUDA(cudaStreamSynchronize(stream));  // scratch buffers die here
    launches += 3;
    built_ = true;
}

void FixedBaseEngine::run(const Fr* scalars, size_t n, G1Affine* out, cudaStream_t stream) {
    if (n == 0) return;
    ensure_table(stream);
    const size_t chunk = (size_t)1 << 22;  // bound the XYZZ scratch (192 B / point)
    G1XYZZ* xyzz = xyzz_.as<G1XYZZ>(n < chunk ? n : chunk);
    for (size_t lo = 0; lo < n; lo += chunk) {
        size_t cnt = (n - lo) < chunk ? (n - lo) : chunk;
        k_fixed_base<<<ceil_div(cnt, 128), 128, 0, stream>>>(table_.get<G1Affine>(), scalars + lo, cnt, xyzz);
        PM_LAUNCH_CHECK();
        launch_batch_to_affine(xyzz, cnt, out + lo, stream);
        launches += 2;
    }
}

}  // namespace pm
