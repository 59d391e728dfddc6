#include "synth.cuh"

namespace pm {
namespace {
__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void k_fill_fr(Fr* out, size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed ^ (i * 0xd1342543de82ef95ull);
    Fr v;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint64_t w = splitmix(s);
        v.v[2 * k] = (uint32_t)w;
        v.v[2 * k + 1] = (uint32_t)(w >> 32);
    }
    v.v[7] &= 0x3fffffffu;
    out[i] = v;
}
// SURVEY.md 8d skew: ~90 % of the scalars take ONE repeated value, ~10 % are zero, the rest stay uniform; 1 % of
// the bases become the point at infinity (S-dummy keys: benches/bench.rs:38-61 leaves ~n infinity bases).
__global__ void k_skew_msm_inputs(Fr* scalars, G1Affine* bases, size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed ^ (i * 0x9e3779b97f4a7c15ull);
    const uint32_t h = (uint32_t)(splitmix(s) % 1000u);
    if (h < 890u) {
        Fr v;
#pragma unroll
        for (int k = 0; k < 8; k++) v.v[k] = 0x9d2c5680u + 0x01000193u * (uint32_t)k;
        v.v[7] &= 0x3fffffffu;
        scalars[i] = v;
    } else if (h < 990u) {
        scalars[i] = Fr::zero();
    }
    if ((uint32_t)(splitmix(s) % 100u) == 0u) {
        G1Affine inf;
        inf.x = Fq::zero();
        inf.y = Fq::zero();
        bases[i] = inf;
    }
}
}  // namespace
void launch_skew_msm_inputs(Fr* scalars, G1Affine* bases, size_t n, uint64_t seed, cudaStream_t stream) {
    if (n == 0) return;
    k_skew_msm_inputs<<<ceil_div(n, 256), 256, 0, stream>>>(scalars, bases, n, seed);
    PM_LAUNCH_CHECK();
}
void launch_fill_fr(Fr* out, size_t n, uint64_t seed, cudaStream_t stream) {
    if (n == 0) return;
    k_fill_fr<<<ceil_div(n, 256), 256, 0, stream>>>(out, n, seed);
    PM_LAUNCH_CHECK();
}
}  // namespace pm
