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
}  // namespace
void launch_fill_fr(Fr* out, size_t n, uint64_t seed, cudaStream_t stream) {
    if (n == 0) return;
    k_fill_fr<<<ceil_div(n, 256), 256, 0, stream>>>(out, n, seed);
    PM_LAUNCH_CHECK();
}
}  // namespace pm
