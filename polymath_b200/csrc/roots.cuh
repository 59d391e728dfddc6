// Roots of unity of Fr on the device (arkworks FrConfig: GENERATOR = 7, TWO_ADICITY = 32).
#pragma once
#include "field.cuh"

namespace pm {

// 2^32-th primitive root of unity of Fr (7^((r-1)/2^32)) and its inverse, Montgomery form
static __device__ __constant__ uint32_t ROOT32[8] = {0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u,
                                             0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau};
static __device__ __constant__ uint32_t ROOT32_INV[8] = {0xdcf3219au, 0x4256481au, 0x96b6cad3u, 0x45f37b7fu,
                                                 0x5f7a3b27u, 0xf9c3f1d7u, 0x658afd43u, 0x2d2fc049u};

// w_{2^log_size} = `Radix2EvaluationDomain::group_gen` (or its inverse)
__device__ __forceinline__ Fr root_of_unity(int log_size, bool inverse) {
    Fr w;
#pragma unroll
    for (int i = 0; i < 8; i++) w.v[i] = inverse ? ROOT32_INV[i] : ROOT32[i];
    for (int i = log_size; i < 32; i++) w = w.sqr();
    return w;
}


}  // namespace pm
