// Synthetic device-resident inputs for the kernel sweep (bench only; never used by prove/setup).
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace pm {
// out[i] = pseudo-random Montgomery-form Fr derived from (seed, i); limbs < 2^254 so always reduced.
void launch_fill_fr(Fr* out, size_t n, uint64_t seed, cudaStream_t stream);
}  // namespace pm
