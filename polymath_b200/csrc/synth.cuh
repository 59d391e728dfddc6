// Synthetic device-resident inputs for the kernel sweep (bench only; never used by prove/setup).
#pragma once
#include "common.cuh"
#include "field.cuh"
#include "ec.cuh"

namespace pm {
// out[i] = pseudo-random Montgomery-form Fr derived from (seed, i); limbs < 2^254 so always reduced.
void launch_fill_fr(Fr* out, size_t n, uint64_t seed, cudaStream_t stream);
// Overwrites uniform MSM inputs with the skewed distribution of SURVEY.md 8d: 89 % one repeated scalar, 10 % zero,
// 1 % uniform; 1 % of the bases at infinity.
void launch_skew_msm_inputs(Fr* scalars, G1Affine* bases, size_t n, uint64_t seed, cudaStream_t stream);
}  // namespace pm
