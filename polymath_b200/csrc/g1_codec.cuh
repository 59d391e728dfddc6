// Device-side G1 compression / decompression (zcash encoding used by ark-bls12-381).  See g1_codec.cu.
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace pm {

// out[i] = decode(in[48 i .. 48 i + 48)); invalid encodings decode to infinity and record
// *first_bad = min over failures of (index << 3 | status); initialise *first_bad to ~0ull.
// validate: also require the point to lie in the prime-order subgroup (`deserialize_compressed`); without it the
// check stops at "on the curve" (`deserialize_compressed_unchecked` checks nothing).
void launch_g1_decompress(const uint8_t* in_dev, size_t n, bool validate, G1Affine* out_dev, unsigned long long* first_bad_dev,
                          cudaStream_t stream);
void launch_g1_compress(const G1Affine* in_dev, size_t n, uint8_t* out_dev, cudaStream_t stream);
const char* g1_decode_status_name(unsigned status);

}  // namespace pm
