// K5 — fixed-base batch scalar multiplication [s_i]G and batch affine normalisation (interface).
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace pm {

class FixedBaseEngine {
public:
    // out[i] = canonical affine image of scalars[i] * G; scalars are Montgomery-form Fr on the device.
    void run(const Fr* scalars, size_t n, G1Affine* out, cudaStream_t stream);
    size_t launches = 0;

private:
    void ensure_table(cudaStream_t stream);
    DevBuf table_;      // [windows][2^wbits - 1] affine multiples of G
    DevBuf xyzz_;       // per-call scratch
    bool built_ = false;
};

// out[i] = affine(in[i]) with one field inversion per `kBatch` points (Montgomery's trick).
void launch_batch_to_affine(const G1XYZZ* in, size_t n, G1Affine* out, cudaStream_t stream);

}  // namespace pm
