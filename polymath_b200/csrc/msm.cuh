// K4 — G1 Pippenger MSM engine (interface).  See msm.cu.
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace pm {

struct MsmConfig {
    int c = 0;          // window bits (0 = choose from n)
    int heavy = 0;      // heavy-bucket threshold (0 = auto)
};

// Reusable workspace + launch sequence.  One engine per context / stream.
class MsmEngine {
public:
    // out_xyzz[0] = sum_{i<n} scalars[i] * bases[i].
    //  bases   : device, 96-byte affine points, Montgomery limbs, (0,0) = infinity
    //  scalars : device, Fr in Montgomery form (arkworks in-memory form)
    // All work is enqueued on `stream`; nothing is synchronised.
    void run(const G1Affine* bases, const Fr* scalars, size_t n, G1XYZZ* out_xyzz, cudaStream_t stream,
             MsmConfig cfg = {});
    static int choose_window(size_t n);
    size_t launches = 0;   // kernels launched so far (bench accounting)
    // timing hook for bench.py's roofline: CUDA events around the bucket-accumulation kernel of the last run
    cudaEvent_t ev_acc_begin = nullptr, ev_acc_end = nullptr;
    bool time_accumulate = false;
    ~MsmEngine();

private:
    DevBuf counts_, offsets_, cursors_, sorted_, buckets_, segs_, winsums_, heavy_list_, heavy_count_;
};

// out_affine = canonical affine image of sum of `k` XYZZ partials (k small); single-thread kernel.
void launch_xyzz_sum_to_affine(const G1XYZZ* parts, int k, G1Affine* out_affine, cudaStream_t stream);

}  // namespace pm
