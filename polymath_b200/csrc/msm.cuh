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
    // Enqueues sum_{i<n} scalars[i*scalar_stride + scalar_offset] * bases[i] up to the per-window sums:
    // on completion winsums_out[w] (device, >= kMaxMsmWindows entries) holds W_w and the result is
    // sum_w 2^(c*w) * W_w, finished on the host (host/g1_host.hpp: combine_windows).
    //  bases   : device, 96-byte affine points, Montgomery limbs, (0,0) = infinity
    //  scalars : device, Fr in Montgomery form (arkworks in-memory form)
    // All work is enqueued on `stream`; nothing is synchronised.
    struct Shape { int c; int nwin; };
    Shape run(const G1Affine* bases, const Fr* scalars, size_t n, G1XYZZ* winsums_out, cudaStream_t stream,
              MsmConfig cfg = {}, size_t scalar_stride = 1, size_t scalar_offset = 0);
    static int choose_window(size_t n);
    size_t launches = 0;   // kernels launched so far (bench accounting)
    // timing hook for bench.py's roofline: CUDA events around the bucket-accumulation kernel of the last run
    cudaEvent_t ev_acc_begin = nullptr, ev_acc_end = nullptr;
    bool time_accumulate = false;
    ~MsmEngine();

private:
    DevBuf counts_, offsets_, cursors_, sorted_, buckets_, segs_, heavy_list_, heavy_count_;
};

constexpr int kMaxMsmWindows = 64;

}  // namespace pm
