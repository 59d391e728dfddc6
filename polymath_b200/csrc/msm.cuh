// K4 — G1 Pippenger MSM engine (interface).  See msm.cu.
#pragma once
#include <map>
#include <tuple>
#include <utility>
#include "common.cuh"
#include "ec.cuh"

namespace pm {

struct MsmConfig {
    int c = 0;          // window bits (0 = choose from n)
    int heavy = 0;      // heavy-bucket threshold (0 = auto)
    // Precomputed levels (fixed bases): the base array is [levels][level_stride] with
    // level l holding 2^(c*l) * P_i.  Window w = g*levels + l then uses level l and the bucket set of
    // group g, so `levels` windows share one bucket set and the final Horner only spans the groups.
    int levels = 1;
    size_t level_stride = 0;
    // Batched-affine pair rounds before the XYZZ walk (msm.cu: k_pairs_*): -1 = choose from the mean run
    // length, 0 = none, k = exactly k rounds.  rounds_bias shifts the automatic choice (an MSM that runs beside
    // another one on a second stream hides the fixed per-round latency and can afford one more round).
    int rounds = -1;
    int rounds_bias = 0;
    // Bucket sets handled per pass (0 = as many as address space and free device memory allow); test hook.
    int sets_per_pass = 0;
};

// Reusable workspace + launch sequence.  One engine per context / stream.
class MsmEngine {
public:
    // Enqueues sum_{i<n} scalars[i*scalar_stride + scalar_offset] * bases[i] up to the per-window sums:
    // on completion winsums_out (device, >= kMaxMsmSums entries) holds the partial sums described by Shape.
    //  bases   : device, 96-byte affine points, Montgomery limbs, (0,0) = infinity
    //  scalars : device, Fr in Montgomery form (arkworks in-memory form)
    // All work is enqueued on `stream`; nothing is synchronised.
    // Result layout: winsums_out[g * nsum + s] = S_{g,s};  result = sum_g 2^(c*g) * sum_s 2^(shift[s]) * S_{g,s}
    // (g < nwin bucket sets; s < nsum partial sums of the hierarchical bucket reduction; c already includes
    // the precomputed-levels factor).  Finished on the host: host/g1_host.hpp combine_shifted().
    struct Shape {
        int c = 1, nwin = 1, nsum = 1;
        uint8_t shift[32] = {0};
        int count() const { return nwin * nsum; }
    };
    Shape run(const G1Affine* bases, const Fr* scalars, size_t n, G1XYZZ* winsums_out, cudaStream_t stream,
              MsmConfig cfg = {}, size_t scalar_stride = 1, size_t scalar_offset = 0);
    static int choose_window(size_t n);
    // process-wide default for MsmConfig::rounds when a call leaves it automatic (tests, tuning sweeps)
    static void set_tuning(int rounds);
    size_t launches = 0;   // kernels launched so far (bench accounting)
    // timing hook for bench.py's roofline: CUDA events around the bucket-accumulation kernel of the last run
    cudaEvent_t ev_acc_begin = nullptr, ev_acc_end = nullptr;
    // ... and around the first-round k_pairs_backward launch inside it (the heaviest single kernel), with the
    // geometry of the last run: pair rounds used, slot pairs of the first round (upper bound), entries (n * windows)
    cudaEvent_t ev_bwd_begin = nullptr, ev_bwd_end = nullptr;
    int last_rounds = 0;
    size_t last_entries = 0;
    bool time_accumulate = false;
    ~MsmEngine();

private:
    // (n, c, levels, forced rounds, rounds bias, forced sets) -> (pair rounds, bucket sets per pass)
    using PlanKey = std::tuple<size_t, int, int, int, int, int>;
    std::map<PlanKey, std::pair<int, int>> plans_;
    DevBuf counts_, offsets_, cursors_, sorted_, buckets_, segs_, heavy_list_, heavy_count_, order_;
    DevBuf pairs_a_, pairs_b_, prefix_, tvals_, tpre_;   // pair rounds
    void ensure_side_stream();
    cudaStream_t side_stream_ = nullptr;                   // second span of the pair rounds; level sums of the reduction
    cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;   // pair rounds
};

constexpr int kMaxMsmWindows = 64;
constexpr int kMaxMsmSums = 512;   // capacity (XYZZ records) of a winsums_out buffer: bucket sets x reduction levels

// Fill levels 1..levels-1 of a [levels][stride] base array from level 0: level l = 2^c * level (l-1).
void launch_build_levels(G1Affine* bases, size_t count, int levels, size_t stride, int c, cudaStream_t stream);

}  // namespace pm
