// K3 — Fr NTT engine (interface).  See ntt.cu.
#pragma once
#include <map>
#include <memory>
#include "common.cuh"
#include "field.cuh"

namespace pm {

class NttEngine {
public:
    // In-place transform of data[0 .. 2^log_n), natural order in and out, on `stream`.
    //   forward: out[i] = sum_j in[j] * w^(i*j), w = Radix2EvaluationDomain::group_gen
    //   inverse: out = n^-1 * (transform with w^-1)
    void run(Fr* data, int log_n, bool inverse, cudaStream_t stream);
    // Sharded transform of size 2^log_n over 2^log_g ranks, rank g holding the interleaved subsequence x[j*G + g]
    // (2^(log_n - log_g) elements).  dist_local: local transform in place + twiddle, packed destination-major into
    // `send` for ONE all-to-all (equal blocks of 2^(log_n - 2 log_g) elements); dist_combine: G-point transform across
    // the received blocks -> X[k], k = rank (mod G), at local index (k - rank)/G.  The exchange itself belongs to
    // the caller (NCCL all-to-all over NVLink, polymath_b200/sharded.py).
    void dist_local(Fr* data, Fr* send, int log_n, int log_g, uint32_t rank, bool inverse, cudaStream_t stream);
    void dist_combine(const Fr* recv, Fr* out, int log_n, int log_g, bool inverse, cudaStream_t stream);
    size_t launches = 0;
    // CUDA events around the passes of the last run (bench.py roofline)
    bool time_passes = false;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    ~NttEngine();

    struct Tables {
        DevBuf core_fwd, core_inv;     // w_{2^12}^k and its inverse, k < 2^11
        DevBuf tw_fwd[3], tw_inv[3];   // w_N^(k * 2^(b*level)), k < 2^b per level; b = 11 up to 2^22, ceil(log_n / 2) above
        DevBuf n_inv;                  // n^-1
    };

private:
    const Tables& tables(int log_n, cudaStream_t stream);
    std::map<int, std::unique_ptr<Tables>> tables_;
    DevBuf scratch_;
};

// data[i] *= g^i (coset shift) — elementwise helper used by pm_ntt_fr's coset variant
void launch_scale_by_powers(Fr* data, size_t n, const Fr* g_dev, cudaStream_t stream);

}  // namespace pm
