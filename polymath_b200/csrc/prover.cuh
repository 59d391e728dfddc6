// Device-resident prover context: proving key + per-proof work buffers + the three phases.
#pragma once
#include <vector>
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"
#include "poly_kernels.cuh"
#include "runtime.cuh"

namespace pm {

// CSR (and optionally CSC) of one R1CS matrix, deduplicated on the host, resident on the device.
struct DevMatrix {
    DevBuf row_ptr, col, val;          // CSR
    DevBuf col_ptr, row, cval;         // CSC (setup only)
    uint64_t nnz = 0;
    DevCsr csr() const { return {row_ptr.get<uint32_t>(), col.get<uint32_t>(), val.get<Fr>()}; }
};

struct ProverCtx {
    // dimensions
    uint64_t m0 = 0, mw = 0, nr = 0, n = 0, sigma = 0, cols = 0;
    int log_n = 0;
    // MSM sharding: this process holds the points g = k*world + rank of both base
    // arrays; the polynomial work is replicated.  world == 1: the whole key.
    int rank = 0, world = 1;
    // compressed key vectors (point_stride 48): subgroup-checked like `deserialize_compressed` unless the caller opts
    // out (`deserialize_compressed_unchecked`); the on-curve and encoding checks always run
    bool validate_key = true;
    uint64_t local_count(uint64_t total) const { return total > (uint64_t)rank ? (total - rank + world - 1) / world : 0; }
    // Fixed-base tables for the big MSMs: each base array is [levels][stride] with level l = 2^(c*l) * P
    // (msm.cuh MsmConfig); levels == 1 means no precomputation.  Chosen from the size and free memory.
    struct MsmPlan { int c = 0; int levels = 1; size_t stride = 1; };
    MsmPlan plan_c, plan_d;
    void plan_tables();    // before the base arrays are allocated
    void build_tables();   // after level 0 of both arrays is filled
    MsmConfig cfg_c() const { MsmConfig m; m.c = plan_c.c; m.levels = plan_c.levels; m.level_stride = plan_c.stride; return m; }
    MsmConfig cfg_d() const { MsmConfig m; m.c = plan_d.c; m.levels = plan_d.levels; m.level_stride = plan_d.stride; return m; }
    // key
    DevMatrix A, B, C;
    DevBuf bases_c;   // [x_powers (n+1) | x_powers_y_alpha (3) | x_powers_y_gamma (2) | zh (n-1) | lcs (cols-m0)]
    DevBuf bases_d;   // x_powers_y_gamma_z (10n + 23)
    uint64_t len_c() const { return (n + 1) + 3 + 2 + (n - 1) + (cols - m0); }
    uint64_t len_d() const { return 2 * (n - 1) + 8 * sigma + 1; }
    // per-proof buffers
    DevBuf ztail, u, w, wu, u2, scal_a, scal_c, q, chunk_vals, carries, small;
    DevBuf status, acc;
    void* host_stage = nullptr;   // pinned staging for results
    int phase = 0;                // 0 idle, 1 after phase 1, 2 after phase 2
    bool assignment_set = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double phase_ms[3] = {0, 0, 0};

    // Optional NCCL communicator of the sharded flow (nccl_dyn.cuh): with it the per-rank partial sums are
    // all-gathered on the device, stream-ordered behind the MSM kernels, and every rank finishes on its host.
    void* nccl_comm = nullptr;
    DevBuf gathered;              // [rank][sums] receive buffers of the all-gathers
    void* host_gather = nullptr;  // pinned copy of `gathered`
    size_t host_gather_bytes = 0;
    void attach_nccl(const char* libnccl_path, const uint8_t id[128]);
    // Transform of a replicated array through the SHARDED NTT: every rank transforms its interleaved
    // slice (local (N/G)-point NTT, twiddle/pack, ONE all-to-all over NCCL, G-point combine) and the slices are
    // all-gathered back into the full array.  Used for domains of at least 2^sharded_ntt_min_log elements when a
    // communicator is attached; smaller ones run replicated (the exchange latency would exceed the saving).
    DevBuf ntt_loc, ntt_send, ntt_recv, ntt_gather;
    int sharded_ntt_min_log = 22;
    void ntt_full(Fr* data, int log_size, bool inverse, cudaStream_t s);
    void phase1_collective(const uint8_t* ra, uint8_t* a_out, uint8_t* c_out);
    void phase3_collective(const uint8_t* x2, const uint8_t* c_at_x1, uint8_t* d_out);

    ~ProverCtx();
    void allocate_work();
    void upload_matrix(DevMatrix& dst, const uint64_t* row_ptr, const uint32_t* col, const uint8_t* val, bool want_csc);
    void set_assignment(const uint8_t* x, const uint8_t* w);
    // partial = this rank's XYZZ sums (192 B each): [a-side, c-side] for phase 1, [d] for phase 3
    void phase1_partial(const uint8_t* ra, uint8_t* partials_out);
    void phase1_finish(const uint8_t* gathered, int count, uint8_t* a_out, uint8_t* c_out);
    void phase2(const uint8_t* x1, const uint8_t* y1_alpha, uint8_t* a_at_x1_out);
    void phase3_partial(const uint8_t* x2, const uint8_t* c_at_x1, uint8_t* partial_out);
    void phase3_finish(const uint8_t* gathered, int count, uint8_t* d_out);
    NumeratorSrc numerator_src() const;

private:
    struct Phase1Shapes { MsmEngine::Shape sa, sc; };
    // device work of a phase up to and including its MSM launches, joined on the main stream; nothing synchronised
    Phase1Shapes phase1_enqueue(const uint8_t* ra);
    MsmEngine::Shape phase3_enqueue(const uint8_t* x2, const uint8_t* c_at_x1);
    uint8_t* gather_stage(size_t bytes);
};

// setup.cu: fill bases_c / bases_d of a context whose matrices are uploaded (with CSC), and the G2 images.
void run_setup(ProverCtx& ctx, const uint8_t* x, const uint8_t* z, uint8_t* x_g2, uint8_t* z_g2);

}  // namespace pm
