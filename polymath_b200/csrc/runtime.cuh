// Process-wide device runtime shared by the C-ABI entry points: one stream, one instance of
// each engine (grow-only workspaces).  One process drives one GPU (the CUDA device current
// at first use).
#pragma once
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"
#include "fixed_base.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace pm {

struct StatusError : std::runtime_error {
    int code;
    StatusError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct Runtime {
    Runtime();
    cudaStream_t stream = nullptr;
    // side stream + second MSM workspace: independent MSMs of one phase overlap their latency-bound tails
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    NttEngine ntt;
    MsmEngine msm, msm2;
    FixedBaseEngine fixed_base;
    uint64_t extra_launches = 0;
    uint64_t total_launches() const;
};

Runtime& runtime();
// All contexts of a process share one stream and one set of engine workspaces, so device work of
// different contexts is serialised: every C-ABI entry point holds this lock for its duration.
std::recursive_mutex& api_mutex();

template <class F>
int guarded(F&& f) {
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    try {
        f();
        return PM_OK;
    } catch (const StatusError& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const CudaError& e) {
        set_last_error(e.what());
        return PM_ERR_CUDA;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return PM_ERR_CUDA;
    }
}

void pack_points_host(const uint8_t* src, size_t stride, size_t n, std::vector<uint8_t>& dst);

// out[0] = in[0]^-1 (single thread)
void launch_fr_inverse(const Fr* in, Fr* out, cudaStream_t stream);

}  // namespace pm
