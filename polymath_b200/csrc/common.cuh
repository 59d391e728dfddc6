// Shared host-side plumbing for the C-ABI library: error reporting, grow-only device
// buffers, launch helpers.  No torch types; plain CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/polymath_b200.h"   // PM_OK / PM_ERR_* status codes

namespace pm {

void set_last_error(const std::string& msg);

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define PM_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _buf[512];                                                                    \
            snprintf(_buf, sizeof _buf, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,   \
                     cudaGetErrorString(_e));                                                  \
            throw ::pm::CudaError(_buf);                                                       \
        }                                                                                      \
    } while (0)

#define PM_LAUNCH_CHECK() PM_CUDA(cudaGetLastError())

// Grow-only device allocation; contents are not preserved across growth.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void* reserve(size_t bytes) {
        if (bytes > cap) {
            if (p) PM_CUDA(cudaFree(p));
            p = nullptr; cap = 0;
            PM_CUDA(cudaMalloc(&p, bytes));
            cap = bytes;
        }
        return p;
    }
    template <class T> T* as(size_t count) { return static_cast<T*>(reserve(count * sizeof(T))); }
    template <class T> T* get() const { return static_cast<T*>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

int sm_count();

}  // namespace pm
