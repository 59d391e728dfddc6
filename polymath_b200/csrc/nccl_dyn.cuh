// NCCL bound at run time (dlopen): the library has no link-time dependency on NCCL, the caller names the
// libnccl.so.2 its process already uses (e.g. the one bundled with torch) so that one NCCL serves the process.
// Only the handful of entry points of the sharded flow are bound (all-gather of the per-GPU
// partial sums of an MSM split by point range; group addition is not an NCCL reduction).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>
#include "common.cuh"

namespace pm {

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    static constexpr int kUint8 = 1;            // ncclUint8

    void load(const char* path) {
        if (lib) return;
        const char* p = (path && *path) ? path : "libnccl.so.2";
        void* h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw CudaError(std::string("cannot load NCCL from ") + p + ": " + dlerror());
        auto sym = [&](const char* name) {
            void* s = dlsym(h, name);
            if (!s) throw CudaError(std::string("NCCL symbol missing: ") + name);
            return s;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
        Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        lib = h;
    }
    void check(int rc, const char* what) const {
        if (rc != 0) throw CudaError(std::string(what) + " failed: " + (GetErrorString ? GetErrorString(rc) : "NCCL error"));
    }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace pm
