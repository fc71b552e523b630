// nccl_dyn.h -- NCCL bound at run time (dlopen), so that libakaze_b200.so loads on hosts without NCCL and
// shares the NCCL a host process may already have loaded (torch ships its own libnccl.so.2).
//
// Only the multi-GPU matcher (akz_context_comm_init*, akz_match_top2_sharded*) needs it: the per-shard top-2
// records of the brute-force matcher are all-gathered over NVLink and merged (the exchange
// step that follows akaze/src/ops/feature_matching.rs:37-50 when the database is sharded).
// The handful of NCCL 2.x ABI items used here (stable since 2.0) are declared below instead of including nccl.h.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>
#include <string>

typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;    // ncclSuccess == 0
typedef int ncclDataType_t;  // ncclInt8 = 0, ncclUint8 = 1

namespace nccl_dyn {

constexpr ncclDataType_t kUint8 = 1;

struct Api {
    void* handle = nullptr;
    std::string error;  // why loading failed ("" when loaded)
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return handle != nullptr; }
};

inline Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[4] = {getenv("AKZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; i < 3 && !a.handle; i++)
            if (names[i] && names[i][0]) a.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) {
            const char* e = dlerror();
            a.error = std::string("NCCL not found (set AKZ_NCCL_LIB to libnccl.so.2): ") + (e ? e : "dlopen failed");
            return;
        }
        bool all = true;
        auto sym = [&](const char* n) {
            void* p = dlsym(a.handle, n);
            if (!p) all = false;
            return p;
        };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        if (!all) {
            a.error = "the NCCL library lacks a required symbol";
            a.handle = nullptr;
        }
    });
    return a;
}

}  // namespace nccl_dyn
