// nccl_dyn.hpp — NCCL resolved at run time (dlopen), so that libmcphylo_b200.so itself has no link-time
// dependency beyond libc/libdl: a single-GPU user never needs NCCL installed, and a multi-GPU context
// (mcp_create_multi) that asks for the NCCL reduction fails with a clear message when it is absent.
// Only the handful of entry points the [logL, gradient] all-reduce needs are bound; the declarations
// restate the public NCCL 2.x C API (stable since 2.0: opaque communicator, 128-byte unique id).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <mutex>
#include <string>

namespace mcpnccl {

typedef struct ncclComm* comm_t;
typedef int result_t;                       // ncclResult_t; 0 = ncclSuccess
struct unique_id { char internal[128]; };   // ncclUniqueId
constexpr int kFloat64 = 8;                 // ncclFloat64 / ncclDouble
constexpr int kSum = 0;                     // ncclSum

struct Api {
    void* handle = nullptr;
    std::string where;                      // soname that resolved
    result_t (*GetVersion)(int*) = nullptr;
    result_t (*GetUniqueId)(unique_id*) = nullptr;
    result_t (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    result_t (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    result_t (*CommDestroy)(comm_t) = nullptr;
    result_t (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    result_t (*GroupStart)() = nullptr;
    result_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(result_t) = nullptr;
    bool ok() const { return handle != nullptr; }
};

// Returns the process-wide binding; api.ok() is false (and `why` says so) when no NCCL could be loaded.
// MCPHYLO_B200_NCCL names a specific library; otherwise the soname libnccl.so.2 is searched the usual
// way (a libnccl already loaded into the process, e.g. the one a Python host's torch brought, is reused).
inline const Api& api(std::string* why = nullptr) {
    static Api a;
    static std::string err;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[3] = {getenv("MCPHYLO_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* n : names) {
            if (!n || !*n) continue;
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) { a.where = n; break; }
            err = dlerror();
        }
        if (!h) return;
        bool all = true;
        auto bind = [&](auto& fn, const char* sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, sym));
            if (!fn) { all = false; err = std::string("symbol ") + sym + " not found in " + a.where; }
        };
        bind(a.GetVersion, "ncclGetVersion");
        bind(a.GetUniqueId, "ncclGetUniqueId");
        bind(a.CommInitAll, "ncclCommInitAll");
        bind(a.CommInitRank, "ncclCommInitRank");
        bind(a.CommDestroy, "ncclCommDestroy");
        bind(a.AllReduce, "ncclAllReduce");
        bind(a.GroupStart, "ncclGroupStart");
        bind(a.GroupEnd, "ncclGroupEnd");
        bind(a.GetErrorString, "ncclGetErrorString");
        if (all) a.handle = h;
    });
    if (why) *why = err;
    return a;
}

}  // namespace mcpnccl
