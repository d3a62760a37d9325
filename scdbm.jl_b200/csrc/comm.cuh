// NCCL binding for the data-parallel exchange steps of the path: one allreduce(sum) of the gradient
// block per PCD step, one scalar allreduce(max) per mean-field iteration (the reference's global stopping rule,
// src/dbmtraining.jl:194-225), and the scalar reductions of logmeanexp over particle shards (src/evaluating.jl:935-943).
//
// libnccl.so.2 is resolved at run time (dlopen): a process that already carries NCCL (PyTorch bundles its own copy
// under the same SONAME) keeps exactly one copy; a Julia process picks up the system library.  Only the types of
// <nccl.h> are used at compile time, nothing is linked.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace scdbm {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load(std::string& err) {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            err = std::string("cannot load libnccl.so.2: ") + dlerror();
            return false;
        }
        auto sym = [&](const char* name) {
            void* f = dlsym(handle, name);
            if (!f && err.empty()) err = std::string("libnccl is missing ") + name;
            return f;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!err.empty()) {
            handle = nullptr;
            return false;
        }
        return true;
    }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace scdbm
