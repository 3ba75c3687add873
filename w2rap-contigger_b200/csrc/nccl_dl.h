// nccl_dl.h — NCCL bound at run time (dlopen), so that the single-GPU path has no link dependency on it.
// Only the handful of entry points the sharded step 2 needs.  Types mirror nccl.h (NCCL 2.x ABI).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <mutex>

namespace w2r {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* error = nullptr;

    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            const char* names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
            if (!api.lib) { api.error = "libnccl.so.2 not found (multi-GPU step 2 needs NCCL)"; return; }
#define W2R_NCCL_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym)); if (!api.field) { api.error = "missing NCCL symbol " sym; return; }
            W2R_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
            W2R_NCCL_SYM(CommInitRank, "ncclCommInitRank")
            W2R_NCCL_SYM(CommDestroy, "ncclCommDestroy")
            W2R_NCCL_SYM(AllReduce, "ncclAllReduce")
            W2R_NCCL_SYM(Broadcast, "ncclBroadcast")
            W2R_NCCL_SYM(AllGather, "ncclAllGather")
            W2R_NCCL_SYM(Send, "ncclSend")
            W2R_NCCL_SYM(Recv, "ncclRecv")
            W2R_NCCL_SYM(GroupStart, "ncclGroupStart")
            W2R_NCCL_SYM(GroupEnd, "ncclGroupEnd")
            W2R_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef W2R_NCCL_SYM
        });
        return api;
    }
};

}  // namespace w2r
