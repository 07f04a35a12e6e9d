// NCCL group state shared by hb_comm.cu and hb_cache.cu (multi-GPU exchange).
#pragma once

#include "hb_common.cuh"

namespace hb {

// Minimal NCCL ABI (nccl.h): opaque comm, 128-byte unique id, enum values fixed by the ABI.
struct NcclUniqueId {
    char internal[128];
};
using NcclComm = void *;
constexpr int kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclUint32 = 3, kNcclInt64 = 4,
              kNcclUint64 = 5, kNcclFloat32 = 7;
constexpr int kNcclSum = 0, kNcclMax = 2;

struct NcclApi {
    bool loaded = false;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
};

struct Comm {
    int rank = 0, world = 1, device = 0;
    NcclComm comm = nullptr;
    cudaStream_t stream = nullptr;
    int *scratch = nullptr; // 64 ints of device memory
    NcclApi api;
};

extern Comm g_comm;
NcclApi &nccl();
void nccl_check(int rc, const char *what);

} // namespace hb
