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

constexpr int kMaxWorld = 8; // one NVSwitch domain

struct Comm {
    int rank = 0, world = 1, device = 0;
    NcclComm comm = nullptr;
    cudaStream_t stream = nullptr;
    int *scratch = nullptr; // 64 ints of device memory
    NcclApi api;
    // device-side barrier over peer-mapped memory: rank r stores its epoch into flags[r] of every
    // peer and waits until its own flags[0..world) have all reached that epoch
    u64 *flags = nullptr;                // [kMaxWorld] in this rank's memory
    u64 *peer_flags[kMaxWorld] = {};     // the same array of every rank, mapped here
    u32 *barrier_err = nullptr;          // device word: a barrier timed out
    u64 epoch = 0;                       // barriers enqueued so far (same on every rank)
};

extern Comm g_comm;
NcclApi &nccl();
void nccl_check(int rc, const char *what);

// Map `local` (a cudaMalloc allocation of this rank) into every rank of the group: peers[r] is
// rank r's allocation as seen from here (peers[rank] == local).  Collective; same call order on
// every rank.  CUDA IPC: the ranks are processes on one node, the mapping goes over NVLink.
void ipc_share(void *local, void **peers);
void ipc_unshare(void **peers);
// Enqueue a barrier of the whole group on `st` (no host synchronisation).  Collective.
void device_barrier(cudaStream_t st);

} // namespace hb
