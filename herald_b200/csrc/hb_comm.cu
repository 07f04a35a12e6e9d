// Multi-GPU group: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces ps-lite's worker<->server transport (ZMQ / ibverbs vans) for the embedding path.
// NCCL is bound at run time with dlopen so that a process that already carries a libnccl
// (PyTorch's bundled one in the bench harness) shares that copy instead of loading a second.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "hb_comm.cuh"

namespace hb {

Comm g_comm;

namespace {

std::once_flag g_load_once;
std::string g_load_error;

template <typename F>
bool bind(void *lib, const char *name, F &fn) {
    fn = reinterpret_cast<F>(dlsym(lib, name));
    return fn != nullptr;
}

void load_nccl() {
    std::vector<std::string> names;
    if (const char *env = getenv("HERALD_NCCL_LIB"))
        names.push_back(env);
    names.push_back("libnccl.so.2"); // already-loaded copy (e.g. torch's) or the system one
    names.push_back("libnccl.so");
    void *lib = nullptr;
    for (auto &n : names) {
        lib = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (lib)
            break;
    }
    if (!lib) {
        g_load_error = std::string("cannot load libnccl: ") + dlerror();
        return;
    }
    NcclApi &a = g_comm.api;
    bool ok = bind(lib, "ncclGetUniqueId", a.GetUniqueId) &&
              bind(lib, "ncclCommInitRank", a.CommInitRank) &&
              bind(lib, "ncclCommDestroy", a.CommDestroy) &&
              bind(lib, "ncclGetErrorString", a.GetErrorString) &&
              bind(lib, "ncclGroupStart", a.GroupStart) && bind(lib, "ncclGroupEnd", a.GroupEnd) &&
              bind(lib, "ncclSend", a.Send) && bind(lib, "ncclRecv", a.Recv) &&
              bind(lib, "ncclAllReduce", a.AllReduce) && bind(lib, "ncclAllGather", a.AllGather);
    if (!ok)
        g_load_error = "libnccl lacks a required symbol";
    else
        a.loaded = true;
}

} // namespace

void nccl_check(int rc, const char *what) {
    if (rc != 0) {
        const char *msg = g_comm.api.GetErrorString ? g_comm.api.GetErrorString(rc) : "?";
        throw Error(std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + msg + ")");
    }
}

NcclApi &nccl() {
    std::call_once(g_load_once, load_nccl);
    if (!g_comm.api.loaded)
        throw Error(g_load_error.empty() ? "NCCL not available" : g_load_error);
    return g_comm.api;
}

void ipc_share(void *local, void **peers) {
    const int world = g_comm.world, rank = g_comm.rank;
    for (int r = 0; r < kMaxWorld; r++)
        peers[r] = nullptr;
    peers[rank] = local;
    if (world == 1)
        return;
    HB_CHECK(g_comm.comm, "group not initialised");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t mine;
    HB_CUDA(cudaIpcGetMemHandle(&mine, local));
    char *dev = nullptr;
    HB_CUDA(cudaMalloc((void **)&dev, 64 * (size_t)(world + 1)));
    HB_CUDA(cudaMemcpy(dev, &mine, 64, cudaMemcpyHostToDevice));
    nccl_check(nccl().AllGather(dev, dev + 64, 64, kNcclInt8, g_comm.comm, g_comm.stream),
               "ncclAllGather(ipc handles)");
    HB_CUDA(cudaStreamSynchronize(g_comm.stream));
    std::vector<cudaIpcMemHandle_t> all(world);
    HB_CUDA(cudaMemcpy(all.data(), dev + 64, 64 * (size_t)world, cudaMemcpyDeviceToHost));
    cudaFree(dev);
    for (int r = 0; r < world; r++) {
        if (r == rank)
            continue;
        HB_CUDA(cudaIpcOpenMemHandle(&peers[r], all[r], cudaIpcMemLazyEnablePeerAccess));
    }
}

void ipc_unshare(void **peers) {
    for (int r = 0; r < kMaxWorld; r++) {
        if (peers[r] && r != g_comm.rank)
            cudaIpcCloseMemHandle(peers[r]);
        peers[r] = nullptr;
    }
}

namespace {
struct PeerFlags {
    u64 *peer[kMaxWorld];
    u64 *local;
};

__global__ void peer_barrier_kernel(PeerFlags pf, int rank, int world, u64 epoch, u32 *err) {
    pdl_enter();
    const int r = threadIdx.x;
    if (r >= world)
        return;
    __threadfence_system(); // everything this GPU wrote to peer memory before the barrier
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pf.peer[r] + rank), "l"(epoch) : "memory");
    const long long t0 = clock64();
    u64 seen = 0;
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(pf.local + r) : "memory");
        if (seen >= epoch)
            break;
        if (clock64() - t0 > 60000000000ll) { // ~30 s: a peer is gone; fail the call, do not hang
            *err = 1;
            break;
        }
    }
}
} // namespace

void device_barrier(cudaStream_t st) {
    if (g_comm.world == 1)
        return;
    PeerFlags pf;
    for (int r = 0; r < kMaxWorld; r++)
        pf.peer[r] = g_comm.peer_flags[r];
    pf.local = g_comm.flags;
    g_comm.epoch++;
    HB_LAUNCH(peer_barrier_kernel, 1, 32, 0, st, pf, g_comm.rank, g_comm.world, g_comm.epoch,
                                          g_comm.barrier_err);
    HB_LAUNCHED();
}

} // namespace hb

using namespace hb;

extern "C" {

int hb_comm_unique_id(void *id128) {
    HB_API_BEGIN();
    NcclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, sizeof(id));
    HB_API_END();
}

int hb_comm_init(const void *id128, int rank, int world, int device) {
    HB_API_BEGIN();
    HB_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    HB_CHECK(!g_comm.comm, "group already initialised");
    HB_CUDA(cudaSetDevice(device));
    g_comm.rank = rank;
    g_comm.world = world;
    g_comm.device = device;
    if (world > 1) {
        NcclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        nccl_check(nccl().CommInitRank(&g_comm.comm, world, id, rank), "ncclCommInitRank");
        HB_CUDA(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
        HB_CUDA(cudaMalloc((void **)&g_comm.scratch, 256));
        HB_CUDA(cudaMemset(g_comm.scratch, 0, 256));
        HB_CHECK(world <= kMaxWorld, "at most 8 ranks (one NVSwitch domain)");
        HB_CUDA(cudaMalloc((void **)&g_comm.flags, 256));
        HB_CUDA(cudaMemset(g_comm.flags, 0, 256));
        g_comm.barrier_err = reinterpret_cast<u32 *>(g_comm.flags + 16);
        g_comm.epoch = 0;
        ipc_share(g_comm.flags, reinterpret_cast<void **>(g_comm.peer_flags));
    }
    HB_API_END();
}

int hb_comm_rank(int *rank, int *world) {
    if (rank)
        *rank = g_comm.rank;
    if (world)
        *world = g_comm.world;
    return 0;
}

int hb_comm_barrier(void) {
    HB_API_BEGIN();
    if (g_comm.world > 1) {
        HB_CHECK(g_comm.comm, "group not initialised");
        nccl_check(nccl().AllReduce(g_comm.scratch, g_comm.scratch + 16, 1, kNcclInt32, kNcclSum,
                                    g_comm.comm, g_comm.stream),
                   "ncclAllReduce(barrier)");
        HB_CUDA(cudaStreamSynchronize(g_comm.stream));
    }
    HB_API_END();
}

int hb_comm_finalize(void) {
    HB_API_BEGIN();
    if (g_comm.comm) {
        cudaDeviceSynchronize();
        hb_comm_barrier(); // nobody still reads or writes a peer's memory
        ipc_unshare(reinterpret_cast<void **>(g_comm.peer_flags));
        cudaFree(g_comm.flags);
        g_comm.flags = nullptr;
        nccl().CommDestroy(g_comm.comm);
        g_comm.comm = nullptr;
        cudaStreamDestroy(g_comm.stream);
        g_comm.stream = nullptr;
        cudaFree(g_comm.scratch);
        g_comm.scratch = nullptr;
    }
    g_comm.rank = 0;
    g_comm.world = 1;
    HB_API_END();
}

} // extern "C"
