// Micro-benchmark (not part of the product): random 512 B row reads from a peer GPU's memory owned
// by ANOTHER PROCESS, mapped (a) with the legacy cudaIpc handles, (b) with the VMM API
// (cuMemCreate + POSIX fd + cuMemMap).  Two processes (fork), GPU 0 reads GPU 1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ipcbench scripts/ipcbench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
#define CU(x) do { CUresult e_ = (x); if (e_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(e_, &s_); printf("%s: %s\n", #x, s_); exit(1);} } while (0)

__global__ void __launch_bounds__(256) gather_rows(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                   const unsigned *__restrict__ idx, unsigned n) {
    const unsigned lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (size_t)gridDim.x * 8;
    for (size_t base = w * 32; base < n; base += nw * 32) {
        const unsigned mi = base + lane < n ? idx[base + lane] : 0;
        const int rows = (int)min((size_t)32, n - base);
        for (int g0 = 0; g0 < rows; g0 += 4) {
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                unsigned s = __shfl_sync(~0u, mi, (g0 + r) & 31);
                if (g0 + r < rows) v[r] = src[(size_t)s * 32 + lane];
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (g0 + r < rows) dst[(base + g0 + r) * 32 + lane] = v[r];
        }
    }
}

// same loads through the coherent path (what a kernel that also writes the table gets)
__global__ void __launch_bounds__(256) gather_rows_coh(float4 *src, float4 *dst, const unsigned *idx, unsigned n, int never) {
    const unsigned lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (size_t)gridDim.x * 8;
    if (never) src[lane] = make_float4(0, 0, 0, 0);
    for (size_t base = w * 32; base < n; base += nw * 32) {
        const unsigned mi = base + lane < n ? idx[base + lane] : 0;
        const int rows = (int)min((size_t)32, n - base);
        for (int g0 = 0; g0 < rows; g0 += 4) {
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                unsigned s = __shfl_sync(~0u, mi, (g0 + r) & 31);
                if (g0 + r < rows) v[r] = src[(size_t)s * 32 + lane];
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (g0 + r < rows) dst[(base + g0 + r) * 32 + lane] = v[r];
        }
    }
}

static void send_fd(int sock, int fd) {
    char buf[CMSG_SPACE(sizeof(int))]; memset(buf, 0, sizeof(buf));
    char dummy = 'x'; iovec io{&dummy, 1};
    msghdr msg{}; msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = buf; msg.msg_controllen = sizeof(buf);
    cmsghdr *c = CMSG_FIRSTHDR(&msg); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    if (sendmsg(sock, &msg, 0) < 0) { perror("sendmsg"); exit(1); }
}
static int recv_fd(int sock) {
    char buf[CMSG_SPACE(sizeof(int))]; memset(buf, 0, sizeof(buf));
    char dummy; iovec io{&dummy, 1};
    msghdr msg{}; msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = buf; msg.msg_controllen = sizeof(buf);
    if (recvmsg(sock, &msg, 0) < 0) { perror("recvmsg"); exit(1); }
    int fd; memcpy(&fd, CMSG_DATA(CMSG_FIRSTHDR(&msg)), sizeof(int)); return fd;
}

int main(int argc, char **argv) {
    const size_t BYTES = argc > 1 ? strtoull(argv[1], 0, 10) : 8643219968ull, ROWS = BYTES / 512; // default: the shard size of the 2-GPU bench
    printf("bytes %zu (%.3f x 2 MiB)\n", BYTES, BYTES / 2097152.0);
    const unsigned N = 52000;
    int sp[2]; socketpair(AF_UNIX, SOCK_STREAM, 0, sp);
    pid_t pid = fork();
    if (pid == 0) { // owner process: GPU 1
        CK(cudaSetDevice(1)); CK(cudaFree(0));
        void *legacy; CK(cudaMalloc(&legacy, BYTES)); CK(cudaMemset(legacy, 0, BYTES));
        cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, legacy));
        write(sp[1], &h, sizeof(h));
        CUmemAllocationProp prop{}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = 1; prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t gran; CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        printf("owner: granularity %zu\n", gran);
        const size_t VB = (BYTES + gran - 1) / gran * gran;
        CUmemGenericAllocationHandle mh; CU(cuMemCreate(&mh, VB, &prop, 0));
        int fd; CU(cuMemExportToShareableHandle(&fd, mh, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        send_fd(sp[1], fd);
        char done; read(sp[1], &done, 1);
        return 0;
    }
    CK(cudaSetDevice(0)); CK(cudaFree(0));
    float4 *dst; unsigned *idx;
    CK(cudaMalloc(&dst, (size_t)N * 512)); CK(cudaMalloc(&idx, N * 4 * 4));
    std::mt19937_64 rng(1); std::vector<unsigned> hidx(N * 4); for (auto &x : hidx) x = rng() % ROWS;
    CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, const float4 *src) {
        for (int i = 0; i < 3; i++) gather_rows<<<832, 256>>>(src, dst, idx + (i % 4) * N, N);
        CK(cudaDeviceSynchronize());
        float tot = 0;
        for (int i = 0; i < 10; i++) {
            cudaEventRecord(e0); gather_rows<<<832, 256>>>(src, dst, idx + (i % 4) * N, N); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
        }
        printf("%-36s avg %8.1f us  %7.1f GB/s\n", name, tot / 10 * 1e3, (double)N * 512 / (tot / 10 * 1e-3) / 1e9);
    };
    cudaIpcMemHandle_t h; read(sp[0], &h, sizeof(h));
    void *legacy; CK(cudaIpcOpenMemHandle(&legacy, h, cudaIpcMemLazyEnablePeerAccess));
    timeit("legacy cudaIpc mapping, random rows", (const float4 *)legacy);
    {
        float tot = 0;
        for (int i = 0; i < 13; i++) {
            cudaEventRecord(e0); gather_rows_coh<<<832, 256>>>((float4 *)legacy, dst, idx + (i % 4) * N, N, 0); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 3) tot += ms;
        }
        printf("%-36s avg %8.1f us\n", "legacy mapping, coherent ld.global", tot / 10 * 1e3);
        // with a big local working set in the reading process (shard + cache rows)
        void *big; CK(cudaMalloc(&big, 12ull << 30)); CK(cudaMemset(big, 0, 12ull << 30));
        timeit("legacy mapping, after 12 GB local alloc", (const float4 *)legacy);
    }
    int fd = recv_fd(sp[0]);
    CUmemGenericAllocationHandle mh; CU(cuMemImportFromShareableHandle(&mh, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    const size_t VB = (BYTES + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    CUdeviceptr va; CU(cuMemAddressReserve(&va, VB, 0, 0, 0)); CU(cuMemMap(va, VB, 0, mh, 0));
    CUmemAccessDesc acc{}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(va, VB, &acc, 1));
    timeit("VMM (cuMemMap of a POSIX fd), random rows", (const float4 *)va);
    char done = 1; write(sp[0], &done, 1);
    int st; waitpid(pid, &st, 0);
    return 0;
}
