// Micro-benchmark (not part of the product): random 512 B row reads and 8 B word reads from a PEER
// GPU's memory over NVLink, as sync_kernel does for remote shards.  One process, two GPUs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/p2pbench scripts/p2pbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

template <int R>
__global__ void __launch_bounds__(256) gather_rows(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                   const unsigned *__restrict__ idx, unsigned n) {
    const unsigned lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (size_t)gridDim.x * 8;
    for (size_t base = w * 32; base < n; base += nw * 32) {
        const unsigned mi = base + lane < n ? idx[base + lane] : 0;
        const int rows = (int)min((size_t)32, n - base);
        for (int g0 = 0; g0 < rows; g0 += R) {
            float4 v[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                unsigned s = __shfl_sync(~0u, mi, (g0 + r) & 31);
                if (g0 + r < rows) v[r] = src[(size_t)s * 32 + lane];
            }
#pragma unroll
            for (int r = 0; r < R; r++)
                if (g0 + r < rows) dst[(base + g0 + r) * 32 + lane] = v[r];
        }
    }
}
__global__ void gather_words(const long long *__restrict__ src, long long *__restrict__ dst,
                             const unsigned *__restrict__ idx, unsigned n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}
// push direction: write rows into the peer's memory
__global__ void __launch_bounds__(256) scatter_rows(const float4 *__restrict__ src, float4 *__restrict__ dst, unsigned n) {
    const unsigned lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (size_t)gridDim.x * 8;
    for (size_t r = w; r < n; r += nw) dst[r * 32 + lane] = src[r * 32 + lane];
}

int main() {
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t ROWS = 16u << 20; // 8 GB of 512 B rows on the peer
    const unsigned N = 52000;
    CK(cudaSetDevice(1));
    float4 *remote; long long *rver;
    CK(cudaMalloc(&remote, ROWS * 512)); CK(cudaMemset(remote, 0, ROWS * 512));
    CK(cudaMalloc(&rver, ROWS * 8)); CK(cudaMemset(rver, 0, ROWS * 8));
    float4 *rbox; CK(cudaMalloc(&rbox, (size_t)N * 512));
    CK(cudaSetDevice(0));
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("peer access 0->1: %d\n", can);
    CK(cudaDeviceEnablePeerAccess(1, 0));
    float4 *local, *dst; long long *lver, *dver; unsigned *idx;
    CK(cudaMalloc(&local, ROWS * 512)); CK(cudaMemset(local, 0, ROWS * 512));
    CK(cudaMalloc(&lver, ROWS * 8)); CK(cudaMemset(lver, 0, ROWS * 8));
    CK(cudaMalloc(&dst, (size_t)N * 512)); CK(cudaMalloc(&dver, (size_t)N * 8)); CK(cudaMalloc(&idx, N * 4 * 4));
    std::mt19937_64 rng(1);
    std::vector<unsigned> h(N * 4);
    for (auto &x : h) x = rng() % ROWS;
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, double bytes, auto launch) {
        for (int i = 0; i < 3; i++) launch(i % 4);
        CK(cudaDeviceSynchronize());
        float tot = 0, best = 1e9;
        for (int i = 0; i < 10; i++) {
            cudaEventRecord(e0); launch(i % 4); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms; best = ms < best ? ms : best;
        }
        printf("%-40s avg %8.1f us best %8.1f us  %7.1f GB/s\n", name, tot / 10 * 1e3, best * 1e3, bytes / (tot / 10 * 1e-3) / 1e9);
    };
    const double rb = (double)N * 512, wb = (double)N * 8;
    timeit("local rows R=4 grid 832", rb, [&](int s) { gather_rows<4><<<832, 256>>>(local, dst, idx + s * N, N); });
    timeit("remote rows R=4 grid 832", rb, [&](int s) { gather_rows<4><<<832, 256>>>(remote, dst, idx + s * N, N); });
    timeit("remote rows R=8 grid 1184", rb, [&](int s) { gather_rows<8><<<1184, 256>>>(remote, dst, idx + s * N, N); });
    timeit("remote rows R=1 grid 1184", rb, [&](int s) { gather_rows<1><<<1184, 256>>>(remote, dst, idx + s * N, N); });
    timeit("remote rows R=4 grid 148", rb, [&](int s) { gather_rows<4><<<148, 256>>>(remote, dst, idx + s * N, N); });
    timeit("local words", wb, [&](int s) { gather_words<<<208, 256>>>(lver, dver, idx + s * N, N); });
    timeit("remote words", wb, [&](int s) { gather_words<<<208, 256>>>(rver, dver, idx + s * N, N); });
    timeit("remote row writes (contiguous)", rb, [&](int s) { scatter_rows<<<832, 256>>>(local + (size_t)s * N * 32, rbox, N); });
    return 0;
}
