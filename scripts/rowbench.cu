// Micro-benchmark behind the design of the segment-reduce cold path (not part of the product):
// what HBM bandwidth does the access pattern "per unique id: read cache row + gradient row +
// owner row, write cache row + owner row" (512 B rows, random addresses over 17 GB / 1.7 GB /
// 109 MB regions) reach as a function of rows in flight per warp and resident warps?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowbench scripts/rowbench.cu && ./rowbench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1);} } while (0)

constexpr int D4 = 32; // float4 per row (D = 128)

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// register variant: R row-triples in flight per warp
template <int R, int MINB>
__global__ void __launch_bounds__(256, MINB)
rmw_reg(float4 *__restrict__ cache, const float4 *__restrict__ grads, float4 *__restrict__ table,
        const unsigned *__restrict__ slot, const unsigned *__restrict__ gidx,
        const unsigned *__restrict__ trow, unsigned U, unsigned *ticket) {
    const unsigned lane = threadIdx.x & 31;
    while (true) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(~0u, t, 0);
        if (t * 32 >= U) break;
        const unsigned u = t * 32 + lane;
        const unsigned ms = u < U ? slot[u] : 0, mg = u < U ? gidx[u] : 0, mt = u < U ? trow[u] : 0;
        const int nrows = min(32u, U - t * 32);
        for (int g0 = 0; g0 < nrows; g0 += R) {
            float4 d[R], g[R], tt[R];
            unsigned s[R], tr[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                s[r] = __shfl_sync(~0u, ms, (g0 + r) & 31);
                unsigned gi = __shfl_sync(~0u, mg, (g0 + r) & 31);
                tr[r] = __shfl_sync(~0u, mt, (g0 + r) & 31);
                if (g0 + r < nrows) {
                    d[r] = cache[(size_t)s[r] * D4 + lane];
                    g[r] = __ldcs(grads + (size_t)gi * D4 + lane);
                    tt[r] = table[(size_t)tr[r] * D4 + lane];
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++)
                if (g0 + r < nrows) {
                    cache[(size_t)s[r] * D4 + lane] = add4(d[r], g[r]);
                    table[(size_t)tr[r] * D4 + lane] = add4(tt[r], g[r]);
                }
        }
    }
}

// cp.async variant: every warp owns a ring of S stages x 3 rows x 512 B in shared memory
__device__ __forceinline__ void cp16(void *smem, const void *g) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
}
template <int S, int MINB>
__global__ void __launch_bounds__(256, MINB)
rmw_ring(float4 *__restrict__ cache, const float4 *__restrict__ grads, float4 *__restrict__ table,
         const unsigned *__restrict__ slot, const unsigned *__restrict__ gidx,
         const unsigned *__restrict__ trow, unsigned U, unsigned *ticket) {
    extern __shared__ float4 ring[]; // [8 warps][S][3][32]
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *my = ring + (size_t)warp * S * 3 * 32;
    while (true) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(~0u, t, 0);
        if (t * 32 >= U) break;
        const unsigned u = t * 32 + lane;
        const unsigned ms = u < U ? slot[u] : 0, mg = u < U ? gidx[u] : 0, mt = u < U ? trow[u] : 0;
        const int nrows = min(32u, U - t * 32);
        auto issue = [&](int i) {
            if (i < nrows) {
                unsigned s = __shfl_sync(~0u, ms, i), gi = __shfl_sync(~0u, mg, i), tr = __shfl_sync(~0u, mt, i);
                float4 *st = my + (size_t)(i % S) * 96;
                cp16(st + lane, cache + (size_t)s * D4 + lane);
                cp16(st + 32 + lane, grads + (size_t)gi * D4 + lane);
                cp16(st + 64 + lane, table + (size_t)tr * D4 + lane);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll 1
        for (int i = 0; i < S - 1; i++) issue(i);
#pragma unroll 1
        for (int i = 0; i < nrows; i++) {
            issue(i + S - 1); // into the stage consumed in iteration i - 1
            asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
            const float4 *st = my + (size_t)(i % S) * 96; // each lane reads back its own 16 B
            float4 d = st[lane], g = st[32 + lane], tt = st[64 + lane];
            unsigned s = __shfl_sync(~0u, ms, i), tr = __shfl_sync(~0u, mt, i);
            cache[(size_t)s * D4 + lane] = add4(d, g);
            table[(size_t)tr * D4 + lane] = add4(tt, g);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
}

// plain gather: dst[n] = cache[slot[n]], R rows in flight
template <int R>
__global__ void __launch_bounds__(256)
gather(const float4 *__restrict__ cache, float4 *__restrict__ dst, const unsigned *__restrict__ slot, unsigned N) {
    const unsigned lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (size_t)gridDim.x * 8;
    for (size_t base = w * 32; base < N; base += nw * 32) {
        const unsigned ms = base + lane < N ? slot[base + lane] : 0;
        const int nrows = (int)min((size_t)32, N - base);
        for (int g0 = 0; g0 < nrows; g0 += R) {
            float4 v[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                unsigned s = __shfl_sync(~0u, ms, (g0 + r) & 31);
                if (g0 + r < nrows) v[r] = cache[(size_t)s * D4 + lane];
            }
#pragma unroll
            for (int r = 0; r < R; r++)
                if (g0 + r < nrows) __stcs(dst + (base + g0 + r) * D4 + lane, v[r]);
        }
    }
}

int main() {
    const size_t V = 33762577, SLOTS = 3400000, N = 212992;
    const unsigned U = 129000;
    const int SETS = 4;
    float4 *cache, *grads, *table, *dst;
    CK(cudaMalloc(&cache, SLOTS * 512));
    CK(cudaMalloc(&grads, N * 512 * (size_t)SETS));
    CK(cudaMalloc(&table, V * 512));
    CK(cudaMalloc(&dst, N * 512 * (size_t)2));
    CK(cudaMemset(cache, 0, SLOTS * 512));
    CK(cudaMemset(grads, 0, N * 512 * (size_t)SETS));
    CK(cudaMemset(table, 0, V * 512));
    std::mt19937_64 rng(1);
    unsigned *slot, *gidx, *trow, *gslot, *ticket;
    CK(cudaMalloc(&slot, U * 4 * SETS)); CK(cudaMalloc(&gidx, U * 4 * SETS)); CK(cudaMalloc(&trow, U * 4 * SETS));
    CK(cudaMalloc(&gslot, N * 4 * SETS)); CK(cudaMalloc(&ticket, 4));
    {
        std::vector<unsigned> a(U * SETS), b(U * SETS), c(U * SETS), e(N * SETS);
        for (int s = 0; s < SETS; s++) {
            // distinct rows per set (a permutation prefix), sorted owner rows like sorted uniques
            std::vector<unsigned> perm(N);
            for (unsigned i = 0; i < N; i++) perm[i] = i;
            for (unsigned i = 0; i < U; i++) { std::swap(perm[i], perm[i + rng() % (N - i)]); }
            std::vector<unsigned> tr(U);
            for (unsigned i = 0; i < U; i++) tr[i] = (unsigned)((double)i / U * V) + rng() % 200;
            for (unsigned i = 0; i < U; i++) {
                a[s * U + i] = (unsigned)((rng() % (SLOTS / U)) + (size_t)i * (SLOTS / U)); // distinct slots
                std::swap(a[s * U + i], a[s * U + rng() % (i + 1)]);
                b[s * U + i] = perm[i] + s * N;
                c[s * U + i] = tr[i];
            }
            for (unsigned i = 0; i < N; i++) e[s * N + i] = rng() % SLOTS;
        }
        CK(cudaMemcpy(slot, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(gidx, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(trow, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(gslot, e.data(), e.size() * 4, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double rmw_bytes = (double)U * 512 * 5, gather_bytes = (double)N * 512 * 2;
    auto timeit = [&](const char *name, double bytes, auto launch) {
        for (int i = 0; i < 4; i++) launch(i % SETS);
        CK(cudaDeviceSynchronize());
        float best = 1e9, tot = 0;
        for (int i = 0; i < 12; i++) {
            CK(cudaMemsetAsync(ticket, 0, 4));
            cudaEventRecord(e0);
            launch(i % SETS);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best; tot += ms;
        }
        printf("%-28s avg %7.1f us  best %7.1f us  %7.0f GB/s (avg)\n", name, tot / 12 * 1e3, best * 1e3, bytes / (tot / 12 * 1e-3) / 1e9);
    };
#define RMW_REG(R, MINB, CTAS) timeit("rmw_reg R=" #R " ctas/sm=" #CTAS, rmw_bytes, [&](int s) { \
        cudaMemsetAsync(ticket, 0, 4); \
        rmw_reg<R, MINB><<<148 * CTAS, 256>>>(cache, grads, table, slot + s * U, gidx + s * U, trow + s * U, U, ticket); })
    RMW_REG(1, 8, 8); RMW_REG(2, 4, 4); RMW_REG(2, 6, 6); RMW_REG(2, 8, 8);
    RMW_REG(4, 2, 2); RMW_REG(4, 3, 3); RMW_REG(4, 4, 4); RMW_REG(8, 2, 2);
#define RMW_RING(S, MINB, CTAS) do { \
        size_t smem = (size_t)8 * S * 3 * 512; \
        CK(cudaFuncSetAttribute(rmw_ring<S, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        timeit("rmw_ring S=" #S " ctas/sm=" #CTAS, rmw_bytes, [&](int s) { \
            cudaMemsetAsync(ticket, 0, 4); \
            rmw_ring<S, MINB><<<148 * CTAS, 256, smem>>>(cache, grads, table, slot + s * U, gidx + s * U, trow + s * U, U, ticket); }); } while (0)
    RMW_RING(3, 4, 4); RMW_RING(4, 4, 4); RMW_RING(5, 3, 3); RMW_RING(6, 2, 2); RMW_RING(8, 2, 2); RMW_RING(3, 6, 6);
#define GATHER(R, CTAS) timeit("gather R=" #R " ctas/sm=" #CTAS, gather_bytes, [&](int s) { \
        gather<R><<<148 * CTAS, 256>>>(cache, dst + (size_t)(s & 1) * N * D4, gslot + s * N, (unsigned)N); })
    GATHER(2, 8); GATHER(4, 4); GATHER(4, 8); GATHER(8, 4); GATHER(8, 8);
    return 0;
}
