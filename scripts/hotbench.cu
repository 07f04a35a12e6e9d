// Micro-benchmark of the hot chain of segment_reduce in isolation (not part of the product): the
// real producer / adder device functions of herald_b200/csrc/hb_rows.cuh on one hot item per CTA
// (11k occurrences, 16-column chunk), with the other warps idle or streaming rows like the cold
// phase.  Prints SM cycles per occurrence of the adder.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iinclude \
//        -o hotbench scripts/hotbench.cu && ./hotbench
#include <cstdio>
#include <vector>
#include "../herald_b200/csrc/hb_rows.cuh"

using namespace hb;

struct Fn { // the accumulate functor's data phase, reduced to registers
    struct Acc { float d, g, t; };
    struct Ctx { float *out; unsigned pad[2]; };
    __device__ Acc load(const Ctx &, size_t) const { return Acc{1.f, 0.f, 0.f}; }
    __device__ Acc step(const Acc &a, float v) const { return Acc{__fadd_rn(a.d, v), __fadd_rn(a.g, v), a.t}; }
    __device__ void store(const Ctx &x, size_t c, const Acc &a) const { x.out[c] = a.d + a.g; }
};

template <int W>
__global__ void __launch_bounds__(256, 2)
    hot_kernel(const u32 *perm, const float *vals, size_t D, u32 cnt, float *out, long long *cycles, int S,
               int kind, float *scratch) {
    extern __shared__ __align__(16) float s_ring[];
    u64 *s_full = reinterpret_cast<u64 *>(s_ring + (size_t)S * kHotStageFloats);
    u64 *s_empty = s_full + S;
    __shared__ volatile int s_done;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&s_full[i], 32); mbar_init(&s_empty[i], 1); }
        s_done = 0;
    }
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    HotRing hr{s_ring, s_full, s_empty, (u32)S, 0u, 0u};
    const u32 q = blockIdx.x % (u32)(D / W);
    if (warp == 0) {
        Fn f; Fn::Ctx ctx{out + (size_t)blockIdx.x * D, {0, 0}};
        u64 waited = 0;
        const long long c0 = clock64();
        hot_add<W>(hr, f, ctx, D, cnt, q, waited);
        const long long c1 = clock64();
        if (lane == 0) { cycles[2 * blockIdx.x] = c1 - c0; cycles[2 * blockIdx.x + 1] = (long long)waited; }
        s_done = 1;
    } else if (warp < (unsigned)kHotWarps) {
        hot_produce<W, true>(hr, warp - 1, perm, vals, D, cnt, q);
    } else if (kind == 3) {
        float4 *p = reinterpret_cast<float4 *>(scratch);
        for (int it = 0; !s_done; it++) {
            float4 v[4]; size_t idx[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                idx[r] = ((size_t)(blockIdx.x * 8 + warp) * 7919 + (size_t)(it * 4 + r) * 104729) % (1u << 22);
                v[r] = p[idx[r] * 32 + lane];
            }
#pragma unroll
            for (int r = 0; r < 4; r++) { v[r].x += 1.f; p[idx[r] * 32 + lane] = v[r]; }
        }
    }
}

int main() {
    const size_t N = 212992, D = 128; const u32 cnt = 11322;
    std::vector<u32> perm(cnt);
    for (u32 i = 0; i < cnt; i++) perm[i] = (u32)(((size_t)i * 7919 + 13) % N);
    std::sort(perm.begin(), perm.end());
    u32 *dperm; float *vals, *out, *scratch; long long *cyc;
    cudaMalloc(&dperm, cnt * 4); cudaMemcpy(dperm, perm.data(), cnt * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&vals, N * D * 4); cudaMemset(vals, 0, N * D * 4);
    cudaMalloc(&out, 4096 * D * 4); cudaMalloc(&cyc, 8192 * 8);
    cudaMalloc(&scratch, ((size_t)1 << 22) * 512 + 4096); cudaMemset(scratch, 0, ((size_t)1 << 22) * 512 + 4096);
    for (int S : {3, 6})
        for (int grid : {8, 148, 296})
            for (int kind : {0, 3}) {
                size_t smem = hot_smem_bytes(S);
                cudaFuncSetAttribute(hot_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                for (int rep = 0; rep < 2; rep++) {
                    hot_kernel<16><<<grid, 256, smem>>>(dperm, vals, D, cnt, out, cyc, S, kind, scratch);
                    cudaDeviceSynchronize();
                }
                long long h[4];
                cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                printf("W=16 stages %d grid %3d others %d: %6.2f cycles/occurrence, of which waiting %5.2f  (cta1 %6.2f)  %s\n", S, grid,
                       kind, (double)h[0] / cnt, (double)h[1] / cnt, (double)h[2] / cnt, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
