// Micro-benchmark behind the hot-chain adder of segment_reduce (not part of the product): how many
// SM cycles per occurrence does ONE warp need to run the two dependent FADD chains
// (grad += g; data += g, src/hetu_cache/include/embedding.h:78-91) over values staged in shared
// memory, as a function of how the shared-memory loads are scheduled around the adds?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o addbench scripts/addbench.cu && ./addbench
#include <cuda_runtime.h>
#include <cstdio>

constexpr int W = 16, TILE = 128, STAGES = 4;

// V0: batch of B loaded one batch ahead of the adds (what hot_add does)
template <int B>
__device__ __forceinline__ void tile_pipelined(const float *st, float &g, float &d) {
    constexpr int NB = TILE / B;
    float cur[B], nxt[B];
#pragma unroll
    for (int j = 0; j < B; j++) cur[j] = st[j * W];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        if (b + 1 < NB) {
#pragma unroll
            for (int j = 0; j < B; j++) nxt[j] = st[((b + 1) * B + j) * W];
        }
#pragma unroll
        for (int j = 0; j < B; j++) { g = __fadd_rn(g, cur[j]); d = __fadd_rn(d, cur[j]); }
        if (b + 1 < NB) {
#pragma unroll
            for (int j = 0; j < B; j++) cur[j] = nxt[j];
        }
    }
}
// V1: whole tile loaded, then added (registers: TILE)
__device__ __forceinline__ void tile_bulk(const float *st, float &g, float &d) {
    float v[TILE];
#pragma unroll
    for (int j = 0; j < TILE; j++) v[j] = st[j * W];
#pragma unroll
    for (int j = 0; j < TILE; j++) { g = __fadd_rn(g, v[j]); d = __fadd_rn(d, v[j]); }
}
// V2: rolling window: value j + AHEAD is loaded right before value j is added, fully unrolled
// (no register copies: the window lives in a compile-time-indexed array)
template <int AHEAD>
__device__ __forceinline__ void tile_window(const float *st, float &g, float &d) {
    float v[TILE];
#pragma unroll
    for (int j = 0; j < AHEAD; j++) v[j] = st[j * W];
#pragma unroll
    for (int j = 0; j < TILE; j++) {
        if (j + AHEAD < TILE) v[j + AHEAD] = st[(j + AHEAD) * W];
        g = __fadd_rn(g, v[j]); d = __fadd_rn(d, v[j]);
    }
}
// V3: like V2 but volatile asm loads (the compiler cannot regroup them)
template <int AHEAD>
__device__ __forceinline__ void tile_window_asm(const float *st, float &g, float &d) {
    float v[TILE];
    const unsigned a = (unsigned)__cvta_generic_to_shared(st);
#pragma unroll
    for (int j = 0; j < AHEAD; j++)
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[j]) : "r"(a + j * W * 4));
#pragma unroll
    for (int j = 0; j < TILE; j++) {
        if (j + AHEAD < TILE)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[j + AHEAD]) : "r"(a + (j + AHEAD) * W * 4));
        g = __fadd_rn(g, v[j]); d = __fadd_rn(d, v[j]);
    }
}
// V4: one chain only (is the second chain free?)
__device__ __forceinline__ void tile_one_chain(const float *st, float &g, float &d) {
    float v[TILE];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = st[j * W];
#pragma unroll
    for (int j = 0; j < TILE; j++) {
        if (j + 16 < TILE) v[j + 16] = st[(j + 16) * W];
        g = __fadd_rn(g, v[j]);
    }
    d = g;
}

// kind: 0 = the other warps exit, 1 = dependent global loads, 2 = independent FADD/FFMA work,
// 3 = rows streamed like the cold phase (4 x 512 B loads in flight per warp, adds, stores)
template <int VARIANT>
__global__ void __launch_bounds__(256, 2) bench(float *out, long long *cycles, int ntiles, int kind, int adder) {
    extern __shared__ float ring[]; // [STAGES][TILE][W]
    for (int i = threadIdx.x; i < STAGES * TILE * W; i += blockDim.x) ring[i] = 1e-3f * (float)(i % 97);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ volatile int s_done;
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    if ((int)warp == adder) {
        float g = 0.f, d = 1.f;
        const bool active = lane < W;
        const long long c0 = clock64();
        for (int k = 0; k < ntiles; k++) {
            const float *st = ring + (k % STAGES) * TILE * W + lane;
            if (active) {
                if (VARIANT == 0) tile_pipelined<16>(st, g, d);
                if (VARIANT == 2) tile_window<16>(st, g, d);
                if (VARIANT == 4) tile_one_chain(st, g, d);
            }
            __syncwarp();
        }
        const long long c1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = c1 - c0;
        if (active) out[blockIdx.x * 32 + lane] = g + d;
        s_done = 1;
    } else if (kind == 1) {
        float acc = 0.f;
        const float4 *p = reinterpret_cast<const float4 *>(out) + 1024;
        for (int it = 0; !s_done; it++) {
            size_t idx = ((size_t)(blockIdx.x * 8 + warp) * 7919 + (size_t)it * 104729) % (1u << 22);
            float4 v = p[idx * 32 + lane];
            acc += v.x;
        }
        if (acc == 123.456f) out[0] = acc;
    } else if (kind == 2) {
        float a0 = lane, a1 = 1.f, a2 = 2.f, a3 = 3.f;
        while (!s_done) {
#pragma unroll
            for (int i = 0; i < 64; i++) { a0 = a0 * 1.0001f + a1; a1 = a1 * 0.9999f + a2; a2 = a2 * 1.0002f + a3; a3 = a3 * 0.9998f + a0; }
        }
        if (a0 + a1 + a2 + a3 == 123.456f) out[0] = a0;
    } else if (kind == 3) {
        float4 *p = reinterpret_cast<float4 *>(out) + 1024;
        for (int it = 0; !s_done; it++) {
            float4 v[4];
            size_t idx[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                idx[r] = ((size_t)(blockIdx.x * 8 + warp) * 7919 + (size_t)(it * 4 + r) * 104729) % (1u << 22);
                v[r] = p[idx[r] * 32 + lane];
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                v[r].x += 1.f; v[r].y += 1.f; v[r].z += 1.f; v[r].w += 1.f;
                p[idx[r] * 32 + lane] = v[r];
            }
        }
    }
}

template <int V>
void run(const char *name, int grid, int kind, int adder, float *out, long long *cyc) {
    const int ntiles = 88; // the hottest row of a WDL batch: 11k occurrences
    size_t smem = STAGES * TILE * W * 4;
    cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<V><<<grid, 256, smem>>>(out, cyc, ntiles, kind, adder);
    cudaDeviceSynchronize();
    bench<V><<<grid, 256, smem>>>(out, cyc, ntiles, kind, adder);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-26s grid %4d others %d adder warp %d  %6.2f cycles/occurrence (cta0) %6.2f (cta1)  %s\n", name, grid, kind, adder,
           (double)h[0] / (ntiles * TILE), (double)h[1] / (ntiles * TILE), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, ((size_t)1 << 22) * 512 + (1 << 20));
    cudaMemset(out, 0, ((size_t)1 << 22) * 512 + (1 << 20));
    cudaMalloc(&cyc, 4096 * sizeof(long long));
    const int grids[3] = {1, 148, 296};
    for (int gi = 0; gi < 3; gi++)
        for (int kind = 0; kind < 4; kind++)
            for (int adder = 0; adder < 8; adder += 7) {
                run<0>("pipelined batches of 16", grids[gi], kind, adder, out, cyc);
                if (kind == 0 || kind == 3) {
                    run<2>("rolling window 16", grids[gi], kind, adder, out, cyc);
                    run<4>("one chain", grids[gi], kind, adder, out, cyc);
                }
            }
    return 0;
}
