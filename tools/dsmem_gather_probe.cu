// dsmem_gather_probe.cu — micro-benchmark for the next encode design: how fast are random 2-byte gathers
// from the PARTNER CTA's shared memory (distributed shared memory in a 2-CTA cluster) compared with local ones?
//
// The encode kernel keeps one 128 KB (level, feature) table slice per CTA and spends its index / weight arithmetic twice per
// level (once per feature).  A 2-CTA cluster per level could compute the 8 corner indices once per point and gather feature 0
// from its own slice and feature 1 from the partner's — if remote gathers sustain about the local rate.  This probe measures
// exactly that access pattern: 1024 threads per CTA, 65536-entry fp16 table per CTA, 8 independent hashed gathers per "point".
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/dsmem_probe tools/dsmem_gather_probe.cu
//   gpurun_out/dsmem_probe            # prints G gathers/s for: local, remote, half local + half remote, and L2-resident 4-byte gathers
//
// MEASUREMENT TOOL ONLY: not linked into the product.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

constexpr int THREADS = 1024;
constexpr uint32_t ENTRIES = 65536;   // 128 KB of fp16
constexpr int POINTS_PER_THREAD = 256;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE 0: all gathers local; 1: all from the partner CTA; 2: per point 8 local + 8 remote (the proposed encode)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_probe(float* __restrict__ sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half* table = reinterpret_cast<__half*>(smem);
    cg::cluster_group cluster = cg::this_cluster();
    for (uint32_t i = threadIdx.x; i < ENTRIES; i += THREADS) table[i] = __float2half_rn((float)((i * 2654435761u) >> 24) * (1.0f / 256.0f));
    cluster.sync();
    const __half* remote = cluster.map_shared_rank(table, cluster.block_rank() ^ 1u);
    float acc0 = 0.0f, acc1 = 0.0f;
    uint32_t seed = blockIdx.x * THREADS + threadIdx.x;
#pragma unroll 2
    for (int p = 0; p < POINTS_PER_THREAD; ++p) {
        // a "cell": three hashed coordinates, 8 corners by XOR like the hash grid's index
        const uint32_t h = mix(seed + (uint32_t)p * 0x9e3779b9u);
        const uint32_t x0 = h, x1 = h + 1u, y0 = (h >> 7) * 2654435761u, y1 = ((h >> 7) + 1u) * 2654435761u, z0 = (h >> 13) * 805459861u,
                       z1 = ((h >> 13) + 1u) * 805459861u;
        uint32_t idx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = (((k & 1) ? x1 : x0) ^ ((k & 2) ? y1 : y0) ^ ((k & 4) ? z1 : z0)) & (ENTRIES - 1u);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) acc0 += __half2float(table[idx[k]]);
            if (MODE == 1) acc0 += __half2float(remote[idx[k]]);
            if (MODE == 2) { acc0 += __half2float(table[idx[k]]); acc1 += __half2float(remote[idx[k]]); }
        }
    }
    cluster.sync();   // the partner may still be reading this CTA's table
    if (acc0 + acc1 == -1.0f) sink[0] = acc0;
}

// MODE 3 of the comparison: no shared-memory table at all — both features of an entry as one 4-byte word in an L2-resident global
// table of the real size (16 levels x 65536 entries x 4 B = 4 MB), 8 gathers per point instead of 16, many more warps per SM.
__global__ void __launch_bounds__(256) k_probe_l2(const uint32_t* __restrict__ table, uint32_t entries_mask, int points_per_thread, float* __restrict__ sink) {
    float acc0 = 0.0f, acc1 = 0.0f;
    const uint32_t seed = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t level_base = (blockIdx.x & 15u) << 16;   // a CTA works on one level's 65536 entries, like one encode job
#pragma unroll 2
    for (int p = 0; p < points_per_thread; ++p) {
        const uint32_t h = mix(seed + (uint32_t)p * 0x9e3779b9u);
        const uint32_t x0 = h, x1 = h + 1u, y0 = (h >> 7) * 2654435761u, y1 = ((h >> 7) + 1u) * 2654435761u, z0 = (h >> 13) * 805459861u,
                       z1 = ((h >> 13) + 1u) * 805459861u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t idx = (((k & 1) ? x1 : x0) ^ ((k & 2) ? y1 : y0) ^ ((k & 4) ? z1 : z0)) & entries_mask;
            const uint32_t w = __ldg(table + level_base + idx);
            const __half2 v = *reinterpret_cast<const __half2*>(&w);
            acc0 += __low2float(v); acc1 += __high2float(v);
        }
    }
    if (acc0 + acc1 == -1.0f) sink[0] = acc0;
}

static void run_l2(int sms, float* sink) {
    const size_t n = (size_t)16 << 16;
    uint32_t* table;
    CK(cudaMalloc(&table, n * 4));
    CK(cudaMemset(table, 0x3c, n * 4));
    const int ctas = sms * 8, ppt = POINTS_PER_THREAD;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) k_probe_l2<<<ctas, 256>>>(table, 65535u, ppt, sink);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) k_probe_l2<<<ctas, 256>>>(table, 65535u, ppt, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double words = (double)ctas * 256 * ppt * 8;
    printf("%-28s %8.2f us/launch  %8.1f G 4-byte gathers/s = %.1f G feature gathers/s  (%d CTAs x 256 threads)\n", "L2-resident half2 table", ms / reps * 1e3,
           words / (ms / reps * 1e-3) * 1e-9, 2 * words / (ms / reps * 1e-3) * 1e-9, ctas);
    CK(cudaFree(table));
}

template <int MODE>
static void run(const char* name, int ctas, float* sink) {
    CK(cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ENTRIES * 2)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) k_probe<MODE><<<ctas, THREADS, ENTRIES * 2>>>(sink);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) k_probe<MODE><<<ctas, THREADS, ENTRIES * 2>>>(sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double gathers = (double)ctas * THREADS * POINTS_PER_THREAD * (MODE == 2 ? 16 : 8);
    printf("%-28s %8.2f us/launch  %8.1f G gathers/s  (%d CTAs, incl. the %.0f KB table fill per launch)\n", name, ms / reps * 1e3,
           gathers / (ms / reps * 1e-3) * 1e-9, ctas, ENTRIES * 2 / 1024.0);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int ctas = prop.multiProcessorCount & ~1;   // whole 2-CTA clusters
    float* sink;
    CK(cudaMalloc(&sink, 4));
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    run<0>("local gathers", ctas, sink);
    run<1>("remote (DSMEM) gathers", ctas, sink);
    run<2>("8 local + 8 remote / point", ctas, sink);
    run_l2(prop.multiProcessorCount, sink);
    // the encode kernel today: 131072 points x 32 (level, feature) jobs x 8 gathers = 33.6 M gathers in 22 us = 1525 G gathers/s
    printf("reference point: the encode kernel sustains 1525 G gathers/s today (33.6 M gathers in 22.0 us)\n");
    CK(cudaFree(sink));
    return 0;
}
