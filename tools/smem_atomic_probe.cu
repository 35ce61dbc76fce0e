// smem_atomic_probe.cu — how fast can one SM accumulate scattered contributions in its own shared memory?
// Decides the accumulator format of the fused scatter + Adam kernel (kernels_scatter_adam.cu):
//   cas_f16x2 : atomicAdd(__half2*) on shared memory = LDS + HADD2 + ATOMS.CAST.SPIN loop (what nvcc emits for every
//               floating-point shared-memory atomic on sm_100a)
//   add_i32   : atomicAdd(int*) = native ATOMS.ADD
//   add_i32x2 : two native ATOMS.ADD on adjacent words (one table entry = two fixed-point features)
//   lds_u16   : plain 2-byte gathers (the hash-encode kernel's inner operation), for scale
//   red_global: red.global.add.noftz.f16x2 into a 256 KB table in L2 (what round 1 shipped)
// Random indices (LCG per thread) into a 128 KB table, 1024 threads per CTA, one CTA per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_atomic_probe tools/smem_atomic_probe.cu && ./smem_atomic_probe
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define THREADS 1024
#define OPS 512
extern __shared__ unsigned char sm[];

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE, int COARSE>
__global__ void __launch_bounds__(THREADS, 1) k_probe(uint32_t* out, uint32_t* gtab) {
    uint32_t* w = reinterpret_cast<uint32_t*>(sm);
    for (int i = threadIdx.x; i < 32768; i += THREADS) w[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u, acc = 0;
    // COARSE: 32 lanes of a warp fall into a 64-entry window (coarse levels: neighbouring samples share cells)
    const uint32_t win = (threadIdx.x >> 5) * 977u;
#pragma unroll 4
    for (int k = 0; k < OPS; ++k) {
        const uint32_t r = lcg(s);
        const uint32_t e = COARSE ? ((win + k * 131u + (r & 63u)) & 32767u) : (r & 32767u);
        if (MODE == 0) atomicAdd(reinterpret_cast<__half2*>(sm) + e, __floats2half2_rn(1.0f, 0.5f));
        else if (MODE == 1) atomicAdd(reinterpret_cast<int*>(sm) + e, (int)(r | 1u));
        else if (MODE == 2) { int* p = reinterpret_cast<int*>(sm) + (e & 16383u) * 2; atomicAdd(p, (int)r); atomicAdd(p + 1, (int)(r >> 3)); }
        else if (MODE == 3) acc += reinterpret_cast<const unsigned short*>(sm)[(r & 65535u)];
        else if (MODE == 4) { const __half2 v = __floats2half2_rn(1.0f, 0.5f); asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(gtab + (r & 65535u)), "r"(*reinterpret_cast<const uint32_t*>(&v)) : "memory"); }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = w[blockIdx.x & 1023] + acc;
}

// distributed shared memory: every update goes to a random CTA of the cluster (own rank included) — can the accumulators of
// one table be spread over the shared memories of a cluster, with the index computed once per sample?
//   MODE 0: red.shared::cluster.add.s32 (native remote integer atomic)   MODE 1: st.shared::cluster.u32 (plain remote store, for scale)
template <int MODE, int CL>
__global__ void __launch_bounds__(THREADS, 1) k_dsmem(uint32_t* out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    uint32_t* w = reinterpret_cast<uint32_t*>(sm);
    for (int i = threadIdx.x; i < 32768; i += THREADS) w[i] = 0;
    cluster.sync();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
#pragma unroll 4
    for (int k = 0; k < OPS; ++k) {
        const uint32_t r = lcg(s);
        const uint32_t rank = (r >> 16) % CL;
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + (r & 32767u) * 4u), "r"(rank));
        if (MODE == 0) asm volatile("red.relaxed.cluster.shared::cluster.add.s32 [%0], %1;" ::"r"(remote), "r"((int)(r | 1u)) : "memory");
        else asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(r) : "memory");
    }
    cluster.sync();
    if (threadIdx.x == 0) out[blockIdx.x] = w[blockIdx.x & 1023];
}

template <int MODE, int CL>
static void run_dsmem(const char* name, int sms, uint32_t* out) {
    cudaFuncSetAttribute(k_dsmem<MODE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    if (CL > 8) cudaFuncSetAttribute(k_dsmem<MODE, CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    const int ctas = sms / CL * CL >= 128 ? 128 / CL * CL : sms / CL * CL;
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = 131072;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, k_dsmem<MODE, CL>, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) cudaLaunchKernelEx(&cfg, k_dsmem<MODE, CL>, out);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double ops = (double)ctas * THREADS * OPS;
    printf("{\"mode\": \"%s\", \"cluster\": %d, \"ctas\": %d, \"us\": %.1f, \"updates_per_s_G\": %.1f, \"cycles_per_update_per_sm\": %.3f, \"err\": \"%s\"}\n", name, CL, ctas, ms * 1e3,
           ops / (ms * 1e-3) / 1e9, ms * 1e-3 * 1.965e9 / (THREADS * OPS), cudaGetErrorString(cudaGetLastError()));
}

template <int MODE, int COARSE>
static void run(const char* name, int sms, uint32_t* out, uint32_t* gtab) {
    cudaFuncSetAttribute(k_probe<MODE, COARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<MODE, COARSE><<<sms, THREADS, 131072>>>(out, gtab);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k_probe<MODE, COARSE><<<sms, THREADS, 131072>>>(out, gtab);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double ops = (double)sms * THREADS * OPS * (MODE == 2 ? 1 : 1);
    printf("{\"mode\": \"%s\", \"coarse\": %d, \"us\": %.1f, \"updates_per_s_G\": %.1f, \"cycles_per_update_per_sm\": %.3f, \"err\": \"%s\"}\n", name, COARSE, ms * 1e3,
           ops / (ms * 1e-3) / 1e9, ms * 1e-3 * 1.965e9 / (THREADS * OPS), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out, *gtab;
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&gtab, 65536 * 4);
    cudaMemset(gtab, 0, 65536 * 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"threads\": %d, \"updates_per_thread\": %d}\n", p.name, sms, THREADS, OPS);
    run<0, 0>("cas_f16x2", sms, out, gtab);
    run<0, 1>("cas_f16x2", sms, out, gtab);
    run<1, 0>("add_i32", sms, out, gtab);
    run<1, 1>("add_i32", sms, out, gtab);
    run<2, 0>("add_i32x2 (one entry = two words)", sms, out, gtab);
    run<2, 1>("add_i32x2 (one entry = two words)", sms, out, gtab);
    run<3, 0>("lds_u16", sms, out, gtab);
    run<4, 0>("red_global_f16x2", sms, out, gtab);
    run_dsmem<0, 2>("dsmem_red_add_s32", sms, out);
    run_dsmem<0, 4>("dsmem_red_add_s32", sms, out);
    run_dsmem<0, 8>("dsmem_red_add_s32", sms, out);
    run_dsmem<0, 16>("dsmem_red_add_s32", sms, out);
    run_dsmem<1, 8>("dsmem_store_u32", sms, out);
    return 0;
}
