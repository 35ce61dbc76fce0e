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
    return 0;
}
