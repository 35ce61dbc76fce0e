// tc05.cuh — thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path: tcgen05.mma with
// shared-memory operand descriptors, TMEM allocation / loads, mbarrier completion.  Only what the fused
// MLP kernels need (cta_group::1, kind::f16, fp16 operands, fp32 accumulators in TMEM).
//
// Shared-memory operand layouts (16-byte units; T = 8 fp16 per unit), as the hardware defines them:
//   K-major  SWIZZLE_128B : rows of 128 B (64 fp16 along K); 8-row groups SBO bytes apart; the 16-B chunk index
//                           inside a row is XORed with (row & 7)            [address bits 4-6 ^= bits 7-9]
//   K-major  SWIZZLE_64B  : rows of 64 B; chunk index (2 bits) ^= ((row >> 1) & 3)   [bits 4-5 ^= bits 7-8]
//   K-major  no swizzle   : 8x8 "core matrices" (8 rows x 16 B, contiguous 128 B); SBO between 8-row groups,
//                           LBO between the two K chunks of one MMA
//   MN-major SWIZZLE_128B : 64 fp16 along MN contiguous (128 B) per K row; rows 128 B apart; 8-K-row groups SBO
//                           apart; the next 64 MN elements LBO apart  — physically the SAME bytes as a K-major
//                           SW128 tile whose rows are the K index, which is what lets one activation tile feed a
//                           forward GEMM (K-major) and the weight-gradient GEMM (MN-major) without a transpose
//   MN-major SWIZZLE_64B / no swizzle: analogous with 32 / 8 elements per row.
// Tiles must be aligned to the swizzle period (1024 B for SW128, 512 B for SW64).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared memory matrix descriptor (64 bit)
enum : uint64_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t swizzle) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell), bits [46,48)
    d |= swizzle << 61;                                    // layout type, bits [61,64)
    return d;
}

// The same descriptor as two 32-bit halves.  The high word (SBO, version, swizzle) is a compile-time constant per operand; the low
// word is `sm16 + constant` with sm16 = (1024-aligned shared base) >> 4 computed once per kernel: an operand descriptor costs the
// issuing thread ONE uniform add instead of the shift / mask / or chain of make_desc (20 MMAs per tile were 1750 cycles of issue,
// profiles/r1z_mlp_phase_times.txt).  Shared addresses are below 256 KB, so base + offset cannot carry out of the 14-bit field.
struct Desc2 { uint32_t lo, hi; };
__device__ __forceinline__ Desc2 make_desc2(uint32_t sm16, uint32_t off_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t swizzle) {
    Desc2 d;
    d.lo = sm16 + ((off_bytes >> 4) | ((lbo_bytes >> 4) << 16));
    d.hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)swizzle << 29);
    return d;
}

// ---- instruction descriptor (32 bit): fp16 x fp16 -> fp32, dense
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4)                 // c_format = F32
         | (0u << 7) | (0u << 10)    // a_format = b_format = F16
         | (a_mn_major << 15) | (b_mn_major << 16)
         | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, Desc2 a, Desc2 b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// one lane of a converged warp (the tcgen05 issue / commit instructions are single-thread)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t n_cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(n_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t n_cols) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit: thread i of the warp receives N consecutive fp32 columns of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(addr)
        : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// ---- swizzled tile addressing helpers (byte offset of the 16-byte chunk `chunk` of row `row`)
__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }
__device__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }
// no-swizzle core-matrix tile with `kchunks` 16-byte chunks per row: core (row/8, chunk) at (row/8)*kchunks*128 + chunk*128
__device__ __forceinline__ uint32_t core_off(uint32_t row, uint32_t chunk, uint32_t kchunks) {
    return (row >> 3) * (kchunks * 128u) + chunk * 128u + (row & 7u) * 16u;
}

}  // namespace tc05
