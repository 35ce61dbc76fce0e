// scatter_global.cuh — one (sample, level) of the hash-grid gradient scatter with GLOBAL f16x2 reductions, the reference's own form:
// grad[idx_c] += half2(d_enc * w_c) over the 8 corners (TCNN encodings/grid.h:386-509, atomicAdd(__half2) at :427-431).  Shared by
// the stand-alone kernel (kernels_encode.cu k_encode_backward) and the sparse path of the unified scatter kernel
// (kernels_scatter_smem.cu).
#pragma once
#include "mon_device.cuh"

struct EncCorner { uint32_t idx[8]; float w[8]; };

__device__ __forceinline__ void level_corners(const MonGrid& g, uint32_t l, const float* u, EncCorner& c) {
    float f[3]; uint32_t p[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) mon_pos_fract(u[d], g.scale[l], f[d], p[d]);
    const bool hashed = g.hashed[l] != 0;
    const uint32_t size = g.size[l], res = g.res[l];
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        float w = 1.0f;
        w = __fmul_rn(w, (k & 1) ? f[0] : __fsub_rn(1.0f, f[0]));
        w = __fmul_rn(w, (k & 2) ? f[1] : __fsub_rn(1.0f, f[1]));
        w = __fmul_rn(w, (k & 4) ? f[2] : __fsub_rn(1.0f, f[2]));
        c.w[k] = w;
        c.idx[k] = mon_grid_index(hashed, size, res, p[0] + (k & 1), p[1] + ((k >> 1) & 1), p[2] + ((k >> 2) & 1));
    }
}

// explicit global-space reduction: a generic-pointer atomicAdd(__half2*) makes the compiler query the address space
// and branch around every one of the 32 atomics of a thread
__device__ __forceinline__ void red_add_f16x2(__half2* addr, __half2 v) {
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(*reinterpret_cast<const uint32_t*>(&v)) : "memory");
}
// two adjacent table entries (one aligned 8-byte word) in ONE reduction: the scatter is bound by the number of
// reduction lane-operations the LSU can issue (REDG ~1.3 cycles per lane), not by bytes
__device__ __forceinline__ void red_add_f16x2_pair(void* addr8, __half2 lo, __half2 hi) {
    asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(addr8), "r"(*reinterpret_cast<const uint32_t*>(&lo)),
                 "r"(*reinterpret_cast<const uint32_t*>(&hi)) : "memory");
}

// one (sample, level) on a power-of-two table, the level's constants as scalars (the caller may hold them in registers: the fused MLP
// kernel's lanes own one level each for the whole kernel).  tab = the level's slice of the entry-ordered gradient table.
__device__ __forceinline__ void scatter_level_pow2(float scale, uint32_t res, bool hashed, uint32_t size, char* __restrict__ tab,
                                                   float g0, float g1, const float (&u)[3]) {
                // same index arithmetic as the forward kernel, directly in byte offsets of the 4-byte entries
                const uint32_t bmask = 4u * size - 4u;
                float fr[3]; uint32_t cell[3];
                mon_pos_fract(u[0], scale, fr[0], cell[0]);
                mon_pos_fract(u[1], scale, fr[1], cell[1]);
                mon_pos_fract(u[2], scale, fr[2], cell[2]);
                const float h0 = __fsub_rn(1.0f, fr[0]), h1 = __fsub_rn(1.0f, fr[1]), h2 = __fsub_rn(1.0f, fr[2]);
                const float wxy[4] = {__fmul_rn(h0, h1), __fmul_rn(fr[0], h1), __fmul_rn(h0, fr[1]), __fmul_rn(fr[0], fr[1])};
                uint32_t ax[2], ay[2], az[2];
                ax[0] = cell[0] << 2; ax[1] = (cell[0] + 1u) << 2;
                if (hashed) {
                    ay[0] = (cell[1] * 2654435761u) << 2; ay[1] = ((cell[1] + 1u) * 2654435761u) << 2;
                    az[0] = (cell[2] * 805459861u) << 2; az[1] = ((cell[2] + 1u) * 805459861u) << 2;
                } else {
                    ay[0] = (cell[1] * res) << 2; ay[1] = ((cell[1] + 1u) * res) << 2;
                    az[0] = (cell[2] * res * res) << 2; az[1] = ((cell[2] + 1u) * res * res) << 2;
                }
                if ((cell[0] & 1u) == 0u && size >= 2u && (hashed || (res & 1u) == 0u)) {
                    // even x: the corners x and x+1 are entries 2j and 2j+1 (in either order) of one aligned 8-byte word,
                    // for the hash (x enters by XOR, bit 0 of x is clear) and for the dense index (the y/z strides are even)
    #pragma unroll
                    for (uint32_t q = 0; q < 4; ++q) {
                        const float wz = (q & 2) ? fr[2] : h2;
                        const float w0 = __fmul_rn(wxy[(q & 1) * 2], wz), w1 = __fmul_rn(wxy[(q & 1) * 2 + 1], wz);
                        const uint32_t off = (hashed ? (ax[0] ^ ay[q & 1] ^ az[q >> 1]) : (ax[0] + ay[q & 1] + az[q >> 1])) & bmask;
                        const __half2 v0 = __floats2half2_rn(__fmul_rn(g0, w0), __fmul_rn(g1, w0));
                        const __half2 v1 = __floats2half2_rn(__fmul_rn(g0, w1), __fmul_rn(g1, w1));
                        const bool swap = (off & 4u) != 0u;     // corner x sits in the upper half of the word
                        red_add_f16x2_pair(tab + (off & ~7u), swap ? v1 : v0, swap ? v0 : v1);
                    }
                } else {
    #pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        const float wgt = __fmul_rn(wxy[k & 3], (k & 4) ? fr[2] : h2);
                        const uint32_t off = (hashed ? (ax[k & 1] ^ ay[(k >> 1) & 1] ^ az[k >> 2]) : (ax[k & 1] + ay[(k >> 1) & 1] + az[k >> 2])) & bmask;
                        const __half2 v = __floats2half2_rn(__fmul_rn(g0, wgt), __fmul_rn(g1, wgt));
                        red_add_f16x2(reinterpret_cast<__half2*>(tab + off), v);
                    }
                }
}

// Zero d_enc pairs (samples after the early stop) are skipped: adding +0 is an identity, so the result is unchanged.
// one (sample, level): grad[idx_c] += half2(d_enc * w_c) over the 8 corners.  gwj = the level's two fp16 gradients.
__device__ __forceinline__ void scatter_level(const MonGrid& g, uint32_t l, uint32_t gwj, const float (&u)[3], __half* __restrict__ grid_grad) {
            const float g0 = __half2float(__ushort_as_half((unsigned short)(gwj & 0xffffu)));
            const float g1 = __half2float(__ushort_as_half((unsigned short)(gwj >> 16)));
            const uint32_t size = g.size[l];
            char* tab = reinterpret_cast<char*>(reinterpret_cast<__half2*>(grid_grad) + g.offset[l]);
            if ((size & (size - 1)) == 0) {
                scatter_level_pow2(g.scale[l], g.res[l], g.hashed[l] != 0, size, tab, g0, g1, u);
            } else {
                EncCorner c;
                level_corners(g, l, u, c);
    #pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const __half2 v = __halves2half2(__float2half_rn(__fmul_rn(g0, c.w[k])), __float2half_rn(__fmul_rn(g1, c.w[k])));
                    red_add_f16x2(reinterpret_cast<__half2*>(tab) + c.idx[k], v);
                }
            }
}

