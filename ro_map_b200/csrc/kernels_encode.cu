// kernels_encode.cu — multiresolution hash-grid lookup (forward) and gradient scatter (backward).
//
// Replaces GenerateInputPoints (MON/Core/src/nerf_model.cu:536-566), tcnn kernel_grid
// (TCNN/include/tiny-cuda-nn/encodings/grid.h:220-384) and kernel_grid_backward (:386-509).
// Sample positions are never materialised: each thread regenerates its sample from the 36-byte
// ray and the jitter (A3), so the 1.57 MB PointsInput / 0.52 MB SamplesDistances round trips of
// the reference disappear.
//
// Mapping: a CTA owns a tile of 128 consecutive samples (= 4 rays) for ALL levels; thread t
// handles sample t%128 and the 4 levels of group t/128.  Consecutive lanes are consecutive
// samples of one ray, so at coarse levels a warp's 8-corner gathers fall into a handful of
// 128-byte lines, and the coherent-prime hash (x multiplier 1) keeps the x/x+1 corner pair in
// one 32-byte sector at hashed levels.  Output is point-major [N][32] fp16 (64 B per sample, one
// 16-byte store per thread), the layout the fused MLP kernel stages with one bulk copy per tile.
//
// Arithmetic is the reference's: weights in fp32, rounded to fp16, one fp32 FMA per corner whose
// result is rounded back to fp16 after every corner (grid.h:334 via common.h:539-559).
#include "mon_device.cuh"
#include "mon_kernels.h"

#define ENC_THREADS 512
#define ENC_TILE 128

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

// rays: one per S samples.  in_box: optional per-ray validity (render); jitter: injected array or RNG.
__global__ void __launch_bounds__(ENC_THREADS)
k_encode_forward(MonGrid g, uint32_t n_points, uint32_t S, const MonRay* __restrict__ rays,
                 const int* __restrict__ in_box, const float* __restrict__ jitter, uint32_t seed,
                 const MonCtrl* __restrict__ ctrl, uint32_t rng_stream, uint32_t iter_fixed, float bmin0, float bmin1, float bmin2,
                 float bmax0, float bmax1, float bmax2, const __half* __restrict__ grid, __half* __restrict__ enc) {
    if (ctrl && ctrl->skip) return;
    const uint32_t p = threadIdx.x & (ENC_TILE - 1), lg = threadIdx.x >> 7;
    const uint32_t pt = blockIdx.x * ENC_TILE + p;
    if (pt >= n_points) return;
    const uint32_t ray = pt / S, n = pt - ray * S;
    uint4 outv = make_uint4(0, 0, 0, 0);
    if (!in_box || in_box[ray]) {
        const MonRay r = rays[ray];
        // the batch kernel already advanced ctrl->iter; this iteration's counter is iter-1
        const uint32_t iter = ctrl ? ctrl->iter - 1 : iter_fixed;
        const float xi = mon_rand(jitter, seed, iter, rng_stream, pt);
        const float t = mon_sample_t(r, n, xi, (float)S);
        const float bmin[3] = {bmin0, bmin1, bmin2}, bmax[3] = {bmax0, bmax1, bmax2};
        float u[3];
        mon_sample_point(r, t, bmin, bmax, u);
        uint32_t packed[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t l = lg * 4 + j;
            uint32_t res2 = 0;
            if (l < g.n_levels) {
                EncCorner c;
                level_corners(g, l, u, c);
                const __half2* tab = reinterpret_cast<const __half2*>(grid) + g.offset[l];
                __half2 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(tab + c.idx[k]);
                __half a0 = __float2half_rn(0.0f), a1 = a0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float wh = __half2float(__float2half_rn(c.w[k]));
                    a0 = __float2half_rn(__fmaf_rn(wh, __low2float(v[k]), __half2float(a0)));
                    a1 = __float2half_rn(__fmaf_rn(wh, __high2float(v[k]), __half2float(a1)));
                }
                res2 = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(a1) << 16);
            }
            packed[j] = res2;
        }
        outv = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    reinterpret_cast<uint4*>(enc + (size_t)pt * MON_IN)[lg] = outv;
}

// gradient scatter: grad[idx] += half2(d_enc * w) with f16x2 reductions, exactly the reference's
// atomicAdd(__half2) (grid.h:427-431).  Zero d_enc pairs (samples after the early stop) are skipped:
// adding +0 is an identity, so the result is unchanged.
__global__ void __launch_bounds__(ENC_THREADS)
k_encode_backward(MonGrid g, uint32_t n_points, uint32_t S, const MonRay* __restrict__ rays,
                  const float* __restrict__ jitter, uint32_t seed, const MonCtrl* __restrict__ ctrl,
                  float bmin0, float bmin1, float bmin2, float bmax0, float bmax1, float bmax2,
                  const __half* __restrict__ d_enc, __half* __restrict__ grid_grad) {
    if (ctrl->skip) return;
    const uint32_t p = threadIdx.x & (ENC_TILE - 1), lg = threadIdx.x >> 7;
    const uint32_t pt = blockIdx.x * ENC_TILE + p;
    if (pt >= n_points) return;
    const uint4 gv = reinterpret_cast<const uint4*>(d_enc + (size_t)pt * MON_IN)[lg];
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
    if (((gv.x | gv.y | gv.z | gv.w) & 0x7fff7fffu) == 0) return;
    const uint32_t ray = pt / S, n = pt - ray * S;
    const MonRay r = rays[ray];
    const uint32_t iter = ctrl->iter - 1;
    const float xi = mon_rand(jitter, seed, iter, 2, pt);
    const float t = mon_sample_t(r, n, xi, (float)S);
    const float bmin[3] = {bmin0, bmin1, bmin2}, bmax[3] = {bmax0, bmax1, bmax2};
    float u[3];
    mon_sample_point(r, t, bmin, bmax, u);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t l = lg * 4 + j;
        if (l >= g.n_levels || (gw[j] & 0x7fff7fffu) == 0) continue;
        const float g0 = __half2float(__ushort_as_half((unsigned short)(gw[j] & 0xffffu)));
        const float g1 = __half2float(__ushort_as_half((unsigned short)(gw[j] >> 16)));
        EncCorner c;
        level_corners(g, l, u, c);
        __half2* tab = reinterpret_cast<__half2*>(grid_grad) + g.offset[l];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const __half2 v = __halves2half2(__float2half_rn(__fmul_rn(g0, c.w[k])), __float2half_rn(__fmul_rn(g1, c.w[k])));
            atomicAdd(tab + c.idx[k], v);
        }
    }
}

void mon_launch_encode_forward(const MonGrid& g, uint32_t n_points, uint32_t S, const MonRay* rays, const int* in_box,
                               const float* jitter, uint32_t seed, const MonCtrl* ctrl, uint32_t rng_stream, uint32_t iter_fixed,
                               const float* bmin, const float* bmax, const __half* grid, __half* enc, cudaStream_t st) {
    const uint32_t blocks = (n_points + ENC_TILE - 1) / ENC_TILE;
    k_encode_forward<<<blocks, ENC_THREADS, 0, st>>>(g, n_points, S, rays, in_box, jitter, seed, ctrl, rng_stream, iter_fixed,
                                                    bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2], grid, enc);
}
void mon_launch_encode_backward(const MonGrid& g, uint32_t n_points, uint32_t S, const MonRay* rays,
                                const float* jitter, uint32_t seed, const MonCtrl* ctrl,
                                const float* bmin, const float* bmax, const __half* d_enc, __half* grid_grad, cudaStream_t st) {
    const uint32_t blocks = (n_points + ENC_TILE - 1) / ENC_TILE;
    k_encode_backward<<<blocks, ENC_THREADS, 0, st>>>(g, n_points, S, rays, jitter, seed, ctrl,
                                                     bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2], d_enc, grid_grad);
}

// stand-alone encode of explicit unit-cube positions (parity hook mon_stage_encode)
__global__ void k_encode_points(MonGrid g, uint32_t n_points, const float* __restrict__ pts,
                                const __half* __restrict__ grid, __half* __restrict__ enc) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pt = gid / g.n_levels, l = gid - pt * g.n_levels;
    if (pt >= n_points) return;
    const float u[3] = {pts[pt * 3], pts[pt * 3 + 1], pts[pt * 3 + 2]};
    EncCorner c;
    level_corners(g, l, u, c);
    const __half2* tab = reinterpret_cast<const __half2*>(grid) + g.offset[l];
    __half a0 = __float2half_rn(0.0f), a1 = a0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const __half2 v = __ldg(tab + c.idx[k]);
        const float wh = __half2float(__float2half_rn(c.w[k]));
        a0 = __float2half_rn(__fmaf_rn(wh, __low2float(v), __half2float(a0)));
        a1 = __float2half_rn(__fmaf_rn(wh, __high2float(v), __half2float(a1)));
    }
    enc[(size_t)pt * (2 * g.n_levels) + 2 * l] = a0;
    enc[(size_t)pt * (2 * g.n_levels) + 2 * l + 1] = a1;
}
void mon_launch_encode_points(const MonGrid& g, uint32_t n_points, const float* pts, const __half* grid, __half* enc, cudaStream_t st) {
    const uint32_t total = n_points * g.n_levels;
    k_encode_points<<<(total + 255) / 256, 256, 0, st>>>(g, n_points, pts, grid, enc);
}
