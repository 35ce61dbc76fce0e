// mon_device.cuh — device-side arithmetic shared by the kernels.  Every floating-point
// expression whose rounding decides an integer (pixel, grid cell, hash index) is written with
// explicit _rn intrinsics so that nvcc's contraction cannot change it (the parity tests state the
// same operation order on the CPU with contraction disabled).
#pragma once
#include "mon_types.h"
#include <float.h>

#define MON_DEV __device__ __forceinline__

// ---- counter-based RNG replacing the reference's 3 host cuRAND calls per iteration
// (nerf_model.cu:1432,1434,1468).  Streams: 0 = pixel xy, 1 = background colour, 2 = sample jitter.
MON_DEV uint32_t mon_hash4(uint32_t seed, uint32_t iter, uint32_t stream, uint32_t idx) {
    uint64_t x = ((uint64_t)seed << 32) ^ ((uint64_t)iter * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)stream << 59) ^ (uint64_t)idx;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (uint32_t)x;
}
// uniform in (0,1], the interval curandGenerateUniform documents
MON_DEV float mon_u01(uint32_t r) { return __fmul_rn((float)(r >> 8) + 1.0f, 1.0f / 16777216.0f); }

MON_DEV float mon_rand(const float* injected, uint32_t seed, uint32_t iter, uint32_t stream, uint32_t idx) {
    return injected ? __ldg(injected + idx) : mon_u01(mon_hash4(seed, iter, stream, idx));
}

// ---- A1: slab test (nerf_model.cu:87-138), IEEE division, no zero guard
MON_DEV bool mon_ray_box(const float* bmin, const float* bmax, const float* o, const float* d, float& t0, float& t1) {
    float tmin = __fdiv_rn(bmin[0] - o[0], d[0]);
    float tmax = __fdiv_rn(bmax[0] - o[0], d[0]);
    if (tmin > tmax) { float s = tmin; tmin = tmax; tmax = s; }
    float tymin = __fdiv_rn(bmin[1] - o[1], d[1]);
    float tymax = __fdiv_rn(bmax[1] - o[1], d[1]);
    if (tymin > tymax) { float s = tymin; tymin = tymax; tymax = s; }
    if (tmin > tymax || tymin > tmax) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = __fdiv_rn(bmin[2] - o[2], d[2]);
    float tzmax = __fdiv_rn(bmax[2] - o[2], d[2]);
    if (tzmin > tzmax) { float s = tzmin; tzmin = tzmax; tzmax = s; }
    if (tmin > tzmax || tzmin > tmax) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    t0 = tmin; t1 = tmax;
    return t0 != FLT_MAX;
}

// 3x3 block (column-major 4x4) times vector as the reference build evaluates Eigen's fixed-size product: the row reduction is
// x0 + (x1 + x2) (Eigen's unrolled redux splits the range in halves) and nvcc fuses every multiply-add of it —
// fma(m0, v0, fma(m1, v1, m2 * v2)).  Bit-identical to the rays of the reference's GenerateRays / GenerateRenderRays
// (tests/golden/romap_golden.npz, tests/test_device_math_host.py).
MON_DEV void mon_rot3(const float* M, const float* v, float* out) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out[r] = __fmaf_rn(M[r], v[0], __fmaf_rn(M[4 + r], v[1], __fmul_rn(M[8 + r], v[2])));
}

// pixel -> object-space ray (nerf_model.cu:403-413)
MON_DEV void mon_pixel_ray(float x, float y, const float* K, const float* Twc, const float* Tow,
                           float* o, float* d, float& d_norm) {
    float dir[3] = {__fdiv_rn(x - K[2], K[0]), __fdiv_rn(y - K[3], K[1]), 1.0f};
    d_norm = __fsqrt_rn(__fmaf_rn(dir[0], dir[0], __fmaf_rn(dir[1], dir[1], __fmul_rn(dir[2], dir[2]))));   // squaredNorm: the same reduction
    float dn[3] = {__fdiv_rn(dir[0], d_norm), __fdiv_rn(dir[1], d_norm), __fdiv_rn(dir[2], d_norm)};
    float dw[3]; mon_rot3(Twc, dn, dw);
    float ow[3] = {Twc[12], Twc[13], Twc[14]};
    mon_rot3(Tow, dw, d);
    float oo[3]; mon_rot3(Tow, ow, oo);
    o[0] = __fadd_rn(oo[0], Tow[12]); o[1] = __fadd_rn(oo[1], Tow[13]); o[2] = __fadd_rn(oo[2], Tow[14]);
}

// ---- A3: stratified sample n of S along the ray (nerf_model.cu:545-565)
MON_DEV float mon_sample_t(const MonRay& r, uint32_t n, float xi, float S) {
    const float dt = __fdiv_rn(r.tmax - r.tmin, S);
    return __fmaf_rn(dt, __fadd_rn((float)n, xi), r.tmin);
}
MON_DEV void mon_sample_point(const MonRay& r, float t, const float* bmin, const float* bmax, float* u) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float p = __fmaf_rn(t, r.d[k], r.o[k]);
        u[k] = __fdiv_rn(p - bmin[k], bmax[k] - bmin[k]);
    }
}

// ---- A4: grid cell + fractional position (common_device.h:485-495), index (grid.h:170-187)
MON_DEV void mon_pos_fract(float input, float scale, float& frac, uint32_t& cell) {
    const float pos = __fmaf_rn(input, scale, 0.5f);
    const int tmp = (int)floorf(pos);
    cell = (uint32_t)tmp;
    frac = __fsub_rn(pos, (float)tmp);
}

MON_DEV uint32_t mon_grid_index(bool hashed, uint32_t size, uint32_t res, uint32_t x, uint32_t y, uint32_t z) {
    // the dense form wraps in uint32 exactly like the reference's running stride (res == 65536: z*res*res == 0)
    const uint32_t index = hashed ? (x ^ (y * 2654435761u) ^ (z * 805459861u)) : (x + y * res + z * res * res);
    // every table of the supported configurations is a power of two (res^3 padded to 8, capped at 2^log2_hashmap_size);
    // the general modulo is kept for the ones that are not
    return (size & (size - 1)) == 0 ? (index & (size - 1)) : (index % size);
}
