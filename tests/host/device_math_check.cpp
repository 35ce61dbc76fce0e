// device_math_check.cpp — ro_map_b200/csrc/mon_device.cuh (the arithmetic the kernels share: pixel -> ray, slab test,
// sampling, grid cell / hash index) compiled FOR THE HOST: the CUDA _rn intrinsics are IEEE single operations, restated
// below one to one (compile with -ffp-contract=off so that g++ fuses nothing on its own).  Prints the rays of a render
// window and the samples / cells of a few of them; tests/test_device_math_host.py holds the output against the CPU
// oracle bit for bit (and, through it, against the reference's golden vectors).  No GPU involved.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T> static inline T __ldg(const T* p) { return *p; }

#include "mon_device.cuh"

int main(int argc, char** argv) {
    // argv: bx by h w  K[4]  Twc[16]  Tow[16]  bmin[3] bmax[3]   (column-major matrices)
    if (argc != 1 + 4 + 4 + 16 + 16 + 6) { fprintf(stderr, "usage: device_math_check bx by h w K4 Twc16 Tow16 bmin3 bmax3\n"); return 2; }
    int a = 1;
    const int bx = atoi(argv[a++]), by = atoi(argv[a++]), h = atoi(argv[a++]), w = atoi(argv[a++]);
    float K[4], Twc[16], Tow[16], bmin[3], bmax[3];
    auto f = [&](const char* s) { uint32_t bits = (uint32_t)strtoul(s, nullptr, 16); float v; memcpy(&v, &bits, 4); return v; };   // floats travel as hex bit patterns
    for (float& v : K) v = f(argv[a++]);
    for (float& v : Twc) v = f(argv[a++]);
    for (float& v : Tow) v = f(argv[a++]);
    for (float& v : bmin) v = f(argv[a++]);
    for (float& v : bmax) v = f(argv[a++]);
    auto bits = [](float v) { uint32_t b; memcpy(&b, &v, 4); return b; };
    for (int i = 0; i < h * w; ++i) {
        MonRay r;
        memset(&r, 0, sizeof r);
        mon_pixel_ray((float)(bx + i % w), (float)(by + i / w), K, Twc, Tow, r.o, r.d, r.d_norm);
        float t0 = 0, t1 = 0;
        const bool hit = mon_ray_box(bmin, bmax, r.o, r.d, t0, t1);
        if (hit) { r.tmin = fmaxf(t0, 0.0f); r.tmax = t1; }
        printf("%d %08x %08x %08x %08x %08x %08x %08x %08x %08x", hit ? 1 : 0, bits(r.o[0]), bits(r.o[1]), bits(r.o[2]), bits(r.d[0]), bits(r.d[1]), bits(r.d[2]),
               bits(r.d_norm), bits(r.tmin), bits(r.tmax));
        if (hit) {   // sample 5 of 64 with xi = 0.625, its unit-cube position, and its cell / fraction / hash index on level 9 (scale 8191)
            const float t = mon_sample_t(r, 5, 0.625f, 64.0f);
            float u[3];
            mon_sample_point(r, t, bmin, bmax, u);
            float fr[3]; uint32_t cell[3];
            for (int k = 0; k < 3; ++k) mon_pos_fract(u[k], 8191.0f, fr[k], cell[k]);
            printf(" %08x %08x %08x %08x %08x %08x %08x %u", bits(t), bits(u[0]), bits(u[1]), bits(u[2]), bits(fr[0]), bits(fr[1]), bits(fr[2]),
                   mon_grid_index(true, 65536u, 8192u, cell[0] + 1, cell[1], cell[2] + 1));
        }
        printf("\n");
    }
    return 0;
}
