// mon_oracle.cpp — CPU ORACLE (test infrastructure only; see mon_oracle.h for the contract,
// the reference file:line map and the pinning status).  Build: make -C oracle
// Compile with -ffp-contract=off: all fused multiply-adds are explicit fmaf() calls.
#include "mon_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- fp16 (IEEE binary16, RNE)
inline uint16_t f2h(float f) {
    _Float16 h = (_Float16)f;  // gcc: round-to-nearest-even, same as __float2half_rn
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
}
inline float h2f(uint16_t u) {
    _Float16 h;
    std::memcpy(&h, &u, 2);
    return (float)h;
}
inline float rh(float f) { return h2f(f2h(f)); }  // round through fp16

// ---------------------------------------------------------------- pcg32 (dependencies/pcg32/pcg32.h:46-170)
struct Pcg32 {
    uint64_t state, inc;
    static constexpr uint64_t MULT = 0x5851f42d4c957f2dULL;
    explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) {
        state = 0U;
        inc = (initseq << 1u) | 1u;
        next_uint();
        state += initstate;
        next_uint();
    }
    uint32_t next_uint() {
        uint64_t old = state;
        state = old * MULT + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    float next_float() {
        union { uint32_t u; float f; } x;
        x.u = (next_uint() >> 9) | 0x3f800000u;
        return x.f - 1.0f;
    }
    void advance(int64_t delta_) {
        uint64_t cur_mult = MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        uint64_t delta = (uint64_t)delta_;
        while (delta > 0) {
            if (delta & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
};

// ---------------------------------------------------------------- grid geometry (grid.h:195-204, 964-997)
struct Layout {
    uint32_t L = 0;
    uint32_t offsets[ORC_MAX_LEVELS + 1];
    float scale[ORC_MAX_LEVELS];
    uint32_t res[ORC_MAX_LEVELS];
    uint32_t n_grid_params = 0;
    uint32_t n_mlp = 0;
    uint32_t width = 64, in_w = 32, out_w = 16, n_hidden = 1;
};

inline uint32_t next_multiple(uint32_t v, uint32_t d) { return ((v + d - 1) / d) * d; }

Layout make_layout(const orc_config& c) {
    Layout ly;
    ly.L = c.n_levels;
    const float log2_pls = std::log2(c.per_level_scale);
    uint32_t offset = 0;
    for (uint32_t i = 0; i < ly.L; ++i) {
        // grid_scale(): exp2f(level * log2_per_level_scale) * base_resolution - 1.0f
        const float scale = exp2f((float)i * log2_pls) * (float)c.base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1;
        const uint32_t max_params = UINT32_MAX / 2;
        uint32_t params_in_level = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
        params_in_level = next_multiple(params_in_level, 8u);
        params_in_level = std::min(params_in_level, 1u << c.log2_hashmap_size);
        ly.offsets[i] = offset;
        ly.scale[i] = scale;
        ly.res[i] = res;
        offset += params_in_level;
    }
    ly.offsets[ly.L] = offset;
    ly.n_grid_params = offset * c.n_features;
    ly.width = c.n_neurons;
    ly.in_w = c.n_levels * c.n_features;
    ly.out_w = c.padded_output_width;
    ly.n_hidden = c.n_hidden_layers;
    ly.n_mlp = ly.width * ly.in_w + (ly.n_hidden - 1) * ly.width * ly.width + ly.out_w * ly.width;
    return ly;
}

// grid_index() + coherent_prime_hash() (grid.h:131-135, 170-187)
inline uint32_t grid_index(uint32_t hashmap_size, uint32_t res, const uint32_t p[3]) {
    uint32_t stride = 1, index = 0;
    for (uint32_t dim = 0; dim < 3 && stride <= hashmap_size; ++dim) {
        index += p[dim] * stride;
        stride *= res;
    }
    if (hashmap_size < stride) {
        index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
    }
    return index % hashmap_size;
}

// pos_fract() (common_device.h:485-495); nvcc fuses input*scale+0.5f
inline void pos_fract(float input, float scale, float* frac, uint32_t* cell) {
    float pos = fmaf(input, scale, 0.5f);
    int tmp = (int)floorf(pos);
    *cell = (uint32_t)tmp;
    *frac = pos - (float)tmp;
}

struct Corner { uint32_t idx; float w; };

inline void level_corners(const Layout& ly, uint32_t l, const float u[3], Corner out[8]) {
    const uint32_t size = ly.offsets[l + 1] - ly.offsets[l];
    float fr[3]; uint32_t g[3];
    for (int d = 0; d < 3; ++d) pos_fract(u[d], ly.scale[l], &fr[d], &g[d]);
    for (uint32_t c = 0; c < 8; ++c) {
        float w = 1.0f; uint32_t p[3];
        for (int d = 0; d < 3; ++d) {
            if ((c & (1u << d)) == 0) { w *= 1.0f - fr[d]; p[d] = g[d]; }
            else                      { w *= fr[d];        p[d] = g[d] + 1; }
        }
        out[c].idx = grid_index(size, ly.res[l], p);
        out[c].w = w;
    }
}

// ---------------------------------------------------------------- A1
inline bool ray_box(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float& t0, float& t1) {
    float tmin = (bmin[0] - o[0]) / d[0];
    float tmax = (bmax[0] - o[0]) / d[0];
    if (tmin > tmax) std::swap(tmin, tmax);
    float tymin = (bmin[1] - o[1]) / d[1];
    float tymax = (bmax[1] - o[1]) / d[1];
    if (tymin > tymax) std::swap(tymin, tymax);
    if (tmin > tymax || tymin > tmax) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (bmin[2] - o[2]) / d[2];
    float tzmax = (bmax[2] - o[2]) / d[2];
    if (tzmin > tzmax) std::swap(tzmin, tzmax);
    if (tmin > tzmax || tzmin > tmax) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    t0 = tmin; t1 = tmax;
    // the reference signals a miss with FLT_MAX in slot 0 and tests `!= FLT_MAX` (nerf_model.cu:416)
    return t0 != FLT_MAX;
}

// 3x3 (column-major 4x4 upper-left) times vector.  Eigen evaluates the fixed-size product row by row as the reduction
// x0 + (x1 + x2) (redux_novec_unroller halves the range) and nvcc's default --fmad=true fuses both multiply-adds:
// fma(c0, v0, fma(c1, v1, c2 * v2)).  Pinned bit for bit by the rays of the reference's own kernels (romap_golden.npz).
inline void rot3(const float M[16], const float v[3], float out[3]) {
    for (int r = 0; r < 3; ++r) out[r] = fmaf(M[0 * 4 + r], v[0], fmaf(M[1 * 4 + r], v[1], M[2 * 4 + r] * v[2]));
}

// camera pixel -> object-space ray (nerf_model.cu:403-413)
inline void pixel_ray(float x, float y, const float K[4], const float Twc[16], const float Tow[16],
                      float o[3], float d[3], float& d_norm) {
    float dir[3] = {(x - K[2]) / K[0], (y - K[3]) / K[1], 1.0f};
    d_norm = sqrtf(fmaf(dir[0], dir[0], fmaf(dir[1], dir[1], dir[2] * dir[2])));   // d.norm(): the same reduction, fused the same way
    float dn[3] = {dir[0] / d_norm, dir[1] / d_norm, dir[2] / d_norm};
    float dw[3]; rot3(Twc, dn, dw);
    float ow[3] = {Twc[12], Twc[13], Twc[14]};
    rot3(Tow, dw, d);
    float oo[3]; rot3(Tow, ow, oo);
    o[0] = oo[0] + Tow[12]; o[1] = oo[1] + Tow[13]; o[2] = oo[2] + Tow[14];
}

inline float logistic(float x) { return 1.0f / (1.0f + expf(-x)); }
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

// =====================================================================================
extern "C" {

void orc_default_config(orc_config* c) {
    c->n_levels = 16; c->n_features = 2; c->log2_hashmap_size = 16; c->base_resolution = 16;
    c->per_level_scale = 2.0f;
    c->n_neurons = 64; c->n_hidden_layers = 1; c->padded_output_width = 16;
    c->learning_rate = 1e-2f; c->beta1 = 0.9f; c->beta2 = 0.99f; c->epsilon = 1e-15f; c->l2_reg = 1e-6f;
    c->ema_decay = 0.95f; c->decay_start = 20000; c->decay_interval = 10000; c->decay_base = 0.33f;
    c->loss_scale = 128.0f;
}

uint32_t orc_grid_layout(const orc_config* cfg, uint32_t* offsets, float* scales, uint32_t* resolutions) {
    Layout ly = make_layout(*cfg);
    for (uint32_t i = 0; i <= ly.L; ++i) if (offsets) offsets[i] = ly.offsets[i];
    for (uint32_t i = 0; i < ly.L; ++i) { if (scales) scales[i] = ly.scale[i]; if (resolutions) resolutions[i] = ly.res[i]; }
    return ly.n_grid_params;
}
uint32_t orc_n_mlp_params(const orc_config* cfg) { return make_layout(*cfg).n_mlp; }
uint32_t orc_n_params(const orc_config* cfg) { Layout ly = make_layout(*cfg); return ly.n_mlp + ly.n_grid_params; }

void orc_seed_seq_1(uint32_t seed, uint32_t out[2]) {
    std::seed_seq seq{seed};
    std::vector<uint32_t> seeds(2);
    seq.generate(seeds.begin(), seeds.end());
    out[0] = seeds[0]; out[1] = seeds[1];
}

void orc_pcg32_floats(uint64_t initstate, uint64_t advance, uint32_t n, float* out) {
    Pcg32 r(initstate);
    r.advance((int64_t)advance);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.next_float();
}

// A12: Trainer ctor + FullyFusedMLP::initialize_params + GridEncoding::initialize_params
void orc_init_params(const orc_config* cfg, uint32_t seed, float* params) {
    Layout ly = make_layout(*cfg);
    uint32_t seeds[2]; orc_seed_seq_1(seed, seeds);
    Pcg32 rng(seeds[0]);
    // MLP: xavier uniform, host-side, sequential draws (gpu_matrix.h:291-303)
    float* p = params;
    auto xavier = [&](uint32_t rows, uint32_t cols) {
        float scale = 1.0f;
        scale *= std::sqrt(6.0f / (float)(rows + cols));
        for (uint32_t i = 0; i < rows * cols; ++i) p[i] = rng.next_float() * 2.0f * scale - scale;
        p += rows * cols;
    };
    xavier(ly.width, ly.in_w);
    for (uint32_t i = 0; i + 1 < ly.n_hidden; ++i) xavier(ly.width, ly.width);
    xavier(ly.out_w, ly.width);
    // grid: generate_random_uniform on the device (random.h:66-92): N_TO_GENERATE=4 per thread,
    // thread i copies the rng, advance(4*i), writes out[i + n_threads*j]; device lambda
    // val*(upper-lower)+lower is fused by nvcc into one FMA.
    const size_t n = ly.n_grid_params;
    const size_t n_threads_needed = (n + 3) / 4;
    const size_t n_blocks = (n_threads_needed + 127) / 128;
    const size_t n_threads = n_blocks * 128;
    const float lower = -1e-4f, upper = 1e-4f;
    const float range = upper - lower;
    for (size_t i = 0; i < n_threads; ++i) {
        Pcg32 r = rng;
        r.advance((int64_t)(i * 4));
        for (size_t j = 0; j < 4; ++j) {
            size_t idx = i + n_threads * j;
            if (idx >= n) break;
            p[idx] = fmaf(r.next_float(), range, lower);
        }
    }
}

uint16_t orc_f2h(float f) { return f2h(f); }
float orc_h2f(uint16_t h) { return h2f(h); }
void orc_f2h_array(const float* in, uint16_t* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = f2h(in[i]); }
void orc_h2f_array(const uint16_t* in, float* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = h2f(in[i]); }

int orc_ray_intersect(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float* tmin, float* tmax) {
    float a = FLT_MAX, b = FLT_MAX;
    bool hit = ray_box(bmin, bmax, o, d, a, b);
    *tmin = hit ? a : FLT_MAX; *tmax = hit ? b : FLT_MAX;
    return hit ? 1 : 0;
}

uint32_t orc_generate_rays(uint32_t R, const orc_bbox2d* boxes, uint32_t n_boxes,
                           const orc_frame* frames, int H, int W, const float K[4],
                           const float Tow[16], const float bmin[3], const float bmax[3],
                           uint8_t obj_instance, int use_depth,
                           const float* sample_xy, const float* rand_colors,
                           orc_ray* rays, uint8_t* rays_instance, float* target_rgb, float* target_depth) {
    (void)H;
    uint32_t n_in = 0;
    for (uint32_t i = 0; i < R; ++i) {
        const orc_bbox2d& b = boxes[i % n_boxes];
        const orc_frame& fr = frames[b.FrameId];
        const int h = (int)b.h, w = (int)b.w;
        const uint32_t x = b.x + (uint32_t)(sample_xy[2 * i] * (float)w);
        const uint32_t y = b.y + (uint32_t)(sample_xy[2 * i + 1] * (float)h);
        const uint8_t inst = fr.instance[(size_t)y * W + x];
        if (inst != 0 && inst != obj_instance) continue;  // occluded by another object
        float o[3], d[3], d_norm;
        pixel_ray((float)x, (float)y, K, fr.pose, Tow, o, d, d_norm);
        float t0, t1;
        if (!ray_box(bmin, bmax, o, d, t0, t1)) continue;
        const uint32_t idx = n_in++;
        orc_ray& r = rays[idx];
        for (int k = 0; k < 3; ++k) { r.o[k] = o[k]; r.d[k] = d[k]; }
        r.d_norm = d_norm; r.tmin = fmaxf(t0, 0.0f); r.tmax = t1;
        if (inst != 0) {
            const uint8_t* px = fr.rgb + ((size_t)y * W + x) * 3;
            for (int k = 0; k < 3; ++k) target_rgb[idx * 3 + k] = (float)px[k] * (float)(1.0 / 255.0);
            target_depth[idx] = (use_depth && fr.depth) ? fr.depth[(size_t)y * W + x] * d_norm : 0.0f;
            rays_instance[idx] = 1;
        } else {
            for (int k = 0; k < 3; ++k) target_rgb[idx * 3 + k] = rand_colors[idx * 3 + k];
            target_depth[idx] = 0.0f;
            rays_instance[idx] = 0;
        }
    }
    // fill_rollover_rays (nerf_model.cu:280-294); n_in == 0 is undefined in the reference (mod 0):
    // defined here as "leave outputs untouched, caller skips the iteration".
    if (n_in > 0) {
        for (uint32_t i = n_in; i < R; ++i) {
            const uint32_t s = i % n_in;
            rays[i] = rays[s];
            rays_instance[i] = rays_instance[s];
            for (int k = 0; k < 3; ++k) target_rgb[i * 3 + k] = target_rgb[s * 3 + k];
            target_depth[i] = target_depth[s];
        }
    }
    return n_in;
}

void orc_sample_points(uint32_t R, uint32_t S, const orc_ray* rays, const float bmin[3], const float bmax[3],
                       const float* rand_dt, float* points, float* t_out) {
    for (uint32_t i = 0; i < R; ++i) {
        const orc_ray& r = rays[i];
        const float dt = (r.tmax - r.tmin) / (float)S;
        for (uint32_t n = 0; n < S; ++n) {
            const float t = fmaf(dt, (float)n + rand_dt[(size_t)i * S + n], r.tmin);
            const size_t pi = (size_t)i * S + n;
            for (int k = 0; k < 3; ++k) {
                const float p = fmaf(t, r.d[k], r.o[k]);
                points[pi * 3 + k] = (p - bmin[k]) / (bmax[k] - bmin[k]);
            }
            t_out[pi] = t;
        }
    }
}

void orc_encode(const orc_config* cfg, const uint16_t* grid, const float* points, uint32_t N, uint16_t* enc) {
    Layout ly = make_layout(*cfg);
    const uint32_t C = ly.in_w;
    for (uint32_t i = 0; i < N; ++i) {
        for (uint32_t l = 0; l < ly.L; ++l) {
            Corner cn[8];
            level_corners(ly, l, points + (size_t)i * 3, cn);
            const uint16_t* tab = grid + (size_t)ly.offsets[l] * 2;
            uint16_t acc0 = 0, acc1 = 0;  // vector_t<__half,2> result = {}
            for (int c = 0; c < 8; ++c) {
                const float wh = rh(cn[c].w);  // (T)weight
                acc0 = f2h(fmaf(wh, h2f(tab[cn[c].idx * 2 + 0]), h2f(acc0)));
                acc1 = f2h(fmaf(wh, h2f(tab[cn[c].idx * 2 + 1]), h2f(acc1)));
            }
            enc[(size_t)i * C + 2 * l + 0] = acc0;
            enc[(size_t)i * C + 2 * l + 1] = acc1;
        }
    }
}

void orc_encode_corners(const orc_config* cfg, const float* points, uint32_t N, uint32_t* idx, float* w) {
    Layout ly = make_layout(*cfg);
    for (uint32_t i = 0; i < N; ++i)
        for (uint32_t l = 0; l < ly.L; ++l) {
            Corner cn[8];
            level_corners(ly, l, points + (size_t)i * 3, cn);
            for (int c = 0; c < 8; ++c) {
                idx[((size_t)i * ly.L + l) * 8 + c] = cn[c].idx;
                w[((size_t)i * ly.L + l) * 8 + c] = cn[c].w;
            }
        }
}

void orc_mlp_forward(const orc_config* cfg, const uint16_t* mlp, const uint16_t* enc, uint32_t N,
                     uint16_t* hidden, uint16_t* out) {
    Layout ly = make_layout(*cfg);
    const uint32_t Wd = ly.width, In = ly.in_w, Out = ly.out_w, NH = ly.n_hidden;
    std::vector<float> wf(ly.n_mlp);
    for (uint32_t i = 0; i < ly.n_mlp; ++i) wf[i] = h2f(mlp[i]);
    std::vector<float> a(std::max(Wd, In)), b(Wd);
    for (uint32_t i = 0; i < N; ++i) {
        for (uint32_t k = 0; k < In; ++k) a[k] = h2f(enc[(size_t)i * In + k]);
        const float* Wl = wf.data();
        uint32_t fan_in = In;
        for (uint32_t layer = 0; layer < NH; ++layer) {
            for (uint32_t j = 0; j < Wd; ++j) {
                float acc = 0.0f;
                for (uint32_t k = 0; k < fan_in; ++k) acc = fmaf(Wl[j * fan_in + k], a[k], acc);
                b[j] = rh(fmaxf(acc, 0.0f));  // ReLU, stored fp16
            }
            if (hidden) for (uint32_t j = 0; j < Wd; ++j) hidden[((size_t)layer * N + i) * Wd + j] = f2h(b[j]);
            for (uint32_t j = 0; j < Wd; ++j) a[j] = b[j];
            Wl += Wd * fan_in;
            fan_in = Wd;
        }
        for (uint32_t o = 0; o < Out; ++o) {
            float acc = 0.0f;
            for (uint32_t k = 0; k < Wd; ++k) acc = fmaf(Wl[o * Wd + k], a[k], acc);
            out[(size_t)i * Out + o] = f2h(acc);
        }
    }
}

void orc_volume_render(uint32_t R, uint32_t S, const uint16_t* out, const float* t, const float* bg,
                       float* rgb_rays, float* depth_rays, float* mask_rays) {
    for (uint32_t i = 0; i < R; ++i) {
        float T = 1.0f, C[3] = {0, 0, 0}, D = 0.0f, last = 0.0f;
        for (uint32_t n = 0; n < S; ++n) {
            if (T < 1e-4f) break;
            const uint16_t* o = out + ((size_t)i * S + n) * 16;
            const float cur = t[(size_t)i * S + n];
            const float dt = cur - last;
            const float density = expf(h2f(o[3]));
            const float alpha = 1.0f - expf(-density * dt);
            const float w = alpha * T;
            for (int k = 0; k < 3; ++k) C[k] += w * logistic(h2f(o[k]));
            D += w * cur;
            T *= (1.0f - alpha);
            last = cur;
        }
        for (int k = 0; k < 3; ++k) rgb_rays[i * 3 + k] = C[k] + T * bg[i * 3 + k];
        depth_rays[i] = D;
        mask_rays[i] = 1.0f - T;
    }
}

static void loss_backward_k(uint32_t R, uint32_t S, float k, const uint16_t* out, const float* t,
                       const uint8_t* rays_instance, const float* target_rgb, const float* target_depth,
                       const float* rgb_rays, const float* depth_rays, const float* mask_rays,
                       uint16_t* dout, float* loss) {
    std::memset(dout, 0, (size_t)R * S * 16 * sizeof(uint16_t));  // memset_async(dloss_dout) nerf_model.cu:1578
    for (uint32_t i = 0; i < R; ++i) {
        float g[3], mean_loss = 0.0f;
        for (int c = 0; c < 3; ++c) {
            const float diff = rgb_rays[i * 3 + c] - target_rgb[i * 3 + c];
            g[c] = 2.0f * diff;
            mean_loss += diff * diff;
        }
        mean_loss /= 3.0f;
        const float Dt = target_depth[i], Dr = depth_rays[i], mask = mask_rays[i];
        float dd = 0.0f;
        if (Dt > 0.0f) dd = 0.5f * (Dr - Dt >= 0.0f ? 1.0f : -1.0f);
        const bool is_obj = rays_instance[i] == 1;
        loss[i] = is_obj ? mean_loss + dd * (Dr - Dt) + (1.0f - mask) : mean_loss + mask;

        float T = 1.0f, C2[3] = {0, 0, 0}, D2 = 0.0f, last = 0.0f;
        for (uint32_t n = 0; n < S; ++n) {
            if (T < 1e-4f) break;
            const uint16_t* o = out + ((size_t)i * S + n) * 16;
            uint16_t* go = dout + ((size_t)i * S + n) * 16;
            const float cur = t[(size_t)i * S + n];
            const float dt = cur - last;
            float rgb[3];
            for (int c = 0; c < 3; ++c) rgb[c] = logistic(h2f(o[c]));
            const float density = expf(h2f(o[3]));
            const float alpha = 1.0f - expf(-density * dt);
            const float w = alpha * T;
            for (int c = 0; c < 3; ++c) C2[c] += w * rgb[c];
            D2 += w * cur;
            T *= (1.0f - alpha);
            for (int c = 0; c < 3; ++c) {
                const float s = logistic(h2f(o[c]));
                go[c] = f2h(k * ((w * g[c]) * (s * (1.0f - s))));
            }
            const float dsig = expf(clampf(h2f(o[3]), -15.0f, 15.0f));
            const float depth_sup = dd * (T * cur - (Dr - D2));
            const float dmask_dsig = 1.0f - mask;
            float dmlp;
            if (is_obj) {
                const float dmask = 0.5f * (mask >= 1.0f ? 1.0f : -1.0f);
                float dot = 0.0f;
                for (int c = 0; c < 3; ++c) dot += g[c] * (T * rgb[c] - (rgb_rays[i * 3 + c] - C2[c]));
                dmlp = dsig * dt * (dot + depth_sup + dmask * dmask_dsig);
            } else {
                const float dmask = 0.5f * (mask >= 0.0f ? 1.0f : -1.0f);
                dmlp = dsig * dt * dmask * dmask_dsig + dsig * 0.01f;
            }
            go[3] = f2h(k * dmlp);
            last = cur;
        }
    }
}

void orc_loss_backward(uint32_t R, uint32_t S, float loss_scale, const uint16_t* out, const float* t,
                       const uint8_t* rays_instance, const float* target_rgb, const float* target_depth,
                       const float* rgb_rays, const float* depth_rays, const float* mask_rays,
                       uint16_t* dout, float* loss) {
    // loss_scale /= nRays (nerf_model.cu:882)
    loss_backward_k(R, S, loss_scale / (float)R, out, t, rays_instance, target_rgb, target_depth,
                    rgb_rays, depth_rays, mask_rays, dout, loss);
}

void orc_mlp_backward(const orc_config* cfg, const uint16_t* mlp, const uint16_t* enc,
                      const uint16_t* hidden, const uint16_t* dout, uint32_t N,
                      uint16_t* d_enc, float* dW, int round_fp16) {
    Layout ly = make_layout(*cfg);
    const uint32_t Wd = ly.width, In = ly.in_w, Out = ly.out_w, NH = ly.n_hidden;
    std::vector<float> wf(ly.n_mlp);
    for (uint32_t i = 0; i < ly.n_mlp; ++i) wf[i] = h2f(mlp[i]);
    std::vector<double> acc(ly.n_mlp, 0.0);  // wide accumulator: the "true" sum the tolerance is stated against
    // weight-matrix offsets
    std::vector<uint32_t> woff(NH + 1);
    woff[0] = 0;
    woff[1] = Wd * In;
    for (uint32_t l = 1; l < NH; ++l) woff[l + 1] = woff[l] + Wd * Wd;
    const uint32_t out_off = woff[NH];
    std::vector<float> g(Wd), gprev(Wd), go(Out);
    for (uint32_t i = 0; i < N; ++i) {
        for (uint32_t o = 0; o < Out; ++o) go[o] = h2f(dout[(size_t)i * Out + o]);
        const uint16_t* hl = hidden + ((size_t)(NH - 1) * N + i) * Wd;
        // output layer: dW_out = dout * hid^T ; d_hid = W_out^T dout, masked by ReLU
        for (uint32_t o = 0; o < Out; ++o) {
            if (go[o] == 0.0f) continue;
            for (uint32_t j = 0; j < Wd; ++j) acc[out_off + o * Wd + j] += (double)go[o] * (double)h2f(hl[j]);
        }
        for (uint32_t j = 0; j < Wd; ++j) {
            float s = 0.0f;
            for (uint32_t o = 0; o < Out; ++o) s = fmaf(wf[out_off + o * Wd + j], go[o], s);
            g[j] = rh(h2f(hl[j]) > 0.0f ? s : 0.0f);
        }
        // hidden layers, last to first
        for (int l = (int)NH - 1; l >= 1; --l) {
            const uint16_t* hprev = hidden + ((size_t)(l - 1) * N + i) * Wd;
            const uint32_t off = woff[l];
            for (uint32_t j = 0; j < Wd; ++j) {
                if (g[j] == 0.0f) continue;
                for (uint32_t k2 = 0; k2 < Wd; ++k2) acc[off + j * Wd + k2] += (double)g[j] * (double)h2f(hprev[k2]);
            }
            for (uint32_t k2 = 0; k2 < Wd; ++k2) {
                float s = 0.0f;
                for (uint32_t j = 0; j < Wd; ++j) s = fmaf(wf[off + j * Wd + k2], g[j], s);
                gprev[k2] = rh(h2f(hprev[k2]) > 0.0f ? s : 0.0f);
            }
            g = gprev;
        }
        // input layer
        for (uint32_t j = 0; j < Wd; ++j) {
            if (g[j] == 0.0f) continue;
            for (uint32_t k2 = 0; k2 < In; ++k2) acc[j * In + k2] += (double)g[j] * (double)h2f(enc[(size_t)i * In + k2]);
        }
        if (d_enc) {
            for (uint32_t k2 = 0; k2 < In; ++k2) {
                float s = 0.0f;
                for (uint32_t j = 0; j < Wd; ++j) s = fmaf(wf[j * In + k2], g[j], s);
                d_enc[(size_t)i * In + k2] = f2h(s);
            }
        }
    }
    for (uint32_t i = 0; i < ly.n_mlp; ++i) dW[i] = round_fp16 ? rh((float)acc[i]) : (float)acc[i];
}

void orc_encode_backward(const orc_config* cfg, const float* points, const uint16_t* d_enc, uint32_t N,
                         float* grid_grad, int mode) {
    Layout ly = make_layout(*cfg);
    const uint32_t C = ly.in_w;
    std::memset(grid_grad, 0, (size_t)ly.n_grid_params * sizeof(float));
    for (uint32_t i = 0; i < N; ++i) {
        for (uint32_t l = 0; l < ly.L; ++l) {
            Corner cn[8];
            level_corners(ly, l, points + (size_t)i * 3, cn);
            float* tab = grid_grad + (size_t)ly.offsets[l] * 2;
            const float g0 = h2f(d_enc[(size_t)i * C + 2 * l]), g1 = h2f(d_enc[(size_t)i * C + 2 * l + 1]);
            for (int c = 0; c < 8; ++c) {
                // __half2 v = {(__half)((float)grad[f] * weight), ...}; atomicAdd(__half2*)
                const float v0 = rh(g0 * cn[c].w), v1 = rh(g1 * cn[c].w);
                float* e = tab + (size_t)cn[c].idx * 2;
                if (mode == 0) { e[0] = rh(e[0] + v0); e[1] = rh(e[1] + v1); }
                else           { e[0] += v0;           e[1] += v1; }
            }
        }
    }
}

void orc_optimizer_step(const orc_config* cfg, uint32_t step, const float* grads,
                        float* pf, uint16_t* ph, float* m, float* v, uint32_t* psteps, uint16_t* ema) {
    Layout ly = make_layout(*cfg);
    const uint32_t P = ly.n_mlp + ly.n_grid_params;
    // ExponentialDecayOptimizer::step (exponential_decay.h:60-71): evaluated with the nested
    // step counter BEFORE the Adam increment, i.e. step-1.
    float lr_factor = 1.0f;
    if (step - 1 >= cfg->decay_start) {
        const uint32_t n_decays = (step - 1 - cfg->decay_start) / cfg->decay_interval + 1;
        for (uint32_t s = 0; s < n_decays; ++s) lr_factor *= cfg->decay_base;
    }
    const float base_lr = cfg->learning_rate * lr_factor;
    const float b1 = cfg->beta1, b2 = cfg->beta2;
    for (uint32_t i = 0; i < P; ++i) {
        float gradient = grads[i] / cfg->loss_scale;
        if (i >= ly.n_mlp) { if (gradient == 0.0f) continue; }
        const float w = pf[i];
        if (i < ly.n_mlp) gradient = fmaf(cfg->l2_reg, w, gradient);
        const float gsq = gradient * gradient;
        const float fm = m[i] = fmaf(b1, m[i], (1.0f - b1) * gradient);
        const float sm = v[i] = fmaf(b2, v[i], (1.0f - b2) * gsq);
        float lr = base_lr;  // non_matrix_learning_rate_factor = 1
        const uint32_t cs = ++psteps[i];
        lr *= sqrtf(1.0f - powf(b2, (float)cs)) / (1.0f - powf(b1, (float)cs));
        const float eff = fminf(fmaxf(lr / (sqrtf(sm) + cfg->epsilon), 0.0f), FLT_MAX);
        // weight_decay(0,0,w) = (1-0)*w - copysignf(0,w) = w
        const float nw = fmaf(-eff, fm, w);
        pf[i] = nw;
        ph[i] = f2h(nw);
    }
    // EMA (ema.h:102-136), all params, global step
    const float old_db = 1.0f - (float)std::pow((double)cfg->ema_decay, (double)(step - 1));
    const float new_db = 1.0f / (1.0f - (float)std::pow((double)cfg->ema_decay, (double)step));
    for (uint32_t i = 0; i < P; ++i) {
        const float f = fmaf(h2f(ph[i]), 1.0f - cfg->ema_decay, (h2f(ema[i]) * cfg->ema_decay) * old_db) * new_db;
        ema[i] = f2h(f);
    }
}

void orc_volume_render_test(uint32_t n_rays, uint32_t S2, const float* out4, const float* t,
                            const int* in_box, const float* d_norm, float bg,
                            float* rgb, float* depth, float* mask) {
    for (uint32_t i = 0; i < n_rays; ++i) {
        if (!in_box[i]) { rgb[i*3] = rgb[i*3+1] = rgb[i*3+2] = bg; depth[i] = 0; mask[i] = 0; continue; }
        float T = 1.0f, C[3] = {0,0,0}, D = 0.0f, last = 0.0f;
        for (uint32_t n = 0; n < S2; ++n) {
            if (T < 1e-4f) break;
            const float* o = out4 + ((size_t)i * S2 + n) * 4;
            const float cur = t[(size_t)i * S2 + n];
            const float dt = cur - last;
            const float alpha = 1.0f - expf(-expf(o[3]) * dt);
            const float w = alpha * T;
            for (int k = 0; k < 3; ++k) C[k] += w * logistic(o[k]);
            D += w * cur;
            T *= (1.0f - alpha);
            last = cur;
        }
        if (1.0f - T > 0.5f) {
            for (int k = 0; k < 3; ++k) rgb[i*3+k] = C[k] + T * bg;
            depth[i] = D / d_norm[i]; mask[i] = 1.0f;
        } else { rgb[i*3] = rgb[i*3+1] = rgb[i*3+2] = bg; depth[i] = 0; mask[i] = 0; }
    }
}

// =====================================================================================
// A13: whole object
struct orc_object {
    orc_config cfg; Layout ly;
    uint32_t R, S, N, P;
    float Tow[16], bmin[3], bmax[3];
    uint8_t inst; int use_depth; int n_threads;
    uint32_t step = 0;
    std::vector<float> pf, m, v, grad;
    std::vector<uint16_t> ph, ema;
    std::vector<uint32_t> psteps;
    // batch
    std::vector<orc_ray> rays; std::vector<uint8_t> rinst;
    std::vector<float> tgt, tgtd, points, t, rgb_r, dep_r, mask_r, loss, bg;
    std::vector<uint16_t> enc, hidden, out, dout, denc;
    std::vector<float> sxy, rcol, rdt;
};

}  // extern "C"

static void parallel_rays(int n_threads, uint32_t R, const std::function<void(uint32_t, uint32_t, int)>& fn) {
    if (n_threads <= 1) { fn(0, R, 0); return; }
    std::vector<std::thread> th;
    const uint32_t chunk = (R + n_threads - 1) / n_threads;
    for (int k = 0; k < n_threads; ++k) {
        uint32_t a = std::min(R, k * chunk), b = std::min(R, a + chunk);
        th.emplace_back(fn, a, b, k);
    }
    for (auto& x : th) x.join();
}

extern "C" {

orc_object* orc_object_create(const orc_config* cfg, uint32_t seed, uint32_t R, uint32_t S,
                              const float Tow[16], const float bmin[3], const float bmax[3],
                              uint8_t inst, int use_depth, int n_threads) {
    orc_object* o = new orc_object();
    o->cfg = *cfg; o->ly = make_layout(*cfg);
    o->R = R; o->S = S; o->N = R * S; o->P = o->ly.n_mlp + o->ly.n_grid_params;
    std::memcpy(o->Tow, Tow, 64); std::memcpy(o->bmin, bmin, 12); std::memcpy(o->bmax, bmax, 12);
    o->inst = inst; o->use_depth = use_depth; o->n_threads = std::max(1, n_threads);
    const uint32_t P = o->P, N = o->N;
    o->pf.resize(P); o->m.assign(P, 0.f); o->v.assign(P, 0.f); o->grad.assign(P, 0.f);
    o->ph.resize(P); o->ema.assign(P, 0); o->psteps.assign(P, 0);
    orc_init_params(cfg, seed, o->pf.data());
    orc_f2h_array(o->pf.data(), o->ph.data(), P);
    o->rays.resize(R); o->rinst.resize(R); o->tgt.resize(3 * R); o->tgtd.resize(R);
    o->points.resize(3 * (size_t)N); o->t.resize(N); o->rgb_r.resize(3 * R); o->dep_r.resize(R); o->mask_r.resize(R);
    o->loss.resize(R); o->bg.resize(3 * R);
    o->enc.resize((size_t)N * o->ly.in_w); o->hidden.resize((size_t)N * o->ly.width * o->ly.n_hidden);
    o->out.resize((size_t)N * 16); o->dout.resize((size_t)N * 16); o->denc.resize((size_t)N * o->ly.in_w);
    return o;
}
void orc_object_destroy(orc_object* o) { delete o; }
uint32_t orc_object_n_params(const orc_object* o) { return o->P; }

void orc_object_get(const orc_object* o, int which, float* out) {
    for (uint32_t i = 0; i < o->P; ++i) {
        switch (which) {
            case 0: out[i] = o->pf[i]; break;
            case 1: out[i] = h2f(o->ph[i]); break;
            case 2: out[i] = h2f(o->ema[i]); break;
            case 3: out[i] = o->grad[i]; break;
            case 4: out[i] = o->m[i]; break;
            case 5: out[i] = o->v[i]; break;
            case 6: out[i] = (float)o->psteps[i]; break;
        }
    }
}
void orc_object_set_params(orc_object* o, const float* p) {
    std::memcpy(o->pf.data(), p, (size_t)o->P * 4);
    orc_f2h_array(o->pf.data(), o->ph.data(), o->P);
}

static float train_iter_body(orc_object* o, const orc_bbox2d* boxes, uint32_t n_boxes,
                             const orc_frame* frames, int H, int W, const float K[4],
                             const float* sxy, const float* rcol, const float* rdt, uint32_t* n_in_out) {
    const uint32_t R = o->R, S = o->S, N = o->N;
    const Layout& ly = o->ly;
    const uint32_t n_in = orc_generate_rays(R, boxes, n_boxes, frames, H, W, K, o->Tow, o->bmin, o->bmax, o->inst,
                                            o->use_depth, sxy, rcol, o->rays.data(), o->rinst.data(),
                                            o->tgt.data(), o->tgtd.data());
    if (n_in_out) *n_in_out = n_in;
    if (n_in == 0) return 0.0f;
    // background colour for ray i is RandColors[(i % n_in)*3..] (nerf_model.cu:760)
    for (uint32_t i = 0; i < R; ++i) for (int k = 0; k < 3; ++k) o->bg[i * 3 + k] = rcol[(i % n_in) * 3 + k];
    const uint16_t* mlp = o->ph.data();
    const uint16_t* grid = o->ph.data() + ly.n_mlp;
    const int T = o->n_threads;
    // forward + loss, parallel over rays
    parallel_rays(T, R, [&](uint32_t a, uint32_t b, int) {
        if (a >= b) return;
        const uint32_t n = b - a;
        orc_sample_points(n, S, o->rays.data() + a, o->bmin, o->bmax, rdt + (size_t)a * S,
                          o->points.data() + (size_t)a * S * 3, o->t.data() + (size_t)a * S);
        orc_encode(&o->cfg, grid, o->points.data() + (size_t)a * S * 3, n * S, o->enc.data() + (size_t)a * S * ly.in_w);
    });
    // the hidden buffer is [layer][N][64]; run the MLP per thread-chunk on chunk-local views
    if (ly.n_hidden == 1) {
        parallel_rays(T, R, [&](uint32_t a, uint32_t b, int) {
            if (a >= b) return;
            const uint32_t n = (b - a) * S;
            orc_mlp_forward(&o->cfg, mlp, o->enc.data() + (size_t)a * S * ly.in_w, n,
                            o->hidden.data() + (size_t)a * S * ly.width, o->out.data() + (size_t)a * S * 16);
        });
    } else {
        orc_mlp_forward(&o->cfg, mlp, o->enc.data(), N, o->hidden.data(), o->out.data());
    }
    parallel_rays(T, R, [&](uint32_t a, uint32_t b, int) {
        if (a >= b) return;
        const uint32_t n = b - a;
        orc_volume_render(n, S, o->out.data() + (size_t)a * S * 16, o->t.data() + (size_t)a * S, o->bg.data() + a * 3,
                          o->rgb_r.data() + a * 3, o->dep_r.data() + a, o->mask_r.data() + a);
    });
    // loss/backward uses loss_scale / R with the FULL R (not the chunk) -> call per chunk with scaled loss_scale
    parallel_rays(T, R, [&](uint32_t a, uint32_t b, int) {
        if (a >= b) return;
        const uint32_t n = b - a;
        loss_backward_k(n, S, o->cfg.loss_scale / (float)R, o->out.data() + (size_t)a * S * 16, o->t.data() + (size_t)a * S,
                          o->rinst.data() + a, o->tgt.data() + a * 3, o->tgtd.data() + a,
                          o->rgb_r.data() + a * 3, o->dep_r.data() + a, o->mask_r.data() + a,
                          o->dout.data() + (size_t)a * S * 16, o->loss.data() + a);
    });
    // MLP backward: deterministic single accumulation when T==1; with threads, per-thread partials summed in order
    std::vector<std::vector<float>> dWp(T, std::vector<float>(ly.n_mlp, 0.f));
    if (ly.n_hidden == 1) {
        parallel_rays(T, R, [&](uint32_t a, uint32_t b, int k) {
            if (a >= b) return;
            const uint32_t n = (b - a) * S;
            orc_mlp_backward(&o->cfg, mlp, o->enc.data() + (size_t)a * S * ly.in_w, o->hidden.data() + (size_t)a * S * ly.width,
                             o->dout.data() + (size_t)a * S * 16, n, o->denc.data() + (size_t)a * S * ly.in_w, dWp[k].data(), 0);
        });
    } else {
        orc_mlp_backward(&o->cfg, mlp, o->enc.data(), o->hidden.data(), o->dout.data(), N, o->denc.data(), dWp[0].data(), 0);
    }
    for (uint32_t i = 0; i < ly.n_mlp; ++i) {
        float s = 0.f;
        for (int k = 0; k < T; ++k) s += dWp[k][i];
        o->grad[i] = rh(s);
    }
    // grid backward: T==1 -> reference-faithful fp16 sequential accumulation; T>1 -> fp32 partials
    if (T == 1) {
        orc_encode_backward(&o->cfg, o->points.data(), o->denc.data(), N, o->grad.data() + ly.n_mlp, 0);
    } else {
        std::vector<std::vector<float>> gp(T);
        parallel_rays(T, R, [&](uint32_t a, uint32_t b, int k) {
            gp[k].assign(ly.n_grid_params, 0.f);
            if (a >= b) return;
            orc_encode_backward(&o->cfg, o->points.data() + (size_t)a * S * 3, o->denc.data() + (size_t)a * S * ly.in_w,
                                (b - a) * S, gp[k].data(), 1);
        });
        float* gg = o->grad.data() + ly.n_mlp;
        parallel_rays(T, ly.n_grid_params, [&](uint32_t a, uint32_t b, int) {
            for (uint32_t i = a; i < b; ++i) { float s = 0.f; for (int k = 0; k < T; ++k) s += gp[k][i]; gg[i] = rh(s); }
        });
    }
    o->step += 1;
    orc_optimizer_step(&o->cfg, o->step, o->grad.data(), o->pf.data(), o->ph.data(), o->m.data(), o->v.data(),
                       o->psteps.data(), o->ema.data());
    double sum = 0.0;
    for (uint32_t i = 0; i < R; ++i) sum += o->loss[i];
    return (float)(sum / (double)R);
}

float orc_object_train_iter(orc_object* o, const orc_bbox2d* boxes, uint32_t n_boxes,
                            const orc_frame* frames, int H, int W, const float K[4],
                            const float* sxy, const float* rcol, const float* rdt, uint32_t* n_in_out) {
    return train_iter_body(o, boxes, n_boxes, frames, H, W, K, sxy, rcol, rdt, n_in_out);
}

static inline uint32_t hash32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (uint32_t)x;
}
static inline float u01_open_closed(uint32_t r) { return ((float)(r >> 8) + 1.0f) * (1.0f / 16777216.0f); }  // (0,1]

float orc_object_train_iter_rng(orc_object* o, const orc_bbox2d* boxes, uint32_t n_boxes,
                                const orc_frame* frames, int H, int W, const float K[4], uint64_t iter_seed) {
    const uint32_t R = o->R, S = o->S;
    o->sxy.resize(2 * R); o->rcol.resize(3 * R); o->rdt.resize((size_t)R * S);
    for (uint32_t i = 0; i < 2 * R; ++i) o->sxy[i] = u01_open_closed(hash32(iter_seed * 0x9E3779B97F4A7C15ULL + i));
    for (uint32_t i = 0; i < 3 * R; ++i) o->rcol[i] = u01_open_closed(hash32(iter_seed * 0x9E3779B97F4A7C15ULL + (1ull << 32) + i));
    for (uint32_t i = 0; i < R * S; ++i) o->rdt[i] = u01_open_closed(hash32(iter_seed * 0x9E3779B97F4A7C15ULL + (2ull << 32) + i));
    return train_iter_body(o, boxes, n_boxes, frames, H, W, K, o->sxy.data(), o->rcol.data(), o->rdt.data(), nullptr);
}

size_t orc_object_last(const orc_object* o, int which, float* out, size_t cap) {
    auto copy_f = [&](const std::vector<float>& v) { size_t n = std::min(cap, v.size()); std::memcpy(out, v.data(), n * 4); return v.size(); };
    auto copy_h = [&](const std::vector<uint16_t>& v) { size_t n = std::min(cap, v.size()); for (size_t i = 0; i < n; ++i) out[i] = h2f(v[i]); return v.size(); };
    switch (which) {
        case 0: { size_t n = std::min(cap, o->rays.size() * 9); std::memcpy(out, o->rays.data(), n * 4); return o->rays.size() * 9; }
        case 1: return copy_f(o->points);
        case 2: return copy_f(o->t);
        case 3: return copy_h(o->enc);
        case 4: return copy_h(o->out);
        case 5: return copy_f(o->rgb_r);
        case 6: return copy_f(o->dep_r);
        case 7: return copy_f(o->mask_r);
        case 8: return copy_h(o->dout);
        case 9: return copy_h(o->denc);
        case 10: return copy_f(o->tgt);
        case 11: return copy_f(o->tgtd);
        case 12: { size_t n = std::min(cap, o->rinst.size()); for (size_t i = 0; i < n; ++i) out[i] = (float)o->rinst[i]; return o->rinst.size(); }
        case 13: return copy_f(o->loss);
    }
    return 0;
}

// GenerateRenderRays (nerf_model.cu:448-492): one ray per pixel of the 2-D box, row-major; rays outside the object box keep
// their previous contents in the reference (unwritten), here zeros
void orc_render_rays(uint32_t bx, uint32_t by, uint32_t bh, uint32_t bw, const float Twc[16], const float K[4],
                     const float obj_Tow[16], const float bmin[3], const float bmax[3], orc_ray* rays, int* in_box) {
    for (uint32_t i = 0; i < bh * bw; ++i) {
        const int x = (int)bx + (int)(i % bw), y = (int)by + (int)(i / bw);
        orc_ray r; float t0, t1;
        std::memset(&r, 0, sizeof r);
        pixel_ray((float)x, (float)y, K, Twc, obj_Tow, r.o, r.d, r.d_norm);
        in_box[i] = ray_box(bmin, bmax, r.o, r.d, t0, t1) ? 1 : 0;
        if (in_box[i]) { r.tmin = fmaxf(t0, 0.0f); r.tmax = t1; rays[i] = r; }
        else std::memset(&rays[i], 0, sizeof r);
    }
}

void orc_object_render(const orc_object* o, uint32_t bx, uint32_t by, uint32_t bh, uint32_t bw,
                       const float Twc[16], const float K[4], uint32_t S2,
                       const float* rand_dt, int use_ema, float* rgb, float* depth, float* mask) {
    const Layout& ly = o->ly;
    const uint32_t n_rays = bh * bw;
    const uint16_t* params = use_ema ? o->ema.data() : o->ph.data();
    const int T = o->n_threads;
    parallel_rays(T, n_rays, [&](uint32_t a, uint32_t b, int) {
        std::vector<float> pts(3 * S2), tt(S2), out4(4 * S2);
        std::vector<uint16_t> enc(ly.in_w * S2), out(16 * S2);
        for (uint32_t i = a; i < b; ++i) {
            const int x = (int)bx + (int)(i % bw), y = (int)by + (int)(i / bw);
            orc_ray r; float t0, t1;
            pixel_ray((float)x, (float)y, K, Twc, o->Tow, r.o, r.d, r.d_norm);
            int in_box = ray_box(o->bmin, o->bmax, r.o, r.d, t0, t1) ? 1 : 0;
            if (in_box) {
                r.tmin = fmaxf(t0, 0.0f); r.tmax = t1;
                orc_sample_points(1, S2, &r, o->bmin, o->bmax, rand_dt + (size_t)i * S2, pts.data(), tt.data());
                orc_encode(&o->cfg, params + ly.n_mlp, pts.data(), S2, enc.data());
                orc_mlp_forward(&o->cfg, params, enc.data(), S2, nullptr, out.data());
                for (uint32_t n = 0; n < S2; ++n) for (int k = 0; k < 4; ++k) out4[n * 4 + k] = h2f(out[n * 16 + k]);
            }
            orc_volume_render_test(1, S2, out4.data(), tt.data(), &in_box, &r.d_norm, 1.0f, rgb + (size_t)i * 3, depth + i, mask + i);
        }
    });
}

}  // extern "C"
