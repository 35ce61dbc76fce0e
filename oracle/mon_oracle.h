/*
 * mon_oracle.h — CPU ORACLE for the Multi-Object-NeRF train/render hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * ro_map_b200/ links, imports or calls it; the product path is CUDA-only and fails
 * loudly without its extension.
 *
 * It restates, in plain single-precision C++ with software fp16 rounding, the
 * arithmetic of the reference (paths relative to /root/reference; MON = dependencies/
 * Multi-Object-NeRF, TCNN = MON/Core/third_party/tiny-cuda-nn):
 *
 *   A1  ray_intersect                    MON/Core/src/nerf_model.cu:87-138
 *   A2  GenerateRays + fill_rollover_rays MON/Core/src/nerf_model.cu:369-446, 280-294
 *   A3  GenerateInputPoints / WarpPoint  MON/Core/src/nerf_model.cu:536-566, 140-144
 *   A4  kernel_grid (hash encode fwd)    TCNN/include/tiny-cuda-nn/encodings/grid.h:220-384,
 *                                        170-187, 131-135; common_device.h:485-495
 *   A5  FullyFusedMLP forward            TCNN/src/fully_fused_mlp.cu:315-476, 499-557
 *   A6  VolumeRender                     MON/Core/src/nerf_model.cu:735-815, 22-64
 *   A7  VolumeRenderGradient_No_Compacted MON/Core/src/nerf_model.cu:817-954
 *   A8  FullyFusedMLP backward           TCNN/src/fully_fused_mlp.cu:736-836
 *   A9  kernel_grid_backward             TCNN/include/tiny-cuda-nn/encodings/grid.h:386-509
 *   A10 adam_step (+ExponentialDecay)    TCNN/include/tiny-cuda-nn/optimizers/adam.h:48-118,
 *                                        exponential_decay.h:60-71
 *   A11 ema_step_half_precision          TCNN/include/tiny-cuda-nn/optimizers/ema.h:62-76,102-136
 *   A12 parameter init                   TCNN/include/tiny-cuda-nn/trainer.h:53-90,
 *                                        gpu_matrix.h:291-303, random.h:66-92,
 *                                        encodings/grid.h:1333-1336, dependencies/pcg32/pcg32.h
 *   A13 Train_Step loop                  MON/Core/src/nerf_model.cu:1429-1479,1552-1607,1630-1665
 *   A14 Render                           MON/Core/src/nerf_model.cu:448-493,593-626,1134-1229
 *
 * PINNING STATUS: the reference ships no golden vectors for this path.
 * The oracle is pinned (i) by first-principles known-answer tests (tests/test_oracle_kat.py)
 * and (ii) against outputs of the reference's own vendored tiny-cuda-nn compiled unmodified
 * (oracle/ref/Makefile -> oracle/_ref/libmon_ref.so) and run on a B200; those outputs are
 * committed as tests/golden/tcnn_*.npz together with the generating script
 * (oracle/ref/make_golden.py).  Rows A1-A3, A6, A7, A14 (RO-MAP's own kernels): the
 * reference's nerf_model.cu is compiled unmodified, from where it lies, against stand-in
 * headers for Eigen / OpenCV / GLEW (oracle/ref/shim, those libraries are absent here); its
 * kernels were run on a B200 by oracle/ref/make_golden_romap.py and their inputs and outputs
 * are committed as tests/golden/romap_golden.npz (checked by tests/test_golden_romap.py).
 *
 * Floating-point conventions.  The file is compiled with -ffp-contract=off; every place
 * where nvcc's default --fmad=true would fuse a*b+c in the reference's device code is
 * written as an explicit fmaf() so the choice is visible.  fp16 values travel as uint16_t
 * bit patterns.  The reference's tensor-core MMAs accumulate in fp16 with a hardware-defined
 * order; the oracle accumulates in fp32 in ascending k and rounds once, so MLP parity is a
 * tolerance (stated in the tests), never bit-exactness.  `__expf` is restated as expf.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 32

typedef struct {
    /* encoding (base.json "encoding") */
    uint32_t n_levels;         /* 16 */
    uint32_t n_features;       /* 2 (only 2 is supported, as instantiated by the reference) */
    uint32_t log2_hashmap_size;/* 16 */
    uint32_t base_resolution;  /* 16 */
    float    per_level_scale;  /* 2.0 (tcnn default; grid.h:1437) */
    /* network (base.json "network") */
    uint32_t n_neurons;        /* 64 */
    uint32_t n_hidden_layers;  /* 1 */
    uint32_t padded_output_width; /* 16 */
    /* optimizer (base.json "optimizer") */
    float learning_rate, beta1, beta2, epsilon, l2_reg;
    float ema_decay;
    uint32_t decay_start, decay_interval; float decay_base;
    /* training constants (nerf_model.h:166-175) */
    float loss_scale;          /* 128 */
} orc_config;

/* 20-byte POD, same field order as nerf::FrameIdAndBbox (MON/Core/include/common.h:18-23) */
typedef struct { uint32_t FrameId, x, y, h, w; } orc_bbox2d;
/* 36-byte POD, same field order as nerf::Ray (MON/Core/include/nerf_model.h:34-42) */
typedef struct { float o[3], d[3], d_norm, tmin, tmax; } orc_ray;

typedef struct {
    const uint8_t* rgb;      /* H*W*3 u8, RGB order; pixel value = u8 * (1/255) fp32 */
    const uint8_t* instance; /* H*W u8 */
    const float*   depth;    /* H*W f32 z-depth or NULL */
    float pose[16];          /* camera-to-world, column-major 4x4 */
} orc_frame;

void orc_default_config(orc_config* cfg);

/* ---- geometry of the parameter vector (A4 table, A12) ---- */
/* offsets[l] in entries for l=0..n_levels; scale[l]; resolution[l].  returns total #params */
uint32_t orc_grid_layout(const orc_config* cfg, uint32_t* offsets, float* scales, uint32_t* resolutions);
uint32_t orc_n_mlp_params(const orc_config* cfg);
uint32_t orc_n_params(const orc_config* cfg);

/* ---- A12 ---- */
void orc_seed_seq_1(uint32_t seed, uint32_t out[2]);           /* std::seed_seq{seed}.generate(2 words) */
void orc_pcg32_floats(uint64_t initstate, uint64_t advance, uint32_t n, float* out);
void orc_init_params(const orc_config* cfg, uint32_t seed, float* params_fp32);

/* ---- fp16 helpers ---- */
uint16_t orc_f2h(float f);
float    orc_h2f(uint16_t h);
void     orc_f2h_array(const float* in, uint16_t* out, size_t n);
void     orc_h2f_array(const uint16_t* in, float* out, size_t n);

/* ---- A1 ---- returns 1 on hit */
int orc_ray_intersect(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float* tmin, float* tmax);

/* ---- A2 ---- deterministic (slot order = ascending i, one legal outcome of the reference's
 * atomicAdd compaction).  Returns n_in (rays inside the box before roll-over padding).
 * sample_xy: 2R, rand_colors: 3R.  Outputs are padded to R by modular replication. */
uint32_t orc_generate_rays(uint32_t R, const orc_bbox2d* boxes, uint32_t n_boxes,
                           const orc_frame* frames, int H, int W, const float fxfycxcy[4],
                           const float obj_Tow[16], const float bmin[3], const float bmax[3],
                           uint8_t obj_instance, int use_depth,
                           const float* sample_xy, const float* rand_colors,
                           orc_ray* rays, uint8_t* rays_instance, float* target_rgb, float* target_depth);

/* ---- A3 ---- points: [R*S][3] warped to the unit cube, t: [R*S] */
void orc_sample_points(uint32_t R, uint32_t S, const orc_ray* rays, const float bmin[3], const float bmax[3],
                       const float* rand_dt, float* points, float* t);

/* ---- A4 ---- enc: [N][2*n_levels] fp16, point-major */
void orc_encode(const orc_config* cfg, const uint16_t* grid_fp16, const float* points, uint32_t N, uint16_t* enc);
/* indices + fp32 weights of the 8 corners, for KATs: idx/w [N][n_levels][8] */
void orc_encode_corners(const orc_config* cfg, const float* points, uint32_t N, uint32_t* idx, float* w);

/* ---- A5 ---- mlp_fp16: [W_in 64x32 | (W_h 64x64)* | W_out 16x64], row-major.
 * hidden: [n_hidden_layers][N][64] fp16 (post-ReLU) or NULL; out: [N][16] fp16 */
void orc_mlp_forward(const orc_config* cfg, const uint16_t* mlp_fp16, const uint16_t* enc, uint32_t N,
                     uint16_t* hidden, uint16_t* out);

/* ---- A6 ---- out: [R*S][16] fp16 (cols 0-2 rgb logits, 3 log-density); bg: [R][3] */
void orc_volume_render(uint32_t R, uint32_t S, const uint16_t* out, const float* t, const float* bg_rgb,
                       float* rgb_rays, float* depth_rays, float* mask_rays);

/* ---- A7 ---- dout: [R*S][16] fp16 (pre-zeroed by the callee), loss: [R] */
void orc_loss_backward(uint32_t R, uint32_t S, float loss_scale, const uint16_t* out, const float* t,
                       const uint8_t* rays_instance, const float* target_rgb, const float* target_depth,
                       const float* rgb_rays, const float* depth_rays, const float* mask_rays,
                       uint16_t* dout, float* loss);

/* ---- A8 ---- d_enc: [N][32] fp16; dW: same layout as mlp params, fp32 accumulated then
 * (round_fp16) rounded to fp16-representable values, as the reference stores fp16 gradients */
void orc_mlp_backward(const orc_config* cfg, const uint16_t* mlp_fp16, const uint16_t* enc,
                      const uint16_t* hidden, const uint16_t* dout, uint32_t N,
                      uint16_t* d_enc, float* dW, int round_fp16);

/* ---- A9 ---- grad: [n_grid_params]; mode 0: fp16 sequential accumulation in point order
 * (one legal order of the reference's atomicAdd(__half2)); mode 1: fp32 accumulation */
void orc_encode_backward(const orc_config* cfg, const float* points, const uint16_t* d_enc, uint32_t N,
                         float* grid_grad, int mode);

/* ---- A10 + A11 ---- one optimizer step over all P params; step is the global step AFTER
 * increment (1-based).  grads are the loss-scaled fp16-representable gradient values. */
void orc_optimizer_step(const orc_config* cfg, uint32_t step, const float* grads,
                        float* params_fp32, uint16_t* params_fp16, float* m, float* v,
                        uint32_t* param_steps, uint16_t* ema_fp16);

/* ---- A13 ---- a whole object, for end-to-end parity and the CPU baseline ---- */
typedef struct orc_object orc_object;
orc_object* orc_object_create(const orc_config* cfg, uint32_t seed, uint32_t R, uint32_t S,
                              const float obj_Tow[16], const float bmin[3], const float bmax[3],
                              uint8_t obj_instance, int use_depth, int n_threads);
void orc_object_destroy(orc_object*);
uint32_t orc_object_n_params(const orc_object*);
/* copy out state: which = 0 fp32 master, 1 fp16 params (as float), 2 ema (as float),
 * 3 last gradient (loss-scaled), 4 adam m, 5 adam v, 6 per-param step (as float) */
void orc_object_get(const orc_object*, int which, float* out);
void orc_object_set_params(orc_object*, const float* params_fp32);
/* one iteration with injected randoms (sample_xy 2R, rand_colors 3R, rand_dt R*S).
 * Returns the logged loss (sum over rays / R).  n_in_out receives the in-box ray count. */
float orc_object_train_iter(orc_object*, const orc_bbox2d* boxes, uint32_t n_boxes,
                            const orc_frame* frames, int H, int W, const float fxfycxcy[4],
                            const float* sample_xy, const float* rand_colors, const float* rand_dt,
                            uint32_t* n_in_out);
/* same, randoms drawn from an internal counter-based generator (CPU baseline timing) */
float orc_object_train_iter_rng(orc_object*, const orc_bbox2d* boxes, uint32_t n_boxes,
                                const orc_frame* frames, int H, int W, const float fxfycxcy[4],
                                uint64_t iter_seed);
/* intermediate buffers of the last iteration, for per-stage parity:
 * 0 rays (9 floats each), 1 points (3N), 2 t (N), 3 enc (32N as float), 4 out (16N as float),
 * 5 rgb_rays(3R), 6 depth_rays(R), 7 mask_rays(R), 8 dout (16N as float), 9 d_enc (32N as float),
 * 10 target rgb (3R), 11 target depth (R), 12 rays_instance (R as float), 13 loss per ray (R) */
size_t orc_object_last(const orc_object*, int which, float* out, size_t cap);

/* ---- A14 ---- render a 2-D box with EMA (use_ema=1) or training weights; rand_dt: [h*w*S2]
 * outputs rgb [h*w*3], depth [h*w], mask [h*w] */
void orc_object_render(const orc_object*, uint32_t bx, uint32_t by, uint32_t bh, uint32_t bw,
                       const float Twc[16], const float fxfycxcy[4], uint32_t S2,
                       const float* rand_dt, int use_ema, float* rgb, float* depth, float* mask);

/* A14 rays of a 2-D render box (GenerateRenderRays): rays [h*w] (zeros where in_box == 0), in_box [h*w] */
void orc_render_rays(uint32_t bx, uint32_t by, uint32_t bh, uint32_t bw, const float Twc[16], const float fxfycxcy[4],
                     const float obj_Tow[16], const float bmin[3], const float bmax[3], orc_ray* rays, int* in_box);

/* A14/A6 without an object: raw network output [N][4] fp32 -> pixels */
void orc_volume_render_test(uint32_t n_rays, uint32_t S2, const float* out4, const float* t,
                            const int* in_box, const float* d_norm, float bg,
                            float* rgb, float* depth, float* mask);

#ifdef __cplusplus
}
#endif
