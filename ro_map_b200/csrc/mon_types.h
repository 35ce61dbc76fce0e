// mon_types.h — device/host shared data model of the B200 Multi-Object-NeRF core.
// Reference counterparts: Ray (MON/Core/include/nerf_model.h:34-42), MetaData
// (nerf_data.h:11-17), BatchData (nerf_model.h:45-65), GridOffsetTable (TCNN grid.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mon_c.h"

#define MON_MAX_LEVELS 16
#define MON_S 32          // training samples per ray == warp size (common.h:12 SampleNum)
#define MON_WIDTH 64      // hidden width (base.json n_neurons)
#define MON_IN 32         // encoding width = n_levels * 2
#define MON_OUT 16        // padded output width (tcnn pads 4 -> 16)
#define MON_MAX_MLP_CTAS 592

// 36-byte POD; same field order as nerf::Ray
struct MonRay { float o[3], d[3], d_norm, tmin, tmax; };

// per-frame device pointers; u8 pixels instead of the reference's float pixels
struct MonFrame {
    const uint8_t* rgb;       // H*W*3, RGB
    const uint8_t* instance;  // H*W
    const float* depth;       // H*W f32 — or, if depth_factor != 0, H*W raw 16-bit samples — or nullptr
    float pose[16];           // camera-to-world, column-major
    uint32_t bgr;             // 1: the pixels are stored B, G, R as the SLAM frontend hands them over (cv::imread order);
                              // the batch kernel swaps on read instead of the host swizzling every keyframe
    float depth_factor;       // != 0: the depth plane holds the raw u16 samples of the depth image and the metric value is
                              // (float)u16 * depth_factor, converted on read — the float the reference stores after
                              // depthImg.convertTo(CV_32FC1, mfDepthScale) (nerf_data.cu:182); 0: the plane holds f32 metres
};

// geometry of the multiresolution table, precomputed on the host (grid.h:195-204,964-997)
struct MonGrid {
    uint32_t n_levels;
    uint32_t offset[MON_MAX_LEVELS + 1];  // in entries (1 entry = 2 fp16 features)
    uint32_t size[MON_MAX_LEVELS];        // entries in level
    uint32_t res[MON_MAX_LEVELS];
    uint32_t hashed[MON_MAX_LEVELS];      // 1 if coherent-prime hash is used, 0 if dense
    float scale[MON_MAX_LEVELS];
    // entry -> level without a search: every level >= first_full has exactly 2^log2_cap entries
    // (first_full == n_levels or log2_cap == 0: fall back to the offset search)
    uint32_t first_full, log2_cap;
};

// constant scene description of one object
struct MonScene {
    float Tow[16];      // world -> object
    float bmin[3], bmax[3];
    float K[4];         // fx fy cx cy
    int H, W;
    uint32_t instance_id;
    int use_depth;
};

// device-resident iteration control block (read by every kernel of the iteration graph)
struct MonCtrl {
    uint32_t step;      // optimizer step of the CURRENT iteration (1-based), bumped by the batch kernel
    uint32_t n_in;      // rays inside the box before roll-over padding
    uint32_t skip;      // 1: no ray hit the box this iteration -> all later kernels are no-ops
    uint32_t iter;      // RNG iteration counter (advances even when skipped)
    float loss_mean;    // filled by the optimizer sweep (logged loss of the iteration)
    uint32_t n_boxes;   // live number of 2-D boxes (host-updated; keeps the captured graph valid)
    // per-step optimizer scalars, computed once by the batch kernel instead of once per optimizer CTA
    float lr_base;      // learning_rate * decay_base^(#decays)            (exponential_decay.h:60-71)
    float ema_old;      // 1 - ema_decay^(step-1)                          (ema.h:107)
    float ema_new;      // 1 / (1 - ema_decay^step)                        (ema.h:108)
    uint32_t pad[3];
};

// optimizer hyper-parameters
struct MonOpt {
    float lr, beta1, beta2, eps, l2_reg, ema_decay, loss_scale;
    uint32_t decay_start, decay_interval; float decay_base;
    uint32_t n_mlp, n_params;
    uint32_t n_partials;   // number of per-CTA MLP gradient partials to sum
    float log2_beta1, log2_beta2;   // beta^s is evaluated as exp2f(s * log2(beta))
    float inv_loss_scale;           // 1 / loss_scale, used when loss_scale is a power of two (exact)
    uint32_t loss_scale_pow2;
    const float* debias_lut;        // [n_debias_lut] Adam bias correction by per-parameter step count (kernels_optim.cu)
    uint32_t n_debias_lut;
};

struct MonLossCfg { float loss_scale, depth_lambda, mask_lambda, bg_density_reg; };

// Opt-in occupancy grid (OFF in every parity run: bits == nullptr).  res^3 cells over the object's unit cube, one bit per cell;
// samples in cells whose bit is clear are not encoded and count as empty space (density 0, no gradient).
struct MonOcc {
    const uint32_t* bits;    // [res^3 / 32], cell (x, y, z) = bit x + res * (y + res * z)
    uint32_t res;
    uint32_t* ray_mask;      // [R] bit s = sample s of the ray lies in an occupied cell (written by the sample-points kernel)
    uint32_t* list;          // [N] indices of the occupied samples of the iteration, compacted (any order)
    uint32_t* count;         // [1] number of entries of list; zeroed by the batch kernel
};

// everything a training iteration touches, as raw device pointers
struct MonBatch {
    uint32_t R;                 // rays per batch
    const mon_bbox2d* boxes;
    const MonFrame* frames;
    MonCtrl* state;     // persistent counters (iter, step, n_boxes): read and advanced by the batch kernels only, which run
                        // strictly one after the other
    MonCtrl* ctrl;      // control block of THIS iteration, written by its batch kernel (one per batch buffer)
    MonCtrl* late;      // copy taken by the fused MLP kernel for the scatter / optimizer kernels, so that the batch
                        // kernel of the NEXT iteration may run concurrently with them
    uint32_t seed;
    // learning-rate schedule / EMA constants for the per-step scalars
    float opt_lr, decay_base, ema_decay; uint32_t decay_start, decay_interval;
    // injected randoms (nullptr -> internal counter-based generator)
    const float* inj_xy; const float* inj_col; const float* inj_dt;
    // per-ray
    MonRay* rays; uint8_t* ray_inst; float* target; float* target_depth; float* bg;
    float* rgb_rays; float* depth_rays; float* mask_rays; float* loss;
    // per-point
    __half* enc;      // level-major pairs [16][N][2]
    __half* d_enc;    // [N][32] point-major dL/dencoding: parity hook only (nullptr in production)
    const float* pts; // [N][3] unit-cube sample positions of this iteration (read by the fused MLP kernel for the compaction)
    // compacted live samples (non-zero dL/dencoding row), written by the fused MLP kernel for the scatter+Adam kernel:
    // slot k holds position pts_c[k][4] (x, y, z, unused: one 16-byte load) and, per level l, the level's two fp16 gradients as one word genc[l * N + k]
    float* pts_c;
    uint32_t* genc;
    uint32_t* live_cnt;   // [2]: number of live samples of the iteration, indexed by (iteration & 1); zeroed by the batch kernel
    MonOcc occ;           // opt-in occupancy grid of the object (bits == nullptr: off)
    // debug dumps (nullptr in production): out [N][4], dout [N][4]
    float* dbg_out; float* dbg_dout;
    // parameters
    const __half* params;   // fp16 working copy [P]: W_in | (W_h) | W_out | grid
    __half* grads;          // fp16 gradient [P] (grid part used; MLP part written by optimizer for inspection)
    float* mlp_partials;    // [n_partials][n_mlp] fp32
};
