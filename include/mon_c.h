/*
 * mon_c.h — C ABI of libmon_b200.so: the B200-native Multi-Object-NeRF train/render core.
 *
 * This is the drop-in boundary for the per-object NeRF hot path of RO-MAP
 * (/root/reference/dependencies/Multi-Object-NeRF/Core, "MON/Core" below).  The reference
 * boundary is a C++ class ABI (nerf::NerfManagerOffline / nerf::NerfManagerOnline /
 * nerf::NeRF, MON/Core/include/nerf_manager.h:21-91, nerf.h:19-88); the C++ facade in
 * ro_map_b200/host/ (same class names and signatures) is a thin shim over THIS interface, and
 * INTEGRATION.md shows the binding a RO-MAP maintainer adds.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every call returns
 * MON_OK (0) or a negative error and records a message readable with mon_last_error()
 * (thread-local).  Nothing here throws or calls exit() (the reference prints to cerr and
 * exit(0)s, nerf_manager.cu:21-25).  Handles are thread-compatible: one object <-> one CUDA
 * stream; distinct objects may be driven from distinct host threads concurrently, as the
 * reference does with one std::thread per object (nerf_manager.cu:89,259).
 * Matrices are float[16], column-major 4x4, exactly Eigen::Matrix4f::data().
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MON_C_H_
#define MON_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MON_OK 0
#define MON_ERR_CUDA -1     /* CUDA runtime error (message has the cudaError string) */
#define MON_ERR_ARG -2      /* invalid argument */
#define MON_ERR_IO -3       /* file / JSON error */
#define MON_ERR_STATE -4    /* call made in the wrong state (e.g. train without bboxes) */
#define MON_ERR_NO_DEVICE -5

/* ---- POD types crossing the boundary -------------------------------------------------- */

/* == nerf::FrameIdAndBbox (MON/Core/include/common.h:18-23): 5 x u32, field order x,y,h,w */
typedef struct { uint32_t FrameId, x, y, h, w; } mon_bbox2d;

/* Network / optimizer / batch configuration.  Mirrors MON/Core/configs/base.json plus the
 * compile-time constants of the reference (nerf_model.h:166-175, common.h:12). */
typedef struct {
    uint32_t n_levels;            /* encoding.n_levels               16 */
    uint32_t n_features_per_level;/* encoding.n_features_per_level    2 (only 2 supported) */
    uint32_t log2_hashmap_size;   /* encoding.log2_hashmap_size      16 */
    uint32_t base_resolution;     /* encoding.base_resolution        16 */
    float    per_level_scale;     /* encoding.per_level_scale       2.0 (tcnn default, grid.h:1437) */
    uint32_t n_neurons;           /* network.n_neurons               64 (only 64 supported) */
    uint32_t n_hidden_layers;     /* network.n_hidden_layers          1 (1 or 2) */
    float learning_rate, beta1, beta2, epsilon, l2_reg;   /* optimizer...Adam */
    float ema_decay;                                      /* optimizer.decay */
    uint32_t decay_start, decay_interval; float decay_base; /* ExponentialDecay */
    float loss_scale;             /* 128  (nerf_model.h:166) */
    uint32_t rays_per_batch;      /* 4096 (nerf_model.h:173); multiple of 4 */
    uint32_t samples_per_ray;     /* 32   (common.h:12);  fixed at 32 */
    uint32_t render_samples_per_ray; /* 64 (nerf_model.h:175); multiple of 32 */
    float depth_lambda, mask_lambda, bg_density_reg;      /* 0.5, 0.5, 0.01 (nerf_model.cu:869,927,940) */
} mon_config;

typedef struct mon_dataset mon_dataset;  /* == nerf::NeRF_Dataset, one per GPU (nerf_data.h:19-72) */
typedef struct mon_object mon_object;    /* == nerf::NeRF + nerf::NeRF_Model (nerf.h, nerf_model.h:81-185) */

/* ---- library ---------------------------------------------------------------------------- */
const char* mon_last_error(void);
const char* mon_version(void);
/* NerfManagerOffline::Init / NerfManagerOnline::Init device probe (nerf_manager.cu:16-38,137-160) */
int mon_device_count(int* count);

/* NeRF_Model::ReadNetworkConfig + the parsing half of ResetNetwork (nerf_model.cu:1272-1318).
 * Accepts JSON with comments, like the reference's json::parse(..., ignore_comments=true). */
int mon_config_default(mon_config* cfg);
int mon_config_from_json(const char* path, mon_config* cfg);
/* number of parameters: MLP weights, grid params (n_params = sum of both) */
int mon_config_param_counts(const mon_config* cfg, uint32_t* n_mlp, uint32_t* n_grid);
/* per-level table geometry (grid.h:964-997): offsets[n_levels+1] in entries, scale, resolution */
int mon_config_grid_layout(const mon_config* cfg, uint32_t* offsets, float* scales, uint32_t* resolutions);

/* ---- dataset (keyframes resident on one GPU) -------------------------------------------- */
/* NerfManagerOnline::DatasetInit / NeRF_Dataset::InitDataToGPU (nerf_manager.cu:162-187,
 * nerf_data.cu:232-271) and the allocation half of DataToGPU (nerf_data.cu:123-230). */
int mon_dataset_create(int gpu, float fx, float fy, float cx, float cy, int H, int W,
                       uint32_t max_frames, int use_depth, mon_dataset** out);
/* NeRF_Dataset::FrameDataToGPU (nerf_data.cu:273-339) / one iteration of DataToGPU's loop.
 * rgb: H*W*3 u8 (channel order BGR if is_bgr, as cv::imread delivers, else RGB); stored as u8
 * on the device and converted in-kernel (value = u8 * (1/255), the float the reference
 * stores, nerf_data.cu:163-164).  instance: H*W u8.  depth: H*W f32 metres (already multiplied
 * by DepthMapFactor) or NULL.  pose: camera-to-world.  Host buffers may be pageable (copied through a
 * pinned staging buffer, synchronous) or pinned (see mon_dataset_sync). */
int mon_dataset_add_frame(mon_dataset* ds, uint32_t frame_id, const uint8_t* rgb, int is_bgr,
                          const uint8_t* instance, const float* depth, const float pose[16]);
/* n consecutive keyframes first_id .. first_id + n - 1 at once — the whole of DataToGPU's loop (nerf_data.cu:123-230) in one call.
 * rgb / instance / depth: blocks of n planes each ([n][H][W][3] u8, [n][H][W] u8, [n][H][W] f32), poses16: [n][16].  The device
 * storage is plane-major per slab of 32 frames, so page-locked host blocks (or, with on_device != 0, blocks already in the
 * dataset GPU's memory) arrive in three asynchronous copies per slab instead of three per frame (154 MB: 55 instead of 49 GB/s);
 * pageable host blocks go frame by frame through the staging buffer like mon_dataset_add_frame. */
int mon_dataset_add_frames(mon_dataset* ds, uint32_t first_id, uint32_t n, const uint8_t* rgb, int is_bgr, const uint8_t* instance,
                           const float* depth, const float* poses16, int on_device);
/* The same for a keyframe that is ALREADY in the memory of the dataset's GPU (written by a decoder, or received from another
 * GPU by an NCCL broadcast — how bench.py replicates the keyframe set at N > 1): device-to-device copies on the dataset's
 * upload stream, asynchronous.  The caller makes sure the source buffers are complete before the call and keeps them valid
 * until mon_dataset_sync(). */
int mon_dataset_add_frame_device(mon_dataset* ds, uint32_t frame_id, const uint8_t* d_rgb, int is_bgr,
                                 const uint8_t* d_instance, const float* d_depth, const float pose[16]);
/* Depth as the depth image holds it.  DataToGPU reads a 16-bit PNG, converts it on the host — depthImg.convertTo(CV_32FC1,
 * mfDepthScale) (nerf_data.cu:176-186) — and uploads 4 bytes per pixel.  After this call (allowed until the first keyframe is
 * added) the dataset takes and stores the RAW u16 samples and the batch kernel converts the pixels it picks, (float)u16 *
 * depth_factor = the float the reference stores: a quarter less keyframe traffic over PCIe and in HBM (6 instead of 8 bytes
 * per pixel) and no host-side conversion pass.  Frames are then added through the _d16 entries below (same arguments as their
 * f32 counterparts, depth16: H*W uint16); the f32 entries are refused, and so are the _d16 ones on a dataset in f32 mode. */
int mon_dataset_set_depth_u16(mon_dataset* ds, float depth_factor);
int mon_dataset_add_frame_d16(mon_dataset* ds, uint32_t frame_id, const uint8_t* rgb, int is_bgr,
                              const uint8_t* instance, const uint16_t* depth16, const float pose[16]);
int mon_dataset_add_frames_d16(mon_dataset* ds, uint32_t first_id, uint32_t n, const uint8_t* rgb, int is_bgr, const uint8_t* instance,
                               const uint16_t* depth16, const float* poses16, int on_device);
int mon_dataset_add_frame_device_d16(mon_dataset* ds, uint32_t frame_id, const uint8_t* d_rgb, int is_bgr,
                                     const uint8_t* d_instance, const uint16_t* d_depth16, const float pose[16]);
/* Page-locked (cudaHostAlloc / cudaHostRegister) RGB-order buffers are uploaded asynchronously without a staging copy;
 * they must stay valid until mon_dataset_sync() returns or a blocking call on an object of this dataset completes. */
int mon_dataset_sync(mon_dataset* ds);
/* NeRF_Dataset::UpdateDataGPU (nerf_data.cu:341-353) */
int mon_dataset_update_poses(mon_dataset* ds, uint32_t first_frame, uint32_t n, const float* poses16);
int mon_dataset_frame_count(const mon_dataset* ds, uint32_t* n);
/* next-row: replicate every frame of src (another GPU) into dst over NVLink peer
 * copies instead of re-uploading from the host (nerf_manager.cu:203-216 uploads once per GPU) */
int mon_dataset_clone_from_peer(mon_dataset* dst, const mon_dataset* src);
/* online variant of the same: ONE keyframe, already added to src (upload possibly still in flight), is copied device to
 * device into dst on dst's upload stream; returns without waiting.  NerfManagerOnline::NewFrameToDataset uploads a frame
 * once over PCIe and replicates it with this call (the reference: one host thread and one PCIe upload per GPU,
 * nerf_manager.cu:189-218). */
int mon_dataset_copy_frame_from_peer(mon_dataset* dst, const mon_dataset* src, uint32_t frame_id);
int mon_dataset_destroy(mon_dataset* ds);

/* ---- object -------------------------------------------------------------------------------- */
/* NeRF::CreateModelOffline / CreateModelOnline + NeRF_Model ctor + ResetNetwork +
 * AllocateBatchWorkspace (nerf.cu:37-56,179-185; nerf_model.cu:1259-1427).  obj_Tow: world-to-
 * object; bmin/bmax: object-space AABB; instance_id = uint8(class).  seed 1337 reproduces the
 * reference's parameter initialisation (trainer.h:53-90). */
int mon_object_create(mon_dataset* ds, const mon_config* cfg, uint32_t seed, uint8_t instance_id,
                      const float obj_Tow[16], const float bmin[3], const float bmax[3],
                      mon_object** out);
int mon_object_destroy(mon_object* obj);
/* NeRF_Model::UpdateFrameIdAndBbox (replace, nerf_model.cu:1609) / ...Online (append, :1616) */
int mon_object_set_bboxes(mon_object* obj, const mon_bbox2d* boxes, uint32_t n);
int mon_object_add_bboxes(mon_object* obj, const mon_bbox2d* boxes, uint32_t n);

/* NeRF_Model::Train_Step / Train_Step_Online (nerf_model.cu:1630-1699): `iters` iterations of
 * GenerateBatch -> Step_No_Compacted -> optimizer_step, replayed as CUDA graphs (64 iterations each + one graph of
 * exactly the remainder, captured on first use of that length) on the object's stream with no host synchronisation in between.
 * mon_object_train blocks until done and returns the logged loss of the last iteration
 * (sum over rays / R, nerf_model.cu:1650-1658) if loss != NULL. */
int mon_object_train(mon_object* obj, uint32_t iters, float* loss);
int mon_object_train_async(mon_object* obj, uint32_t iters);
/* the allocation half of AllocateBatchWorkspace (nerf_model.cu:1344-1427) for calls of `iters` iterations: captures and
 * instantiates the iteration graphs now, so that the first mon_object_train[_async](iters) does not pay for it */
int mon_object_prepare_train(mon_object* obj, uint32_t iters);
int mon_object_sync(mon_object* obj);
/* device time (ms, CUDA events on the object's stream) of the last train / train_async call;
 * valid after the call completed */
int mon_object_last_train_ms(mon_object* obj, float* ms);
int mon_object_step_count(mon_object* obj, uint32_t* step);
/* Measurement hook: `iters` iterations launched kernel by kernel (no graph) with a CUDA event between the
 * stages on the object's stream; stage_ms[k] = mean device time of stage k per iteration.
 * Stages: 0 batch (ray generation + compaction), 1 sample positions, 2 hash-grid encode, 3 fused MLP forward +
 * volume render + loss + MLP backward (hands the live samples over compacted), 4 hash-grid gradient scatter, 5 optimizer
 * sweep (MLP-gradient reduction, Adam + EMA + gradient zeroing over all parameters, logged-loss reduction).  In the opt-in
 * stage 4 is two kernels of which one exits at once: shared-memory resident scatter (many live samples) or global f16x2 reductions.  n_stages must be MON_N_STAGES. */
#define MON_N_STAGES 6
int mon_object_train_profiled(mon_object* obj, uint32_t iters, float* stage_ms, uint32_t n_stages);
/* samples of the last iteration that carried gradient (non-zero dL/dencoding row: not behind the early stop T < 1e-4 of
 * VolumeRenderGradient_No_Compacted, nerf_model.cu:889, and not underflown in fp16) out of rays_per_batch * 32: the work of
 * the gradient scatter is proportional to it */
int mon_object_live_samples(mon_object* obj, uint32_t* n_live, uint32_t* n_points);
/* number of CUDA kernels this library launched on behalf of the object so far */
int mon_object_launch_count(mon_object* obj, uint64_t* n);

/* NeRF_Model::Render (nerf_model.cu:1702-1830): render the 2-D box from camera pose Twc with the
 * EMA weights (use_ema=1, what the reference does) into host buffers rgb[h*w*3], depth[h*w]
 * (z-depth), mask[h*w].  rand_dt: h*w*render_samples floats in (0,1] or NULL (internal RNG). */
int mon_object_render(mon_object* obj, mon_bbox2d box, const float Twc[16], int use_ema,
                      const float* rand_dt, float* rgb, float* depth, float* mask);
/* One view of NeRF_Model::RenderVideo (nerf_model.cu:1832-1991; rays as GenerateRenderVideoRays :495-533): same
 * renderer, but the pose is camera -> OBJECT frame (Toc, e.g. from GenerateToc :2186-2205): ObjTow is not applied. */
int mon_object_render_object_centric(mon_object* obj, mon_bbox2d box, const float Toc[16], int use_ema,
                                     const float* rand_dt, float* rgb, float* depth, float* mask);
/* GetDensityOnGrid (nerf_model.cu:2007-2043): raw sigma logit on a res^3 lattice of the AABB, inference (EMA)
 * weights like the reference; x fastest.  out: res[0]*res[1]*res[2] floats on the host.  Input of marching cubes. */
int mon_object_density_grid(mon_object* obj, const uint32_t res[3], float* out);
/* Network logits (r, g, b, sigma; fp16 network output widened to float) at n unit-cube positions [n][3], as
 * compute_mesh_vertex_colors queries them (nerf_model.cu:2045-2067).  use_ema = 1: inference weights. */
int mon_object_query_points(mon_object* obj, const float* points_unit, uint32_t n, int use_ema, float* out4);

/* GenerateMesh as a whole (nerf_model.cu:1993-2043): GetDensityOnGrid + MarchingCubes + compute_mesh_1ring (marching_cubes.cu:41-510)
 * + compute_mesh_vertex_colors (:2045-2067) on the GPU.  res^3 lattice over the object's box (the reference: 64), iso-value thresh on
 * the raw density logit (the reference: 2.0), EMA weights.  Vertex and triangle slots come from exclusive scans over the lattice
 * instead of the reference's atomicAdd race: the order is the lattice order (x fastest; per point the +x, +y, +z edge; per cell the
 * table's triangle order), reproducible.  The vertex count is padded to a multiple of 128 with zero vertices like the reference's
 * (:499).  The mesh stays in device memory; mon_mesh_counts sizes the caller's arrays, mon_mesh_read copies out verts [n_verts][3],
 * unit 1-ring normals [n_verts][3], u8 colours [n_verts][3], indices [n_indices] (the reference's internal winding; its PLY writer
 * reverses it); any of the four may be NULL.  A mesh of an object is read and destroyed before its object is destroyed. */
typedef struct mon_mesh mon_mesh;
int mon_object_extract_mesh(mon_object* obj, uint32_t res, float thresh, mon_mesh** out);
/* the same surface extraction on a caller's lattice (host memory, [z][y][x], box bmin..bmax); colours are zero */
int mon_mesh_from_lattice(int gpu, const float* sigma, uint32_t res, const float bmin[3], const float bmax[3], float thresh, mon_mesh** out);
int mon_mesh_counts(const mon_mesh* mesh, uint32_t* n_verts, uint32_t* n_surface_verts, uint32_t* n_indices);
int mon_mesh_read(const mon_mesh* mesh, float* verts, float* normals, uint8_t* colors, uint32_t* indices);
int mon_mesh_destroy(mon_mesh* mesh);

/* ---- parity / test hooks ------------------------------------------------------------------- */
/* One training iteration with host-provided random numbers instead of the internal generator:
 * sample_xy[2R], rand_colors[3R], rand_dt[R*S], all in (0,1] like curandGenerateUniform
 * (nerf_model.cu:1432-1468).  Keeps every intermediate of the iteration for mon_object_last. */
int mon_object_train_injected(mon_object* obj, const float* sample_xy, const float* rand_colors,
                              const float* rand_dt, float* loss, uint32_t* n_in_box);
/* state vectors, as float[n_params]: 0 fp32 master, 1 fp16 params, 2 EMA (fp16), 3 last
 * gradient (loss-scaled, fp16-representable), 4 Adam m, 5 Adam v, 6 per-param step */
int mon_object_get_state(mon_object* obj, int which, float* out, size_t n);
int mon_object_set_params(mon_object* obj, const float* params_fp32, size_t n);
/* intermediates of the last injected iteration (float), returns count via n_out:
 * 0 rays(9/ray) 1 points(3/pt, unit cube) 3 enc(32/pt) 4 out(4/pt: r,g,b,sigma logits) 5 rgb_rays 6 depth_rays
 * 7 mask_rays 8 dout(4/pt) 9 d_enc(32/pt) 10 target rgb 11 target depth 12 ray instance flag
 * 13 per-ray loss */
int mon_object_last(mon_object* obj, int which, float* out, size_t cap, size_t* n_out);
/* OPT-IN occupancy grid with warp-ballot sample compaction — OFF by default and in every parity run: it changes which samples
 * contribute, hence the results.  The reference carries instant-ngp's accelerators as dead code (NeRF_Model::Step,
 * VolumeRenderGradient with compaction, nerf_model.cu:957-1132,1504-1550); BASELINE.json's north_star names them.
 * grid_res^3 cells over the object's box (0 switches the mode off; multiple of 4 in 8..256).  After warmup_iters iterations the
 * grid is refreshed every update_interval iterations (at graph-replay boundaries) from the network's density on the cell
 * corners (running maximum with decay 0.95); a cell is empty while the opacity of one average sample interval inside it is below
 * alpha_threshold.  Samples in empty cells are not encoded and count as empty space (density 0, no gradient). */
int mon_object_set_occupancy(mon_object* obj, uint32_t grid_res, uint32_t warmup_iters, uint32_t update_interval, float alpha_threshold);
/* fraction of occupied cells, and of the last iteration's samples that fell into occupied cells */
int mon_object_occupancy_stats(mon_object* obj, float* occupied_cell_fraction, float* occupied_sample_fraction);
/* host-only: the work split of the hash-encode kernel (kernels_encode.cu) for n_ctas CTAs over levels
 * [level_begin, level_end): out4[4*b..] = first (job, point) and end (job, point) of piece b, job = 2*level + feature.
 * No device needed; lets a CPU test check that the pieces tile the [job][point] space exactly. */
int mon_debug_encode_pieces(const mon_config* cfg, uint32_t n_points, uint32_t n_ctas, uint32_t level_begin,
                            uint32_t level_end, uint32_t* out4);
/* host-only: the work split of the shared-memory resident gradient scatter (kernels_scatter_smem.cu) for n_live live samples
 * over n_ctas CTAs: out4[4*b..] = first (job, sample) and end (job, sample) of piece b, job = 4*level + 2*(entry index & 1) + feature. */
int mon_debug_scatter_pieces(const mon_config* cfg, uint32_t n_live, uint32_t n_ctas, uint32_t* out4);
/* stand-alone stage entry points on host buffers (copies inside), for kernel-level parity */
int mon_stage_encode(const mon_config* cfg, const uint16_t* grid_fp16, size_t n_grid_params,
                     const float* points_unit, uint32_t n_points, uint16_t* enc_out);

#ifdef __cplusplus
}
#endif
#endif /* MON_C_H_ */
