// mon_kernels.h — launch prototypes of the sm_100a kernels (one .cu per kernel family).
#pragma once
#include <atomic>

#include "mon_types.h"

// Function attributes (dynamic shared memory size, carve-out) are per DEVICE: run `f` once for every device a kernel is
// launched on (one process may drive several GPUs through the C++ facade).  `f` is idempotent, so a benign race
// between two host threads only repeats it.
template <typename F>
inline cudaError_t mon_once_per_device(std::atomic<uint64_t>& done_mask, F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = f();
    if (e == cudaSuccess) done_mask.fetch_or(bit, std::memory_order_release);
    return e;
}

// Programmatic dependent launch.  The kernels of one iteration form a chain on the object's stream; each is launched
// with cudaLaunchAttributeProgrammaticStreamSerialization so that its CTAs may become resident (and run their
// prologue: barrier init, TMEM allocation, weight staging) while the previous kernel drains, and each executes
//     mon_pdl_wait();      // griddepcontrol.wait: the previous kernel of the chain has COMPLETED, memory visible
//     mon_pdl_trigger();   // griddepcontrol.launch_dependents: the next kernel may start its prologue
// in that order, on every thread, before any early return and before touching anything the predecessor wrote.
// Because the trigger comes after the wait, a kernel's pre-wait prologue only ever overlaps its immediate
// predecessor; everything older is complete and may be read there.  Captured into the iteration graphs as
// programmatic edges; without the attribute (MON_NO_PDL=1) both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void mon_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mon_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
// mon_core.cu: bit set = that kernel is launched with the attribute.  All on by default; MON_NO_PDL=1 clears all,
// MON_PDL_MASK=<int> selects (A/B measurements).
enum { MON_PDL_POINTS = 1, MON_PDL_ENCODE = 2, MON_PDL_MLP = 4, MON_PDL_SCATTER = 8, MON_PDL_OPTIM = 16 };
unsigned mon_pdl_mask();

// Per-launch options of the iteration graphs.  pdl = false: no programmatic edge even if the mask allows one (a kernel
// whose predecessor in the stream is not the producer it waits for).  priority: CUDA stream-priority value recorded on
// the kernel node (numerically lower = dispatched first).
struct MonLaunchOpt {
    bool pdl = true;
    bool set_priority = false;
    int priority = 0;
};

template <typename... KArgs, typename... Args>
inline cudaError_t mon_launch_chain(unsigned which, const MonLaunchOpt& lo, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (lo.pdl && (mon_pdl_mask() & which)) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (lo.set_priority) {
        attr[n].id = cudaLaunchAttributePriority;
        attr[n].val.priority = lo.priority;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kernels_batch.cu
// slim: the 16 x 256-thread shape that fits beside the hash-encode kernel's CTAs (kernels_batch.cu)
void mon_launch_generate_batch(const MonBatch& b, const MonScene& sc, cudaStream_t st, const MonLaunchOpt& lo = MonLaunchOpt(), bool slim = false);
void mon_launch_render_rays(uint32_t n_rays, mon_bbox2d box, const MonScene& sc, const float* Twc_dev, float bgc,
                            MonRay* rays_hit, uint32_t* orig, uint32_t* n_hit, float* rgb, float* depth, float* mask, cudaStream_t st);

// kernels_encode.cu
void mon_launch_sample_points(uint32_t n_points, uint32_t S, const MonRay* rays, const int* in_box, const float* jitter,
                              uint32_t seed, const MonCtrl* ctrl, uint32_t rng_stream, uint32_t iter_fixed,
                              const float* bmin, const float* bmax, float* pts, cudaStream_t st, const MonLaunchOpt& lo = MonLaunchOpt(),
                              const uint32_t* orig_ray = nullptr,    // render: compacted ray -> pixel, indexes the jitter
                              const MonOcc* occ = nullptr);          // training, opt-in occupancy mode: ray masks + compacted list of occupied samples
// levels [level_begin, level_end) only (default: all)
cudaError_t mon_launch_encode_forward(const MonGrid& g, uint32_t n_points, const float* pts, const __half* planar, __half* enc_soa,
                                      const MonCtrl* ctrl, uint32_t sm_count, cudaStream_t st, uint32_t level_begin = 0,
                                      uint32_t level_end = 0xffffffffu, const MonLaunchOpt& lo = MonLaunchOpt(),
                                      const uint32_t* occ_list = nullptr, const uint32_t* occ_count = nullptr);   // opt-in: encode the listed samples only
void mon_launch_planarize(const MonGrid& g, const __half* inter, __half* planar, cudaStream_t st);
void mon_launch_deplanarize(const MonGrid& g, const __half* planar, __half* inter, cudaStream_t st);   // the inverse (state getter)
void mon_encode_pieces_host(const MonGrid& g, uint32_t n_points, uint32_t n_ctas, uint32_t level_begin, uint32_t level_end, uint32_t* out4);
// stand-alone gradient scatter with global f16x2 reductions over the compacted live samples: configurations the unified kernel
// (kernels_scatter_smem.cu) does not cover; returns at once in iterations with >= resident_min_live live samples
void mon_launch_encode_backward(const MonGrid& g, uint32_t n_points, uint32_t resident_min_live, const uint32_t* live_cnt, const float* pts_c,
                                const uint32_t* genc, const MonCtrl* ctrl, __half* grid_grad, cudaStream_t st, const MonLaunchOpt& lo = MonLaunchOpt());

// kernels_mlp_tc.cu (tcgen05 / TMEM product family)
cudaError_t mon_launch_mlp_train_tc(const MonBatch& b, const MonLossCfg& lc, uint32_t n_hidden, uint32_t n_mlp,
                                    uint32_t n_ctas, cudaStream_t st, const MonLaunchOpt& lo = MonLaunchOpt(),
                                    const MonGrid* fuse_grid = nullptr, __half* gh_grid = nullptr);   // fuse_grid: scatter the grid gradients in-kernel
cudaError_t mon_launch_mlp_infer_tc(uint32_t n_points, uint32_t n_hidden, const __half* params, const __half* enc,
                                    float* out4, cudaStream_t st);
cudaError_t mon_launch_mlp_render_tc(uint32_t n_rays, uint32_t S2, uint32_t n_hidden, const MonRay* rays, const int* in_box, const float* jitter,
                                     uint32_t seed, uint32_t iter, const __half* params, const __half* enc, float bgc,
                                     float* rgb, float* depth, float* mask, cudaStream_t st, const uint32_t* orig_ray = nullptr);

// kernels_optim.cu
void mon_launch_init_grid(uint64_t state, uint64_t inc, uint32_t n, float* out, cudaStream_t st);
void mon_launch_cast_params(uint32_t n, const float* pf, __half* ph, cudaStream_t st);
void mon_launch_optimizer(const MonOpt& o, MonCtrl* ctrl, float* pf, __half* ph, __half* gh, const float* partials,
                          float* m, float* v, uint16_t* ps, __half* ema, const float* loss, uint32_t R, const MonGrid& grid,
                          __half* planar, cudaStream_t st, int part = 0, uint32_t level_begin = 0, uint32_t level_end = 0xffffffffu,
                          const MonLaunchOpt& lo = MonLaunchOpt(), __half* gcls = nullptr, const uint32_t* live_cnt = nullptr,
                          uint32_t resident_min_live = 0xffffffffu, uint32_t sm_count = 148);
enum { MON_OPT_ALL = 0, MON_OPT_MLP = 1, MON_OPT_GRID = 2, MON_OPT_MLP_GRID = 3 };
// kernels_scatter_smem.cu: THE gradient scatter of the iteration graph, one launch, two paths chosen on the device by the
// iteration's live-sample count: >= min_live -> shared-memory resident fixed-point slices flushed with TMA bulk reductions into the
// class-planar table gcls; fewer -> global f16x2 reductions into the entry-ordered table gh_grid (min_live = 0xffffffff: always)
bool mon_scatter_resident_supported(const MonGrid& g);   // power-of-two tables of >= 16 entries, even dense-level resolutions
void mon_scatter_resident_pieces_host(const MonGrid& g, uint32_t n_live, uint32_t n_ctas, uint32_t* out4);
cudaError_t mon_launch_scatter(const MonGrid& g, uint32_t n_points, uint32_t min_live, const uint32_t* live_cnt, const float* pts_c,
                               const uint32_t* genc, const MonCtrl* ctrl, __half* gcls, __half* gh_grid, uint32_t sm_count, cudaStream_t st,
                               const MonLaunchOpt& lo = MonLaunchOpt(), bool leave_spare_sms = false);
void mon_launch_snapshot_grad(uint32_t n, uint32_t n_mlp, uint32_t n_partials, const __half* gh, const float* partials,
                              float* out, cudaStream_t st, const MonGrid& grid, const __half* gcls);

// kernels_mesh.cu: marching cubes on the device (count + exclusive scans, then vertices / faces / 1-ring normals), vertex colours
size_t mon_mesh_scan_scratch_words(uint32_t res);
cudaError_t mon_launch_mc_count(const float* sigma, uint32_t res, float thresh, uint32_t* v_off, uint32_t* i_off, uint32_t* totals, uint32_t* sums,
                                cudaStream_t st);
cudaError_t mon_launch_mc_build(const float* sigma, uint32_t res, float thresh, const float bmin[3], const float bmax[3], const uint32_t* v_off,
                                const uint32_t* i_off, uint32_t n_surface, uint32_t n_verts_padded, uint32_t n_indices, uint32_t* vid, float* verts,
                                float* normals, uint32_t* indices, cudaStream_t st);
void mon_launch_mesh_unit_points(uint32_t n_verts, const float* verts, const float bmin[3], const float bmax[3], float* unit, cudaStream_t st);
void mon_launch_mesh_colors(uint32_t n_verts, const float* out4, uint8_t* colors, cudaStream_t st);
