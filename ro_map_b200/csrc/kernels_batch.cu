// kernels_batch.cu — ray generation for training batches and for test renders.
//
// Replaces, in ONE launch and with no host synchronisation, the reference sequence
//   curandGenerateUniform x2 -> GenerateRays -> [cudaStreamSynchronize] -> fill_rollover_rays
// (MON/Core/src/nerf_model.cu:1429-1465; kernels at :369-446 and :280-294).
// The reference compacts in-box rays with a global atomicAdd (order is run-to-run
// nondeterministic, :419); here a single CTA compacts with a block-wide exclusive scan, so the
// slot order is "ascending batch index" — one of the reference's legal outcomes, and
// reproducible.  The iteration control block (step counter, skip flag) is advanced here, which
// is what lets the whole iteration replay as a static CUDA graph.
#include "mon_device.cuh"
#include "mon_kernels.h"
#include "mon_timeline.cuh"
MON_TL_DEFINE(batch)

#include <cooperative_groups.h>

// Two shapes of the one cluster (mon_core.cu capture_graph says which graph uses which):
//   WIDE  8 CTAs x 512 threads, 80 registers: needs SMs of its own — the graphs WITH a scatter kernel run it beside that kernel,
//         which leaves 8 SMs free;
//   SLIM  16 CTAs x 256 threads, <= 64 registers: such a CTA fits beside a CTA of the hash-encode kernel (1024 threads x 46
//         registers, 128 KB of shared memory), so the graphs WITHOUT a scatter kernel generate the batch of iteration i + 1 while
//         iteration i is encoded.  16 is a non-portable cluster size; where the device cannot co-schedule it the launcher falls
//         back to 8 CTAs (two candidates per thread).
#define BATCH_THREADS_WIDE 512
#define BATCH_THREADS_SLIM 256
#define BATCH_CTAS 16
#define BATCH_CTAS_PORTABLE 8

struct RayCand {
    bool valid; uint32_t x, y; uint8_t inst; uint32_t frame; size_t pix;
    float o[3], d[3], d_norm, t0, t1;
};

__device__ __forceinline__ RayCand make_candidate(const MonBatch& b, const MonScene& sc, uint32_t n_boxes, uint32_t i, uint32_t iter) {
    RayCand c; c.valid = false;
    const mon_bbox2d box = b.boxes[i % n_boxes];
    const MonFrame* fr = b.frames + box.FrameId;
    const float sx = mon_rand(b.inj_xy, b.seed, iter, 0, 2 * i);
    const float sy = mon_rand(b.inj_xy, b.seed, iter, 0, 2 * i + 1);
    const int h = (int)box.h, w = (int)box.w;
    c.x = box.x + (uint32_t)__fmul_rn(sx, (float)w);
    c.y = box.y + (uint32_t)__fmul_rn(sy, (float)h);
    c.frame = box.FrameId;
    // linear pixel index exactly as the reference forms it (y*W+x, nerf_model.cu:398); x may equal
    // box.x+box.w because curand's interval is (0,1] - clamped only so that a box touching the last
    // image row can never read past the frame
    const size_t last_px = (size_t)sc.H * sc.W - 1;
    size_t pix = (size_t)c.y * sc.W + c.x;
    c.pix = pix < last_px ? pix : last_px;
    c.inst = fr->instance[c.pix];
    if (c.inst != 0 && c.inst != (uint8_t)sc.instance_id) return c;  // occluded by another object
    mon_pixel_ray((float)c.x, (float)c.y, sc.K, fr->pose, sc.Tow, c.o, c.d, c.d_norm);
    c.valid = mon_ray_box(sc.bmin, sc.bmax, c.o, c.d, c.t0, c.t1);
    return c;
}

// One thread-block cluster of BATCH_CTAS CTAs (all co-resident, distributed shared memory) owns the whole batch:
// thread g of the cluster evaluates candidates [g*per, (g+1)*per).  The compaction order is the ascending batch
// index: CTA-local exclusive scan, then every CTA reads the totals of the lower-ranked CTAs out of their shared
// memory (DSMEM) — no global atomics, no second kernel, no host round trip.  A second cluster barrier makes the
// compacted rays visible cluster-wide before the roll-over padding reads them.
__device__ __forceinline__ void write_ray(const MonBatch& b, const MonScene& sc, const RayCand& c, uint32_t idx, uint32_t iter) {
    MonRay r;
#pragma unroll
    for (int q = 0; q < 3; ++q) { r.o[q] = c.o[q]; r.d[q] = c.d[q]; }
    r.d_norm = c.d_norm; r.tmin = fmaxf(c.t0, 0.0f); r.tmax = c.t1;
    b.rays[idx] = r;
    const MonFrame* fr = b.frames + c.frame;
    const size_t pix = c.pix;
    if (c.inst != 0) {
        const uint8_t* px = fr->rgb + pix * 3;
        // the reference stores float pixels = u8 * (1/255) (nerf_data.cu:163-164); same value here
        const uint32_t ir = fr->bgr ? 2u : 0u;    // cv::cvtColor(BGR2RGB) of the reference (nerf_data.cu:286), done on read
        b.target[idx * 3 + 0] = __fmul_rn((float)px[ir], (float)(1.0 / 255.0));
        b.target[idx * 3 + 1] = __fmul_rn((float)px[1], (float)(1.0 / 255.0));
        b.target[idx * 3 + 2] = __fmul_rn((float)px[2u - ir], (float)(1.0 / 255.0));
        float z = 0.0f;
        if (sc.use_depth && fr->depth)
            z = fr->depth_factor != 0.0f ? __fmul_rn((float)reinterpret_cast<const uint16_t*>(fr->depth)[pix], fr->depth_factor) : fr->depth[pix];
        b.target_depth[idx] = __fmul_rn(z, c.d_norm);
        b.ray_inst[idx] = 1;
    } else {
        b.target[idx * 3 + 0] = mon_rand(b.inj_col, b.seed, iter, 1, idx * 3 + 0);
        b.target[idx * 3 + 1] = mon_rand(b.inj_col, b.seed, iter, 1, idx * 3 + 1);
        b.target[idx * 3 + 2] = mon_rand(b.inj_col, b.seed, iter, 1, idx * 3 + 2);
        b.target_depth[idx] = 0.0f;
        b.ray_inst[idx] = 0;
    }
}

template <uint32_t BATCH_THREADS>
__global__ void __launch_bounds__(BATCH_THREADS, BATCH_THREADS == BATCH_THREADS_SLIM ? 4 : 1)
k_generate_batch(MonBatch b, MonScene sc) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t warp_sums[BATCH_THREADS / 32];
    __shared__ uint32_t s_total;          // read by the other CTAs of the cluster
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster.block_rank(), n_cta = cluster.num_blocks();
    const uint32_t iter = b.state->iter;
    const uint32_t n_boxes = b.state->n_boxes;
    MON_TL(MON_TL_B, iter);
    const uint32_t R = b.R;
    const uint32_t n_threads = n_cta * BATCH_THREADS;
    const uint32_t per_thread = (R + n_threads - 1) / n_threads;
    const uint32_t i0 = (rank * BATCH_THREADS + tid) * per_thread;

    // pass 1: evaluate; the first candidate stays in registers, further ones (R > 4096) are recomputed in pass 2
    RayCand first; first.valid = false;
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < per_thread; ++k) {
        const uint32_t i = i0 + k;
        if (i >= R) break;
        const RayCand c = make_candidate(b, sc, n_boxes, i, iter);
        if (k == 0) first = c;
        if (c.valid) ++cnt;
    }
    // CTA-wide exclusive scan of cnt
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < BATCH_THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, vi, o); if (lane >= o) vi += u; }
        if (lane < BATCH_THREADS / 32) warp_sums[lane] = vi - v;
        if (lane == 31) s_total = vi;
    }
    cluster.sync();                        // every CTA's s_total is published
    uint32_t base = 0, n_in = 0;
    for (uint32_t r = 0; r < n_cta; ++r) {
        const uint32_t t = *cluster.map_shared_rank(&s_total, r);
        if (r < rank) base += t;
        n_in += t;
    }
    uint32_t slot = base + warp_sums[warp] + incl - cnt;

    // pass 2: write compacted rays + targets
    for (uint32_t k = 0; k < per_thread; ++k) {
        const uint32_t i = i0 + k;
        if (i >= R) break;
        const RayCand c = (k == 0) ? first : make_candidate(b, sc, n_boxes, i, iter);
        if (c.valid) write_ray(b, sc, c, slot++, iter);
    }
    cluster.sync();                        // compacted rays visible to the whole cluster; also: nobody exits while its s_total is still read

    // fill_rollover_rays (:280-294) + the background colour VolumeRender indexes as RandColors[(i % n_in) * 3] (:760)
    if (n_in > 0) {
        for (uint32_t i = rank * BATCH_THREADS + tid; i < R; i += n_threads) {
            const uint32_t s = i % n_in;
            if (i >= n_in) {
                b.rays[i] = b.rays[s];
                b.ray_inst[i] = b.ray_inst[s];
                b.target[i * 3 + 0] = b.target[s * 3 + 0];
                b.target[i * 3 + 1] = b.target[s * 3 + 1];
                b.target[i * 3 + 2] = b.target[s * 3 + 2];
                b.target_depth[i] = b.target_depth[s];
            }
            b.bg[i * 3 + 0] = mon_rand(b.inj_col, b.seed, iter, 1, s * 3 + 0);
            b.bg[i * 3 + 1] = mon_rand(b.inj_col, b.seed, iter, 1, s * 3 + 1);
            b.bg[i * 3 + 2] = mon_rand(b.inj_col, b.seed, iter, 1, s * 3 + 2);
        }
    }
    if (rank == 0 && tid == 0) {
        // the iteration's own control block (b.ctrl, one per batch buffer) is written completely; the counters that
        // survive the iteration live in b.state
        b.ctrl->n_in = n_in;
        b.ctrl->n_boxes = n_boxes;
        b.ctrl->skip = (n_in == 0) ? 1u : 0u;  // reference: modulo by zero (undefined); here: skip the iteration
        b.ctrl->step = b.state->step;
        if (n_in > 0) {
            const uint32_t step = b.state->step + 1;
            b.state->step = step;
            b.ctrl->step = step;
            // ExponentialDecay evaluates its condition with the nested step BEFORE Adam increments it
            float factor = 1.0f;
            if (step - 1 >= b.decay_start) {
                const uint32_t n_decays = (step - 1 - b.decay_start) / b.decay_interval + 1;
                for (uint32_t s = 0; s < n_decays; ++s) factor = __fmul_rn(factor, b.decay_base);
            }
            b.ctrl->lr_base = __fmul_rn(b.opt_lr, factor);
            // host code in the reference: float debias from a double pow (ema.h:107-108)
            b.ctrl->ema_old = 1.0f - (float)pow((double)b.ema_decay, (double)(step - 1));
            b.ctrl->ema_new = 1.0f / (1.0f - (float)pow((double)b.ema_decay, (double)step));
        }
        b.ctrl->iter = iter + 1;
        b.state->iter = iter + 1;
        // live-sample counter of THIS iteration (filled by the fused MLP kernel, read by the scatter + Adam kernel): two
        // counters alternate, because this kernel runs one iteration ahead, beside the previous iteration's scatter
        if (b.live_cnt) b.live_cnt[iter & 1u] = 0u;
        if (b.occ.count) *b.occ.count = 0u;       // opt-in occupancy mode: the sample-points kernel appends this iteration's occupied samples
    }
}

// ---- test-render rays (GenerateRenderRays, nerf_model.cu:448-493): one ray per box pixel.  Pixels whose ray misses the
// object's box get their final value here (background, depth 0, mask 0: VolumeRender_Render's else branch, :1217-1226)
// and are dropped: the hits are compacted (ray + original pixel index), so that sampling, hash encoding and the MLP —
// which the reference runs for every pixel of the view — only see rays that can contribute.
__global__ void k_render_rays(uint32_t n_rays, mon_bbox2d box, MonScene sc, const float* __restrict__ Twc_dev, float bgc,
                              MonRay* __restrict__ rays_hit, uint32_t* __restrict__ orig, uint32_t* __restrict__ n_hit,
                              float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    MonRay r;
    if (i < n_rays) {
        float Twc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) Twc[k] = Twc_dev[k];
        const int x = (int)box.x + (int)(i % box.w);
        const int y = (int)box.y + (int)(i / box.w);
        float t0, t1;
        mon_pixel_ray((float)x, (float)y, sc.K, Twc, sc.Tow, r.o, r.d, r.d_norm);
        hit = mon_ray_box(sc.bmin, sc.bmax, r.o, r.d, t0, t1);
        r.tmin = hit ? fmaxf(t0, 0.0f) : 0.0f;
        r.tmax = hit ? t1 : 0.0f;
        if (!hit) { rgb[i * 3] = rgb[i * 3 + 1] = rgb[i * 3 + 2] = bgc; depth[i] = 0.0f; mask[i] = 0.0f; }
    }
    // warp-aggregated append: one atomic per warp, lanes take consecutive slots
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (m == 0) return;
    const uint32_t lane = threadIdx.x & 31u, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(n_hit, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (hit) {
        const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
        rays_hit[slot] = r;
        orig[slot] = i;
    }
}

// cluster size of the slim batch kernel on the current device: 16 where such a cluster can be resident, else the portable 8
static uint32_t batch_cluster_size_slim() {
    static std::atomic<uint64_t> prepared{0};
    static std::atomic<uint32_t> size_by_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    mon_once_per_device(prepared, [&] {
        uint32_t n = BATCH_CTAS_PORTABLE;
        // the hash-encode kernel's carve-out: an SM keeps its shared-memory split while CTAs are resident, and with another
        // preference whichever of the two kernels came second to an SM waited for the first one's CTA to leave (measured: the
        // encode CTAs of the batch cluster's 16 SMs started 11 us late)
        cudaFuncSetAttribute(k_generate_batch<BATCH_THREADS_SLIM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaFuncSetAttribute(k_generate_batch<BATCH_THREADS_SLIM>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(BATCH_CTAS); cfg.blockDim = dim3(BATCH_THREADS_SLIM);
            cudaLaunchAttribute a;
            a.id = cudaLaunchAttributeClusterDimension;
            a.val.clusterDim.x = BATCH_CTAS; a.val.clusterDim.y = 1; a.val.clusterDim.z = 1;
            cfg.attrs = &a; cfg.numAttrs = 1;
            int n_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&n_clusters, k_generate_batch<BATCH_THREADS_SLIM>, &cfg) == cudaSuccess && n_clusters >= 1) n = BATCH_CTAS;
        }
        cudaGetLastError();
        if (const char* e = getenv("MON_BATCH_CTAS")) { if (atoi(e) < BATCH_CTAS) n = BATCH_CTAS_PORTABLE; }      // A/B measurements
        size_by_dev[dev & 63].store(n);
        return cudaSuccess;
    });
    return size_by_dev[dev & 63].load();
}

void mon_launch_generate_batch(const MonBatch& b, const MonScene& sc, cudaStream_t st, const MonLaunchOpt& lo, bool slim) {
    const uint32_t n_ctas = slim ? batch_cluster_size_slim() : BATCH_CTAS_PORTABLE;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas);
    cfg.blockDim = dim3(slim ? BATCH_THREADS_SLIM : BATCH_THREADS_WIDE);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = n_ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (lo.set_priority) {
        attr[1].id = cudaLaunchAttributePriority;
        attr[1].val.priority = lo.priority;
        cfg.numAttrs = 2;
    }
    if (slim) cudaLaunchKernelEx(&cfg, k_generate_batch<BATCH_THREADS_SLIM>, b, sc);
    else cudaLaunchKernelEx(&cfg, k_generate_batch<BATCH_THREADS_WIDE>, b, sc);
}
void mon_launch_render_rays(uint32_t n_rays, mon_bbox2d box, const MonScene& sc, const float* Twc_dev, float bgc,
                            MonRay* rays_hit, uint32_t* orig, uint32_t* n_hit, float* rgb, float* depth, float* mask, cudaStream_t st) {
    cudaMemsetAsync(n_hit, 0, sizeof(uint32_t), st);
    k_render_rays<<<(n_rays + 127) / 128, 128, 0, st>>>(n_rays, box, sc, Twc_dev, bgc, rays_hit, orig, n_hit, rgb, depth, mask);
}
