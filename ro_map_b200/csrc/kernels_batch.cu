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

#define BATCH_THREADS 1024

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

__global__ void __launch_bounds__(BATCH_THREADS, 1)
k_generate_batch(MonBatch b, MonScene sc) {
    __shared__ uint32_t warp_sums[BATCH_THREADS / 32];
    __shared__ uint32_t s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t iter = b.ctrl->iter;
    const uint32_t n_boxes = b.ctrl->n_boxes;
    const uint32_t R = b.R;
    const uint32_t per_thread = (R + BATCH_THREADS - 1) / BATCH_THREADS;
    const uint32_t i0 = tid * per_thread;

    // pass 1: count this thread's in-box rays
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < per_thread; ++k) {
        const uint32_t i = i0 + k;
        if (i < R && make_candidate(b, sc, n_boxes, i, iter).valid) ++cnt;
    }
    // block-wide exclusive scan of cnt
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = warp_sums[lane];
        uint32_t vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, vi, o); if (lane >= o) vi += u; }
        warp_sums[lane] = vi - v;
        if (lane == 31) s_total = vi;
    }
    __syncthreads();
    uint32_t slot = warp_sums[warp] + incl - cnt;
    const uint32_t n_in = s_total;

    // pass 2: recompute and write compacted rays + targets
    for (uint32_t k = 0; k < per_thread; ++k) {
        const uint32_t i = i0 + k;
        if (i >= R) break;
        RayCand c = make_candidate(b, sc, n_boxes, i, iter);
        if (!c.valid) continue;
        const uint32_t idx = slot++;
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
            b.target[idx * 3 + 0] = __fmul_rn((float)px[0], (float)(1.0 / 255.0));
            b.target[idx * 3 + 1] = __fmul_rn((float)px[1], (float)(1.0 / 255.0));
            b.target[idx * 3 + 2] = __fmul_rn((float)px[2], (float)(1.0 / 255.0));
            b.target_depth[idx] = (sc.use_depth && fr->depth) ? __fmul_rn(fr->depth[pix], c.d_norm) : 0.0f;
            b.ray_inst[idx] = 1;
        } else {
            b.target[idx * 3 + 0] = mon_rand(b.inj_col, b.seed, iter, 1, idx * 3 + 0);
            b.target[idx * 3 + 1] = mon_rand(b.inj_col, b.seed, iter, 1, idx * 3 + 1);
            b.target[idx * 3 + 2] = mon_rand(b.inj_col, b.seed, iter, 1, idx * 3 + 2);
            b.target_depth[idx] = 0.0f;
            b.ray_inst[idx] = 0;
        }
    }
    __syncthreads();  // global writes of this CTA are visible to the CTA after the barrier

    // fill_rollover_rays (:280-294) + the background colour VolumeRender indexes as
    // RandColors[(i % n_in) * 3] (:760)
    if (n_in > 0) {
        for (uint32_t i = tid; i < R; i += BATCH_THREADS) {
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
    if (tid == 0) {
        b.ctrl->n_in = n_in;
        b.ctrl->skip = (n_in == 0) ? 1u : 0u;  // reference: modulo by zero (undefined); here: skip the iteration
        if (n_in > 0) b.ctrl->step += 1;
        b.ctrl->iter = iter + 1;
    }
}

// ---- test-render rays (GenerateRenderRays, nerf_model.cu:448-493): one ray per box pixel
__global__ void k_render_rays(uint32_t n_rays, mon_bbox2d box, MonScene sc, const float* __restrict__ Twc_dev,
                              MonRay* __restrict__ rays, int* __restrict__ in_box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    float Twc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) Twc[k] = Twc_dev[k];
    const int x = (int)box.x + (int)(i % box.w);
    const int y = (int)box.y + (int)(i / box.w);
    MonRay r; float t0, t1;
    mon_pixel_ray((float)x, (float)y, sc.K, Twc, sc.Tow, r.o, r.d, r.d_norm);
    const bool hit = mon_ray_box(sc.bmin, sc.bmax, r.o, r.d, t0, t1);
    r.tmin = hit ? fmaxf(t0, 0.0f) : 0.0f;
    r.tmax = hit ? t1 : 0.0f;
    rays[i] = r;
    in_box[i] = hit ? 1 : 0;
}

void mon_launch_generate_batch(const MonBatch& b, const MonScene& sc, cudaStream_t st) {
    k_generate_batch<<<1, BATCH_THREADS, 0, st>>>(b, sc);
}
void mon_launch_render_rays(uint32_t n_rays, mon_bbox2d box, const MonScene& sc, const float* Twc_dev,
                            MonRay* rays, int* in_box, cudaStream_t st) {
    k_render_rays<<<(n_rays + 127) / 128, 128, 0, st>>>(n_rays, box, sc, Twc_dev, rays, in_box);
}
