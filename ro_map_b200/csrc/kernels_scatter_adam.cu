// kernels_scatter_adam.cu — hash-grid gradient scatter FUSED with the per-parameter Adam / EMA update.
//
// Replaces, in one launch, cudaMemsetAsync(grid gradients) + kernel_grid_backward (TCNN encodings/grid.h:1132,386-509) and
// the grid part of adam_step / ema_step_half_precision (TCNN optimizers/adam.h:48-118, ema.h:62-76).
//
// The reference scatters 8 atomicAdd(__half2) per (sample, level) into a global fp16 gradient table and sweeps the whole
// parameter vector afterwards.  On B200 that scatter is bound by the number of reduction requests an SM can send to L2 (one
// per lane and corner, 124 G/s measured — tools/smem_atomic_probe.cu, profiles/r4b_smem_atomic_probe.txt): 16.8 M of them
// for a fresh object, whose every sample carries gradient — 87-108 us, 5x the rest of the iteration.  Here the GRADIENT
// TABLE is the resident operand, like the weight table in the encode kernel:
//
//   job      = one quarter of one level's table (<= 16384 entries = 128 KB of accumulators in shared memory): 64 jobs for
//              base.json.  Every level is cut in four, the small dense ones too, so that all jobs cost the same.
//   cluster  = 2 CTAs work on one job.  Each takes half of the iteration's LIVE samples (the fused MLP kernel hands over
//              only the samples with a non-zero gradient row, compacted: position + one word per level), computes the 8
//              corner entries and adds the contributions that fall into the job's slice into its PRIVATE copy of the slice.
//   accumulators = two 32-bit FIXED-POINT words per entry with unit 2^-24.  Every contribution is rounded to fp16 first,
//              exactly as the reference rounds it before its atomicAdd(__half2) (grid.h:427-431), and every fp16 value is a
//              multiple of 2^-24, so the integer sum is EXACT and independent of the order: the gradient is bit-reproducible
//              run to run (the reference's fp16 running sum is not), and native integer shared-memory atomics (ATOMS.ADD,
//              2975 G/s) replace the compare-and-swap loop every floating-point shared-memory atomic compiles to (830 G/s).
//              Range +-128 in loss-scaled units, contributions clamped to +-16 (gradients of this loss are ~1e-3).
//   reduce   = after a cluster barrier every CTA owns half of the slice: it adds the partner's copy (distributed shared
//              memory) and rounds the sum to fp16 once: the loss-scaled fp16 gradient the reference keeps
//   update   = and applies Adam + EMA to those entries in place — the complete gradient of an entry is known inside the
//              cluster, which is the grid-wide phase boundary a fused scatter + Adam needs.  Entries with a zero
//              gradient are skipped by Adam exactly as in the reference (adam.h:75-79) and only EMA-filtered.
//
// The fp16 gradient table never exists in global memory (3.8 MB written by atomics, read and zeroed again per iteration),
// and the launch between scatter and optimizer is gone.  fuse == 0 (A/B, MON_SO_FUSE=0) stops after the reduction and
// stores the gradient for the separate optimizer sweep of kernels_optim.cu.
#include <cooperative_groups.h>

#include "mon_device.cuh"
#include "mon_kernels.h"
#include "optim_math.cuh"
#include "mon_timeline.cuh"
MON_TL_DEFINE(scatter_adam)

namespace cg = cooperative_groups;

#define SO_THREADS 1024
#define SO_CLUSTER 2
#define SO_SLICES_PER_LEVEL 4u
#define SO_SLICE_ENTRIES 16384u
#define SO_SMEM_BYTES (SO_SLICE_ENTRIES * 8u)
#define SO_FIXED_ONE 16777216.0f          // 2^24: one unit of the accumulators is 2^-24, the smallest fp16 subnormal
#define SO_CONTRIB_MAX 16.0f

struct SoArgs {
    MonGrid g;
    MonOpt o;
    uint32_t n_points;          // N: stride (in words) between the levels of genc
    uint32_t n_jobs;
    uint32_t fuse;              // 1: Adam + EMA in this kernel; 0: store the reduced gradient into gh
    const uint32_t* live_cnt;   // [2], indexed by iteration parity
    const float* pts_c;         // [n_live][3]
    const uint32_t* genc;       // [n_levels][N]: slot k's two fp16 gradients of the level
    const MonCtrl* ctrl;        // the iteration's control block (copy taken by the fused MLP kernel)
    float* pf; __half* ph; float* m; float* v; uint32_t* ps; __half* ema; __half* planar;
    __half* gh;                 // grid gradients [n_grid] (fuse == 0 only)
    float* grad_snap;           // parity hook: loss-scaled gradient of every parameter as float (nullptr in production)
};

// job -> (level, first entry, entries); slices of a level in ascending order
__host__ __device__ __forceinline__ uint32_t so_slices(uint32_t size) {
    const uint32_t need = (size + SO_SLICE_ENTRIES - 1) / SO_SLICE_ENTRIES;
    return need > SO_SLICES_PER_LEVEL ? need : SO_SLICES_PER_LEVEL;
}
__device__ __forceinline__ void so_job(const MonGrid& g, uint32_t job, uint32_t& level, uint32_t& e0, uint32_t& ne) {
    uint32_t j = job;
    for (uint32_t l = 0; l < g.n_levels; ++l) {
        const uint32_t slices = so_slices(g.size[l]);
        if (j < slices) {
            level = l;
            // equal slices, each a multiple of 8 entries (level sizes are multiples of 8)
            const uint32_t per = ((g.size[l] / 8 + slices - 1) / slices) * 8;
            e0 = j * per;
            ne = e0 < g.size[l] ? min(per, g.size[l] - e0) : 0u;
            return;
        }
        j -= slices;
    }
    level = 0; e0 = 0; ne = 0;
}

uint32_t mon_scatter_adam_jobs(const MonGrid& g) {
    uint32_t n = 0;
    for (uint32_t l = 0; l < g.n_levels; ++l) n += so_slices(g.size[l]);
    return n;
}

// one corner: the contribution is rounded to fp16 like the reference's (grid.h:427-431), then added exactly in fixed point
__device__ __forceinline__ void so_add(uint32_t slice_addr, uint32_t rel, uint32_t ne, float g0, float g1, float w) {
    if (rel >= ne) return;
    const float2 f = __half22float2(__floats2half2_rn(__fmul_rn(g0, w), __fmul_rn(g1, w)));
    const int i0 = __float2int_rn(fminf(fmaxf(f.x, -SO_CONTRIB_MAX), SO_CONTRIB_MAX) * SO_FIXED_ONE);
    const int i1 = __float2int_rn(fminf(fmaxf(f.y, -SO_CONTRIB_MAX), SO_CONTRIB_MAX) * SO_FIXED_ONE);
    const uint32_t addr = slice_addr + rel * 8u;
    if (i0) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(i0) : "memory");
    if (i1) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr + 4u), "r"(i1) : "memory");
}

// slot handled by thread `tid` in trip `k` of a CTA's share [s_begin, s_end).  Hashed levels: consecutive lanes take
// consecutive slots (coalesced loads; their entries are spread by the hash).  Dense levels: consecutive slots are
// neighbouring samples of one ray, which share corner entries at coarse resolution — a warp would hit the same accumulator
// word from many lanes and the hardware serialises those.  There, inside every block of 1024 slots, lane i of warp w takes
// slot 32 * i + w: the lanes of a warp are 32 samples apart, i.e. on 32 different rays.
__device__ __forceinline__ uint32_t so_slot(uint32_t s_begin, uint32_t k, uint32_t tid, bool spread) {
    const uint32_t t = spread ? (((tid & 31u) << 5) | (tid >> 5)) : tid;
    return s_begin + k * SO_THREADS + t;
}

__global__ void __launch_bounds__(SO_THREADS, 1)
k_scatter_adam(const __grid_constant__ SoArgs a) {
    extern __shared__ __align__(16) unsigned char so_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t tid = threadIdx.x;
    const uint32_t rank = cluster.block_rank();
    const uint32_t cluster_id = blockIdx.x / SO_CLUSTER, n_clusters = gridDim.x / SO_CLUSTER;
    const uint32_t slice_addr = (uint32_t)__cvta_generic_to_shared(so_smem);

    // before the dependency wait (overlaps the tail of the fused MLP kernel): clear the first job's accumulators
    {
        uint4* z = reinterpret_cast<uint4*>(so_smem);
        for (uint32_t i = tid; i < SO_SMEM_BYTES / 16; i += SO_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    mon_pdl_wait();       // the fused MLP kernel (live samples, control block) has completed
    mon_pdl_trigger();
    if (a.ctrl->skip) return;                          // uniform over the grid
    MON_TL(MON_TL_S, a.ctrl->iter - 1);
    const uint32_t n_live = a.live_cnt[(a.ctrl->iter - 1) & 1u];
    const float lr_base = a.ctrl->lr_base, old_db = a.ctrl->ema_old, new_db = a.ctrl->ema_new;
    const uint32_t s_begin = (uint32_t)((uint64_t)n_live * rank / SO_CLUSTER), s_end = (uint32_t)((uint64_t)n_live * (rank + 1) / SO_CLUSTER);
    const uint32_t n_trips = (s_end - s_begin + SO_THREADS - 1) / SO_THREADS;
    __syncthreads();

    bool first = true;
    for (uint32_t job = cluster_id; job < a.n_jobs; job += n_clusters) {
        uint32_t l, e0, ne;
        so_job(a.g, job, l, e0, ne);
        if (!first) {
            // the previous job's slice was read by the partner until the cluster barrier at its end
            uint4* z = reinterpret_cast<uint4*>(so_smem);
            for (uint32_t i = tid; i < SO_SMEM_BYTES / 16; i += SO_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }
        first = false;

        // ---- phase 1: this CTA's half of the live samples -> private accumulators
        const uint32_t size = a.g.size[l], res = a.g.res[l];
        const float scale = a.g.scale[l];
        const bool hashed = a.g.hashed[l] != 0;
        const uint32_t* gl = a.genc + (size_t)l * a.n_points;
        const bool pow2 = (size & (size - 1)) == 0;
        const uint32_t emask = size - 1u;
        const uint32_t my = hashed ? 2654435761u : res, mz = hashed ? 805459861u : res * res;
        for (uint32_t k = 0; k < n_trips && ne; ++k) {
            const uint32_t s = so_slot(s_begin, k, tid, !hashed);
            if (s >= s_end) continue;
            const uint32_t gw = __ldg(gl + s);
            if ((gw & 0x7fff7fffu) == 0u) continue;       // adding +0 is an identity
            const float g0 = __half2float(__ushort_as_half((unsigned short)(gw & 0xffffu)));
            const float g1 = __half2float(__ushort_as_half((unsigned short)(gw >> 16)));
            const float u0 = __ldg(a.pts_c + (size_t)s * 3), u1 = __ldg(a.pts_c + (size_t)s * 3 + 1), u2 = __ldg(a.pts_c + (size_t)s * 3 + 2);
            float fr[3]; uint32_t cell[3];
            mon_pos_fract(u0, scale, fr[0], cell[0]);
            mon_pos_fract(u1, scale, fr[1], cell[1]);
            mon_pos_fract(u2, scale, fr[2], cell[2]);
            const float h0 = __fsub_rn(1.0f, fr[0]), h1 = __fsub_rn(1.0f, fr[1]), h2 = __fsub_rn(1.0f, fr[2]);
            // same multiplication order as the forward kernel and the reference: (wx * wy) * wz
            const float wxy[4] = {__fmul_rn(h0, h1), __fmul_rn(fr[0], h1), __fmul_rn(h0, fr[1]), __fmul_rn(fr[0], fr[1])};
            if (pow2) {
                const uint32_t ax[2] = {cell[0], cell[0] + 1u};
                const uint32_t ay[2] = {cell[1] * my, (cell[1] + 1u) * my};
                const uint32_t az[2] = {cell[2] * mz, (cell[2] + 1u) * mz};
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const uint32_t idx = (hashed ? (ax[c & 1] ^ ay[(c >> 1) & 1] ^ az[c >> 2]) : (ax[c & 1] + ay[(c >> 1) & 1] + az[c >> 2])) & emask;
                    so_add(slice_addr, idx - e0, ne, g0, g1, __fmul_rn(wxy[c & 3], (c & 4) ? fr[2] : h2));
                }
            } else {
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const uint32_t idx = mon_grid_index(hashed, size, res, cell[0] + (c & 1), cell[1] + ((c >> 1) & 1), cell[2] + ((c >> 2) & 1));
                    so_add(slice_addr, idx - e0, ne, g0, g1, __fmul_rn(wxy[c & 3], (c & 4) ? fr[2] : h2));
                }
            }
        }
        cluster.sync();        // both private copies of the slice are complete and visible cluster-wide
        if (job == cluster_id) MON_TL_MARK(MON_TL_S + 1, a.ctrl->iter - 1);

        // ---- phase 2: this CTA's half of the slice: add the partner's copy, round to fp16, then Adam + EMA in place
        const int4* mine = reinterpret_cast<const int4*>(so_smem);
        const int4* theirs = reinterpret_cast<const int4*>(cluster.map_shared_rank(so_smem, rank ^ 1u));
        const uint32_t n_chunks = ne / 4;                                  // 4 entries (8 accumulator words) per chunk (ne % 8 == 0)
        const uint32_t c_begin = (uint32_t)((uint64_t)n_chunks * rank / SO_CLUSTER), c_end = (uint32_t)((uint64_t)n_chunks * (rank + 1) / SO_CLUSTER);
        const OptimPtrs ptrs = {a.pf, a.ph, a.m, a.v, a.ps, a.ema};
        for (uint32_t c = c_begin + tid; c < c_end; c += SO_THREADS) {
            const int4 m0 = mine[2 * c], m1 = mine[2 * c + 1], t0 = theirs[2 * c], t1 = theirs[2 * c + 1];
            const int sum[8] = {m0.x + t0.x, m0.y + t0.y, m0.z + t0.z, m0.w + t0.w, m1.x + t1.x, m1.y + t1.y, m1.z + t1.z, m1.w + t1.w};
            // the gradient the reference keeps is fp16 (loss-scaled): one rounding here
            uint32_t gp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(__fmul_rn((float)sum[2 * q], 1.0f / SO_FIXED_ONE), __fmul_rn((float)sum[2 * q + 1], 1.0f / SO_FIXED_ONE));
                gp[q] = *reinterpret_cast<const uint32_t*>(&h);
            }
            const uint32_t entry = e0 + 4 * c;                              // within the level
            const uint32_t i8 = a.o.n_mlp + 2u * (a.g.offset[l] + entry);   // first of the 8 parameters
            if (a.grad_snap) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&gp[q]));
                    a.grad_snap[i8 + 2 * q] = f.x; a.grad_snap[i8 + 2 * q + 1] = f.y;
                }
            }
            if (!a.fuse) {
                *reinterpret_cast<uint4*>(a.gh + (i8 - a.o.n_mlp)) = make_uint4(gp[0], gp[1], gp[2], gp[3]);
                continue;
            }
            const uint4 wraw = *reinterpret_cast<const uint4*>(a.ph + i8);
            const uint4 eraw = *reinterpret_cast<const uint4*>(a.ema + i8);
            __half* f0 = a.planar + (size_t)a.g.offset[l] * 2 + entry;
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {                                 // two quads of 4 parameters = 2 entries each
                const __half2 ga = *reinterpret_cast<const __half2*>(&gp[2 * hq]), gb = *reinterpret_cast<const __half2*>(&gp[2 * hq + 1]);
                float g[4] = {__low2float(ga), __high2float(ga), __low2float(gb), __high2float(gb)};
                const uint2 wq = hq ? make_uint2(wraw.z, wraw.w) : make_uint2(wraw.x, wraw.y);
                const uint2 eq = hq ? make_uint2(eraw.z, eraw.w) : make_uint2(eraw.x, eraw.y);
                optim_quad(a.o, lr_base, old_db, new_db, false, i8 + 4u * hq, g, wq, eq, ptrs, f0 + 2 * hq, size);
            }
        }
        cluster.sync();        // nobody leaves (or clears its slice for the next job) while the partner still reads it
        if (job == cluster_id) MON_TL_MARK(MON_TL_S + 2, a.ctrl->iter - 1);
    }
}

cudaError_t mon_launch_scatter_adam(const MonGrid& g, const MonOpt& o, uint32_t n_points, const uint32_t* live_cnt, const float* pts_c,
                                    const uint32_t* genc, const MonCtrl* ctrl, float* pf, __half* ph, float* m, float* v, uint32_t* ps,
                                    __half* ema, __half* planar, __half* gh_grid, float* grad_snap, bool fuse, uint32_t sm_count,
                                    cudaStream_t st, const MonLaunchOpt& lo) {
    static std::atomic<uint64_t> prepared{0};
    const cudaError_t prep = mon_once_per_device(prepared, [] {
        cudaError_t e = cudaFuncSetAttribute(k_scatter_adam, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_scatter_adam, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SO_SMEM_BYTES);
    });
    if (prep != cudaSuccess) return prep;
    SoArgs a;
    a.g = g; a.o = o; a.n_points = n_points; a.n_jobs = mon_scatter_adam_jobs(g); a.fuse = fuse ? 1u : 0u;
    a.live_cnt = live_cnt; a.pts_c = pts_c; a.genc = genc; a.ctrl = ctrl;
    a.pf = pf; a.ph = ph; a.m = m; a.v = v; a.ps = ps; a.ema = ema; a.planar = planar; a.gh = gh_grid; a.grad_snap = grad_snap;
    if (a.n_jobs == 0) return cudaSuccess;
    // one cluster per job while the chip has room for them (one CTA per SM: 128 KB of shared memory), else clusters loop
    uint32_t n_clusters = a.n_jobs;
    const uint32_t max_clusters = sm_count / SO_CLUSTER;
    if (n_clusters > max_clusters) n_clusters = max_clusters ? max_clusters : 1u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * SO_CLUSTER);
    cfg.blockDim = dim3(SO_THREADS);
    cfg.dynamicSmemBytes = SO_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[3];
    unsigned n = 0;
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = SO_CLUSTER; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
    if (lo.pdl && (mon_pdl_mask() & MON_PDL_SCATTER)) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, k_scatter_adam, a);
}
