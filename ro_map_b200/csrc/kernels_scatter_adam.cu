// kernels_scatter_adam.cu — hash-grid gradient scatter FUSED with the per-parameter Adam / EMA update.
//
// Replaces, in one launch, cudaMemsetAsync(grid gradients) + kernel_grid_backward (TCNN encodings/grid.h:1132,386-509) and
// the grid part of adam_step / ema_step_half_precision (TCNN optimizers/adam.h:48-118, ema.h:62-76).
//
// The reference scatters 8 atomicAdd(__half2) per (sample, level) into a global fp16 gradient table and sweeps the whole
// parameter vector afterwards.  On B200 that scatter is bound by the number of reduction requests an SM can send to L2: one
// per lane and corner, 120 G/s measured for the whole chip (tools/smem_atomic_probe.cu, profiles/r4b_smem_atomic_probe.txt).
// A fresh object, whose every sample carries gradient, issues 16.8 M of them: 87-108 us, 5x the rest of the iteration.
// Shared memory takes the same updates 7-25x faster (compare-and-swap on an f16x2 word 830 G/s, native integer add 2940 G/s)
// — but only the SM's OWN shared memory: atomics into a cluster partner's (distributed shared memory) run at 47-159 G/s,
// no better than L2.  So the gradient table is made the resident operand, like the weight table in the encode kernel, and
// the job structure is chosen so that a CTA finds the corners that belong to its slice WITHOUT testing all eight:
//
//   job      = one PARITY class of one level's table: the entries with (index & 1) == q, <= 32768 f16x2 accumulators = 128 KB
//              of shared memory; 2 jobs per level, 32 for base.json.  Both hash primes are odd, so the parity of a corner's
//              entry is (x ^ y ^ z) & 1 of the corner's lattice coordinates (x & 1 for the dense levels, whose y / z strides are
//              even): exactly FOUR of the eight corners of every cell fall into each class, and which four follows from the
//              cell's own parity — branch-free, no index is computed for a corner of the other class.
//   cluster  = 4 CTAs work on one job.  Each takes a quarter of the iteration's LIVE samples (the fused MLP kernel hands
//              over only the samples with a non-zero gradient row, compacted: position + one word per level) and adds its
//              four contributions per sample into its PRIVATE copy of the slice: fp16 products accumulated in fp16, exactly
//              the arithmetic of the reference's atomicAdd(__half2), in an order that is not defined there either.
//   reduce   = after a cluster barrier every CTA owns a quarter of the slice: it sums the four private copies through
//              distributed shared memory (bulk, coalesced reads: fast, unlike scattered remote atomics) in a fixed order
//              and rounds to fp16: the loss-scaled fp16 gradient the reference keeps
//   update   = and applies Adam + EMA to those entries in place — the complete gradient of an entry is known inside the
//              cluster, which is the grid-wide phase boundary a fused scatter + Adam needs.  Entries with a zero
//              gradient are skipped by Adam exactly as in the reference (adam.h:75-79) and only EMA-filtered.
//
// Measured and rejected on the way (profiles/r4c_*): slices by index RANGE (a quarter of a level per job, int32 fixed-point
// accumulators with native ATOMS.ADD, 2-CTA clusters) — every job tests all eight corners of every sample, three quarters of
// them for nothing, and the kernel is instruction-bound: 207 us for a fresh object, 47 us in steady state.
//
// The fp16 gradient table never exists in global memory (3.8 MB written by atomics, read and zeroed again per iteration),
// and the launch between scatter and optimizer is gone.
//
// STATUS: opt-in (MON_SCATTER_SMEM=1), parity-tested, NOT the default: on B200 it measured slower than the global-reduction
// scatter + optimizer sweep in both phases of training (profiles/r4e_*, r4f_*: steady state 70 vs 35 us, fresh object 230 vs
// 107 us).  Two costs the design does not remove: owning every other entry makes the Adam phase read half-used sectors of
// the fp32 state (55 us instead of 13), and on the coarse hashed levels of a fresh object the compare-and-swap loops of 1024
// threads collide (75-146 us for levels 2-10 against 36 us for levels 13-15).
#include <cooperative_groups.h>

#include "mon_device.cuh"
#include "mon_kernels.h"
#include "optim_math.cuh"
#include "mon_timeline.cuh"
MON_TL_DEFINE(scatter_adam)
#ifdef MON_TIMELINE
// per CTA [start, scatter phase done, reduce + Adam phase done] of its first job, for the iteration with (iter % 64) == 20
static __device__ unsigned long long mon_tl_so_cta[256 * 3];
extern "C" int mon_debug_tl_so_cta_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, mon_tl_so_cta, sizeof(mon_tl_so_cta)); }
#define SO_CTA_STAMP(slot) do { if (threadIdx.x == 0 && blockIdx.x < 256 && (a.ctrl->iter - 1) % 64 == 20) { \
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); mon_tl_so_cta[blockIdx.x * 3 + (slot)] = t_; } } while (0)
#else
#define SO_CTA_STAMP(slot) do { } while (0)
#endif

namespace cg = cooperative_groups;

#define SO_THREADS 1024
#define SO_CLUSTER 4
#define SO_SLICE_ENTRIES 32768u                  // f16x2 accumulators per job
#define SO_SMEM_BYTES (SO_SLICE_ENTRIES * 4u)

struct SoArgs {
    MonGrid g;
    MonOpt o;
    uint32_t n_points;          // N: stride (in words) between the levels of genc
    uint32_t n_jobs;            // 2 * n_levels
    uint32_t fuse;              // 1: Adam + EMA in this kernel; 0: store the reduced gradient into gh
    const uint32_t* live_cnt;   // [2], indexed by iteration parity
    const float* pts_c;         // [n_live][3]
    const uint32_t* genc;       // [n_levels][N]: slot k's two fp16 gradients of the level
    const MonCtrl* ctrl;        // the iteration's control block (copy taken by the fused MLP kernel)
    float* pf; __half* ph; float* m; float* v; uint32_t* ps; __half* ema; __half* planar;
    __half* gh;                 // grid gradients [n_grid] (fuse == 0 only)
    float* grad_snap;           // parity hook: loss-scaled gradient of every parameter as float (nullptr in production)
};

uint32_t mon_scatter_adam_jobs(const MonGrid& g) { return 2u * g.n_levels; }

// the kernel's parity rule needs: power-of-two table (index = hash & (size - 1)), even size, and for dense levels an even
// resolution (y / z strides even).  True for every level of every supported configuration; checked on the host.
bool mon_scatter_adam_supported(const MonGrid& g) {
    for (uint32_t l = 0; l < g.n_levels; ++l) {
        const uint32_t size = g.size[l];
        if (size < 8 || (size & (size - 1)) != 0 || size / 2 > SO_SLICE_ENTRIES) return false;
        if (!g.hashed[l] && (g.res[l] & 1u)) return false;
    }
    return true;
}

// Four accumulator words += fp16 pairs: compare-and-swap on the CTA's own shared memory, all four in flight together (four
// loads, four adds, four swaps; only a swap that lost against another thread is retried on its own) — the latency of one
// read-modify-write per sample instead of four.
__device__ __forceinline__ void so_add4_f16x2(const uint32_t (&addr)[4], const __half2 (&v)[4]) {
    uint32_t old[4], got[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old[c]) : "r"(addr[c]) : "memory");
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&old[c]), v[c]);
        asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(got[c]) : "r"(addr[c]), "r"(old[c]), "r"(*reinterpret_cast<const uint32_t*>(&sum)) : "memory");
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        while (got[c] != old[c]) {      // lost: another thread (or an earlier corner of this one) changed the word in between
            old[c] = got[c];
            const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&old[c]), v[c]);
            asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(got[c]) : "r"(addr[c]), "r"(old[c]), "r"(*reinterpret_cast<const uint32_t*>(&sum)) : "memory");
        }
    }
}

// Coarse (dense) levels: thousands of samples land on a few hundred entries and a compare-and-swap loop would spin.  Their
// slices are small enough (<= 16384 entries per parity class) for two 32-bit FIXED-POINT words per entry, unit 2^-24: every
// contribution is rounded to fp16 first, exactly as the reference rounds it before its atomicAdd(__half2) (grid.h:427-431),
// and every fp16 value is a multiple of 2^-24, so the integer sum is exact and native shared-memory integer atomics
// (ATOMS.ADD: no retry, same-address updates are serialised by the hardware) apply.  Range +-128 in loss-scaled units
// (cvt.rni.s32.f32 saturates a single contribution; gradients of this loss are ~1e-3).
#define SO_FIXED_ONE 16777216.0f
#define SO_INT_ENTRIES 16384u
__device__ __forceinline__ void so_add_fixed(uint32_t addr8, __half2 v) {
    const float2 f = __half22float2(v);
    const int i0 = __float2int_rn(f.x * SO_FIXED_ONE), i1 = __float2int_rn(f.y * SO_FIXED_ONE);
    if (i0) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr8), "r"(i0) : "memory");
    if (i1) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr8 + 4u), "r"(i1) : "memory");
}

// slot handled by thread `tid` in trip `k` of a CTA's share.  Hashed levels: consecutive lanes take consecutive slots
// (coalesced loads; their entries are spread by the hash).  Dense levels: consecutive slots are neighbouring samples of one
// ray, which share corner entries at coarse resolution — a warp would hit the same accumulator word from many lanes and the
// compare-and-swap loops would serialise.  There, inside every block of 1024 slots, lane i of warp w takes slot 32 * i + w:
// the lanes of a warp are 32 samples apart, i.e. on 32 different rays.
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
    SO_CTA_STAMP(0);
    const uint32_t n_live = a.live_cnt[(a.ctrl->iter - 1) & 1u];
    const float lr_base = a.ctrl->lr_base, old_db = a.ctrl->ema_old, new_db = a.ctrl->ema_new;
    const uint32_t s_begin = (uint32_t)((uint64_t)n_live * rank / SO_CLUSTER), s_end = (uint32_t)((uint64_t)n_live * (rank + 1) / SO_CLUSTER);
    const uint32_t n_trips = (s_end - s_begin + SO_THREADS - 1) / SO_THREADS;
    __syncthreads();

    bool first = true;
    for (uint32_t job = cluster_id; job < a.n_jobs; job += n_clusters) {
        const uint32_t l = job >> 1, q = job & 1u;
        if (!first) {
            // the previous job's slice was read by the partners until the cluster barrier at its end
            uint4* z = reinterpret_cast<uint4*>(so_smem);
            for (uint32_t i = tid; i < SO_SMEM_BYTES / 16; i += SO_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }
        first = false;

        // ---- phase 1: this CTA's quarter of the live samples -> private accumulators (entry index >> 1 inside the class)
        const uint32_t size = a.g.size[l], res = a.g.res[l];
        const float scale = a.g.scale[l];
        const bool hashed = a.g.hashed[l] != 0;
        const uint32_t* gl = a.genc + (size_t)l * a.n_points;
        const uint32_t emask = size - 1u;
        const uint32_t my = hashed ? 2654435761u : res, mz = hashed ? 805459861u : res * res;
        const bool fixed = !hashed && size / 2 <= SO_INT_ENTRIES;   // small dense level: fixed-point accumulators, native integer atomics
        for (uint32_t k = 0; k < n_trips; ++k) {
            const uint32_t s = so_slot(s_begin, k, tid, !hashed);
            if (s >= s_end) continue;
            const uint32_t gw = __ldg(gl + s);
            const float u0 = __ldg(a.pts_c + (size_t)s * 3), u1 = __ldg(a.pts_c + (size_t)s * 3 + 1), u2 = __ldg(a.pts_c + (size_t)s * 3 + 2);
            if ((gw & 0x7fff7fffu) == 0u) continue;       // adding +0 is an identity
            const float g0 = __half2float(__ushort_as_half((unsigned short)(gw & 0xffffu)));
            const float g1 = __half2float(__ushort_as_half((unsigned short)(gw >> 16)));
            float fr[3]; uint32_t cell[3];
            mon_pos_fract(u0, scale, fr[0], cell[0]);
            mon_pos_fract(u1, scale, fr[1], cell[1]);
            mon_pos_fract(u2, scale, fr[2], cell[2]);
            // x offset of the class's corner for (dy, dz) = (0, 0): hashed  dx = q ^ parity(x ^ y ^ z) [^ dy ^ dz],  dense  dx = q ^ (x & 1)
            const uint32_t dx00 = (q ^ cell[0] ^ (hashed ? (cell[1] ^ cell[2]) : 0u)) & 1u;
            const float wx[2] = {__fsub_rn(1.0f, fr[0]), fr[0]}, wy[2] = {__fsub_rn(1.0f, fr[1]), fr[1]}, wz[2] = {__fsub_rn(1.0f, fr[2]), fr[2]};
            const uint32_t ay[2] = {cell[1] * my, (cell[1] + 1u) * my};
            const uint32_t az[2] = {cell[2] * mz, (cell[2] + 1u) * mz};
            uint32_t addr[4];
            __half2 val[4];
#pragma unroll
            for (uint32_t c = 0; c < 4; ++c) {
                const uint32_t dy = c & 1u, dz = c >> 1;
                const uint32_t dx = hashed ? (dx00 ^ dy ^ dz) : dx00;
                const uint32_t cx = cell[0] + dx;
                const uint32_t idx = (hashed ? (cx ^ ay[dy] ^ az[dz]) : (cx + ay[dy] + az[dz])) & emask;
                // same multiplication order as the forward kernel and the reference: (wx * wy) * wz
                const float wgt = __fmul_rn(__fmul_rn(dx ? wx[1] : wx[0], wy[dy]), wz[dz]);
                addr[c] = slice_addr + (idx >> 1) * (fixed ? 8u : 4u);
                val[c] = __floats2half2_rn(__fmul_rn(g0, wgt), __fmul_rn(g1, wgt));
            }
            if (fixed) {
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c) so_add_fixed(addr[c], val[c]);
            } else {
                so_add4_f16x2(addr, val);
            }
        }
        cluster.sync();        // every private copy of the slice is complete and visible cluster-wide
        if (job == cluster_id) { MON_TL_MARK(MON_TL_S + 1, a.ctrl->iter - 1); SO_CTA_STAMP(1); }

        // ---- phase 2: this CTA's quarter of the slice: sum the four copies, round to fp16, then Adam + EMA in place.
        // A thread takes 4 consecutive accumulator words = the class's entries 2j+q .. 2(j+3)+q of the level.
        const uint4* copies[SO_CLUSTER];
#pragma unroll
        for (uint32_t r = 0; r < SO_CLUSTER; ++r) copies[r] = reinterpret_cast<const uint4*>(cluster.map_shared_rank(so_smem, r));
        const uint32_t n_chunks = size / 8;                                // slice entries = size / 2, 4 per chunk (size % 8 == 0)
        const uint32_t c_begin = (uint32_t)((uint64_t)n_chunks * rank / SO_CLUSTER), c_end = (uint32_t)((uint64_t)n_chunks * (rank + 1) / SO_CLUSTER);
        const OptimPtrs ptrs = {a.pf, a.ph, a.m, a.v, a.ps, a.ema};
        for (uint32_t c = c_begin + tid; c < c_end; c += SO_THREADS) {
            float acc[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (fixed) {
                int isum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
                for (uint32_t r = 0; r < SO_CLUSTER; ++r) {
                    const int4 w0 = reinterpret_cast<const int4*>(copies[r])[2 * c], w1 = reinterpret_cast<const int4*>(copies[r])[2 * c + 1];
                    isum[0] += w0.x; isum[1] += w0.y; isum[2] += w0.z; isum[3] += w0.w;
                    isum[4] += w1.x; isum[5] += w1.y; isum[6] += w1.z; isum[7] += w1.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = __fmul_rn((float)isum[e], 1.0f / SO_FIXED_ONE);
            } else {
#pragma unroll
                for (uint32_t r = 0; r < SO_CLUSTER; ++r) {
                    const uint4 w = copies[r][c];
                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ww[e]));
                        acc[2 * e] = __fadd_rn(acc[2 * e], f.x);
                        acc[2 * e + 1] = __fadd_rn(acc[2 * e + 1], f.y);
                    }
                }
            }
#pragma unroll
            for (uint32_t e = 0; e < 4; ++e) {
                // the gradient the reference keeps is fp16 (loss-scaled): one rounding here
                const __half2 gh2 = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
                const uint32_t entry = 2u * (4u * c + e) + q;                   // within the level
                const uint32_t i2 = a.o.n_mlp + 2u * (a.g.offset[l] + entry);   // the entry's two parameters
                const float2 gf = __half22float2(gh2);
                if (a.grad_snap) { a.grad_snap[i2] = gf.x; a.grad_snap[i2 + 1] = gf.y; }
                if (!a.fuse) {
                    *reinterpret_cast<__half2*>(a.gh + (i2 - a.o.n_mlp)) = gh2;
                    continue;
                }
                optim_pair(a.o, lr_base, old_db, new_db, i2, gf.x, gf.y, ptrs, a.planar + (size_t)a.g.offset[l] * 2 + entry, size);
            }
        }
        cluster.sync();        // nobody leaves (or clears its slice for the next job) while a partner still reads it
        if (job == cluster_id) { MON_TL_MARK(MON_TL_S + 2, a.ctrl->iter - 1); SO_CTA_STAMP(2); }
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
    if (!mon_scatter_adam_supported(g)) return cudaErrorNotSupported;
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
