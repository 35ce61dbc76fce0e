// kernels_scatter_smem.cu — hash-grid gradient scatter with the GRADIENT TABLE as the shared-memory resident operand.
//
// Replaces cudaMemsetAsync(grid gradients) + kernel_grid_backward (TCNN encodings/grid.h:1132,386-509) for iterations in
// which many samples carry gradient (every sample of a fresh object does: 16.8 M corner updates per iteration).
//
// The reference issues one atomicAdd(__half2) per (sample, level, corner) into a global table.  On B200 that form is bound
// by the number of reduction lane-operations the SMs can send to L2 (REDG: ~1.3 cycles per lane and SM, ~200 G/s for the
// chip): 75-110 us for a fresh object, more than the rest of the iteration together.  Shared memory takes integer atomics
// 13x faster (ATOMS.ADD: ~0.1 cycle per lane, profiles/r4b_smem_atomic_probe.txt) — floating-point shared-memory atomics
// do not exist in hardware (f32 / f16x2 / u64 compile to compare-and-swap loops).  So:
//
//   job        = (level, parity class q, feature f): the entries of the level with (index & 1) == q, one feature:
//                <= 32768 accumulators of 32 bits = 128 KB of shared memory.  Both hash primes are odd, so the parity of a
//                corner's entry is (x ^ y ^ z) & 1 of its lattice coordinates (x & 1 on the dense levels, whose y / z strides
//                are even): exactly FOUR of the eight corners of every cell fall into each class, and which four follows
//                from the cell's own parity — no index is computed for a corner of the other class, no membership test, no
//                divergence.  (Measured alternative, profiles/r5b_*: classes index & 3 with both features per job — four
//                candidate corners of which two pass on average — costs 140 instead of 81 instructions per visit.)
//   accumulate = every contribution is rounded to fp16 exactly as the reference rounds it before its atomicAdd(__half2)
//                (grid.h:427-431), then added as a 32-bit FIXED-POINT number, unit 2^-24: every fp16 value is a multiple
//                of 2^-24, so the integer sum is exact and independent of the order (native ATOMS.ADD, no retry loop).
//                Range +-128 in loss-scaled units per private slice (a single contribution saturates in cvt.rni; measured
//                gradients of this loss stay below 0.4, tools/grad_range_probe.py).
//   work split = the flattened [job][live sample] space is cut into one contiguous piece per CTA, equal in modelled cost,
//                so a CTA holds one slice at a time and touches at most two or three jobs.
//   flush      = the CTA converts its accumulators to fp16 pairs in place (one rounding of the exact partial sum) and adds
//                the slice to the global gradient table with ONE TMA bulk reduction per 16 KB
//                (cp.reduce.async.bulk.global.shared::cta.add.noftz.f16 -> SASS UBLKRED.G.S.ADD.F16.RN): the L2 does the
//                element-wise fp16 additions at copy speed instead of one REDG lane-operation per entry.  The two or three
//                partial sums of an entry meet in L2 in fp16, the reference's own accumulation type.
//
// The global table this kernel adds into is CLASS-PLANAR: per level [q = 0: f0 | f1][q = 1: f0 | f1], each size/2 fp16 —
// entry i, feature f of a level sits at ((i & 1) * 2 + f) * size/2 + (i >> 1), a job's slice is contiguous.  The optimizer
// sweep reads and zeroes it (kernels_optim.cu) in iterations that took this path; iterations with few live samples (steady
// state: the early stop leaves ~1 sample in 12) take the OTHER PATH OF THIS KERNEL — global f16x2 reductions, the reference's
// own form (scatter_global.cuh), into the entry-ordered table — because there the fixed costs of the resident form (clearing,
// converting and flushing 128 KB slices) exceed the reductions saved.  The iteration's live-sample count decides on the device,
// uniformly for the grid; one launch per iteration either way.
#include "mon_device.cuh"
#include "mon_kernels.h"
#include "scatter_global.cuh"
#include "mon_timeline.cuh"
MON_TL_DEFINE(scatter_smem)
#ifdef MON_TIMELINE
// per CTA [start, first job accumulated, done] for the iteration with (iter % 64) == 20
static __device__ unsigned long long mon_tl_sr_cta[256 * 3];
extern "C" int mon_debug_tl_sr_cta_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, mon_tl_sr_cta, sizeof(mon_tl_sr_cta)); }
#define SR_CTA_STAMP(slot) do { if (threadIdx.x == 0 && blockIdx.x < 256 && (a.ctrl->iter - 1) % 64 == 20) { \
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); mon_tl_sr_cta[blockIdx.x * 3 + (slot)] = t_; } } while (0)
#else
#define SR_CTA_STAMP(slot) do { } while (0)
#endif

#define SR_THREADS 1024
#define SR_SLICE_WORDS 32768u                    // 32-bit accumulators per job
#define SR_SMEM_ACC_BYTES (SR_SLICE_WORDS * 4u)
#define SR_FIXED_ONE 16777216.0f                 // 2^24: 1 / (smallest fp16 subnormal)
#define SR_BULK_CHUNK 16384u
// software pipeline: a thread's loads run SR_AHEAD samples ahead of its arithmetic (the plain load-then-use form spent 20 % of
// its issue slots stalled on the first use of a loaded position).  Measured alternative (profiles/r5d_*, r5e_*): TMA bulk
// copies into a 2- or 4-stage shared-memory ring with full / empty mbarriers — slower (47 us against 40): the producer thread
// sits in a warp that also computes, the arbiter favours high warp ids, and every warp ends up waiting on the full barrier.
// Also measured and dropped (r5l): floor by the 1.5 * 2^23 add trick instead of F2I.FLOOR + I2FP — same time, the conversion
// pipe is not what binds the loop.
#define SR_AHEAD 2
#define SR_TILE 64                               // slots per tile of the global-reduction path
#define SR_SMEM_BYTES SR_SMEM_ACC_BYTES
// work split: cost of one (sample, job) in 1/64 of a hashed-level one.  On the coarse dense levels neighbouring samples of a
// ray (= neighbouring lanes) share corner entries and their shared-memory atomics are serialised (measured per-CTA times).
#define SR_W_HASH 64u
#define SR_W_DENSE 98u
// every job a piece touches ends with a flush (convert + bulk reduction + clearing the slice for the next job): ~3.6 us, the time
// of ~7500 hashed-level samples.  Charged at each job START: a piece that spans a job boundary gets that much less sample work.
#define SR_FLUSH 480000ull
// opt-in occupancy mode only: there the batch kernel of the NEXT iteration (one cluster) runs on the forked branch beside this
// kernel, and a full-chip grid would leave some of its one-per-SM CTAs waiting ~13 us for those SMs; the caller passes the number
// of SMs to leave free (default graphs: 0 — the next batch is generated beside the hash encode instead)
#define SR_SPARE_SMS 8u

struct SrArgs {
    MonGrid g;
    uint32_t n_points;          // N: stride (in words) between the levels of genc
    uint32_t n_jobs;            // 4 * n_levels
    uint32_t n_slow;            // jobs [0, n_slow) are the coarse dense levels (cost SR_W_DENSE)
    uint32_t min_live;          // iterations with fewer live samples are left to the global-reduction kernel
    const uint32_t* live_cnt;   // [2], indexed by iteration parity
    const float* pts_c;         // [n_live][4]: x, y, z, unused
    const uint32_t* genc;       // [n_levels][N]: slot k's two fp16 gradients of the level
    const MonCtrl* ctrl;        // the iteration's control block (copy taken by the fused MLP kernel)
    __half* gcls;               // class-planar gradient table [n_grid] fp16 (resident path)
    __half* gh_grid;            // entry-ordered gradient table [n_grid] fp16 (global-reduction path)
};

// needs: power-of-two tables (index = hash & (size - 1)) of >= 16 entries (16-byte bulk granularity of a half-size fp16
// slice), at most 2 * SR_SLICE_WORDS entries, and an even resolution on the dense levels (even y / z strides: the parity of a
// dense index is x & 1).  True for base_resolution = 2^k with per_level_scale = 2 (the reference's base.json).
bool mon_scatter_resident_supported(const MonGrid& g) {
    for (uint32_t l = 0; l < g.n_levels; ++l) {
        const uint32_t size = g.size[l];
        if (size < 16 || (size & (size - 1)) != 0 || size / 2 > SR_SLICE_WORDS) return false;
        if (!g.hashed[l] && (g.res[l] & 1u)) return false;
    }
    return true;
}

struct SrSample { uint32_t gw; float u0, u1, u2; };

// slot s of the run; slots beyond the end read as a zero gradient at the origin (adds four integer zeros)
__device__ __forceinline__ SrSample sr_load(const float* __restrict__ pts_c, const uint32_t* __restrict__ gl, uint32_t s, uint32_t s_end) {
    SrSample r;
    r.gw = 0u; r.u0 = r.u1 = r.u2 = 0.0f;
    if (s < s_end) {
        r.gw = __ldg(gl + s);
        const float4 u = __ldg(reinterpret_cast<const float4*>(pts_c) + s);
        r.u0 = u.x; r.u1 = u.y; r.u2 = u.z;
    }
    return r;
}

// one (sample, job): the four corners of the job's parity class, one feature
template <bool HASHED>
__device__ __forceinline__ void sr_item(const SrSample& sm, uint32_t f, uint32_t q, float scale, uint32_t bmask, uint32_t my, uint32_t mz, unsigned char* acc) {
    // branch-free: a zero gradient (rare among the compacted live samples) adds four integer zeros
    const float g = __half2float(__ushort_as_half((unsigned short)(f ? (sm.gw >> 16) : (sm.gw & 0xffffu))));
    float fr[3]; uint32_t cell[3];
    mon_pos_fract(sm.u0, scale, fr[0], cell[0]);
    mon_pos_fract(sm.u1, scale, fr[1], cell[1]);
    mon_pos_fract(sm.u2, scale, fr[2], cell[2]);
    // x offset of the class's corner for (dy, dz) = (0, 0): hashed  dx = q ^ parity(x ^ y ^ z) [^ dy ^ dz],  dense  dx = q ^ (x & 1)
    const uint32_t dx00 = (q ^ cell[0] ^ (HASHED ? (cell[1] ^ cell[2]) : 0u)) & 1u;
    const float wx[2] = {__fsub_rn(1.0f, fr[0]), fr[0]}, wy[2] = {__fsub_rn(1.0f, fr[1]), fr[1]}, wz[2] = {__fsub_rn(1.0f, fr[2]), fr[2]};
    // per-axis terms of 2 * index: the byte offset of accumulator (index >> 1) is (2 * index) & (2 * size - 4)
    const uint32_t ax[2] = {cell[0] << 1, (cell[0] + 1u) << 1};
    const uint32_t ay[2] = {(cell[1] * my) << 1, ((cell[1] + 1u) * my) << 1};
    const uint32_t az[2] = {(cell[2] * mz) << 1, ((cell[2] + 1u) * mz) << 1};
    const float wxa = dx00 ? wx[1] : wx[0], wxb = dx00 ? wx[0] : wx[1];       // dx = dx00 / dx00 ^ 1
    const uint32_t axa = dx00 ? ax[1] : ax[0], axb = dx00 ? ax[0] : ax[1];
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        const uint32_t dy = k & 1u, dz = k >> 1;
        const bool flip = HASHED && ((dy ^ dz) != 0u);
        const uint32_t axc = flip ? axb : axa;
        const uint32_t off = (HASHED ? (axc ^ ay[dy] ^ az[dz]) : (axc + ay[dy] + az[dz])) & bmask;
        // same multiplication order as the forward kernel and the reference: (wx * wy) * wz; the product with the gradient
        // is rounded to fp16 (the reference's (T)(weight * grad)), which makes it an exact multiple of 2^-24
        const float wgt = __fmul_rn(__fmul_rn(flip ? wxb : wxa, wy[dy]), wz[dz]);
        // times 2^24 as an integer add on the exponent field (ALU pipe instead of one more FMUL): a widened fp16 value is zero
        // or a normal float below 2^16; zero becomes 2^-103, which converts to 0 as well
        const float v = __uint_as_float(__float_as_uint(__half2float(__float2half_rn(__fmul_rn(g, wgt)))) + (24u << 23));
        atomicAdd(reinterpret_cast<int*>(acc + off), __float2int_rn(v));   // result unused: ATOMS.ADD without a return value
    }
}

// accumulate slots s_first, s_first + SR_THREADS, ... < s_end of level row gl
template <bool HASHED>
__device__ __forceinline__ void sr_samples(const float* __restrict__ pts_c, const uint32_t* __restrict__ gl, uint32_t s_first, uint32_t s_end,
                                           uint32_t f, uint32_t q, float scale, uint32_t size, uint32_t res, unsigned char* acc) {
    const uint32_t bmask = 2u * size - 4u;
    const uint32_t my = HASHED ? 2654435761u : res, mz = HASHED ? 805459861u : res * res;
    // a ring of SR_AHEAD + 1 register sets, fully unrolled: while sample k is computed, the loads of k+1 .. k+SR_AHEAD are in flight
    SrSample r[SR_AHEAD + 1];
#pragma unroll
    for (uint32_t j = 0; j < SR_AHEAD; ++j) r[j] = sr_load(pts_c, gl, s_first + j * SR_THREADS, s_end);
    for (uint32_t s = s_first; s < s_end; s += (SR_AHEAD + 1) * SR_THREADS) {
#pragma unroll
        for (uint32_t j = 0; j <= SR_AHEAD; ++j) {
            r[(j + SR_AHEAD) % (SR_AHEAD + 1)] = sr_load(pts_c, gl, s + (j + SR_AHEAD) * SR_THREADS, s_end);
            if (s + j * SR_THREADS < s_end) sr_item<HASHED>(r[j], f, q, scale, bmask, my, mz, acc);
        }
    }
}

// position in the flattened [job][sample] space at cumulative cost x (monotone; the same expression gives one CTA's end and
// the next one's begin, so the pieces tile the space exactly)
struct SrPos { uint32_t job, p; };
__host__ __device__ __forceinline__ uint64_t sr_cost_total(uint32_t n, uint32_t n_slow, uint32_t n_jobs) {
    return (uint64_t)n_slow * ((uint64_t)n * SR_W_DENSE + SR_FLUSH) + (uint64_t)(n_jobs - n_slow) * ((uint64_t)n * SR_W_HASH + SR_FLUSH);
}
__host__ __device__ __forceinline__ SrPos sr_cost_to_pos(uint64_t x, uint32_t n, uint32_t n_slow) {
    const uint64_t slow = (uint64_t)n * SR_W_DENSE + SR_FLUSH, fast = (uint64_t)n * SR_W_HASH + SR_FLUSH, c0 = slow * n_slow;
    uint64_t per, w; uint32_t base;
    if (x < c0) { per = slow; w = SR_W_DENSE; base = 0; }
    else { x -= c0; per = fast; w = SR_W_HASH; base = n_slow; }
    const uint64_t j = x / per, rem = x - j * per;
    SrPos r;
    r.job = base + (uint32_t)j;
    r.p = rem <= SR_FLUSH ? 0u : (uint32_t)((rem - SR_FLUSH) / w);     // < n because rem < per
    return r;
}

__global__ void __launch_bounds__(SR_THREADS, 1)
k_scatter(const __grid_constant__ SrArgs a) {
    extern __shared__ __align__(128) unsigned char sr_smem[];
    const uint32_t tid = threadIdx.x;
    const uint32_t acc_addr = (uint32_t)__cvta_generic_to_shared(sr_smem);

    mon_pdl_wait();       // the fused MLP kernel (live samples, control block) has completed
    mon_pdl_trigger();
    if (a.ctrl->skip) return;                          // uniform over the grid
    const uint32_t n_live = a.live_cnt[(a.ctrl->iter - 1) & 1u];
    if (n_live == 0u) return;
    if (n_live < a.min_live) {
        // ---- few live samples (steady state): global f16x2 reductions, the reference's form.  Each half of the CTA (16 warps)
        // takes tiles of 64 consecutive slots: positions staged in shared memory, warp w scatters level w for the tile — every
        // lane has real work and the level (table base, scale, hash or dense) is uniform across a warp.
        MON_TL(MON_TL_S, a.ctrl->iter - 1);
        float (*s_u)[SR_TILE] = reinterpret_cast<float (*)[SR_TILE]>(sr_smem) + 3 * (tid >> 9);      // [axis][slot] of this half
        const uint32_t gt = tid & 511u, lane = tid & 31u, warp = gt >> 5, bar_id = 1u + (tid >> 9);
        const uint32_t n_tiles = (n_live + SR_TILE - 1) / SR_TILE;
        for (uint32_t tile = 2u * blockIdx.x + (tid >> 9); tile < n_tiles; tile += 2u * gridDim.x) {
            const uint32_t s0 = tile * SR_TILE, n_tile = min((uint32_t)SR_TILE, n_live - s0);
            for (uint32_t i = gt; i < 3 * n_tile; i += 512u) s_u[i % 3][i / 3] = __ldg(a.pts_c + (size_t)(s0 + i / 3) * 4 + i % 3);   // slots are (x, y, z, -)
            asm volatile("bar.sync %0, 512;" ::"r"(bar_id) : "memory");
            for (uint32_t l = warp; l < a.g.n_levels; l += 16u) {
                const uint32_t* gl = a.genc + (size_t)l * a.n_points + s0;
                for (uint32_t r = lane; r < n_tile; r += 32) {
                    const uint32_t gwj = __ldg(gl + r);
                    if ((gwj & 0x7fff7fffu) == 0u) continue;   // adding +0 is an identity
                    const float u[3] = {s_u[0][r], s_u[1][r], s_u[2][r]};
                    scatter_level(a.g, l, gwj, u, a.gh_grid);
                }
            }
            asm volatile("bar.sync %0, 512;" ::"r"(bar_id) : "memory");      // the tile's positions are no longer read
        }
        return;
    }
    // ---- many live samples (a fresh object: all of them): shared-memory resident slices
    {
        uint4* z = reinterpret_cast<uint4*>(sr_smem);
        for (uint32_t i = tid; i < SR_SMEM_ACC_BYTES / 16; i += SR_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    MON_TL(MON_TL_S + 1, a.ctrl->iter - 1);
    SR_CTA_STAMP(0);
    const uint64_t total = sr_cost_total(n_live, a.n_slow, a.n_jobs);
    const SrPos pos_b = sr_cost_to_pos(total * blockIdx.x / gridDim.x, n_live, a.n_slow);
    const SrPos pos_e = sr_cost_to_pos(total * (blockIdx.x + 1) / gridDim.x, n_live, a.n_slow);
    __syncthreads();

    bool first = true;
    for (uint32_t job = pos_b.job; job < a.n_jobs && (job < pos_e.job || (job == pos_e.job && pos_e.p > 0)); ++job) {
        const uint32_t p0 = job == pos_b.job ? pos_b.p : 0u;
        const uint32_t p1 = job == pos_e.job ? pos_e.p : n_live;
        if (p1 <= p0) continue;
        const uint32_t l = job >> 2, q = (job >> 1) & 1u, f = job & 1u;
        const uint32_t size = a.g.size[l], n_e = size >> 1;
        if (!first) {
            // the previous job's slice has been read by its bulk reductions (waited for below)
            uint4* z = reinterpret_cast<uint4*>(sr_smem);
            for (uint32_t i = tid; i < n_e / 4; i += SR_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }

        // ---- accumulate this CTA's samples of the job
        const uint32_t* gl = a.genc + (size_t)l * a.n_points;
        if (a.g.hashed[l]) sr_samples<true>(a.pts_c, gl, p0 + tid, p1, f, q, a.g.scale[l], size, a.g.res[l], sr_smem);
        else sr_samples<false>(a.pts_c, gl, p0 + tid, p1, f, q, a.g.scale[l], size, a.g.res[l], sr_smem);
        __syncthreads();
        if (first) SR_CTA_STAMP(1);
        first = false;

        // ---- exact partial sums -> fp16, in place (accumulator i: bytes [4i, 4i+4) -> [2i, 2i+2): read all, then write)
        int vals[SR_SLICE_WORDS / SR_THREADS];
#pragma unroll
        for (uint32_t k = 0; k < SR_SLICE_WORDS / SR_THREADS; ++k) {
            const uint32_t i = k * SR_THREADS + tid;
            vals[k] = i < n_e ? reinterpret_cast<const int*>(sr_smem)[i] : 0;
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < SR_SLICE_WORDS / SR_THREADS; ++k) {
            const uint32_t i = k * SR_THREADS + tid;
            if (i < n_e) reinterpret_cast<__half*>(sr_smem)[i] = __float2half_rn(__fmul_rn((float)vals[k], 1.0f / SR_FIXED_ONE));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk (async-proxy) reads
        __syncthreads();

        // ---- flush: the L2 adds the slice to the class-planar table
        const uint32_t bytes = n_e * 2u;
        const uint32_t off = tid * SR_BULK_CHUNK;
        if (off < bytes) {
            char* dst = reinterpret_cast<char*>(a.gcls + (size_t)a.g.offset[l] * 2 + (size_t)(q * 2u + f) * n_e) + off;
            const uint32_t n = min(SR_BULK_CHUNK, bytes - off);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.f16 [%0], [%1], %2;" ::"l"(dst), "r"(acc_addr + off), "r"(n) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the buffer may be reused
        }
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");                   // every reduction has been performed before the kernel ends
    SR_CTA_STAMP(2);
}

// host mirror of the kernel's work split: piece b of n_ctas as (first job, first sample, end job, end sample); used by the CPU
// test that checks the pieces tile the [job][sample] space exactly
void mon_scatter_resident_pieces_host(const MonGrid& g, uint32_t n_live, uint32_t n_ctas, uint32_t* out4) {
    uint32_t n_slow = 0;
    while (n_slow < g.n_levels && !g.hashed[n_slow] && g.size[n_slow] <= 32768u) ++n_slow;
    n_slow *= 4u;
    const uint64_t total = sr_cost_total(n_live, n_slow, 4u * g.n_levels);
    for (uint32_t b = 0; b < n_ctas; ++b) {
        const SrPos pb = sr_cost_to_pos(total * b / n_ctas, n_live, n_slow), pe = sr_cost_to_pos(total * (b + 1) / n_ctas, n_live, n_slow);
        out4[4 * b + 0] = pb.job; out4[4 * b + 1] = pb.p; out4[4 * b + 2] = pe.job; out4[4 * b + 3] = pe.p;
    }
}

cudaError_t mon_launch_scatter(const MonGrid& g, uint32_t n_points, uint32_t min_live, const uint32_t* live_cnt, const float* pts_c,
                               const uint32_t* genc, const MonCtrl* ctrl, __half* gcls, __half* gh_grid, uint32_t sm_count, cudaStream_t st,
                               const MonLaunchOpt& lo, bool leave_spare_sms) {
    static std::atomic<uint64_t> prepared{0};
    const cudaError_t prep = mon_once_per_device(prepared, [] {
        // 132 of the SM's 228 KB as shared memory, the rest stays L1: the optimizer sweep follows through a programmatic edge on SMs
        // that keep this split, and with the maximum carve-out its streaming loads had too few L1 lines in flight (27 instead of 16 us)
        // (the maximum carve-out here changes nothing now that the sweep follows through a plain edge: profiles/r9e_timeline_carve100.txt)
        cudaError_t e = cudaFuncSetAttribute(k_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, 58);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_SMEM_BYTES);
    });
    if (prep != cudaSuccess) return prep;
    if (!mon_scatter_resident_supported(g)) return cudaErrorNotSupported;
    if (n_points == 0 || g.n_levels == 0) return cudaSuccess;
    SrArgs a;
    a.g = g; a.n_points = n_points; a.n_jobs = 4u * g.n_levels; a.min_live = min_live;
    a.n_slow = 0;
    while (a.n_slow < g.n_levels && !g.hashed[a.n_slow] && g.size[a.n_slow] <= 32768u) ++a.n_slow;
    a.n_slow *= 4u;
    a.live_cnt = live_cnt; a.pts_c = pts_c; a.genc = genc; a.ctrl = ctrl; a.gcls = gcls; a.gh_grid = gh_grid;
    static const uint32_t spare = [] { const char* e = getenv("MON_SCATTER_SPARE_SMS"); return e ? (uint32_t)atoi(e) : SR_SPARE_SMS; }();
    const uint32_t ctas = leave_spare_sms && sm_count > 4u * spare ? sm_count - spare : sm_count;
    return mon_launch_chain(MON_PDL_SCATTER, lo, k_scatter, dim3(ctas), dim3(SR_THREADS), SR_SMEM_BYTES, st, a);
}
