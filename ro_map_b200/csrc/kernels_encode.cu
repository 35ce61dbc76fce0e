// kernels_encode.cu — sample positions, multiresolution hash-grid lookup (forward) and gradient scatter (backward).
//
// Replaces GenerateInputPoints (MON/Core/src/nerf_model.cu:536-566), tcnn kernel_grid
// (TCNN/include/tiny-cuda-nn/encodings/grid.h:220-384) and kernel_grid_backward (:386-509).
//
// FORWARD.  The reference launches one thread per (point, level) and gathers 8 x 4 bytes through L1/L2; on B200 that
// shape is bound by L1 wavefronts (one per distinct 128-B line: ~16-32 per warp instruction), not by HBM or L2.
// Here the table is the resident operand instead: the work is cut into (level, feature) JOBS whose table slice
// (<= 65536 fp16 = 128 KB, from a planar copy of the fp16 weights the optimizer maintains) is staged into shared
// memory with TMA bulk copies (cp.async.bulk -> mbarrier), and the CTA then streams its share of the points past
// it: 8 two-byte gathers per point hit the 32 shared-memory banks (~3.5-way conflicts for random indices) instead of
// 8 L1 wavefronts.  The flattened [job][point] space is cut into one piece per CTA, equal in modelled cost (the
// conflict degree differs between dense and hashed levels; staging a slice costs too), so each of the 148 CTAs loads
// at most two table slices and they all finish together.  Output is level-major: enc[level][point][feature] (the two features of a level form one
// 32-bit word per point, which is what the fused MLP kernel stages), written with 2-byte stores at 4-byte stride.
// Arithmetic is the reference's: weights in fp32, rounded to fp16, one fp32 FMA per corner whose result is rounded
// back to fp16 after every corner (grid.h:334 via common.h:539-559) — bit-exact against tiny-cuda-nn, including
// the 32-bit stride wrap that turns level 12 into a table indexed by x alone (mon_core.cu make_grid).
//
// BACKWARD.  grad[idx] += half2(d_enc * w) with f16x2 reductions, exactly the reference's atomicAdd(__half2)
// (grid.h:427-431), for the live samples the fused MLP kernel compacted, one level per warp (k_encode_backward).
#include "mon_device.cuh"
#include "mon_kernels.h"
#include "scatter_global.cuh"
#include "tc05.cuh"
#include "mon_timeline.cuh"
MON_TL_DEFINE(encode)
#ifdef MON_TIMELINE
// per-CTA [table resident, done] of the encode kernel for one iteration (iter % 64 == 20): load balance of the split
static __device__ unsigned long long mon_tl_enc_cta[256 * 2];
extern "C" int mon_debug_tl_enc_cta_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, mon_tl_enc_cta, sizeof(mon_tl_enc_cta)); }
#define ENC_CTA_STAMP(slot) do { if (threadIdx.x == 0 && blockIdx.x < 256 && ctrl && (ctrl->iter - 1) % 64 == 20) { \
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); mon_tl_enc_cta[blockIdx.x * 2 + (slot)] = t_; } } while (0)
#else
#define ENC_CTA_STAMP(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------- sample points
// ---- opt-in occupancy grid: is the unit-cube position u in an occupied cell?  (positions on or slightly outside the faces clamp)
MON_DEV bool mon_occ_test(const uint32_t* __restrict__ bits, uint32_t res, const float* u) {
    const float r = (float)res;
    const uint32_t x = (uint32_t)min(max((int)floorf(u[0] * r), 0), (int)res - 1);
    const uint32_t y = (uint32_t)min(max((int)floorf(u[1] * r), 0), (int)res - 1);
    const uint32_t z = (uint32_t)min(max((int)floorf(u[2] * r), 0), (int)res - 1);
    const uint32_t idx = x + res * (y + res * z);
    return (__ldg(bits + (idx >> 5)) >> (idx & 31u)) & 1u;
}

// A3: t_n = tmin + dt*(n + xi), p = o + t*d, u = (p - bmin) / (bmax - bmin)  (nerf_model.cu:545-565,140-144).
// One thread per sample; rays of the batch are already compacted and padded.  pts: [N][3] fp32, the reference's
// PointsInput.  in_box (render only): rays that miss the box get the cube centre (never composited).
__global__ void __launch_bounds__(256)
k_sample_points(uint32_t n_points, uint32_t S, const MonRay* __restrict__ rays, const int* __restrict__ in_box,
                const float* __restrict__ jitter, uint32_t seed, const MonCtrl* __restrict__ ctrl, uint32_t rng_stream,
                uint32_t iter_fixed, float bmin0, float bmin1, float bmin2, float bmax0, float bmax1, float bmax2,
                float* __restrict__ pts, const uint32_t* __restrict__ orig_ray, MonOcc occ) {
    mon_pdl_wait();
    mon_pdl_trigger();
    if (ctrl && ctrl->skip) return;
    MON_TL(MON_TL_P, ctrl ? ctrl->iter - 1 : 0u);
    const uint32_t pt = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = pt < n_points;
    const uint32_t ray = pt / S, n = pt - ray * S;
    float u[3] = {0.5f, 0.5f, 0.5f};
    if (in_range && (!in_box || in_box[ray])) {
        const MonRay r = rays[ray];
        // the batch kernel already advanced ctrl->iter; this iteration's counter is iter-1
        const uint32_t iter = ctrl ? ctrl->iter - 1 : iter_fixed;
        // render: the rays are compacted (hits only); the jitter belongs to the pixel, not to the slot
        const float xi = mon_rand(jitter, seed, iter, rng_stream, orig_ray ? orig_ray[ray] * S + n : pt);
        const float t = mon_sample_t(r, n, xi, (float)S);
        const float bmin[3] = {bmin0, bmin1, bmin2}, bmax[3] = {bmax0, bmax1, bmax2};
        mon_sample_point(r, t, bmin, bmax, u);
    }
    if (in_range) {
        pts[(size_t)pt * 3 + 0] = u[0];
        pts[(size_t)pt * 3 + 1] = u[1];
        pts[(size_t)pt * 3 + 2] = u[2];
    }
    if (occ.bits) {
        // opt-in occupancy mode (training, S == 32: a warp is one ray): the ray's occupancy mask for the fused MLP kernel and the
        // compacted list of occupied samples for the encode kernel; one atomicAdd per warp reserves the slots of its lanes
        const bool keep = in_range && mon_occ_test(occ.bits, occ.res, u);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t base = 0;
        if (lane == 0) {
            if (in_range) occ.ray_mask[ray] = m;
            if (m) base = atomicAdd(occ.count, (uint32_t)__popc(m));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) occ.list[base + __popc(m & ((1u << lane) - 1u))] = pt;
    }
}

void mon_launch_sample_points(uint32_t n_points, uint32_t S, const MonRay* rays, const int* in_box, const float* jitter,
                              uint32_t seed, const MonCtrl* ctrl, uint32_t rng_stream, uint32_t iter_fixed,
                              const float* bmin, const float* bmax, float* pts, cudaStream_t st, const MonLaunchOpt& lo, const uint32_t* orig_ray,
                              const MonOcc* occ) {
    MonOcc none = {nullptr, 0u, nullptr, nullptr, nullptr};
    // The same shared-memory carve-out as the hash-encode kernel, beside whose CTAs this kernel runs inside the iteration graphs: an
    // SM does not change its carve-out while CTAs are resident, so a kernel that prefers another split waits for the encode CTAs to
    // leave (measured: the sample points of the next iteration started only after the encode kernel had ended).
    static std::atomic<uint64_t> prepared{0};
    mon_once_per_device(prepared, [] { return cudaFuncSetAttribute(k_sample_points, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); });
    mon_launch_chain(MON_PDL_POINTS, lo, k_sample_points, dim3((n_points + 255) / 256), dim3(256), 0, st, n_points, S, rays, in_box, jitter, seed, ctrl, rng_stream,
                     iter_fixed, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2], pts, orig_ray, occ ? *occ : none);
}

// ---------------------------------------------------------------------------------------------- forward
#define ENC_THREADS 1024
#define ENC_TABLE_BYTES (65536 * 2)
#define ENC_BULK_CHUNK 16384u
#define ENC_UNROLL 2   // points in flight per thread (default; MON_ENC_UNROLL=1|2|3|4 selects another instantiation for A/B)

__device__ __forceinline__ void bulk_load_table(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    // one thread: arm the barrier with the byte count, then issue the TMA bulk copies (global -> shared)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc05::smem_u32(bar)), "r"(bytes) : "memory");
    const char* src = static_cast<const char*>(gsrc);
    for (uint32_t off = 0; off < bytes; off += ENC_BULK_CHUNK) {
        const uint32_t n = min(ENC_BULK_CHUNK, bytes - off);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst + off),
                     "l"(src + off), "r"(n), "r"(tc05::smem_u32(bar))
                     : "memory");
    }
}

// LIST != nullptr (opt-in occupancy mode): the kernel walks the compacted list of occupied samples, p is a position in that list
template <bool HASHED>
__device__ __forceinline__ __half enc_one_pow2(const float* __restrict__ pts, uint32_t p, float scale, uint32_t bmask, uint32_t my, uint32_t mz,
                                               const unsigned char* __restrict__ table) {
    const float u0 = __ldg(pts + (size_t)p * 3), u1 = __ldg(pts + (size_t)p * 3 + 1), u2 = __ldg(pts + (size_t)p * 3 + 2);
    float fr[3]; uint32_t cell[3];
    mon_pos_fract(u0, scale, fr[0], cell[0]);
    mon_pos_fract(u1, scale, fr[1], cell[1]);
    mon_pos_fract(u2, scale, fr[2], cell[2]);
    const float g0 = __fsub_rn(1.0f, fr[0]), g1 = __fsub_rn(1.0f, fr[1]), g2 = __fsub_rn(1.0f, fr[2]);
    // (1*fx)*fy shared by the two z corners; same multiplication order as the reference
    const float wxy[4] = {__fmul_rn(g0, g1), __fmul_rn(fr[0], g1), __fmul_rn(g0, fr[1]), __fmul_rn(fr[0], fr[1])};
    // per-axis contributions to the byte offset
    const uint32_t ax[2] = {cell[0] << 1, (cell[0] + 1u) << 1};
    const uint32_t ay[2] = {(cell[1] * my) << 1, ((cell[1] + 1u) * my) << 1};
    const uint32_t az[2] = {(cell[2] * mz) << 1, ((cell[2] + 1u) * mz) << 1};
    __half acc = __float2half_rn(0.0f);
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        const float wgt = __fmul_rn(wxy[k & 3], (k & 4) ? fr[2] : g2);
        const uint32_t off = (HASHED ? (ax[k & 1] ^ ay[(k >> 1) & 1] ^ az[k >> 2]) : (ax[k & 1] + ay[(k >> 1) & 1] + az[k >> 2])) & bmask;
        const float wh = __half2float(__float2half_rn(wgt));
        // generic load from a pointer the compiler knows to be shared: LDS.U16 with the base folded in
        const unsigned short tv = *reinterpret_cast<const unsigned short*>(table + off);
        acc = __float2half_rn(__fmaf_rn(wh, __half2float(__ushort_as_half(tv)), __half2float(acc)));
    }
    return acc;
}

// the inner loop for power-of-two tables (every table of the supported configurations): index arithmetic directly in
// byte offsets of the shared-memory slice, ((a ^ b ^ c) & (size-1)) * 2 == (2a ^ 2b ^ 2c) & (2*size-2); HASHED selects
// the coherent-prime hash or the dense (wrapping) linear index at compile time.  U points per trip: the 8-corner
// fp16 rounding chain of a point is serial (FFMA -> F2F -> HADD2 per corner), further independent chains hide it.
template <bool HASHED, int U>
__device__ __forceinline__ void enc_points_pow2(const float* __restrict__ pts, __half* __restrict__ out, uint32_t p_first, uint32_t p_end,
                                                float scale, uint32_t size, uint32_t res, const unsigned char* __restrict__ table,
                                                const uint32_t* __restrict__ list) {
    const uint32_t bmask = 2u * size - 2u;
    const uint32_t my = HASHED ? 2654435761u : res, mz = HASHED ? 805459861u : res * res;
    uint32_t p = p_first;
    if (list) {
        // occupancy mode: position p of the compacted list -> sample index; one sample per trip (the list is short)
        for (; p < p_end; p += ENC_THREADS) {
            const uint32_t pt = __ldg(list + p);
            out[(size_t)pt * 2] = enc_one_pow2<HASHED>(pts, pt, scale, bmask, my, mz, table);
        }
        return;
    }
    if (U > 1) {
        for (; p + (U - 1) * ENC_THREADS < p_end; p += U * ENC_THREADS) {
            __half a[U];
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = enc_one_pow2<HASHED>(pts, p + u * ENC_THREADS, scale, bmask, my, mz, table);
#pragma unroll
            for (int u = 0; u < U; ++u) out[(size_t)(p + u * ENC_THREADS) * 2] = a[u];
        }
    }
    for (; p < p_end; p += ENC_THREADS) out[(size_t)p * 2] = enc_one_pow2<HASHED>(pts, p, scale, bmask, my, mz, table);
}

// Cost model of the forward kernel for the work split, in 1/64 of a hashed-level point: a point costs W_class, and
// starting a job costs L_class (staging its table slice: ~1.5 us for 128 KB, i.e. ~2300 hashed points of one CTA), so a
// CTA whose piece spans a job boundary — and therefore stages two slices — gets that much less point work.
#define ENC_W_SMALL 56u    // dense, <= 4096 entries
#define ENC_W_LARGE 75u    // dense, larger
#define ENC_W_HASH 64u
#define ENC_L_SMALL 9000ull
#define ENC_L_LARGE 87000ull
#define ENC_L_HASH 148000ull
struct EncPos { uint32_t job, p; };
__host__ __device__ __forceinline__ uint64_t enc_cost_total(uint32_t n, uint32_t jb, uint32_t jl, uint32_t jh, uint32_t je) {
    return (uint64_t)(jl - jb) * ((uint64_t)n * ENC_W_SMALL + ENC_L_SMALL) + (uint64_t)(jh - jl) * ((uint64_t)n * ENC_W_LARGE + ENC_L_LARGE) +
           (uint64_t)(je - jh) * ((uint64_t)n * ENC_W_HASH + ENC_L_HASH);
}
// position in the flattened [job][point] space at cumulative cost c (monotone; the same expression gives one CTA's end and
// the next one's begin, so the pieces tile the space exactly)
__host__ __device__ __forceinline__ EncPos enc_cost_to_pos(uint64_t c, uint32_t n, uint32_t jb, uint32_t jl, uint32_t jh) {
    const uint64_t js = (uint64_t)n * ENC_W_SMALL + ENC_L_SMALL, jg = (uint64_t)n * ENC_W_LARGE + ENC_L_LARGE, jhh = (uint64_t)n * ENC_W_HASH + ENC_L_HASH;
    const uint64_t c0 = js * (jl - jb), c1 = jg * (jh - jl);
    uint64_t per, load, w; uint32_t base;
    if (c < c0) { per = js; load = ENC_L_SMALL; w = ENC_W_SMALL; base = jb; }
    else if (c < c0 + c1) { c -= c0; per = jg; load = ENC_L_LARGE; w = ENC_W_LARGE; base = jl; }
    else { c -= c0 + c1; per = jhh; load = ENC_L_HASH; w = ENC_W_HASH; base = jh; }
    const uint64_t j = c / per, rem = c - j * per;
    EncPos r;
    r.job = base + (uint32_t)j;
    r.p = rem <= load ? 0u : (uint32_t)((rem - load) / w);     // < n because rem < per
    return r;
}

// planar: per level [feature 0 table | feature 1 table], each size[l] fp16 (the level starts at 2*offset[l] halves)
template <int U>
__global__ void __launch_bounds__(ENC_THREADS, 1)
k_encode_forward(MonGrid g, uint32_t n_points_all, const float* __restrict__ pts, const __half* __restrict__ planar,
                 __half* __restrict__ enc_soa, const MonCtrl* __restrict__ ctrl, uint32_t job_begin, uint32_t job_end,
                 uint32_t job_large, uint32_t job_hashed, const uint32_t* __restrict__ occ_list, const uint32_t* __restrict__ occ_count) {
    extern __shared__ __align__(128) unsigned char enc_smem[];
    __shared__ __align__(8) uint64_t bar;
    const __half* table = reinterpret_cast<const __half*>(enc_smem);
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { tc05::mbar_init(&bar, 1); tc05::mbar_fence_init(); }
    __syncthreads();
    mon_pdl_wait();       // the optimizer sweep (planar weights) has completed
    mon_pdl_trigger();
    if (ctrl && ctrl->skip) return;
    MON_TL(MON_TL_E + ((job_begin >> 3) & 3u), ctrl ? ctrl->iter - 1 : 0u);
    // opt-in occupancy mode: the point space of the jobs is the compacted list of occupied samples (length on the device); the
    // encodings of the other samples keep stale values the fused MLP kernel masks.  n_points_all stays the stride of enc_soa.
    const uint32_t n_points = occ_list ? *occ_count : n_points_all;
    if (n_points == 0u) return;

    // jobs [job_begin, job_end) of the 2 * n_levels (level, feature) jobs (callers may launch sub-ranges).
    // The flattened [job][point] space is cut into one contiguous piece per CTA, equal in COST: a point of a dense level
    // is not as expensive as one of a hashed level, because the kernel is bound by shared-memory wavefronts and the
    // bank-conflict degree of a warp's 32 gathers differs (measured per-CTA times and a bank simulation of the batch's
    // rays agree: hashed ~3.5-way, 17 us; dense 16^3 table 2.6-way, 15 us; dense 32^3 table 4.2-way, 20 us —
    // profiles/r1p_encode_balance.txt).  Three cost classes in level order — small dense [job_begin, job_large), large
    // dense [job_large, job_hashed), hashed [job_hashed, job_end) — make the cost -> (job, point) map closed-form
    // (a per-job search through the dynamically indexed kernel parameters costs microseconds of dependent LDCU latency).
    const EncPos pos_b = enc_cost_to_pos(enc_cost_total(n_points, job_begin, job_large, job_hashed, job_end) * blockIdx.x / gridDim.x,
                                         n_points, job_begin, job_large, job_hashed);
    const EncPos pos_e = enc_cost_to_pos(enc_cost_total(n_points, job_begin, job_large, job_hashed, job_end) * (blockIdx.x + 1) / gridDim.x,
                                         n_points, job_begin, job_large, job_hashed);
    uint32_t phase = 0;
    for (uint32_t job = pos_b.job; job < job_end && (job < pos_e.job || (job == pos_e.job && pos_e.p > 0)); ++job) {
        const uint32_t p0 = job == pos_b.job ? pos_b.p : 0u;
        const uint32_t p1 = job == pos_e.job ? pos_e.p : n_points;
        if (p1 <= p0) continue;
        const uint32_t l = job >> 1, f = job & 1;
        const uint32_t size = g.size[l];
        __syncthreads();                                   // the previous slice is no longer read by anyone
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // prior generic reads/writes of the buffer vs. the async-proxy write
            bulk_load_table(tc05::smem_u32(enc_smem), planar + (size_t)g.offset[l] * 2 + (size_t)f * size, size * 2, &bar);
        }
        tc05::mbar_wait(&bar, phase);
        // whole-range launches only (the chain graph): E1 = [first, last] CTA whose first slice is resident
        if (phase == 0 && job_begin == 0 && job_end == 2 * g.n_levels) { MON_TL_MARK(MON_TL_E + 1, ctrl ? ctrl->iter - 1 : 0u); ENC_CTA_STAMP(0); }
        phase ^= 1u;

        const float scale = g.scale[l];
        const bool hashed = g.hashed[l] != 0;
        const uint32_t res = g.res[l];
        const bool pow2 = (size & (size - 1)) == 0;
        __half* out = enc_soa + (size_t)l * n_points_all * 2 + f;   // level-major pairs: enc[level][point][feature]
        if (pow2) {
            if (hashed) enc_points_pow2<true, U>(pts, out, p0 + tid, p1, scale, size, res, enc_smem, occ_list);
            else enc_points_pow2<false, U>(pts, out, p0 + tid, p1, scale, size, res, enc_smem, occ_list);
        } else {
            for (uint32_t q = p0 + tid; q < p1; q += ENC_THREADS) {
                const uint32_t p = occ_list ? __ldg(occ_list + q) : q;
                const float u0 = __ldg(pts + (size_t)p * 3), u1 = __ldg(pts + (size_t)p * 3 + 1), u2 = __ldg(pts + (size_t)p * 3 + 2);
                float fr[3]; uint32_t cell[3];
                mon_pos_fract(u0, scale, fr[0], cell[0]);
                mon_pos_fract(u1, scale, fr[1], cell[1]);
                mon_pos_fract(u2, scale, fr[2], cell[2]);
                const float g0 = __fsub_rn(1.0f, fr[0]), g1 = __fsub_rn(1.0f, fr[1]), g2 = __fsub_rn(1.0f, fr[2]);
                __half acc = __float2half_rn(0.0f);
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    float wgt = (k & 1) ? fr[0] : g0;            // == 1.0f * that factor, bit for bit
                    wgt = __fmul_rn(wgt, (k & 2) ? fr[1] : g1);
                    wgt = __fmul_rn(wgt, (k & 4) ? fr[2] : g2);
                    const uint32_t idx = mon_grid_index(hashed, size, res, cell[0] + (k & 1), cell[1] + ((k >> 1) & 1), cell[2] + ((k >> 2) & 1));
                    const float wh = __half2float(__float2half_rn(wgt));
                    acc = __float2half_rn(__fmaf_rn(wh, __half2float(table[idx]), __half2float(acc)));
                }
                out[(size_t)p * 2] = acc;
            }
        }
    }
    // E2 = [first, last] CTA done (load imbalance of the flattened [job][point] split)
    if (job_begin == 0 && job_end == 2 * g.n_levels) { MON_TL_MARK(MON_TL_E + 2, ctrl ? ctrl->iter - 1 : 0u); ENC_CTA_STAMP(1); }
}

template <int U>
static cudaError_t enc_launch(const MonGrid& g, uint32_t n_points, const float* pts, const __half* planar, __half* enc_soa, const MonCtrl* ctrl,
                              uint32_t ctas, cudaStream_t st, uint32_t level_begin, uint32_t level_end, const MonLaunchOpt& lo,
                              const uint32_t* occ_list, const uint32_t* occ_count) {
    static std::atomic<uint64_t> prepared{0};
    const cudaError_t prep = mon_once_per_device(prepared, [] {
        cudaError_t e = cudaFuncSetAttribute(k_encode_forward<U>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_encode_forward<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_TABLE_BYTES);
    });
    if (prep != cudaSuccess) return prep;
    // cost classes in level order (res grows with the level: small dense, then large dense, then hashed)
    uint32_t l_large = level_begin, l_hashed = level_begin;
    while (l_large < level_end && !g.hashed[l_large] && g.size[l_large] <= 4096u) ++l_large;
    l_hashed = l_large;
    while (l_hashed < level_end && !g.hashed[l_hashed]) ++l_hashed;
    return mon_launch_chain(MON_PDL_ENCODE, lo, k_encode_forward<U>, dim3(ctas), dim3(ENC_THREADS), ENC_TABLE_BYTES, st, g, n_points, pts, planar, enc_soa, ctrl,
                            2 * level_begin, 2 * level_end, 2 * l_large, 2 * l_hashed, occ_list, occ_count);
}

cudaError_t mon_launch_encode_forward(const MonGrid& g, uint32_t n_points, const float* pts, const __half* planar, __half* enc_soa,
                                      const MonCtrl* ctrl, uint32_t sm_count, cudaStream_t st, uint32_t level_begin, uint32_t level_end,
                                      const MonLaunchOpt& lo, const uint32_t* occ_list, const uint32_t* occ_count) {
    if (level_end > g.n_levels) level_end = g.n_levels;
    if (n_points == 0 || level_begin >= level_end) return cudaSuccess;
    // every CTA loads up to two 128 KB slices: do not spread tiny batches over the whole chip
    const uint64_t total = (uint64_t)(2 * (level_end - level_begin)) * n_points;
    const uint64_t want = (total + 16383) / 16384;
    uint32_t ctas = want < (uint64_t)sm_count ? (uint32_t)want : sm_count;
    if (ctas == 0) ctas = 1;
    static const int unroll = [] { const char* e = getenv("MON_ENC_UNROLL"); const int u = e ? atoi(e) : ENC_UNROLL; return (u >= 1 && u <= 4) ? u : ENC_UNROLL; }();
    switch (unroll) {
        case 1: return enc_launch<1>(g, n_points, pts, planar, enc_soa, ctrl, ctas, st, level_begin, level_end, lo, occ_list, occ_count);
        case 3: return enc_launch<3>(g, n_points, pts, planar, enc_soa, ctrl, ctas, st, level_begin, level_end, lo, occ_list, occ_count);
        case 4: return enc_launch<4>(g, n_points, pts, planar, enc_soa, ctrl, ctas, st, level_begin, level_end, lo, occ_list, occ_count);
        default: return enc_launch<2>(g, n_points, pts, planar, enc_soa, ctrl, ctas, st, level_begin, level_end, lo, occ_list, occ_count);
    }
}

// host mirror of the kernel's work split (the same two functions): piece b of n_ctas as (first job, first point, end job,
// end point); used by the CPU test that checks the pieces tile the [job][point] space exactly for arbitrary sizes
void mon_encode_pieces_host(const MonGrid& g, uint32_t n_points, uint32_t n_ctas, uint32_t level_begin, uint32_t level_end, uint32_t* out4) {
    uint32_t l_large = level_begin, l_hashed = level_begin;
    while (l_large < level_end && !g.hashed[l_large] && g.size[l_large] <= 4096u) ++l_large;
    l_hashed = l_large;
    while (l_hashed < level_end && !g.hashed[l_hashed]) ++l_hashed;
    const uint32_t jb = 2 * level_begin, jl = 2 * l_large, jh = 2 * l_hashed, je = 2 * level_end;
    const uint64_t total = enc_cost_total(n_points, jb, jl, jh, je);
    for (uint32_t b = 0; b < n_ctas; ++b) {
        const EncPos pb = enc_cost_to_pos(total * b / n_ctas, n_points, jb, jl, jh);
        const EncPos pe = enc_cost_to_pos(total * (b + 1) / n_ctas, n_points, jb, jl, jh);
        out4[4 * b + 0] = pb.job; out4[4 * b + 1] = pb.p; out4[4 * b + 2] = pe.job; out4[4 * b + 3] = pe.p;
    }
}

// interleaved fp16 weights [entry][2] -> planar per level [f0 table | f1 table] (initialisation / set_params).  The planar copy is
// THE fp16 working copy of the grid from then on: the optimizer sweep reads and updates it alone (the interleaved grid part of the
// parameter vector is not maintained; k_deplanarize rebuilds it for the state getter)
__global__ void k_planarize(MonGrid g, uint32_t n_entries, const __half* __restrict__ inter, __half* __restrict__ planar) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    uint32_t l = 0;
    while (l + 1 < g.n_levels && e >= g.offset[l + 1]) ++l;
    const uint32_t local = e - g.offset[l], size = g.size[l];
    const __half2 v = reinterpret_cast<const __half2*>(inter)[e];
    planar[(size_t)g.offset[l] * 2 + local] = __low2half(v);
    planar[(size_t)g.offset[l] * 2 + size + local] = __high2half(v);
}
__global__ void k_deplanarize(MonGrid g, uint32_t n_entries, const __half* __restrict__ planar, __half* __restrict__ inter) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    uint32_t l = 0;
    while (l + 1 < g.n_levels && e >= g.offset[l + 1]) ++l;
    const uint32_t local = e - g.offset[l], size = g.size[l];
    reinterpret_cast<__half2*>(inter)[e] = __halves2half2(planar[(size_t)g.offset[l] * 2 + local], planar[(size_t)g.offset[l] * 2 + size + local]);
}
void mon_launch_deplanarize(const MonGrid& g, const __half* planar, __half* inter, cudaStream_t st) {
    const uint32_t n = g.offset[g.n_levels];
    k_deplanarize<<<(n + 255) / 256, 256, 0, st>>>(g, n, planar, inter);
}
void mon_launch_planarize(const MonGrid& g, const __half* inter, __half* planar, cudaStream_t st) {
    const uint32_t n = g.offset[g.n_levels];
    k_planarize<<<(n + 255) / 256, 256, 0, st>>>(g, n, inter, planar);
}

// ---------------------------------------------------------------------------------------------- backward
#define SCT_THREADS 512
#define SCT_TILE 64

// Scatter of the compacted live samples.  A warp of the fused MLP kernel is one ray and most of its samples sit behind the
// early stop (T < 1e-4) with an all-zero gradient row: in steady state only ~1 sample in 10 carries gradient.  The fused MLP
// kernel hands over the live samples only — slot k: position pts_c[k] (16 bytes: x, y, z, unused), and per level l one word genc[l][k] with the level's
// two fp16 gradients — so every lane here has real work and the level (table base, scale, hash or dense) is uniform across
// a warp: a CTA takes SCT_TILE consecutive slots, warp w scatters level w for them.  CTAs beyond the live count exit at once.
__global__ void __launch_bounds__(SCT_THREADS, 3)
k_encode_backward(MonGrid g, uint32_t n_points, uint32_t resident_min_live, const uint32_t* __restrict__ live_cnt, const float* __restrict__ pts_c,
                  const uint32_t* __restrict__ genc, const MonCtrl* __restrict__ ctrl, __half* __restrict__ grid_grad) {
    __shared__ float s_u[3][SCT_TILE];               // [axis][slot in tile]
    mon_pdl_wait();
    mon_pdl_trigger();
    if (ctrl->skip) return;
    const uint32_t n_live = live_cnt[(ctrl->iter - 1) & 1u];
    if (n_live >= resident_min_live) return;           // grid-uniform: the shared-memory resident scatter takes this iteration (kernels_scatter_smem.cu)
    const uint32_t s0 = blockIdx.x * SCT_TILE;
    if (s0 >= n_live) return;                          // CTA-uniform
    MON_TL(MON_TL_S, ctrl->iter - 1);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_tile = min((uint32_t)SCT_TILE, n_live - s0);
    for (uint32_t i = tid; i < 3 * n_tile; i += SCT_THREADS) s_u[i % 3][i / 3] = __ldg(pts_c + (size_t)(s0 + i / 3) * 4 + i % 3);   // slots are (x, y, z, -)
    __syncthreads();
    for (uint32_t l = warp; l < g.n_levels; l += SCT_THREADS / 32) {
        const uint32_t* gl = genc + (size_t)l * n_points + s0;
        for (uint32_t r = lane; r < n_tile; r += 32) {
            const uint32_t gwj = __ldg(gl + r);
            if ((gwj & 0x7fff7fffu) == 0u) continue;   // adding +0 is an identity
            const float u[3] = {s_u[0][r], s_u[1][r], s_u[2][r]};
            scatter_level(g, l, gwj, u, grid_grad);
        }
    }
}

void mon_launch_encode_backward(const MonGrid& g, uint32_t n_points, uint32_t resident_min_live, const uint32_t* live_cnt, const float* pts_c,
                                const uint32_t* genc, const MonCtrl* ctrl, __half* grid_grad, cudaStream_t st, const MonLaunchOpt& lo) {
    if (n_points == 0) return;
    // iterations with >= resident_min_live live samples are scattered by k_scatter_resident: this grid then only needs the CTAs
    // that can have work
    const uint32_t max_live = resident_min_live < n_points ? resident_min_live : n_points;
    const uint32_t blocks = (max_live + SCT_TILE - 1) / SCT_TILE;
    if (blocks == 0) return;
    mon_launch_chain(MON_PDL_SCATTER, lo, k_encode_backward, dim3(blocks), dim3(SCT_THREADS), 0, st, g, n_points, resident_min_live, live_cnt, pts_c, genc, ctrl, grid_grad);
}
