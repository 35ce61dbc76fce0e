// kernels_optim.cu — parameter initialisation and the fused optimizer sweep.
//
// k_optimizer_sweep fuses, in one pass over the parameter vector, what the reference runs as
// three launches plus a memset per iteration:
//   cudaMemsetAsync(grid gradients)          TCNN encodings/grid.h:1132
//   adam_step<__half>                         TCNN optimizers/adam.h:48-118
//   ExponentialDecay learning-rate schedule   TCNN optimizers/exponential_decay.h:60-71  (scalar from the batch kernel)
//   ema_step_half_precision<__half>           TCNN optimizers/ema.h:62-76,102-136
//   SumLoss + host-side mean                  MON/Core/src/nerf_model.cu:1231-1253,1650-1658
// and, for the first n_mlp parameters, the deterministic reduction of the per-CTA MLP weight-
// gradient partials written by the fused MLP kernel.  Semantics kept exactly: per-parameter
// step counters, grid parameters whose gradient is exactly zero are skipped by Adam
// (adam.h:75-79) but still EMA-filtered, L2 regularisation only on MLP weights, fp32 master +
// fp16 working copy + fp16 EMA (the inference weights).  The gradient is consumed and zeroed in
// the same pass, so the next iteration's scatter starts from zero without a memset; touched grid
// weights are also written to the planar copy the hash-encode kernel stages into shared memory.
// A thread owns 4 consecutive parameters (= 2 table entries): 8-byte gradient / EMA words, 16-byte
// master / moment / step words.
#include <algorithm>

#include "mon_device.cuh"
#include "mon_kernels.h"
#include "optim_math.cuh"
#include "mon_timeline.cuh"
MON_TL_DEFINE(optim)

// ---- pcg32 on the device (TCNN dependencies/pcg32/pcg32.h:46-170), for A12 grid init
struct DevPcg32 {
    uint64_t state, inc;
    __device__ uint32_t next_uint() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    __device__ float next_float() { return __uint_as_float((next_uint() >> 9) | 0x3f800000u) - 1.0f; }
    __device__ void advance(uint64_t delta) {
        uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
};

// generate_random_kernel<float, pcg32, 4> + uniform transform (random.h:66-92): thread i copies
// the generator, advances 4*i, writes out[i + n_threads*j] = val*(upper-lower)+lower (one FMA).
__global__ void k_init_grid(uint64_t state, uint64_t inc, uint32_t n, uint32_t n_threads, float lower, float range, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_threads) return;
    DevPcg32 r{state, inc};
    r.advance((uint64_t)i * 4);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t idx = i + n_threads * j;
        if (idx >= n) return;
        out[idx] = __fmaf_rn(r.next_float(), range, lower);
    }
}

__global__ void k_cast_params(uint32_t n, const float* __restrict__ pf, __half* __restrict__ ph) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ph[i] = __float2half_rn(pf[i]);
}

#define OPT_THREADS 256
#define OPT_PER_THREAD 4
#define OPT_CTAS_PER_SM 4

__device__ __forceinline__ uint32_t level_of_entry(const MonGrid& g, uint32_t e) {
    if (g.log2_cap && g.first_full < g.n_levels && e >= g.offset[g.first_full]) return g.first_full + ((e - g.offset[g.first_full]) >> g.log2_cap);
    uint32_t l = 0;
    while (l + 1 < g.n_levels && e >= g.offset[l + 1]) ++l;
    return l;
}

// what the sweep fetches per grid quad (4 consecutive parameters = table entries e, e + 1) before it knows whether Adam runs
struct GridQuad {
    float g[4];
    uint2 wraw, eraw;          // fp16 weights and EMA weights
    __half* planar_f0;
    uint32_t planar_stride;
};

__device__ __forceinline__ GridQuad load_grid_quad(uint32_t i4, const MonOpt& o, const MonGrid& grid, __half* __restrict__ gh, __half* __restrict__ gcls,
                                                    bool resident, const __half* __restrict__ ema, __half* __restrict__ planar) {
    GridQuad q;
    q.eraw = *reinterpret_cast<const uint2*>(ema + i4);
    // planar copy read by the hash-encode kernel — the one fp16 working copy of the grid: level l holds [feature 0 | feature 1];
    // entries e, e+1 are adjacent in both feature arrays (level sizes are multiples of 8, so a pair never straddles a level)
    const uint32_t e = (i4 - o.n_mlp) >> 1;
    const uint32_t l = level_of_entry(grid, e);
    const uint32_t e_local = e - grid.offset[l];
    q.planar_f0 = planar + (size_t)grid.offset[l] * 2 + e_local;
    q.planar_stride = grid.size[l];
    {
        const uint32_t f0 = *reinterpret_cast<const uint32_t*>(q.planar_f0), f1 = *reinterpret_cast<const uint32_t*>(q.planar_f0 + q.planar_stride);
        q.wraw = make_uint2(__byte_perm(f0, f1, 0x5410), __byte_perm(f0, f1, 0x7632));     // (e.f0, e.f1), (e+1.f0, e+1.f1)
    }
    if (!resident) {
        // entry-ordered table filled by the global f16x2 reductions
        uint2* gw = reinterpret_cast<uint2*>(gh + i4);
        const uint2 raw = *gw;
        const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
        q.g[0] = __low2float(a); q.g[1] = __high2float(a); q.g[2] = __low2float(b); q.g[3] = __high2float(b);
        if ((raw.x | raw.y) & 0x7fff7fffu) *gw = make_uint2(0u, 0u);   // consumed: the next scatter starts from zero
    } else {
        // class-planar table filled by the shared-memory resident scatter — per level [parity 0: f0 | f1][parity 1: f0 | f1], each
        // size/2 fp16 (kernels_scatter_smem.cu): entries e_local (even: class 0) and e_local + 1 (class 1) share the slot e_local / 2
        const uint32_t half_n = q.planar_stride >> 1;
        __half* c0 = gcls + (size_t)grid.offset[l] * 2 + (e_local >> 1);
        __half hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) hv[k] = c0[(size_t)k * half_n];      // parameter k: entry parity k >> 1, feature k & 1 -> class array k
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            q.g[k] = __half2float(hv[k]);
            if (__half_as_ushort(hv[k]) & 0x7fffu) c0[(size_t)k * half_n] = __ushort_as_half((unsigned short)0);
        }
    }
    return q;
}

// Measured and dropped (profiles/r9e_timeline_optspec.txt): two quads per trip with every load of both — the state Adam needs
// included — issued before the first use, 3 CTAs per SM at 80 registers: 15.3 instead of 13.4 us.  The sweep is bound by the bytes
// it moves through L2 (46 B per touched parameter: 6.3 TB/s for a fresh object), not by the latency of its dependent loads.
__global__ void __launch_bounds__(OPT_THREADS, OPT_CTAS_PER_SM)
k_optimizer_sweep(MonOpt o, MonCtrl* __restrict__ ctrl, float* __restrict__ pf, __half* __restrict__ ph,
                  __half* __restrict__ gh, const float* __restrict__ mlp_partials, float* __restrict__ m,
                  float* __restrict__ v, uint16_t* __restrict__ ps, __half* __restrict__ ema,
                  const float* __restrict__ loss, uint32_t R, MonGrid grid, __half* __restrict__ planar,
                  uint32_t n_mlp_ctas, uint32_t do_loss, uint32_t grid_i4_begin, uint32_t grid_i4_end,
                  __half* __restrict__ gcls, const uint32_t* __restrict__ live_cnt, uint32_t resident_min_live) {
    mon_pdl_wait();       // the gradient scatter has completed
    mon_pdl_trigger();
    if (ctrl->skip) return;
    MON_TL(n_mlp_ctas ? MON_TL_OMLP : MON_TL_O + (level_of_entry(grid, (grid_i4_begin - o.n_mlp) >> 1) >> 2 & 3u), ctrl->iter - 1);
    if (do_loss && blockIdx.x == gridDim.x - 1) {
        // logged loss in the sweep's last CTA: fixed summation order -> reproducible
        __shared__ float s_loss[OPT_THREADS];
        float a = 0.0f;
        for (uint32_t i = threadIdx.x; i < R; i += OPT_THREADS) a += loss[i];
        s_loss[threadIdx.x] = a;
        __syncthreads();
        for (int step = OPT_THREADS / 2; step > 0; step >>= 1) {
            if ((int)threadIdx.x < step) s_loss[threadIdx.x] += s_loss[threadIdx.x + step];
            __syncthreads();
        }
        if (threadIdx.x == 0) ctrl->loss_mean = s_loss[0] / (float)R;
    }
    const float lr_base = ctrl->lr_base, old_db = ctrl->ema_old, new_db = ctrl->ema_new;
    const OptimPtrs ptrs = {pf, ph, m, v, ps, ema};
    // CTA layout: the first n_mlp_ctas (n_mlp/32, or 0 in a grid-only launch) CTAs own the MLP weights (one WARP per 4
    // parameters: the lanes split the per-CTA gradient partials of the fused MLP kernel); the other CTAs walk the quads of
    // [grid_i4_begin, grid_i4_end) with a grid stride (one THREAD per 4 parameters per trip).  The launch sizes them so that all
    // CTAs are resident at once and every thread makes the same number of trips (no partial last wave); the loads of the next
    // quad are issued before the current one is processed.  n_params and n_mlp are multiples of 32 resp. 4.
    if (blockIdx.x < n_mlp_ctas) {
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const uint32_t i4 = (blockIdx.x * (OPT_THREADS / 32) + warp) * OPT_PER_THREAD;
        // fixed-order reduction -> bitwise reproducible MLP gradient: lane l sums partial rows l, l+32, ... in
        // ascending order, then a butterfly over the lanes
        float4 s4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (uint32_t c = lane; c < o.n_partials; c += 32) {
            const float4 p = *reinterpret_cast<const float4*>(mlp_partials + (size_t)c * o.n_mlp + i4);
            s4.x = __fadd_rn(s4.x, p.x); s4.y = __fadd_rn(s4.y, p.y); s4.z = __fadd_rn(s4.z, p.z); s4.w = __fadd_rn(s4.w, p.w);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s4.x = __fadd_rn(s4.x, __shfl_xor_sync(0xffffffffu, s4.x, off));
            s4.y = __fadd_rn(s4.y, __shfl_xor_sync(0xffffffffu, s4.y, off));
            s4.z = __fadd_rn(s4.z, __shfl_xor_sync(0xffffffffu, s4.z, off));
            s4.w = __fadd_rn(s4.w, __shfl_xor_sync(0xffffffffu, s4.w, off));
        }
        if (lane != 0) return;
        const uint2 wraw = *reinterpret_cast<const uint2*>(ph + i4);
        const uint2 eraw = *reinterpret_cast<const uint2*>(ema + i4);
        // the reference stores weight gradients in fp16 (loss-scaled); keep that rounding point
        const __half2 a = __floats2half2_rn(s4.x, s4.y), b = __floats2half2_rn(s4.z, s4.w);
        *reinterpret_cast<uint2*>(gh + i4) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));   // kept for inspection
        float g[4] = {__low2float(a), __high2float(a), __low2float(b), __high2float(b)};
        optim_quad(o, lr_base, old_db, new_db, true, i4, g, wraw, eraw, ptrs, nullptr, 0);
        return;
    }
    // which scatter path produced this iteration's gradient (grid-uniform, decided by the live-sample count)
    const bool resident = gcls != nullptr && live_cnt[(ctrl->iter - 1) & 1u] >= resident_min_live;
    const uint32_t stride = (gridDim.x - n_mlp_ctas) * OPT_THREADS * OPT_PER_THREAD;
    uint32_t i4 = grid_i4_begin + ((blockIdx.x - n_mlp_ctas) * OPT_THREADS + threadIdx.x) * OPT_PER_THREAD;
    if (i4 >= grid_i4_end) return;
    GridQuad cur = load_grid_quad(i4, o, grid, gh, gcls, resident, ema, planar);
    while (true) {
        const uint32_t i4n = i4 + stride;
        const bool more = i4n < grid_i4_end;
        GridQuad nxt;
        if (more) nxt = load_grid_quad(i4n, o, grid, gh, gcls, resident, ema, planar);
        optim_quad(o, lr_base, old_db, new_db, false, i4, cur.g, cur.wraw, cur.eraw, ptrs, cur.planar_f0, cur.planar_stride);
        if (!more) break;
        cur = nxt; i4 = i4n;
    }
}

// used only by tests: snapshot of the loss-scaled gradient before the sweep consumes it
__global__ void k_snapshot_grad(uint32_t n, uint32_t n_mlp, uint32_t n_partials, const __half* __restrict__ gh,
                                const float* __restrict__ mlp_partials, float* __restrict__ out, MonGrid grid, const __half* __restrict__ gcls) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i < n_mlp) {
        float s = 0.0f;
        for (uint32_t c = 0; c < n_partials; ++c) s = __fadd_rn(s, mlp_partials[(size_t)c * n_mlp + i]);
        out[i] = __half2float(__float2half_rn(s));
    } else {
        // one of the two gradient tables is all zero (the iteration's scatter kernel filled the other one)
        float gv = __half2float(gh[i]);
        if (gcls) {
            const uint32_t e = (i - n_mlp) >> 1, f = (i - n_mlp) & 1u;
            const uint32_t l = level_of_entry(grid, e);
            const uint32_t e_local = e - grid.offset[l], half_n = grid.size[l] >> 1;
            gv += __half2float(gcls[(size_t)grid.offset[l] * 2 + (size_t)((e_local & 1u) * 2u + f) * half_n + (e_local >> 1)]);
        }
        out[i] = gv;
    }
}

void mon_launch_init_grid(uint64_t state, uint64_t inc, uint32_t n, float* out, cudaStream_t st) {
    const uint32_t n_needed = (n + 3) / 4;
    const uint32_t blocks = (n_needed + 127) / 128;
    const uint32_t n_threads = blocks * 128;
    const float lower = -1e-4f, upper = 1e-4f;
    k_init_grid<<<blocks, 128, 0, st>>>(state, inc, n, n_threads, lower, upper - lower, out);
}
void mon_launch_cast_params(uint32_t n, const float* pf, __half* ph, cudaStream_t st) {
    k_cast_params<<<(n + 255) / 256, 256, 0, st>>>(n, pf, ph);
}
void mon_launch_optimizer(const MonOpt& o, MonCtrl* ctrl, float* pf, __half* ph, __half* gh, const float* partials,
                          float* m, float* v, uint16_t* ps, __half* ema, const float* loss, uint32_t R, const MonGrid& grid,
                          __half* planar, cudaStream_t st, int part, uint32_t level_begin, uint32_t level_end, const MonLaunchOpt& lo,
                          __half* gcls, const uint32_t* live_cnt, uint32_t resident_min_live, uint32_t sm_count) {
    // part: MON_OPT_ALL everything (MLP weights + loss + whole grid); MON_OPT_MLP the MLP weights and the logged loss only;
    // MON_OPT_GRID the grid parameters of levels [level_begin, level_end) only; MON_OPT_MLP_GRID both of these
    if (level_end > grid.n_levels) level_end = grid.n_levels;
    const bool with_mlp = part != MON_OPT_GRID, with_grid = part != MON_OPT_MLP;
    const uint32_t n_mlp_ctas = with_mlp ? o.n_mlp / (OPT_PER_THREAD * (OPT_THREADS / 32)) : 0u;
    uint32_t i4_begin = 0, i4_end = 0;
    if (with_grid) {
        i4_begin = o.n_mlp + 2u * grid.offset[part == MON_OPT_ALL ? 0u : level_begin];
        i4_end = part == MON_OPT_ALL ? o.n_params : o.n_mlp + 2u * grid.offset[level_end];
    }
    const uint32_t grid_quads = (i4_end - i4_begin + OPT_PER_THREAD - 1) / OPT_PER_THREAD;
    // grid part: all CTAs resident at once (OPT_CTAS_PER_SM per SM beside the MLP CTAs) and the same number of trips for every thread
    uint32_t n_grid_ctas = 1u;      // a grid-less launch still needs the (otherwise idle) last CTA that reduces the logged loss
    const uint32_t per_sm = (uint32_t)OPT_CTAS_PER_SM;
    if (with_grid && grid_quads) {
        const uint32_t slots = std::max(1u, per_sm * sm_count > n_mlp_ctas ? per_sm * sm_count - n_mlp_ctas : 1u);
        const uint32_t trips = (grid_quads + slots * OPT_THREADS - 1) / (slots * OPT_THREADS);
        n_grid_ctas = (grid_quads + trips * OPT_THREADS - 1) / (trips * OPT_THREADS);
    }
    if (n_mlp_ctas + n_grid_ctas == 0) return;
    mon_launch_chain(MON_PDL_OPTIM, lo, k_optimizer_sweep, dim3(n_mlp_ctas + n_grid_ctas), dim3(OPT_THREADS), 0, st, o, ctrl, pf, ph, gh, partials, m,
                     v, ps, ema, loss, R, grid, planar, n_mlp_ctas, with_mlp ? 1u : 0u, i4_begin, i4_end, gcls, live_cnt, resident_min_live);
}
void mon_launch_snapshot_grad(uint32_t n, uint32_t n_mlp, uint32_t n_partials, const __half* gh, const float* partials, float* out, cudaStream_t st,
                              const MonGrid& grid, const __half* gcls) {
    k_snapshot_grad<<<(n + 255) / 256, 256, 0, st>>>(n, n_mlp, n_partials, gh, partials, out, grid, gcls);
}
