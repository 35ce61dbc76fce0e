// kernels_mlp_tc.cu — the fused tiny-MLP kernels on Blackwell tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One persistent CTA (128 threads = 4 warps) walks tiles of 128 sample points = 4 rays x 32 samples.  TMEM lane i
// of the accumulator is sample i of the tile, and tcgen05.ld hands TMEM lane (32*warp + lane) to thread `lane` of
// warp `warp`: a warp owns exactly one ray and each lane one of its 32 samples — the same mapping the
// warp-parallel volume renderer (render_math.cuh) wants, so the network output never leaves the SM between the
// last GEMM and compositing.  Per tile, for the training kernel:
//
//   enc tile (global, level-major fp16 pairs) --> smem [128 x 32]  K-major SWIZZLE_64B
//   MMA1  D[128x64]   = enc  . W_in^T            (K = 32)      epilogue: ReLU, fp16  --> hid  smem SW128
//   MMA2  O[128x16]   = hid  . W_out^T           (K = 64)      epilogue: sigmoid/exp, warp-scan compositing,
//                                                              loss, dL/dout fp16    --> dout smem (core layout)
//   MMA3  D[128x64]   = dout . W_out             (K = 16)      epilogue: ReLU mask, fp16 --> dhid smem SW128
//   MMA4  E[128x32]   = dhid . W_in              (K = 64)      epilogue: fp16 --> live rows, compacted, level-major (global)
//   MMA5  G1[128x16] += [hid|dhid]^T . dout      (K = 128 points; rows 0-63   = dW_out^T)
//   MMA6  G2[128x32] += [hid|dhid]^T . enc       (K = 128 points; rows 64-127 = dW_in)
//
// With n_hidden_layers == 2 ("2x64") the 64x64 hidden GEMM H2 = relu(H1 . W_h^T), its dgrad dH1 = (dH2 . W_h) * relu'
// and its weight gradient G3[128x64] += [H2|dH2]^T . H1 (rows 64-127 = dW_h) are inserted; dW_in then comes from
// [H1|dH1]^T . enc.
//
// MMA5/6 read the activation tiles that MMA2/MMA4 consumed K-major through MN-major descriptors (see tc05.cuh), so
// the weight gradients need no transpose and accumulate in TMEM across ALL tiles of the CTA; they are read out
// once at the end into the per-CTA partial row that the optimizer kernel sums in a fixed order.
// Replaces kernel_mlp_fused, VolumeRender, VolumeRenderGradient_No_Compacted, kernel_mlp_fused_backward and the
// three CUTLASS GEMMs of FullyFusedMLP::backward_impl (TCNN/src/fully_fused_mlp.cu:150-259,499-557,736-836;
// MON/Core/src/nerf_model.cu:735-954) and removes their 16.8 + 4.2 + 4.2 + 16.8 MB of activation round trips.
// fp16 operands, fp32 accumulation (the reference accumulates in fp16: tolerances, not bit-exactness).
#include "mon_kernels.h"
#include "render_math.cuh"
#include "tc05.cuh"
#include "scatter_global.cuh"
#include "mon_timeline.cuh"
MON_TL_DEFINE(mlp)

using namespace tc05;

// optional phase-timing instrumentation (build with MON_EXTRA_NVCC_FLAGS=-DMON_TC_STAMPS; tools/tc_phase_times.py)
#ifdef MON_TC_STAMPS
__device__ unsigned long long mon_tc_stamps[64];
#define TC_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) mon_tc_stamps[(i)] = clock64(); } while (0)
extern "C" int mon_debug_tc_stamps(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, mon_tc_stamps, sizeof(mon_tc_stamps)); }
__device__ unsigned long long mon_tc_cta[1024 * 3];   // per CTA: globaltimer at entry, at exit, SM id
extern "C" int mon_debug_tc_ctas(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, mon_tc_cta, sizeof(mon_tc_cta)); }
__device__ __forceinline__ unsigned long long tc_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned tc_smid() { unsigned s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }
#define TC_CTA_MARK(slot) do { if (threadIdx.x == 0 && blockIdx.x < 1024) { mon_tc_cta[blockIdx.x * 3 + (slot)] = tc_gtime(); mon_tc_cta[blockIdx.x * 3 + 2] = tc_smid(); } } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#define TC_CTA_MARK(slot) do { } while (0)
#endif

#define TC_THREADS 128
// TMEM column map.  The per-tile accumulators are never live at the same time (each is drained by its epilogue
// before the next MMA is issued), so they share columns 0-63; only the weight-gradient accumulators persist.
#define TC_COL_D 0     // 64: hidden pre-activation, later dL/dhidden
#define TC_COL_O 0     // 16: network output                      (aliases D)
#define TC_COL_E 0     // 32: dL/dencoding                        (aliases D)
#define TC_COL_G1 64   // 16: [hid|dhid]^T . dout                 (persistent)
#define TC_COL_G2 96   // 32: [hid|dhid]^T . enc                  (persistent)
#define TC_COL_H2 128  // 64: [H2|dH2]^T . H1, n_hidden_layers == 2 (persistent)
// 128 columns (NH == 1) let four CTAs share an SM's 512 TMEM columns; NH == 2 needs 192 -> 256
#define TC_TMEM_COLS(NH) ((NH) == 1 ? 128u : 256u)
#define TC_CTAS_PER_SM(NH) ((NH) == 1 ? 4 : 2)

// shared memory map (bytes from a 1024-aligned base).  Every weight matrix is stored ONCE, in the K-major layout of
// its forward GEMM; the dgrad GEMMs read the same bytes through MN-major descriptors (tc05.cuh), like the
// weight-gradient GEMMs read the activation tiles.
#define SM_ENC 0         //  8192  [128][32]  SW64
#define SM_HID 8192      // 16384  [128][64]  SW128   (hid | dhid must be adjacent: MN-major LBO = 16384)
#define SM_DHID 24576    // 16384
#define SM_DOUT 40960    //  4096  [128][16]  core layout (2 chunks per row)
#define SM_WIN 45056     //  4096  W_in  [64][32] SW64 : B of the input GEMM (K-major) and of dL/denc (MN-major)
#define SM_WOUT 49152    //  2048  W_out [16][64] SW128: B of the output GEMM (K-major) and of dL/dhidden (MN-major)
#define SM_BAR 51200     //  2 mbarriers (16) + tmem base (4)
// n_hidden_layers == 2 only: the first hidden layer's tiles and the 64x64 weight
#define SM_H1 52224      // 16384  [128][64]  SW128   (H1 | dH1 adjacent, like hid | dhid)
#define SM_DH1 68608     // 16384
#define SM_WH 84992      //  8192  W_h [64][64] SW128: B of the hidden GEMM (K-major) and of its dgrad (MN-major)
#define SM_TOTAL 93184
// dynamic request (incl. 1 KB alignment slack), sized so that exactly TC_CTAS_PER_SM CTAs fit in the 227 KB of an SM:
// NH == 1: 4 x 54 KB = all 512 TMEM columns, NH == 2: 2 x 94 KB
#define TC_SMEM_BYTES(NH) ((NH) == 1 ? 54 * 1024 : 94 * 1024)

static constexpr uint32_t IDESC_64_KK = make_idesc(128, 64, 0, 0);
static constexpr uint32_t IDESC_16_KK = make_idesc(128, 16, 0, 0);
static constexpr uint32_t IDESC_32_KK = make_idesc(128, 32, 0, 0);
static constexpr uint32_t IDESC_16_MM = make_idesc(128, 16, 1, 1);
static constexpr uint32_t IDESC_32_MM = make_idesc(128, 32, 1, 1);
static constexpr uint32_t IDESC_64_MM = make_idesc(128, 64, 1, 1);
static constexpr uint32_t IDESC_64_KM = make_idesc(128, 64, 0, 1);   // A K-major, B MN-major (dgrad through the forward weight tile)
static constexpr uint32_t IDESC_32_KM = make_idesc(128, 32, 0, 1);

struct TcCtx {
    unsigned char* sm;     // 1024-aligned shared base
    uint32_t sm_addr;      // its shared-space address
    uint32_t sm16;         // sm_addr >> 4: the base every operand descriptor's low word is an offset from (tc05::make_desc2)
    uint64_t* bar;         // completion of the per-tile GEMM chain (issued by warp 0)
    uint64_t* bar_wg;      // completion of a tile's weight-gradient GEMMs (issued by warps 1 .. n_wg, one arrival each)
    uint32_t tmem;         // TMEM base (lane 0, column 0 of the allocation)
    uint32_t phase;        // mbarrier parity of the next wait on bar
    uint32_t phase_wg;     // ... on bar_wg
    uint32_t tid, lane, warp;
};

__device__ __forceinline__ void tc_setup(TcCtx& c, unsigned char* raw, uint32_t tmem_cols, uint32_t n_wg_issuers = 1) {
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    const uint32_t raw_addr = smem_u32(raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    c.sm = raw + pad;
    c.sm_addr = raw_addr + pad;
    c.sm16 = c.sm_addr >> 4;
    c.bar = reinterpret_cast<uint64_t*>(c.sm + SM_BAR);
    c.bar_wg = reinterpret_cast<uint64_t*>(c.sm + SM_BAR + 8);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.sm + SM_BAR + 16);
    if (c.tid == 0) { mbar_init(c.bar, 1); mbar_init(c.bar_wg, n_wg_issuers); mbar_fence_init(); }
    if (c.warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, tmem_cols); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    c.tmem = *tmem_slot;
    c.phase = 0;
    c.phase_wg = 0;
}

__device__ __forceinline__ void tc_teardown(TcCtx& c, uint32_t tmem_cols) {
    fence_before_sync();
    __syncthreads();
    if (c.warp == 0) { __syncwarp(); tmem_dealloc(c.tmem, tmem_cols); }
}

// weights -> shared memory in the operand layouts (params: W_in [64][32] | (W_h [64][64]) | W_out [16][64], row-major fp16)
template <int NH>
__device__ __forceinline__ void tc_load_weights(const TcCtx& c, const __half* __restrict__ params) {
    const __half* Win = params;
    const __half* Wh = params + 64 * 32;
    const __half* Wout = params + 64 * 32 + (NH - 1) * 64 * 64;
    if (NH == 2) {
        for (uint32_t i = c.tid; i < 64 * 8; i += TC_THREADS) {       // W_h rows (N = 64 neurons), 8 chunks (K = 64)
            const uint32_t n = i >> 3, ch = i & 7;
            *reinterpret_cast<uint4*>(c.sm + SM_WH + sw128_off(n, ch)) = *reinterpret_cast<const uint4*>(Wh + n * 64 + ch * 8);
        }
    }
    for (uint32_t i = c.tid; i < 64 * 4; i += TC_THREADS) {           // W_in rows (N = 64), 4 chunks of 8 (K = 32)
        const uint32_t n = i >> 2, ch = i & 3;
        *reinterpret_cast<uint4*>(c.sm + SM_WIN + sw64_off(n, ch)) = *reinterpret_cast<const uint4*>(Win + n * 32 + ch * 8);
    }
    for (uint32_t i = c.tid; i < 16 * 8; i += TC_THREADS) {           // W_out rows (N = 16), 8 chunks (K = 64)
        const uint32_t n = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(c.sm + SM_WOUT + sw128_off(n, ch)) = *reinterpret_cast<const uint4*>(Wout + n * 64 + ch * 8);
    }
}

// Staging of one row of the 128 x 32 fp16 encoding tile in two halves, so that the global loads of the NEXT tile can
// be in flight while the tensor core and the epilogues work on the current one.  The encoding is level-major:
// enc[level][point] holds the level's two features as one 32-bit word, so a row is 16 word loads, coalesced across the
// warp (consecutive threads = consecutive points), and nothing depends on a loaded value until the row is stored.
// Invalid rows are zero-filled.
__device__ __forceinline__ void tc_load_enc(uint32_t packed[16], const __half* __restrict__ enc_lm, size_t n_total, size_t pt, bool valid) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(enc_lm) + pt;
#pragma unroll
    for (int k = 0; k < 16; ++k) packed[k] = valid ? __ldg(src + (size_t)k * n_total) : 0u;
}
__device__ __forceinline__ void tc_store_enc(const TcCtx& c, const uint32_t packed[16]) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(c.sm + SM_ENC + sw64_off(c.tid, ch)) = make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
}
__device__ __forceinline__ void tc_stage_enc(const TcCtx& c, const __half* __restrict__ enc_soa, size_t n_total, size_t pt, bool valid) {
    uint32_t packed[16];
    tc_load_enc(packed, enc_soa, n_total, pt, valid);
    tc_store_enc(c, packed);
}

// make the CTA's shared-memory writes visible to the tensor core, then let thread 0 issue
#define TC_PUBLISH_AND_SYNC()  \
    do {                       \
        fence_smem_to_async(); \
        fence_before_sync();   \
        __syncthreads();       \
    } while (0)

#define TC_WAIT(c)                      \
    do {                                \
        __syncwarp();                   \
        mbar_wait((c).bar, (c).phase);  \
        (c).phase ^= 1u;                \
        fence_after_sync();             \
    } while (0)

#define TC_WAIT_WG(c)                         \
    do {                                      \
        __syncwarp();                         \
        mbar_wait((c).bar_wg, (c).phase_wg);  \
        (c).phase_wg ^= 1u;                   \
        fence_after_sync();                   \
    } while (0)

// K-major GEMM into TMEM column `col`: A tile at a_off (row pitch 64 B -> SW64, 128 B -> SW128), nk steps of K = 16
template <uint32_t nk>
__device__ __forceinline__ void tc_mma_sw(const TcCtx& c, uint32_t col, uint32_t a_off, uint32_t b_off, bool sw128, uint32_t idesc) {
    const uint32_t sbo = sw128 ? 1024u : 512u;
    const uint64_t swz = sw128 ? SWZ_128B : SWZ_64B;
#pragma unroll
    for (uint32_t k = 0; k < nk; ++k)
        mma_f16_ss(c.tmem + col, make_desc2(c.sm16, a_off + k * 32, 16, sbo, swz), make_desc2(c.sm16, b_off + k * 32, 16, sbo, swz), idesc, k);
}
// layer in: D = enc . W_in^T
__device__ __forceinline__ void tc_issue_layer_in(const TcCtx& c) {
    fence_after_sync();
    tc_mma_sw<2>(c, TC_COL_D, SM_ENC, SM_WIN, false, IDESC_64_KK);
    commit(c.bar);
}
// hidden layer (NH == 2): D = H1 . W_h^T
__device__ __forceinline__ void tc_issue_layer_hidden(const TcCtx& c) {
    fence_after_sync();
    tc_mma_sw<4>(c, TC_COL_D, SM_H1, SM_WH, true, IDESC_64_KK);
    commit(c.bar);
}
// layer out: O = hid . W_out^T
__device__ __forceinline__ void tc_issue_layer_out(const TcCtx& c) {
    fence_after_sync();
    tc_mma_sw<4>(c, TC_COL_O, SM_HID, SM_WOUT, true, IDESC_16_KK);
    commit(c.bar);
}

// epilogue of a hidden GEMM: fp32 accumulator row -> fp16 -> packed ReLU -> activation tile at tile_off
// (rounding to fp16 and clamping at zero commute, so this equals the reference's relu-then-store)
__device__ __forceinline__ void tc_epilogue_hidden(const TcCtx& c, uint32_t tile_off) {
    const uint32_t row = c.tid;
    const uint32_t taddr = c.tmem + ((c.warp * 32u) << 16) + TC_COL_D;
    const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
    for (uint32_t half_i = 0; half_i < 2; ++half_i) {
        float v[32];
        tmem_ld32(taddr + half_i * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (uint32_t ch = 0; ch < 4; ++ch) {
            uint32_t packed[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                const __half2 h = __hmax2(__floats2half2_rn(v[ch * 8 + 2 * q], v[ch * 8 + 2 * q + 1]), zero);
                packed[q] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(c.sm + tile_off + sw128_off(row, half_i * 4 + ch)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    }
}

// 0xffff in every 16-bit half of `h` that is a non-zero fp16 (post-ReLU activations are >= +0, so non-zero == positive)
__device__ __forceinline__ uint32_t tc_pos_mask2(uint32_t h) { return __vcmpne2(h & 0x7fff7fffu, 0u); }

// epilogue of a dgrad GEMM: ReLU mask (re-read from this thread's own row of the forward activation tile at act_off,
// which it wrote itself) -> fp16 -> gradient tile at tile_off
__device__ __forceinline__ void tc_epilogue_dhidden(const TcCtx& c, uint32_t tile_off, uint32_t act_off) {
    const uint32_t taddr = c.tmem + ((c.warp * 32u) << 16) + TC_COL_D;
#pragma unroll
    for (uint32_t half_i = 0; half_i < 2; ++half_i) {
        float v[32];
        tmem_ld32(taddr + half_i * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (uint32_t ch = 0; ch < 4; ++ch) {
            const uint4 act = *reinterpret_cast<const uint4*>(c.sm + act_off + sw128_off(c.tid, half_i * 4 + ch));
            const uint32_t a[4] = {act.x, act.y, act.z, act.w};
            uint32_t packed[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(v[ch * 8 + 2 * q], v[ch * 8 + 2 * q + 1]);
                packed[q] = *reinterpret_cast<const uint32_t*>(&h) & tc_pos_mask2(a[q]);
            }
            *reinterpret_cast<uint4*>(c.sm + tile_off + sw128_off(c.tid, half_i * 4 + ch)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    }
}

// forward through the hidden layers of one staged tile; leaves the last activations in SM_HID and issues nothing after
template <int NH>
__device__ __forceinline__ void tc_forward_hidden(TcCtx& c) {
    if (NH == 1) {
        tc_epilogue_hidden(c, SM_HID);
    } else {
        tc_epilogue_hidden(c, SM_H1);
        TC_PUBLISH_AND_SYNC();
        if (c.tid == 0) tc_issue_layer_hidden(c);
        TC_WAIT(c);
        tc_epilogue_hidden(c, SM_HID);
    }
}

// ------------------------------------------------------------------------------------------ training
struct TileIn { float tmin, tmax, xi, tgt[3], tgt_depth, bg[3]; uint8_t inst; };

// everything the compositing / loss epilogue needs about ray `ray` and this lane's sample; rays beyond the batch
// (never the case for R % 4 == 0) get neutral values
__device__ __forceinline__ TileIn tc_load_tile_inputs(const MonBatch& b, uint32_t ray, uint32_t lane, uint32_t iter) {
    // loads only — nothing here may depend on a loaded value, so the requests stay in flight until the next tile uses them
    TileIn in;
    if (ray < b.R) {
        in.tmin = b.rays[ray].tmin; in.tmax = b.rays[ray].tmax;
#pragma unroll
        for (int k = 0; k < 3; ++k) { in.tgt[k] = b.target[ray * 3 + k]; in.bg[k] = b.bg[ray * 3 + k]; }
        in.tgt_depth = b.target_depth[ray];
        in.inst = b.ray_inst[ray];
        in.xi = b.inj_dt ? __ldg(b.inj_dt + ray * 32 + lane) : mon_u01(mon_hash4(b.seed, iter, 2, ray * 32 + lane));
    } else {
        in.tmin = 0.0f; in.tmax = 1.0f; in.xi = 1.0f; in.tgt_depth = 0.0f; in.inst = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { in.tgt[k] = 0.0f; in.bg[k] = 0.0f; }
    }
    return in;
}

// FUSE: the hash-grid gradient scatter of the iteration happens HERE, in the last epilogue of every tile, instead of in a kernel of
// its own behind this one.  In steady state the early stop leaves ~1 sample in 12 with a non-zero dL/dencoding row (2-3 per ray);
// the warp spreads the (live sample, level) pairs of its ray over its lanes — lane & 15 = level, two samples per round — and each
// lane issues the level's f16x2 reductions (scatter_global.cuh: the reference's own form, kernel_grid_backward's atomicAdd(__half2),
// TCNN encodings/grid.h:386-509) straight into the entry-ordered gradient table the optimizer sweep reads.  The compacted
// gradient / position lists, the scatter launch, its dependency edges and its tail leave the iteration's critical path (13 us
// kernel + 2 x 2 us of edges against ~1 us more in here: this kernel is latency-bound and has the issue slots).  The host picks
// this variant from the live-sample count of the object's previous calls (mon_core.cu); a fresh object, whose every sample is
// live, keeps the shared-memory resident scatter kernel.
struct MonFuse {
    MonGrid g;
    __half* gh_grid;     // entry-ordered fp16 gradient table of the grid (the optimizer sweep consumes and zeroes it)
};

template <int NH, bool FUSE>
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM(NH))
k_mlp_train_tc(MonBatch b, MonLossCfg lc, uint32_t n_mlp, const __grid_constant__ MonFuse fz) {
    extern __shared__ unsigned char smem_raw[];
    // Prologue that overlaps the tail of the hash-encode kernel (programmatic dependent launch, mon_kernels.h):
    // TMEM allocation, barriers, weight tiles, the constant part of the dout tile and the first tile's ray inputs.
    // Everything read here was written by kernels OLDER than the encode kernel (weights: the previous optimizer
    // sweep; rays/targets/control block: this iteration's batch kernel), which have completed.
    TC_STAMP(0);
    TC_CTA_MARK(0);
    TcCtx c;
    tc_setup(c, smem_raw, TC_TMEM_COLS(NH), NH == 1 ? 2u : 3u);
    TC_STAMP(1);
    tc_load_weights<NH>(c, b.params);
    TC_STAMP(2);
    // the upper half of every dout row (outputs 4..15 and the second K chunk) stays zero for the whole kernel
    for (uint32_t i = c.tid; i < 4096 / 16; i += TC_THREADS) reinterpret_cast<uint4*>(c.sm + SM_DOUT)[i] = make_uint4(0, 0, 0, 0);

    // FUSE: this lane's level for the whole kernel (lane & 15) and its constants
    float fz_scale = 0.0f; uint32_t fz_res = 0, fz_size = 0; bool fz_hashed = false; char* fz_tab = nullptr;
    if (FUSE) {
        const uint32_t l = c.lane & 15u;
        if (l < fz.g.n_levels) {
            fz_scale = fz.g.scale[l]; fz_res = fz.g.res[l]; fz_size = fz.g.size[l]; fz_hashed = fz.g.hashed[l] != 0;
            fz_tab = reinterpret_cast<char*>(reinterpret_cast<__half2*>(fz.gh_grid) + fz.g.offset[l]);
        }
    }
    const bool skip = b.ctrl->skip != 0;
    const float kscale = lc.loss_scale / (float)b.R;
    const uint32_t iter = b.ctrl->iter - 1;
    const uint32_t n_tiles = (b.R + 3) / 4;
    uint32_t tiles_done = 0;

    TileIn tin;              // this warp's ray (targets, background, sample distance of this lane), one tile ahead
    if (!skip && blockIdx.x < n_tiles) tin = tc_load_tile_inputs(b, blockIdx.x * 4 + c.warp, c.lane, iter);
    mon_pdl_wait();          // the encodings of this iteration are complete
    mon_pdl_trigger();
    MON_TL(MON_TL_M, iter);
    // hand the iteration's control block to the kernels behind this one (scatter, optimizer): the batch kernel of
    // the next iteration is allowed to overwrite the live block while they run
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        MonCtrl cc = *b.ctrl;
        cc.loss_mean = b.late->loss_mean;   // the logged loss survives a skipped iteration (the optimizer rewrites it otherwise)
        *b.late = cc;
    }
    if (skip) {
        tc_teardown(c, TC_TMEM_COLS(NH));
        return;
    }
    uint32_t enc_regs[16];   // this thread's encoding row of the NEXT tile, prefetched one tile ahead
    if (blockIdx.x < n_tiles) tc_load_enc(enc_regs, b.enc, (size_t)b.R * 32, (size_t)blockIdx.x * 128 + c.tid, blockIdx.x * 4 + c.warp < b.R);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tiles_done) {
        const uint32_t ray = tile * 4 + c.warp;
        const bool ray_ok = ray < b.R;
        const uint32_t pt = ray * 32 + c.lane;
        // ---- stage encodings, MMA1
        const uint32_t sb = 4 + (tiles_done & 1) * 16;
        TC_STAMP(sb + 0);
        if (tiles_done) TC_WAIT_WG(c);      // the previous tile's weight-gradient GEMMs have read enc / hid / dhid / dout
        tc_store_enc(c, enc_regs);
        TC_PUBLISH_AND_SYNC();
        TC_STAMP(sb + 1);
        if (c.tid == 0) tc_issue_layer_in(c);
        TC_STAMP(sb + 2);
        {
            const uint32_t next = tile + gridDim.x;
            if (next < n_tiles) tc_load_enc(enc_regs, b.enc, (size_t)b.R * 32, (size_t)next * 128 + c.tid, next * 4 + c.warp < b.R);
        }
        // per-ray inputs of the renderer: this tile's were fetched one tile ahead, the next tile's are requested now
        MonRay r_cur;
        r_cur.tmin = tin.tmin; r_cur.tmax = tin.tmax;
        const float t = mon_sample_t(r_cur, c.lane, tin.xi, 32.0f);
        RayTargets rt;
#pragma unroll
        for (int k = 0; k < 3; ++k) { rt.tgt[k] = tin.tgt[k]; rt.bg[k] = tin.bg[k]; }
        rt.tgt_depth = tin.tgt_depth;
        rt.is_obj = tin.inst == 1;
        {
            const uint32_t next = tile + gridDim.x;
            if (next < n_tiles) tin = tc_load_tile_inputs(b, next * 4 + c.warp, c.lane, iter);
        }
        // this sample's unit-cube position, copied beside its gradient row if the sample turns out to carry gradient;
        // requested here so that the load is long complete when the last epilogue of the tile needs it
        float pu[3] = {0.0f, 0.0f, 0.0f};
        if (ray_ok) { pu[0] = __ldg(b.pts + (size_t)pt * 3); pu[1] = __ldg(b.pts + (size_t)pt * 3 + 1); pu[2] = __ldg(b.pts + (size_t)pt * 3 + 2); }
        TC_STAMP(sb + 3);
        TC_WAIT(c);
        TC_STAMP(sb + 4);
        // ---- hidden epilogue(s), MMA2
        tc_forward_hidden<NH>(c);
        TC_STAMP(sb + 5);
        TC_PUBLISH_AND_SYNC();
        if (c.tid == 0) tc_issue_layer_out(c);
        TC_STAMP(sb + 6);
        TC_WAIT(c);
        TC_STAMP(sb + 7);
        // ---- output epilogue: compositing, loss, dL/dout
        {
            float o16[16];
            tmem_ld16(c.tmem + ((c.warp * 32u) << 16) + TC_COL_O, o16);
            tmem_wait_ld();
            float o[4], go[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = __half2float(__float2half_rn(o16[k]));   // the network output is fp16
            // opt-in occupancy mode: a sample in an unoccupied cell was not encoded (its row is stale): empty space, no gradient
            const bool empty = b.occ.bits && ray_ok && !((__ldg(b.occ.ray_mask + ray) >> c.lane) & 1u);
            if (empty) o[3] = -1e30f;                                                    // density exp(o[3]) = 0
            const RayResult rr = warp_render_loss_grad(o, t, c.lane, rt, kscale, lc, go);
            if (empty) { go[0] = go[1] = go[2] = go[3] = 0.0f; }
            if (ray_ok && c.lane == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) b.rgb_rays[ray * 3 + k] = rr.rgb[k];
                b.depth_rays[ray] = rr.depth; b.mask_rays[ray] = rr.mask; b.loss[ray] = rr.loss;
            }
            if (!ray_ok) { go[0] = go[1] = go[2] = go[3] = 0.0f; }
            const __half2 g01 = __floats2half2_rn(go[0], go[1]), g23 = __floats2half2_rn(go[2], go[3]);
            *reinterpret_cast<uint2*>(c.sm + SM_DOUT + core_off(c.tid, 0, 2)) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&g01), *reinterpret_cast<const uint32_t*>(&g23));
            if (b.dbg_out && ray_ok) {
#pragma unroll
                for (int k = 0; k < 4; ++k) b.dbg_out[(size_t)pt * 4 + k] = o[k];
                b.dbg_dout[(size_t)pt * 4 + 0] = __low2float(g01); b.dbg_dout[(size_t)pt * 4 + 1] = __high2float(g01);
                b.dbg_dout[(size_t)pt * 4 + 2] = __low2float(g23); b.dbg_dout[(size_t)pt * 4 + 3] = __high2float(g23);
            }
        }
        // ---- MMA3: dL/dhidden = dout . W_out
        TC_STAMP(sb + 8);
        TC_PUBLISH_AND_SYNC();
        if (c.tid == 0) {
            fence_after_sync();
            // B[k = output o][n = neuron j] = W_out[o][j]: the forward tile (16 rows of 128 B) read MN-major
            mma_f16_ss(c.tmem + TC_COL_D, make_desc2(c.sm16, SM_DOUT, 128, 256, SWZ_NONE),
                       make_desc2(c.sm16, SM_WOUT, 16, 1024, SWZ_128B), IDESC_64_KM, 0);
            commit(c.bar);
        }
        TC_STAMP(sb + 9);
        TC_WAIT(c);
        TC_STAMP(sb + 10);
        tc_epilogue_dhidden(c, SM_DHID, SM_HID);
        TC_STAMP(sb + 11);
        if (NH == 2) {   // dH1 = (dH2 . W_h) * relu'(H1)
            TC_PUBLISH_AND_SYNC();
            if (c.tid == 0) {
                fence_after_sync();
                // B[k = neuron j][n = input i] = W_h[j][i]: the forward tile read MN-major, 16 rows (2 KB) per MMA
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    mma_f16_ss(c.tmem + TC_COL_D, make_desc2(c.sm16, SM_DHID + k * 32, 16, 1024, SWZ_128B),
                               make_desc2(c.sm16, SM_WH + k * 2048, 16, 1024, SWZ_128B), IDESC_64_KM, k);
                commit(c.bar);
            }
            TC_WAIT(c);
            tc_epilogue_dhidden(c, SM_DH1, SM_H1);
        }
        // ---- MMA4 (dL/dencoding) + MMA5/6 (weight gradients, accumulated in TMEM across tiles).  Warp 0 issues the four MMAs the
        // tile's last epilogue waits for; warps 1 .. 3 issue the weight-gradient GEMMs, one accumulator each, onto their own
        // barrier: nothing reads G1 / G2 / G3 before the end of the kernel, so they run under the dL/dencoding epilogue and are
        // only waited for before the next tile overwrites the operand tiles (top of the loop).  One thread issuing all 20 MMAs
        // was the longest serial phase of a tile (1750 of 8500 cycles, profiles/r1z_mlp_phase_times.txt).
        TC_PUBLISH_AND_SYNC();
        {
            const uint32_t acc0 = tiles_done ? 1u : 0u;
            if (c.tid == 0) {
                fence_after_sync();
                // B[k = neuron j][n = input k] = W_in[j][k]: the forward tile (64 rows of 64 B, SW64) read MN-major
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    mma_f16_ss(c.tmem + TC_COL_E, make_desc2(c.sm16, (NH == 2 ? SM_DH1 : SM_DHID) + k * 32, 16, 1024, SWZ_128B),
                               make_desc2(c.sm16, SM_WIN + k * 1024, 16, 512, SWZ_64B), IDESC_32_KM, k);
                commit(c.bar);
            } else if (c.tid == 32) {
                fence_after_sync();
                // [last hidden | its gradient]^T: rows 0-63 x dout = dW_out^T     (16 points per MMA)
#pragma unroll
                for (uint32_t kk = 0; kk < 8; ++kk)
                    mma_f16_ss(c.tmem + TC_COL_G1, make_desc2(c.sm16, SM_HID + kk * 2048, 16384, 1024, SWZ_128B),
                               make_desc2(c.sm16, SM_DOUT + kk * 512, 256, 128, SWZ_NONE), IDESC_16_MM, acc0 | kk);
                commit(c.bar_wg);
            } else if (c.tid == 64) {
                fence_after_sync();
                // NH == 1: [hid | dhid]^T x enc, NH == 2: [H1 | dH1]^T x enc: rows 64-127 = dW_in
#pragma unroll
                for (uint32_t kk = 0; kk < 8; ++kk)
                    mma_f16_ss(c.tmem + TC_COL_G2, make_desc2(c.sm16, (NH == 2 ? SM_H1 : SM_HID) + kk * 2048, 16384, 1024, SWZ_128B),
                               make_desc2(c.sm16, SM_ENC + kk * 1024, 16, 512, SWZ_64B), IDESC_32_MM, acc0 | kk);
                commit(c.bar_wg);
            } else if (NH == 2 && c.tid == 96) {
                fence_after_sync();
                // [H2 | dH2]^T x H1: rows 64-127 = dW_h
#pragma unroll
                for (uint32_t kk = 0; kk < 8; ++kk)
                    mma_f16_ss(c.tmem + TC_COL_H2, make_desc2(c.sm16, SM_HID + kk * 2048, 16384, 1024, SWZ_128B),
                               make_desc2(c.sm16, SM_H1 + kk * 2048, 16, 1024, SWZ_128B), IDESC_64_MM, acc0 | kk);
                commit(c.bar_wg);
            }
        }
        TC_STAMP(sb + 12);
        TC_WAIT(c);
        TC_STAMP(sb + 13);
        {   // dL/dencoding: fp16.  Most samples sit behind the early stop (or underflow in fp16) and have an all-zero row:
            // only the LIVE rows are handed to the scatter + Adam kernel, compacted — slot k gets the position and, level by
            // level (stride N words, so that a CTA of that kernel streams one level's words of consecutive slots), the two
            // fp16 gradients of the level.  One atomicAdd per warp reserves the slots of its live lanes.
            float v[32];
            tmem_ld32(c.tmem + ((c.warp * 32u) << 16) + TC_COL_E, v);
            tmem_wait_ld();
            uint32_t packed[16];
            uint32_t any = 0u;
#pragma unroll
            for (uint32_t q = 0; q < 16; ++q) {
                const __half2 h = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                packed[q] = *reinterpret_cast<const uint32_t*>(&h);
                any |= packed[q];
            }
            const bool live = ray_ok && (any & 0x7fff7fffu) != 0u;
            const uint32_t lm = __ballot_sync(0xffffffffu, live);
            if (FUSE) {
                if (lm) {
                    if (c.lane == 0) atomicAdd(b.live_cnt + (iter & 1u), (uint32_t)__popc(lm));     // the host reads it to choose the graph variant
                    const uint32_t half_i = c.lane >> 4, myl = c.lane & 15u;
                    uint32_t rem = lm;
                    while (rem) {            // warp-uniform: two live samples per round, lanes 0-15 / 16-31 take their 16 levels
                        const uint32_t sa = (uint32_t)__ffs((int)rem) - 1u;
                        rem &= rem - 1u;
                        uint32_t sb = 32u;
                        if (rem) { sb = (uint32_t)__ffs((int)rem) - 1u; rem &= rem - 1u; }
                        const uint32_t src = half_i ? sb : sa;
                        uint32_t gw = 0u;
#pragma unroll
                        for (uint32_t q = 0; q < 16; ++q) {
                            const uint32_t v = __shfl_sync(0xffffffffu, packed[q], src & 31u);
                            if (q == myl) gw = v;
                        }
                        const float u[3] = {__shfl_sync(0xffffffffu, pu[0], src & 31u), __shfl_sync(0xffffffffu, pu[1], src & 31u),
                                            __shfl_sync(0xffffffffu, pu[2], src & 31u)};
                        // a zero pair adds nothing (the stand-alone kernel skips it too); lanes beyond the grid's levels idle
                        if (src < 32u && fz_tab && (gw & 0x7fff7fffu) != 0u)
                            scatter_level_pow2(fz_scale, fz_res, fz_hashed, fz_size, fz_tab, __half2float(__ushort_as_half((unsigned short)(gw & 0xffffu))),
                                               __half2float(__ushort_as_half((unsigned short)(gw >> 16))), u);
                    }
                }
            } else if (lm) {
                uint32_t base = 0;
                if (c.lane == 0) base = atomicAdd(b.live_cnt + (iter & 1u), (uint32_t)__popc(lm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (live) {
                    const uint32_t slot = base + __popc(lm & ((1u << c.lane) - 1u));
                    const size_t n_total = (size_t)b.R * 32;
#pragma unroll
                    for (uint32_t q = 0; q < 16; ++q) b.genc[(size_t)q * n_total + slot] = packed[q];
                    reinterpret_cast<float4*>(b.pts_c)[slot] = make_float4(pu[0], pu[1], pu[2], 0.0f);      // one 16-byte word per slot
                }
            }
            if (b.d_enc && ray_ok) {   // parity hook: the uncompacted rows, point-major
                uint4* dst = reinterpret_cast<uint4*>(b.d_enc + (size_t)pt * MON_IN);
#pragma unroll
                for (uint32_t ch = 0; ch < 4; ++ch) dst[ch] = make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
            }
        }
        TC_STAMP(sb + 14);
        // the next tile overwrites enc / hid / dout once the weight-gradient GEMMs that read them have completed (top of the loop)
    }
    if (tiles_done) TC_WAIT_WG(c);
    TC_STAMP(40);

    // ---- weight gradients: TMEM -> this CTA's partial row (W_in [64][32] | (W_h [64][64]) | W_out [16][64])
    float* partial = b.mlp_partials + (size_t)blockIdx.x * n_mlp;
    if (tiles_done == 0) {
        for (uint32_t i = c.tid; i < n_mlp; i += TC_THREADS) partial[i] = 0.0f;
    } else {
        if (c.warp < 2) {        // G1 rows 0..63: (hid^T . dout)[j][o] = dW_out[o][j]
            float v[16];
            tmem_ld16(c.tmem + ((c.warp * 32u) << 16) + TC_COL_G1, v);
            tmem_wait_ld();
            const uint32_t j = c.tid;
#pragma unroll
            for (uint32_t o = 0; o < 16; ++o) partial[64 * 32 + (NH - 1) * 64 * 64 + o * 64 + j] = v[o];
        } else {                 // G2 rows 64..127: (dhid^T . enc)[j][k] = dW_in[j][k]
            float v[32];
            tmem_ld32(c.tmem + ((c.warp * 32u) << 16) + TC_COL_G2, v);
            tmem_wait_ld();
            const uint32_t j = c.tid - 64;
#pragma unroll
            for (uint32_t k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(partial + j * 32 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
            if (NH == 2) {   // G3 rows 64..127: (dH2^T . H1)[j][i] = dW_h[j][i]
#pragma unroll
                for (uint32_t half_i = 0; half_i < 2; ++half_i) {
                    tmem_ld32(c.tmem + ((c.warp * 32u) << 16) + TC_COL_H2 + half_i * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (uint32_t k = 0; k < 32; k += 4)
                        *reinterpret_cast<float4*>(partial + 64 * 32 + j * 64 + half_i * 32 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                }
            }
        }
    }
    TC_STAMP(41);
    tc_teardown(c, TC_TMEM_COLS(NH));
    TC_STAMP(42);
    TC_CTA_MARK(1);
}

// ------------------------------------------------------------------------------------------ inference
// raw network output (4 logits per point) for the density lattice and the parity hooks
template <int NH>
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM(NH))
k_mlp_infer_tc(uint32_t n_points, const __half* __restrict__ params, const __half* __restrict__ enc, float* __restrict__ out4) {
    extern __shared__ unsigned char smem_raw[];
    TcCtx c;
    tc_setup(c, smem_raw, TC_TMEM_COLS(NH));
    tc_load_weights<NH>(c, params);
    const uint32_t n_tiles = (n_points + 127) / 128;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t pt = tile * 128 + c.tid;
        tc_stage_enc(c, enc, n_points, pt, pt < n_points);
        TC_PUBLISH_AND_SYNC();
        if (c.tid == 0) tc_issue_layer_in(c);
        TC_WAIT(c);
        tc_forward_hidden<NH>(c);
        TC_PUBLISH_AND_SYNC();
        if (c.tid == 0) tc_issue_layer_out(c);
        TC_WAIT(c);
        float o16[16];
        tmem_ld16(c.tmem + ((c.warp * 32u) << 16) + TC_COL_O, o16);
        tmem_wait_ld();
        if (pt < n_points) {
            float4 o;
            o.x = __half2float(__float2half_rn(o16[0])); o.y = __half2float(__float2half_rn(o16[1]));
            o.z = __half2float(__float2half_rn(o16[2])); o.w = __half2float(__float2half_rn(o16[3]));
            reinterpret_cast<float4*>(out4)[pt] = o;
        }
    }
    tc_teardown(c, TC_TMEM_COLS(NH));
}

// test render (VolumeRender_Render, nerf_model.cu:1134-1229): per ray S2/32 tiles with a carried compositing state
template <int NH>
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM(NH))
k_mlp_render_tc(uint32_t n_rays, uint32_t S2, const MonRay* __restrict__ rays, const int* __restrict__ in_box,
                const float* __restrict__ jitter, uint32_t seed, uint32_t iter, const __half* __restrict__ params,
                const __half* __restrict__ enc, float bgc, float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ mask,
                const uint32_t* __restrict__ orig_ray) {
    // orig_ray (render path): the rays are the compacted hits; results and jitter are indexed by the original pixel
    extern __shared__ unsigned char smem_raw[];
    TcCtx c;
    tc_setup(c, smem_raw, TC_TMEM_COLS(NH));
    tc_load_weights<NH>(c, params);
    const uint32_t n_groups = (n_rays + 3) / 4, chunks = S2 / 32;
    for (uint32_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint32_t ray = grp * 4 + c.warp;
        const bool ray_ok = ray < n_rays;
        const bool hit = ray_ok && (!in_box || in_box[ray] != 0);
        const uint32_t oray = (ray_ok && orig_ray) ? orig_ray[ray] : ray;
        MonRay r;
        if (ray_ok) r = rays[ray];
        else { r.tmin = 0.0f; r.tmax = 1.0f; r.d_norm = 1.0f; }
        RenderCarry cr; cr.T = 1.0f; cr.C[0] = cr.C[1] = cr.C[2] = 0.0f; cr.D = 0.0f; cr.last_t = 0.0f;
        for (uint32_t chunk = 0; chunk < chunks; ++chunk) {
            // this thread stages the row of its own sample: point (ray, chunk*32 + lane)
            const uint32_t pt = ray * S2 + chunk * 32 + c.lane;
            tc_stage_enc(c, enc, (size_t)n_rays * S2, pt, hit);
            TC_PUBLISH_AND_SYNC();
            if (c.tid == 0) tc_issue_layer_in(c);
            TC_WAIT(c);
            tc_forward_hidden<NH>(c);
            TC_PUBLISH_AND_SYNC();
            if (c.tid == 0) tc_issue_layer_out(c);
            TC_WAIT(c);
            float o16[16];
            tmem_ld16(c.tmem + ((c.warp * 32u) << 16) + TC_COL_O, o16);
            tmem_wait_ld();
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = __half2float(__float2half_rn(o16[k]));
            const uint32_t n = chunk * 32 + c.lane;
            const float xi = hit ? mon_rand(jitter, seed, iter, 3, oray * S2 + n) : 1.0f;
            const float t = mon_sample_t(r, n, xi, (float)S2);
            warp_render_chunk(o, t, c.lane, cr);
        }
        if (ray_ok && c.lane == 0) {
            if (hit && 1.0f - cr.T > 0.5f) {
#pragma unroll
                for (int k = 0; k < 3; ++k) rgb[oray * 3 + k] = cr.C[k] + cr.T * bgc;
                depth[oray] = __fdiv_rn(cr.D, r.d_norm);
                mask[oray] = 1.0f;
            } else {
                rgb[oray * 3] = rgb[oray * 3 + 1] = rgb[oray * 3 + 2] = bgc; depth[oray] = 0.0f; mask[oray] = 0.0f;
            }
        }
    }
    tc_teardown(c, TC_TMEM_COLS(NH));
}

// ------------------------------------------------------------------------------------------ launchers
template <typename K>
static cudaError_t tc_prepare(K kernel, int smem_bytes) {
    // ask for the full shared-memory carve-out: the residency plan (4 x 54 KB or 2 x 94 KB per SM) needs it
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

template <int NH, bool FUSE>
static cudaError_t tc_launch_train(const MonBatch& b, const MonLossCfg& lc, uint32_t n_mlp, uint32_t n_ctas, cudaStream_t st, const MonLaunchOpt& lo, const MonFuse& fz) {
    static std::atomic<uint64_t> prepared{0};
    const cudaError_t prep = mon_once_per_device(prepared, [] { return tc_prepare(k_mlp_train_tc<NH, FUSE>, TC_SMEM_BYTES(NH)); });
    if (prep != cudaSuccess) return prep;
    return mon_launch_chain(MON_PDL_MLP, lo, k_mlp_train_tc<NH, FUSE>, dim3(n_ctas), dim3(TC_THREADS), TC_SMEM_BYTES(NH), st, b, lc, n_mlp, fz);
}

// fuse_grid != nullptr: the kernel scatters the hash-grid gradients itself into gh_grid (power-of-two tables only, see
// mon_scatter_resident_supported) and writes no compacted lists; no scatter kernel may follow it then
cudaError_t mon_launch_mlp_train_tc(const MonBatch& b, const MonLossCfg& lc, uint32_t n_hidden, uint32_t n_mlp, uint32_t n_ctas, cudaStream_t st,
                                    const MonLaunchOpt& lo, const MonGrid* fuse_grid, __half* gh_grid) {
    MonFuse fz;
    memset(&fz, 0, sizeof(fz));
    if (fuse_grid) { fz.g = *fuse_grid; fz.gh_grid = gh_grid; }
    if (n_hidden == 1) return fuse_grid ? tc_launch_train<1, true>(b, lc, n_mlp, n_ctas, st, lo, fz) : tc_launch_train<1, false>(b, lc, n_mlp, n_ctas, st, lo, fz);
    if (n_hidden == 2) return fuse_grid ? tc_launch_train<2, true>(b, lc, n_mlp, n_ctas, st, lo, fz) : tc_launch_train<2, false>(b, lc, n_mlp, n_ctas, st, lo, fz);
    return cudaErrorNotSupported;
}

cudaError_t mon_launch_mlp_infer_tc(uint32_t n_points, uint32_t n_hidden, const __half* params, const __half* enc, float* out4, cudaStream_t st) {
    uint32_t ctas = (n_points + 127) / 128;
    if (ctas > 592) ctas = 592;
    if (ctas == 0) ctas = 1;
    if (n_hidden == 1) {
        static std::atomic<uint64_t> prepared{0};
        const cudaError_t prep = mon_once_per_device(prepared, [] { return tc_prepare(k_mlp_infer_tc<1>, TC_SMEM_BYTES(1)); });
        if (prep != cudaSuccess) return prep;
        k_mlp_infer_tc<1><<<ctas, TC_THREADS, TC_SMEM_BYTES(1), st>>>(n_points, params, enc, out4);
    } else if (n_hidden == 2) {
        static std::atomic<uint64_t> prepared{0};
        const cudaError_t prep = mon_once_per_device(prepared, [] { return tc_prepare(k_mlp_infer_tc<2>, TC_SMEM_BYTES(2)); });
        if (prep != cudaSuccess) return prep;
        k_mlp_infer_tc<2><<<ctas, TC_THREADS, TC_SMEM_BYTES(2), st>>>(n_points, params, enc, out4);
    } else {
        return cudaErrorNotSupported;
    }
    return cudaGetLastError();
}

cudaError_t mon_launch_mlp_render_tc(uint32_t n_rays, uint32_t S2, uint32_t n_hidden, const MonRay* rays, const int* in_box, const float* jitter,
                                     uint32_t seed, uint32_t iter, const __half* params, const __half* enc, float bgc,
                                     float* rgb, float* depth, float* mask, cudaStream_t st, const uint32_t* orig_ray) {
    uint32_t ctas = (n_rays + 3) / 4;
    if (ctas > 592) ctas = 592;
    if (ctas == 0) ctas = 1;
    if (n_hidden == 1) {
        static std::atomic<uint64_t> prepared{0};
        const cudaError_t prep = mon_once_per_device(prepared, [] { return tc_prepare(k_mlp_render_tc<1>, TC_SMEM_BYTES(1)); });
        if (prep != cudaSuccess) return prep;
        k_mlp_render_tc<1><<<ctas, TC_THREADS, TC_SMEM_BYTES(1), st>>>(n_rays, S2, rays, in_box, jitter, seed, iter, params, enc, bgc, rgb, depth, mask, orig_ray);
    } else if (n_hidden == 2) {
        static std::atomic<uint64_t> prepared{0};
        const cudaError_t prep = mon_once_per_device(prepared, [] { return tc_prepare(k_mlp_render_tc<2>, TC_SMEM_BYTES(2)); });
        if (prep != cudaSuccess) return prep;
        k_mlp_render_tc<2><<<ctas, TC_THREADS, TC_SMEM_BYTES(2), st>>>(n_rays, S2, rays, in_box, jitter, seed, iter, params, enc, bgc, rgb, depth, mask, orig_ray);
    } else {
        return cudaErrorNotSupported;
    }
    return cudaGetLastError();
}
