// mon_core.cu — host side of libmon_b200.so: the C ABI declared in include/mon_c.h.
//
// Reference counterparts (MON = /root/reference/dependencies/Multi-Object-NeRF):
//   NeRF_Dataset                MON/Core/src/nerf_data.cu:123-353   -> mon_dataset_*
//   NeRF_Model ctor/ResetNetwork/AllocateBatchWorkspace
//                               MON/Core/src/nerf_model.cu:1259-1427 -> mon_object_create
//   UpdateFrameIdAndBbox[Online]  :1609-1628                        -> mon_object_{set,add}_bboxes
//   Train_Step / Train_Step_Online :1630-1699                       -> mon_object_train*
//   Render                        :1702-1830                        -> mon_object_render
//   GetDensityOnGrid              :2007-2043                        -> mon_object_density_grid
//
// One training iteration is six working kernels (batch, sample points, hash encode, fused MLP + render + loss +
// backward, gradient scatter — shared-memory resident or global reductions, chosen on the device by the live-sample
// count —, optimizer sweep + logged loss), captured as CUDA graphs of exactly the requested number of iterations and replayed; the reference
// issues ~25 launches, 3 cuRAND host calls and 3 blocking stream synchronisations per iteration
// (SURVEY.md §3.1).  There is no CPU fallback anywhere in this
// file: without a CUDA device every compute entry point returns MON_ERR_NO_DEVICE / MON_ERR_CUDA.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "mon_json.h"
#include "mon_kernels.h"

#define MON_DEBIAS_LUT 32768  // steps covered by the Adam bias-correction table (offline jobs run 5000 iterations)
#define MON_FRAMES_PER_SLAB 32
#define MON_FUSE_BELOW 20480u          // live samples below which a call takes the graphs without a scatter kernel (measured crossover 20-21 k: profiles/r9g_fuse_threshold.txt)
#define MON_RESIDENT_MIN_LIVE 16384u   // live samples from which the shared-memory resident scatter takes an iteration (measured crossover)
#define MON_GRAPH_CHUNK 64   // longest captured graph: a call of n iterations replays one chunk length that divides n (32..64) where
                             // there is one — 500 = 10 x 50, a single graph to instantiate — else n / 64 graphs of 64 + one of exactly n % 64
#define MON_BOX_CAP0 1024    // 2-D boxes the box buffer holds from the start (the iteration graphs capture its address)
#define MON_GRAPH_CACHE 6    // distinct remainder lengths kept instantiated per object (least recently used is dropped)

static thread_local std::string g_err;

// CUDA loads kernels lazily by default (CUDA_MODULE_LOADING=LAZY): the FIRST launch of every kernel loads its module under a
// driver-wide lock.  In online mode that happened in the middle of the run — the first mesh update launches the inference
// kernels — and the frontend thread's keyframe upload sat 10-90 ms inside cudaMemcpyAsync / cudaSetDevice waiting for that lock
// (MON_INGEST_TRACE=1, profiles/r5l_facade_runs.txt).  Loading everything when the context is created moves that cost to
// start-up.  Takes effect when the library is loaded before CUDA is initialised (the SLAM frontend links it); a user setting wins.
// The same start-up hook raises the number of hardware work queues: with the default 8, the streams of one process (2 per object
// + 1 per dataset) share queues, and a keyframe copy that lands in the queue of a training stream waits behind its 35 ms graph
// replays (ingest calls of 10-50 ms inside cudaEventSynchronize on the staging slot).
__attribute__((constructor)) static void mon_process_setup() {
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
}

unsigned mon_pdl_mask() {   // A/B switch for the programmatic-dependent-launch chain (mon_kernels.h)
    static const unsigned mask = [] {
        if (std::getenv("MON_NO_PDL")) return 0u;
        const char* m = std::getenv("MON_PDL_MASK");
        // measured on B200 (profiles/r1g_pdl_ab.txt): encode, fused MLP and optimizer gain from the overlap (+3 %), the
        // 1024-CTA scatter loses 7-9 % when its CTAs become resident early — also in its compacted form, and also when
        // the fused MLP kernel triggers only after its last tile (-4 %) —, sample points is neutral
        return m ? (unsigned)std::atoi(m) : (unsigned)(MON_PDL_ENCODE | MON_PDL_MLP | MON_PDL_OPTIM);
    }();
    return mask;
}

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// Device memory comes from the device's default stream-ordered pool (cudaMallocAsync / cudaFreeAsync).  cudaMalloc and cudaFree
// synchronise the WHOLE device: an object created, a scratch buffer grown or a keyframe slab added while other objects train
// stalled their graph replays and the frontend thread's keyframe ingest behind them (online mode: single calls of 14-260 ms
// among 0.3 ms ones, profiles/r5j_facade_runs.txt).  The pool keeps what is freed (release threshold = max), and an allocation
// is usable from any stream once the allocating stream has been synchronised, which the helper does (that stream only).
static cudaError_t mon_dev_malloc(void** p, size_t bytes, cudaStream_t st) {
    static std::atomic<uint64_t> prepared{0};
    cudaError_t e = mon_once_per_device(prepared, [] {
        int dev = 0;
        cudaError_t e2 = cudaGetDevice(&dev);
        cudaMemPool_t pool;
        if (e2 == cudaSuccess) e2 = cudaDeviceGetDefaultMemPool(&pool, dev);
        uint64_t keep = UINT64_MAX;
        if (e2 == cudaSuccess) e2 = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        // pool memory is reachable from the other GPUs of the node (keyframe replication copies peer to peer over NVLink)
        int n_dev = 0;
        if (e2 == cudaSuccess && cudaGetDeviceCount(&n_dev) == cudaSuccess) {
            for (int peer = 0; peer < n_dev; ++peer) {
                int can = 0;
                if (peer == dev || cudaDeviceCanAccessPeer(&can, peer, dev) != cudaSuccess || !can) continue;
                cudaMemAccessDesc desc = {};
                desc.location.type = cudaMemLocationTypeDevice;
                desc.location.id = peer;
                desc.flags = cudaMemAccessFlagsProtReadWrite;
                cudaMemPoolSetAccess(pool, &desc, 1);
            }
            cudaGetLastError();
        }
        return e2;
    });
    if (e != cudaSuccess) return e;
    if ((e = cudaMallocAsync(p, bytes, st)) != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}
static void mon_dev_free(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }
// Growing the pool (physical allocation + mapping) is the slow part of an allocation and holds the driver: reserve room for the
// first objects when a dataset is created, so that objects created while the frontend streams keyframes allocate from memory the
// pool already owns (MON_POOL_RESERVE_MB, default 2048: about 20 objects of base.json at ~80 MB each with their render / mesh scratch
// and the keyframe slabs; 512 was outgrown by four objects + 30 keyframes once the objects carried two batch sets and both graph
// variants — the frontend's keyframe ingest then stalled 20-100 ms behind a pool growth, profiles/r9_facade_runs.txt).
static void mon_pool_reserve(cudaStream_t st) {
    static std::atomic<uint64_t> reserved{0};
    mon_once_per_device(reserved, [st] {
        const char* env = getenv("MON_POOL_RESERVE_MB");
        const size_t mb = env ? (size_t)atol(env) : 2048;
        void* p = nullptr;
        if (mb && mon_dev_malloc(&p, mb << 20, st) == cudaSuccess) { mon_dev_free(p, st); cudaStreamSynchronize(st); }
        cudaGetLastError();     // a failed reservation is not an error: allocations then grow the pool as they come
        return cudaSuccess;
    });
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(MON_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- pcg32 on the host (TCNN dependencies/pcg32/pcg32.h:46-170), used for the A12 MLP init
namespace {
struct HostPcg32 {
    uint64_t state, inc;
    explicit HostPcg32(uint64_t initstate, uint64_t initseq = 1u) {
        state = 0U;
        inc = (initseq << 1u) | 1u;
        next_uint();
        state += initstate;
        next_uint();
    }
    uint32_t next_uint() {
        const uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    float next_float() {
        union { uint32_t u; float f; } x;
        x.u = (next_uint() >> 9) | 0x3f800000u;
        return x.f - 1.0f;
    }
    void advance(uint64_t delta) {
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

uint32_t next_multiple(uint32_t v, uint32_t d) { return ((v + d - 1) / d) * d; }

bool host_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// per-level table geometry, GridEncodingTemplated ctor (TCNN encodings/grid.h:964-997) and
// grid_scale/grid_resolution (:195-204)
bool make_grid(const mon_config& c, MonGrid& g, std::string& why) {
    if (c.n_levels == 0 || c.n_levels > MON_MAX_LEVELS) { why = "n_levels must be in 1..16"; return false; }
    if (c.n_features_per_level != 2) { why = "only n_features_per_level == 2 is supported"; return false; }
    // a (level, feature) slice of the table is staged into 128 KB of shared memory by the encode kernel, and a slice of
    // 32768 f16x2 accumulators by the scatter + Adam kernel: 2^16 entries per level is the largest table they hold
    if (c.log2_hashmap_size < 3 || c.log2_hashmap_size > 16) { why = "log2_hashmap_size must be in 3..16 (table slices are shared-memory resident)"; return false; }
    memset(&g, 0, sizeof(g));
    g.n_levels = c.n_levels;
    const float log2_pls = std::log2(c.per_level_scale);
    uint32_t offset = 0;
    for (uint32_t i = 0; i < c.n_levels; ++i) {
        const float scale = exp2f((float)i * log2_pls) * (float)c.base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1;
        const uint32_t max_params = UINT32_MAX / 2;
        uint32_t n = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
        n = next_multiple(n, 8u);
        n = std::min(n, 1u << c.log2_hashmap_size);
        // grid_index() (grid.h:170-187) decides dense vs hash from a 32-bit running stride.  The wrap-around
        // is part of the reference's behaviour: at res == 65536 (level 12 of base.json) the stride becomes
        // 2^32 == 0, the hash is NOT taken and the index degenerates to (x + y*65536) % 65536 == x % 65536.
        // The kernels' dense formula x + y*res + z*res*res wraps the same way in uint32.
        uint32_t stride = 1;
        for (int d = 0; d < 3 && stride <= n; ++d) stride *= res;
        g.offset[i] = offset;
        g.size[i] = n;
        g.res[i] = res;
        g.scale[i] = scale;
        g.hashed[i] = (n < stride) ? 1u : 0u;
        offset += n;
    }
    g.offset[c.n_levels] = offset;
    // entry -> level shortcut for the optimizer: all levels from first_full on have the capped (power-of-two) size
    const uint32_t cap = 1u << c.log2_hashmap_size;
    g.first_full = c.n_levels;
    for (uint32_t i = c.n_levels; i-- > 0;) { if (g.size[i] == cap) g.first_full = i; else break; }
    g.log2_cap = c.log2_hashmap_size;
    return true;
}

uint32_t n_mlp_params(const mon_config& c) {
    const uint32_t in_w = c.n_levels * c.n_features_per_level;
    return c.n_neurons * in_w + (c.n_hidden_layers - 1) * c.n_neurons * c.n_neurons + MON_OUT * c.n_neurons;
}

bool validate_config(const mon_config& c, std::string& why) {
    if (c.n_levels * c.n_features_per_level != MON_IN) { why = "encoding width must be 32 (n_levels*n_features_per_level)"; return false; }
    if (c.n_neurons != MON_WIDTH) { why = "only n_neurons == 64 is supported"; return false; }
    if (c.n_hidden_layers < 1 || c.n_hidden_layers > 2) { why = "n_hidden_layers must be 1 or 2"; return false; }
    if (c.samples_per_ray != MON_S) { why = "samples_per_ray must be 32"; return false; }
    if (c.rays_per_batch == 0 || c.rays_per_batch % 4 != 0) { why = "rays_per_batch must be a positive multiple of 4"; return false; }
    if (c.render_samples_per_ray == 0 || c.render_samples_per_ray % 32 != 0) { why = "render_samples_per_ray must be a multiple of 32"; return false; }
    return true;
}
}  // namespace

// =========================================================================================
struct mon_dataset {
    int gpu = 0;
    float K[4];
    int H = 0, W = 0;
    uint32_t max_frames = 0;
    int use_depth = 0;
    uint32_t depth_bytes = 4;      // bytes per depth sample on the device and in the caller's blocks: 4 (f32 metres) or 2 (raw u16, mon_dataset_set_depth_u16)
    float depth_factor = 0.0f;     // u16 mode: metres per count (DepthMapFactor)
    uint32_t n_frames = 0;
    MonFrame* d_frames = nullptr;
    std::vector<MonFrame> h_frames;
    MonFrame* h_frames_pinned = nullptr;   // page-locked mirror the frame-table uploads read (a pageable source is staged by the driver)
    cudaStream_t stream = nullptr;
    // pinned staging for pageable caller buffers (rgb | instance | depth), double-buffered: a frame is copied into one
    // half while the previous frame's DMA still reads the other
    uint8_t* staging[2] = {nullptr, nullptr};
    cudaEvent_t ev_staged[2] = {nullptr, nullptr};
    uint32_t stage_next = 0;
    cudaEvent_t ev_uploaded = nullptr;   // recorded after the last frame upload; training streams wait on it
    // keyframe storage comes in slabs of MON_FRAMES_PER_SLAB frames (one cudaMalloc per slab: a per-frame cudaMalloc
    // serialises with the training graphs in flight and made the online ingest take tens of milliseconds per keyframe)
    std::vector<uint8_t*> slabs;
    mutable std::mutex mu;
};

struct mon_object {
    mon_dataset* ds = nullptr;
    mon_config cfg;
    MonGrid grid;
    MonScene scene;
    MonOpt opt;
    MonLossCfg lc;
    uint32_t seed = 1337;
    uint32_t n_mlp = 0, n_grid = 0, P = 0, R = 0, N = 0;
    // parameters + optimizer state
    float *pf = nullptr, *m = nullptr, *v = nullptr;
    __half *ph = nullptr, *gh = nullptr, *ema = nullptr;
    uint16_t* ps = nullptr;       // per-parameter Adam step counters, 16 bits, saturating (optim_math.cuh)
    // control
    MonCtrl* ctrl_state = nullptr; // persistent counters (iter, step, n_boxes), advanced by the batch kernels
    MonCtrl* ctrl = nullptr;       // control block of the current iteration, written by its batch kernel
    MonCtrl* ctrl_late = nullptr;  // per-iteration copy for scatter / optimizer; carries the logged loss
    MonCtrl* h_ctrl = nullptr;  // pinned
    mon_bbox2d* d_boxes = nullptr;
    uint32_t box_cap = 0;
    std::vector<mon_bbox2d> h_boxes;
    // batch buffers
    MonRay* rays = nullptr;
    uint8_t* ray_inst = nullptr;
    float *target = nullptr, *target_depth = nullptr, *bg = nullptr;
    float *rgb_rays = nullptr, *depth_rays = nullptr, *mask_rays = nullptr, *loss = nullptr;
    float* pts = nullptr;             // [N][3] unit-cube sample positions (the reference's PointsInput)
    __half *enc = nullptr, *d_enc = nullptr;   // enc: level-major pairs [16][N][2]; d_enc: point-major [N][32], parity hook only (lazily allocated)
    __half* ph_planar = nullptr;      // fp16 grid weights, per level [feature 0 | feature 1], kept current by the optimizer
    float* partials = nullptr;
    float* debias_lut = nullptr;      // Adam bias correction by step count, MON_DEBIAS_LUT entries
    uint32_t n_ctas = 0;
    // parity hooks (lazily allocated)
    float *dbg_out = nullptr, *dbg_dout = nullptr, *inj_xy = nullptr, *inj_col = nullptr, *inj_dt = nullptr, *grad_snap = nullptr;
    bool have_injected = false;
    // render workspace (lazily allocated)
    MonRay* r_rays = nullptr; uint32_t* r_orig = nullptr; uint32_t* r_nhit = nullptr; __half* r_enc = nullptr; float* r_jit = nullptr; float* r_pts = nullptr;
    __half* r_planar = nullptr;       // planar copy of the EMA weights, refreshed per render call
    float *r_rgb = nullptr, *r_depth = nullptr, *r_mask = nullptr, *r_Twc = nullptr;
    uint32_t r_cap_rays = 0, r_tile = 0; size_t r_jit_cap = 0;
    uint32_t render_count = 0;
    // grow-only scratch of the inference entry points (density lattice, point queries): a cudaMalloc / cudaFree pair per
    // call synchronises the whole device — it stalled every other object training on the GPU and the keyframe ingest
    void* scr[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scr_cap[4] = {0, 0, 0, 0};
    // execution
    cudaStream_t stream = nullptr, aux = nullptr;   // aux: next iteration's batch + sample points, forked inside the graphs
    cudaEvent_t ev_fork_m = nullptr, ev_join = nullptr;
    // compacted live samples of the iteration (fused MLP kernel -> scatter + Adam kernel)
    float* pts_c = nullptr; uint32_t* genc = nullptr; uint32_t* live_cnt = nullptr;
    // gradient scatter: iterations with >= resident_min_live live samples (a fresh object: all of them) go through the
    // shared-memory resident scatter (kernels_scatter_smem.cu) into the class-planar table gcls, the others through global
    // f16x2 reductions into gh; decided on the device per iteration, the optimizer sweep reads the table that was filled
    __half* gcls = nullptr;
    uint32_t resident_min_live = 0xffffffffu;
    bool scatter_unified = true;      // false: a configuration kernels_scatter_smem.cu does not cover (stand-alone global-reduction kernel)
    // Steady state (few live samples): graphs WITHOUT a scatter kernel — the fused MLP kernel issues the f16x2 reductions itself
    // (kernels_mlp_tc.cu, FUSE).  Chosen per call on the host from the live-sample counts the previous calls left in h_live (an
    // 8-byte asynchronous read-back at the end of every call; a hint, never waited for).  fuse_mode: 0 = by the live count,
    // 1 = always, -1 = never (MON_SCATTER_FUSED: A/B measurements and tests).
    int fuse_mode = 0;
    uint32_t fuse_below = MON_FUSE_BELOW;
    bool fuse_supported = false;
    uint32_t* h_live = nullptr;       // pinned [2]; 0xffffffff until the first read-back has landed
    // opt-in occupancy grid (mon_object_set_occupancy; off: occ_res == 0).  The density grid is the running maximum (decayed) of
    // the network's density on the cell corners; a cell is occupied while that exceeds the threshold
    uint32_t occ_res = 0, occ_warmup = 0, occ_interval = 0;
    float occ_sigma_thresh = 0.0f, occ_decay = 0.95f;
    uint32_t* occ_bits = nullptr; float* occ_density = nullptr;
    uint32_t *occ_ray_mask = nullptr, *occ_list = nullptr, *occ_count = nullptr;
    uint64_t iters_enqueued = 0, occ_last_update = 0;
    bool occ_started = false;
    // instantiated iteration graphs by length
    struct GraphSlot { uint32_t iters; bool fused; cudaGraphExec_t exec; uint64_t stamp; };
    std::vector<GraphSlot> graphs;
    uint64_t graph_clock = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing_pending = false;
    float last_ms = 0.0f;
    uint64_t launches = 0;
    int sm_count = 148;
};

// =========================================================================================
extern "C" {

const char* mon_last_error(void) { return g_err.c_str(); }
const char* mon_version(void) { return "mon-b200 0.1 (sm_100a)"; }

int mon_device_count(int* count) {
    if (!count) return fail(MON_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return fail(MON_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = n;
    return MON_OK;
}

int mon_config_default(mon_config* c) {
    if (!c) return fail(MON_ERR_ARG, "cfg is NULL");
    memset(c, 0, sizeof(*c));
    c->n_levels = 16; c->n_features_per_level = 2; c->log2_hashmap_size = 16; c->base_resolution = 16;
    c->per_level_scale = 2.0f;
    c->n_neurons = 64; c->n_hidden_layers = 1;
    c->learning_rate = 1e-2f; c->beta1 = 0.9f; c->beta2 = 0.99f; c->epsilon = 1e-15f; c->l2_reg = 1e-6f;
    c->ema_decay = 0.95f; c->decay_start = 20000; c->decay_interval = 10000; c->decay_base = 0.33f;
    c->loss_scale = 128.0f;
    c->rays_per_batch = 4096; c->samples_per_ray = 32; c->render_samples_per_ray = 64;
    c->depth_lambda = 0.5f; c->mask_lambda = 0.5f; c->bg_density_reg = 0.01f;
    return MON_OK;
}

int mon_config_from_json(const char* path, mon_config* c) {
    if (!path || !c) return fail(MON_ERR_ARG, "NULL argument");
    std::ifstream f(path);
    if (!f) return fail(MON_ERR_IO, "cannot open network config '%s'", path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    monjson::Value root;
    std::string err;
    monjson::Parser p(text);
    if (!p.parse(root, err)) return fail(MON_ERR_IO, "JSON error in '%s': %s", path, err.c_str());
    mon_config_default(c);
    // defaults of the keys follow ResetNetwork (nerf_model.cu:1299-1302) and tcnn's own
    // defaults (grid.h:1419-1440; adam.h / ema.h / exponential_decay.h update_hyperparams)
    if (const monjson::Value* e = root.find("encoding")) {
        const std::string otype = e->string_or("otype", "HashGrid");
        if (otype != "HashGrid" && otype != "Grid") return fail(MON_ERR_IO, "encoding.otype '%s' is not supported (HashGrid only)", otype.c_str());
        c->n_levels = (uint32_t)e->number_or("n_levels", 16);
        c->n_features_per_level = (uint32_t)e->number_or("n_features_per_level", 2);
        c->log2_hashmap_size = (uint32_t)e->number_or("log2_hashmap_size", 19);
        c->base_resolution = (uint32_t)e->number_or("base_resolution", 16);
        c->per_level_scale = (float)e->number_or("per_level_scale", 2.0);
    }
    if (const monjson::Value* n = root.find("network")) {
        const std::string otype = n->string_or("otype", "FullyFusedMLP");
        if (otype != "FullyFusedMLP" && otype != "MegakernelMLP") return fail(MON_ERR_IO, "network.otype '%s' is not supported", otype.c_str());
        const std::string act = n->string_or("activation", "ReLU");
        if (act != "ReLU") return fail(MON_ERR_IO, "network.activation '%s' is not supported (ReLU only)", act.c_str());
        const std::string oact = n->string_or("output_activation", "None");
        if (oact != "None") return fail(MON_ERR_IO, "network.output_activation '%s' is not supported (None only)", oact.c_str());
        c->n_neurons = (uint32_t)n->number_or("n_neurons", 128);
        c->n_hidden_layers = (uint32_t)n->number_or("n_hidden_layers", 5);
    }
    // optimizer chain: walk "nested" and pick the keys by otype
    const monjson::Value* o = root.find("optimizer");
    bool saw_adam = false;
    c->decay_start = 0xffffffffu;  // no ExponentialDecay in the chain -> never decays
    while (o) {
        const std::string otype = o->string_or("otype", "");
        if (otype == "Ema") {
            c->ema_decay = (float)o->number_or("decay", 0.99);
        } else if (otype == "ExponentialDecay") {
            c->decay_start = (uint32_t)o->number_or("decay_start", 0);
            c->decay_interval = (uint32_t)o->number_or("decay_interval", 1);
            c->decay_base = (float)o->number_or("decay_base", 0.33);
        } else if (otype == "Adam") {
            saw_adam = true;
            c->learning_rate = (float)o->number_or("learning_rate", 1e-3);
            c->beta1 = (float)o->number_or("beta1", 0.9);
            c->beta2 = (float)o->number_or("beta2", 0.999);
            c->epsilon = (float)o->number_or("epsilon", 1e-8);
            c->l2_reg = (float)o->number_or("l2_reg", 1e-8);
        } else {
            return fail(MON_ERR_IO, "optimizer.otype '%s' is not supported (Ema / ExponentialDecay / Adam)", otype.c_str());
        }
        o = o->find("nested");
    }
    if (!saw_adam) return fail(MON_ERR_IO, "optimizer chain has no Adam");
    if (c->decay_interval == 0) c->decay_interval = 1;
    std::string why;
    MonGrid g;
    if (!validate_config(*c, why) || !make_grid(*c, g, why)) return fail(MON_ERR_IO, "unsupported network config: %s", why.c_str());
    return MON_OK;
}

int mon_config_param_counts(const mon_config* c, uint32_t* n_mlp, uint32_t* n_grid) {
    if (!c) return fail(MON_ERR_ARG, "cfg is NULL");
    MonGrid g; std::string why;
    if (!validate_config(*c, why) || !make_grid(*c, g, why)) return fail(MON_ERR_ARG, "%s", why.c_str());
    if (n_mlp) *n_mlp = n_mlp_params(*c);
    if (n_grid) *n_grid = g.offset[c->n_levels] * 2;
    return MON_OK;
}

int mon_config_grid_layout(const mon_config* c, uint32_t* offsets, float* scales, uint32_t* resolutions) {
    if (!c) return fail(MON_ERR_ARG, "cfg is NULL");
    MonGrid g; std::string why;
    if (!make_grid(*c, g, why)) return fail(MON_ERR_ARG, "%s", why.c_str());
    for (uint32_t i = 0; i <= c->n_levels; ++i) if (offsets) offsets[i] = g.offset[i];
    for (uint32_t i = 0; i < c->n_levels; ++i) { if (scales) scales[i] = g.scale[i]; if (resolutions) resolutions[i] = g.res[i]; }
    return MON_OK;
}

// ------------------------------------------------------------------------------- dataset
int mon_dataset_create(int gpu, float fx, float fy, float cx, float cy, int H, int W,
                       uint32_t max_frames, int use_depth, mon_dataset** out) {
    if (!out) return fail(MON_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (H <= 0 || W <= 0 || max_frames == 0) return fail(MON_ERR_ARG, "invalid image size or max_frames");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(MON_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    if (gpu < 0 || gpu >= n) return fail(MON_ERR_ARG, "gpu %d out of range (0..%d)", gpu, n - 1);
    CK(cudaSetDevice(gpu));
    mon_dataset* ds = new mon_dataset();
    ds->gpu = gpu; ds->K[0] = fx; ds->K[1] = fy; ds->K[2] = cx; ds->K[3] = cy;
    ds->H = H; ds->W = W; ds->max_frames = max_frames; ds->use_depth = use_depth;
    ds->h_frames.assign(max_frames, MonFrame{nullptr, nullptr, nullptr, {0}, 0u, 0.0f});
    ds->slabs.assign((max_frames + MON_FRAMES_PER_SLAB - 1) / MON_FRAMES_PER_SLAB, nullptr);
    const size_t px = (size_t)H * W;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&ds->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&ds->d_frames), sizeof(MonFrame) * max_frames, ds->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(ds->d_frames, 0, sizeof(MonFrame) * max_frames, ds->stream)) != cudaSuccess ||
        (e = cudaMallocHost(&ds->h_frames_pinned, sizeof(MonFrame) * max_frames)) != cudaSuccess ||
        (e = cudaMallocHost(&ds->staging[0], px * 3 + px + px * 4)) != cudaSuccess ||
        (e = cudaMallocHost(&ds->staging[1], px * 3 + px + px * 4)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ds->ev_staged[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ds->ev_staged[1], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ds->ev_uploaded, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventRecord(ds->ev_uploaded, ds->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ds->stream)) != cudaSuccess) {
        mon_dataset_destroy(ds);
        return fail(MON_ERR_CUDA, "dataset allocation: %s", cudaGetErrorString(e));
    }
    mon_pool_reserve(ds->stream);
    *out = ds;
    return MON_OK;
}

// device storage of one keyframe, carved out of its slab (allocated on first use).  A slab is PLANE-MAJOR — [rgb of its frames |
// instance of its frames | depth of its frames] — so that consecutive frames' planes are contiguous and a block of frames
// arrives in three copies (mon_dataset_add_frames) instead of three per frame.
struct SlabLayout { size_t s_rgb, s_inst, s_depth; };       // per-frame strides of the three regions (256-byte aligned)
static SlabLayout slab_layout(const mon_dataset* ds) {
    const size_t px = (size_t)ds->H * ds->W;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return {up(px * 3), up(px), ds->use_depth ? up(px * ds->depth_bytes) : 0};
}
static int ensure_frame_storage(mon_dataset* ds, uint32_t frame_id) {
    MonFrame& f = ds->h_frames[frame_id];
    if (f.rgb) return MON_OK;
    const SlabLayout L = slab_layout(ds);
    const uint32_t slab = frame_id / MON_FRAMES_PER_SLAB;
    const uint32_t n = std::min<uint32_t>(MON_FRAMES_PER_SLAB, ds->max_frames - slab * MON_FRAMES_PER_SLAB);
    if (!ds->slabs[slab]) CK(mon_dev_malloc(reinterpret_cast<void**>(&ds->slabs[slab]), (L.s_rgb + L.s_inst + L.s_depth) * n, ds->stream));
    uint8_t* base = ds->slabs[slab];
    const uint32_t k = frame_id % MON_FRAMES_PER_SLAB;
    f.rgb = base + L.s_rgb * k;
    f.instance = base + L.s_rgb * n + L.s_inst * k;
    f.depth = ds->use_depth ? reinterpret_cast<const float*>(base + (L.s_rgb + L.s_inst) * n + L.s_depth * k) : nullptr;
    return MON_OK;
}
// frame-table rows [first, first + n) -> device, from the page-locked mirror (asynchronous)
static cudaError_t upload_frame_rows(mon_dataset* ds, uint32_t first, uint32_t n) {
    memcpy(ds->h_frames_pinned + first, ds->h_frames.data() + first, sizeof(MonFrame) * n);
    return cudaMemcpyAsync(ds->d_frames + first, ds->h_frames_pinned + first, sizeof(MonFrame) * n, cudaMemcpyHostToDevice, ds->stream);
}

// the depth plane of the public entries: f32 metres, or raw u16 samples after mon_dataset_set_depth_u16
static int check_depth_format(const mon_dataset* ds, uint32_t bytes_per_sample) {
    if (ds && ds->use_depth && ds->depth_bytes != bytes_per_sample)
        return fail(MON_ERR_ARG, bytes_per_sample == 4 ? "the dataset takes raw 16-bit depth (mon_dataset_set_depth_u16): use the _d16 entry"
                                                       : "the dataset takes f32 depth: call mon_dataset_set_depth_u16 before the first frame");
    return MON_OK;
}

int mon_dataset_set_depth_u16(mon_dataset* ds, float depth_factor) {
    if (!ds) return fail(MON_ERR_ARG, "ds is NULL");
    if (!(depth_factor > 0.0f)) return fail(MON_ERR_ARG, "depth_factor must be positive");
    std::lock_guard<std::mutex> lock(ds->mu);
    if (!ds->use_depth) return fail(MON_ERR_ARG, "dataset was created without depth");
    for (const uint8_t* sl : ds->slabs) if (sl) return fail(MON_ERR_STATE, "the depth format is fixed once a keyframe has been added");
    ds->depth_bytes = 2;
    ds->depth_factor = depth_factor;
    return MON_OK;
}

static int add_frame_impl(mon_dataset* ds, uint32_t frame_id, const uint8_t* rgb, int is_bgr,
                          const uint8_t* instance, const void* depth, const float pose[16]) {
    if (!ds || !rgb || !instance || !pose) return fail(MON_ERR_ARG, "NULL argument");
    if (frame_id >= ds->max_frames) return fail(MON_ERR_ARG, "frame_id %u >= max_frames %u", frame_id, ds->max_frames);
    if (ds->use_depth && !depth) return fail(MON_ERR_ARG, "dataset was created with use_depth but depth is NULL");
    // MON_INGEST_TRACE=1: where a slow call spent its time (lock, storage, staging-slot wait, host copies, enqueue), to stderr
    static const bool trace = getenv("MON_INGEST_TRACE") != nullptr;
    using tclock = std::chrono::steady_clock;
    tclock::time_point t_[6];
    int nt_ = 0;
    auto stamp = [&] { if (trace && nt_ < 6) t_[nt_++] = tclock::now(); };
    stamp();
    std::lock_guard<std::mutex> lock(ds->mu);
    stamp();
    CK(cudaSetDevice(ds->gpu));
    const size_t px = (size_t)ds->H * ds->W;
    int rc = ensure_frame_storage(ds, frame_id);
    if (rc != MON_OK) return rc;
    stamp();
    MonFrame& f = ds->h_frames[frame_id];
    memcpy(f.pose, pose, sizeof(float) * 16);
    f.bgr = is_bgr ? 1u : 0u;
    f.depth_factor = ds->depth_bytes == 2 ? ds->depth_factor : 0.0f;
    if (host_pinned(rgb) && host_pinned(instance) && (!ds->use_depth || host_pinned(depth))) {
        // page-locked caller buffers: DMA straight out of them, no staging copy and no host synchronisation.  The
        // buffers must stay valid until mon_dataset_sync() or the next blocking call on an object of this dataset.
        CK(cudaMemcpyAsync(const_cast<uint8_t*>(f.rgb), rgb, px * 3, cudaMemcpyHostToDevice, ds->stream));
        CK(cudaMemcpyAsync(const_cast<uint8_t*>(f.instance), instance, px, cudaMemcpyHostToDevice, ds->stream));
        if (ds->use_depth) CK(cudaMemcpyAsync(const_cast<float*>(f.depth), depth, px * ds->depth_bytes, cudaMemcpyHostToDevice, ds->stream));
        CK(upload_frame_rows(ds, frame_id, 1));
        CK(cudaEventRecord(ds->ev_uploaded, ds->stream));
    } else {
        // pageable caller buffers (cv::Mat of the SLAM frontend): one memcpy into the free half of the pinned staging
        // area, DMA from there; the call returns without waiting for the DMA
        const uint32_t half = ds->stage_next;
        ds->stage_next ^= 1u;
        CK(cudaEventSynchronize(ds->ev_staged[half]));   // the DMA that last read this half (two frames ago) is done
        stamp();
        uint8_t* st_rgb = ds->staging[half];
        uint8_t* st_inst = ds->staging[half] + px * 3;
        float* st_depth = reinterpret_cast<float*>(ds->staging[half] + px * 4);
        memcpy(st_rgb, rgb, px * 3);
        memcpy(st_inst, instance, px);
        if (ds->use_depth) memcpy(st_depth, depth, px * ds->depth_bytes);
        stamp();
        CK(cudaMemcpyAsync(const_cast<uint8_t*>(f.rgb), st_rgb, px * 3, cudaMemcpyHostToDevice, ds->stream));
        CK(cudaMemcpyAsync(const_cast<uint8_t*>(f.instance), st_inst, px, cudaMemcpyHostToDevice, ds->stream));
        if (ds->use_depth) CK(cudaMemcpyAsync(const_cast<float*>(f.depth), st_depth, px * ds->depth_bytes, cudaMemcpyHostToDevice, ds->stream));
        CK(upload_frame_rows(ds, frame_id, 1));
        CK(cudaEventRecord(ds->ev_staged[half], ds->stream));
        CK(cudaEventRecord(ds->ev_uploaded, ds->stream));
    }
    stamp();
    if (trace && nt_ >= 2) {
        auto ms = [&](int a, int b) { return std::chrono::duration<double, std::milli>(t_[b] - t_[a]).count(); };
        if (ms(0, nt_ - 1) > 2.0) {
            fprintf(stderr, "[mon ingest] frame %u: %.2f ms =", frame_id, ms(0, nt_ - 1));
            static const char* names6[] = {"lock", "storage", "slot-wait", "host-copy", "enqueue"};
            static const char* names4[] = {"lock", "storage", "enqueue"};
            for (int k = 0; k + 1 < nt_; ++k) fprintf(stderr, " %s %.2f", nt_ == 6 ? names6[k] : names4[k], ms(k, k + 1));
            fprintf(stderr, "\n");
        }
    }
    ds->n_frames = std::max(ds->n_frames, frame_id + 1);
    return MON_OK;
}

int mon_dataset_add_frame(mon_dataset* ds, uint32_t frame_id, const uint8_t* rgb, int is_bgr,
                          const uint8_t* instance, const float* depth, const float pose[16]) {
    const int rc = check_depth_format(ds, 4);
    return rc != MON_OK ? rc : add_frame_impl(ds, frame_id, rgb, is_bgr, instance, depth, pose);
}
int mon_dataset_add_frame_d16(mon_dataset* ds, uint32_t frame_id, const uint8_t* rgb, int is_bgr,
                              const uint8_t* instance, const uint16_t* depth16, const float pose[16]) {
    const int rc = check_depth_format(ds, 2);
    return rc != MON_OK ? rc : add_frame_impl(ds, frame_id, rgb, is_bgr, instance, depth16, pose);
}

static int add_frame_device_impl(mon_dataset* ds, uint32_t frame_id, const uint8_t* d_rgb, int is_bgr, const uint8_t* d_instance,
                                 const void* d_depth, const float pose[16]) {
    if (!ds || !d_rgb || !d_instance || !pose) return fail(MON_ERR_ARG, "NULL argument");
    if (frame_id >= ds->max_frames) return fail(MON_ERR_ARG, "frame_id %u >= max_frames %u", frame_id, ds->max_frames);
    if (ds->use_depth && !d_depth) return fail(MON_ERR_ARG, "dataset was created with use_depth but depth is NULL");
    std::lock_guard<std::mutex> lock(ds->mu);
    CK(cudaSetDevice(ds->gpu));
    const size_t px = (size_t)ds->H * ds->W;
    int rc = ensure_frame_storage(ds, frame_id);
    if (rc != MON_OK) return rc;
    MonFrame& f = ds->h_frames[frame_id];
    memcpy(f.pose, pose, sizeof(float) * 16);
    f.bgr = is_bgr ? 1u : 0u;
    f.depth_factor = ds->depth_bytes == 2 ? ds->depth_factor : 0.0f;
    CK(cudaMemcpyAsync(const_cast<uint8_t*>(f.rgb), d_rgb, px * 3, cudaMemcpyDeviceToDevice, ds->stream));
    CK(cudaMemcpyAsync(const_cast<uint8_t*>(f.instance), d_instance, px, cudaMemcpyDeviceToDevice, ds->stream));
    if (ds->use_depth) CK(cudaMemcpyAsync(const_cast<float*>(f.depth), d_depth, px * ds->depth_bytes, cudaMemcpyDeviceToDevice, ds->stream));
    CK(upload_frame_rows(ds, frame_id, 1));
    CK(cudaEventRecord(ds->ev_uploaded, ds->stream));
    ds->n_frames = std::max(ds->n_frames, frame_id + 1);
    return MON_OK;
}
int mon_dataset_add_frame_device(mon_dataset* ds, uint32_t frame_id, const uint8_t* d_rgb, int is_bgr, const uint8_t* d_instance,
                                 const float* d_depth, const float pose[16]) {
    const int rc = check_depth_format(ds, 4);
    return rc != MON_OK ? rc : add_frame_device_impl(ds, frame_id, d_rgb, is_bgr, d_instance, d_depth, pose);
}
int mon_dataset_add_frame_device_d16(mon_dataset* ds, uint32_t frame_id, const uint8_t* d_rgb, int is_bgr, const uint8_t* d_instance,
                                     const uint16_t* d_depth16, const float pose[16]) {
    const int rc = check_depth_format(ds, 2);
    return rc != MON_OK ? rc : add_frame_device_impl(ds, frame_id, d_rgb, is_bgr, d_instance, d_depth16, pose);
}

static int add_frames_impl(mon_dataset* ds, uint32_t first_id, uint32_t n, const uint8_t* rgb, int is_bgr, const uint8_t* instance,
                           const void* depth_v, const float* poses16, int on_device) {
    const uint8_t* depth = static_cast<const uint8_t*>(depth_v);
    if (!ds || !rgb || !instance || !poses16) return fail(MON_ERR_ARG, "NULL argument");
    if ((uint64_t)first_id + n > ds->max_frames) return fail(MON_ERR_ARG, "frames %u..%u exceed max_frames %u", first_id, first_id + n, ds->max_frames);
    if (ds->use_depth && !depth) return fail(MON_ERR_ARG, "dataset was created with use_depth but depth is NULL");
    if (n == 0) return MON_OK;
    const size_t px = (size_t)ds->H * ds->W, dpx = px * ds->depth_bytes;
    if (!on_device && !(host_pinned(rgb) && host_pinned(instance) && (!ds->use_depth || host_pinned(depth)))) {
        // pageable blocks: frame by frame through the pinned staging area
        for (uint32_t i = 0; i < n; ++i) {
            const int rc = add_frame_impl(ds, first_id + i, rgb + px * 3 * i, is_bgr, instance + px * i, depth ? depth + dpx * i : nullptr, poses16 + 16 * (size_t)i);
            if (rc != MON_OK) return rc;
        }
        return MON_OK;
    }
    std::lock_guard<std::mutex> lock(ds->mu);
    CK(cudaSetDevice(ds->gpu));
    for (uint32_t i = 0; i < n; ++i) {
        const int rc = ensure_frame_storage(ds, first_id + i);
        if (rc != MON_OK) return rc;
        MonFrame& f = ds->h_frames[first_id + i];
        memcpy(f.pose, poses16 + 16 * (size_t)i, sizeof(float) * 16);
        f.bgr = is_bgr ? 1u : 0u;
        f.depth_factor = ds->depth_bytes == 2 ? ds->depth_factor : 0.0f;
    }
    const SlabLayout L = slab_layout(ds);
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // runs of frames that share a slab: one copy per plane kind where the device stride equals the block's (image sizes whose
    // planes are multiples of 256 bytes: 800 x 800 is), else one copy per frame and plane
    for (uint32_t i = 0; i < n;) {
        const uint32_t id = first_id + i;
        const uint32_t run = std::min<uint32_t>(n - i, MON_FRAMES_PER_SLAB - id % MON_FRAMES_PER_SLAB);
        const MonFrame& f0 = ds->h_frames[id];
        const uint32_t r_rgb = L.s_rgb == px * 3 ? 1u : run, r_inst = L.s_inst == px ? 1u : run, r_depth = L.s_depth == dpx ? 1u : run;
        for (uint32_t k = 0; k < r_rgb; ++k)
            CK(cudaMemcpyAsync(const_cast<uint8_t*>(f0.rgb) + L.s_rgb * k, rgb + px * 3 * (i + k), px * 3 * (run / r_rgb), kind, ds->stream));
        for (uint32_t k = 0; k < r_inst; ++k)
            CK(cudaMemcpyAsync(const_cast<uint8_t*>(f0.instance) + L.s_inst * k, instance + px * (i + k), px * (run / r_inst), kind, ds->stream));
        if (ds->use_depth)
            for (uint32_t k = 0; k < r_depth; ++k)
                CK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(const_cast<float*>(f0.depth)) + L.s_depth * k, depth + dpx * (i + k), dpx * (run / r_depth), kind, ds->stream));
        i += run;
    }
    CK(upload_frame_rows(ds, first_id, n));
    CK(cudaEventRecord(ds->ev_uploaded, ds->stream));
    ds->n_frames = std::max(ds->n_frames, first_id + n);
    return MON_OK;
}

int mon_dataset_add_frames(mon_dataset* ds, uint32_t first_id, uint32_t n, const uint8_t* rgb, int is_bgr, const uint8_t* instance,
                           const float* depth, const float* poses16, int on_device) {
    const int rc = check_depth_format(ds, 4);
    return rc != MON_OK ? rc : add_frames_impl(ds, first_id, n, rgb, is_bgr, instance, depth, poses16, on_device);
}
int mon_dataset_add_frames_d16(mon_dataset* ds, uint32_t first_id, uint32_t n, const uint8_t* rgb, int is_bgr, const uint8_t* instance,
                               const uint16_t* depth16, const float* poses16, int on_device) {
    const int rc = check_depth_format(ds, 2);
    return rc != MON_OK ? rc : add_frames_impl(ds, first_id, n, rgb, is_bgr, instance, depth16, poses16, on_device);
}

int mon_dataset_sync(mon_dataset* ds) {
    if (!ds) return fail(MON_ERR_ARG, "ds is NULL");
    CK(cudaSetDevice(ds->gpu));
    CK(cudaStreamSynchronize(ds->stream));
    return MON_OK;
}

int mon_dataset_update_poses(mon_dataset* ds, uint32_t first_frame, uint32_t n, const float* poses16) {
    if (!ds || !poses16) return fail(MON_ERR_ARG, "NULL argument");
    if ((uint64_t)first_frame + n > ds->max_frames) return fail(MON_ERR_ARG, "pose range out of bounds");
    std::lock_guard<std::mutex> lock(ds->mu);
    CK(cudaSetDevice(ds->gpu));
    for (uint32_t i = 0; i < n; ++i) {
        MonFrame& f = ds->h_frames[first_frame + i];
        memcpy(f.pose, poses16 + (size_t)i * 16, sizeof(float) * 16);
    }
    if (n) CK(upload_frame_rows(ds, first_frame, n));
    CK(cudaStreamSynchronize(ds->stream));
    return MON_OK;
}

int mon_dataset_frame_count(const mon_dataset* ds, uint32_t* n) {
    if (!ds || !n) return fail(MON_ERR_ARG, "NULL argument");
    *n = ds->n_frames;
    return MON_OK;
}

// frames [first, end) of src -> dst, device to device (NVLink peer copies between GPUs), on dst's upload stream, ordered
// behind src's own uploads.  Both dataset mutexes are held by the caller.
static int copy_frames_from_peer(mon_dataset* dst, const mon_dataset* src, uint32_t first, uint32_t end) {
    CK(cudaSetDevice(dst->gpu));
    if (dst->gpu != src->gpu) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, dst->gpu, src->gpu));
        if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(src->gpu, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); cudaGetLastError(); }
    }
    // the source frames may still be in flight on src's upload stream (asynchronous DMA from pinned or staged buffers)
    CK(cudaStreamWaitEvent(dst->stream, src->ev_uploaded, 0));
    const size_t px = (size_t)dst->H * dst->W;
    for (uint32_t i = first; i < end; ++i) {
        const MonFrame& s = src->h_frames[i];
        if (!s.rgb) continue;
        int rc = ensure_frame_storage(dst, i);
        if (rc != MON_OK) return rc;
        MonFrame& f = dst->h_frames[i];
        f.bgr = s.bgr;
        f.depth_factor = s.depth_factor;
        CK(cudaMemcpyPeerAsync(const_cast<uint8_t*>(f.rgb), dst->gpu, s.rgb, src->gpu, px * 3, dst->stream));
        CK(cudaMemcpyPeerAsync(const_cast<uint8_t*>(f.instance), dst->gpu, s.instance, src->gpu, px, dst->stream));
        if (dst->use_depth) CK(cudaMemcpyPeerAsync(const_cast<float*>(f.depth), dst->gpu, s.depth, src->gpu, px * dst->depth_bytes, dst->stream));
        memcpy(f.pose, s.pose, sizeof(float) * 16);
        CK(upload_frame_rows(dst, i, 1));
        dst->n_frames = std::max(dst->n_frames, i + 1);
    }
    CK(cudaEventRecord(dst->ev_uploaded, dst->stream));
    return MON_OK;
}

static int peer_compatible(const mon_dataset* dst, const mon_dataset* src) {
    if (!dst || !src) return fail(MON_ERR_ARG, "NULL argument");
    if (dst == src) return fail(MON_ERR_ARG, "source and destination are the same dataset");
    if (dst->H != src->H || dst->W != src->W || dst->use_depth != src->use_depth || dst->depth_bytes != src->depth_bytes)
        return fail(MON_ERR_ARG, "datasets are not shape-compatible");
    return MON_OK;
}

int mon_dataset_clone_from_peer(mon_dataset* dst, const mon_dataset* src) {
    int rc = peer_compatible(dst, src);
    if (rc != MON_OK) return rc;
    std::unique_lock<std::mutex> l1(dst->mu, std::defer_lock), l2(src->mu, std::defer_lock);
    std::lock(l1, l2);
    if (dst->max_frames < src->n_frames) return fail(MON_ERR_ARG, "destination holds %u frames, source has %u", dst->max_frames, src->n_frames);
    rc = copy_frames_from_peer(dst, src, 0, src->n_frames);
    if (rc != MON_OK) return rc;
    CK(cudaStreamSynchronize(dst->stream));
    return MON_OK;
}

int mon_dataset_copy_frame_from_peer(mon_dataset* dst, const mon_dataset* src, uint32_t frame_id) {
    int rc = peer_compatible(dst, src);
    if (rc != MON_OK) return rc;
    std::unique_lock<std::mutex> l1(dst->mu, std::defer_lock), l2(src->mu, std::defer_lock);
    std::lock(l1, l2);
    if (frame_id >= src->max_frames || !src->h_frames[frame_id].rgb) return fail(MON_ERR_ARG, "frame %u is not in the source dataset", frame_id);
    if (frame_id >= dst->max_frames) return fail(MON_ERR_ARG, "frame_id %u >= max_frames %u", frame_id, dst->max_frames);
    return copy_frames_from_peer(dst, src, frame_id, frame_id + 1);   // asynchronous: training streams wait on the upload event
}

int mon_dataset_destroy(mon_dataset* ds) {
    if (!ds) return MON_OK;
    cudaSetDevice(ds->gpu);
    if (ds->stream) cudaStreamSynchronize(ds->stream);
    for (uint8_t* slab : ds->slabs) mon_dev_free(slab, ds->stream);
    mon_dev_free(ds->d_frames, ds->stream);
    if (ds->stream) cudaStreamSynchronize(ds->stream);
    if (ds->h_frames_pinned) cudaFreeHost(ds->h_frames_pinned);
    for (int k = 0; k < 2; ++k) {
        if (ds->staging[k]) cudaFreeHost(ds->staging[k]);
        if (ds->ev_staged[k]) cudaEventDestroy(ds->ev_staged[k]);
    }
    if (ds->ev_uploaded) cudaEventDestroy(ds->ev_uploaded);
    if (ds->stream) cudaStreamDestroy(ds->stream);
    delete ds;
    return MON_OK;
}

// ------------------------------------------------------------------------------- object
static void drop_graphs(mon_object* o) {
    for (auto& g : o->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    o->graphs.clear();
}

static MonBatch make_batch(mon_object* o, bool injected, bool debug, uint32_t set = 0) {
    MonBatch b;
    memset(&b, 0, sizeof(b));
    const size_t sR = (size_t)set * o->R;
    b.R = o->R;
    b.boxes = o->d_boxes;
    b.frames = o->ds->d_frames;
    b.state = o->ctrl_state;
    b.ctrl = o->ctrl + set;
    b.late = o->ctrl_late;
    b.seed = o->seed;
    b.opt_lr = o->cfg.learning_rate; b.decay_base = o->cfg.decay_base; b.ema_decay = o->cfg.ema_decay;
    b.decay_start = o->cfg.decay_start; b.decay_interval = o->cfg.decay_interval ? o->cfg.decay_interval : 1;
    if (injected) { b.inj_xy = o->inj_xy; b.inj_col = o->inj_col; b.inj_dt = o->inj_dt; }
    b.rays = o->rays + sR; b.ray_inst = o->ray_inst + sR; b.target = o->target + sR * 3; b.target_depth = o->target_depth + sR; b.bg = o->bg + sR * 3;
    b.rgb_rays = o->rgb_rays; b.depth_rays = o->depth_rays; b.mask_rays = o->mask_rays; b.loss = o->loss;
    b.enc = o->enc;
    b.pts = o->pts + (size_t)set * o->N * 3; b.pts_c = o->pts_c; b.genc = o->genc; b.live_cnt = o->live_cnt;
    b.occ = MonOcc{o->occ_bits, o->occ_res, o->occ_ray_mask, o->occ_list, o->occ_count};
    if (debug) { b.dbg_out = o->dbg_out; b.dbg_dout = o->dbg_dout; b.d_enc = o->d_enc; }
    b.params = o->ph; b.grads = o->gh; b.mlp_partials = o->partials;
    return b;
}

// One training iteration = batch (B) -> sample points (P) -> hash encode (E) -> fused MLP (M) -> gradient scatter (S) ->
// optimizer sweep (O).  B and P do not depend on the training state, so inside a captured graph the B/P of iteration i+1 run
// on a forked branch beside S/O of iteration i:
//     main:  E(i)  M(i) ---------> S(i) -----> O(i) -------join--> E(i+1) ...
//     aux :          \--> B(i+1) -> P(i+1) ----------------/
// Everything on the branch only needs M(i): the batch kernel rewrites rays/targets (last read by M, which also copied the
// control block for S/O), and the sample positions are last read by M too (it hands the positions of the live samples to S
// in compacted form).  In the opt-in fused mode S also updates the grid and O (MLP weights + logged loss only) runs on a
// second branch beside it.
// Where the next iteration's batch + sample points run inside a graph.  Graphs WITH a scatter kernel (a fresh object): behind the
// MLP kernel, beside the scatter, which leaves the batch cluster 8 SMs.  Graphs WITHOUT one (steady state, scatter fused into the
// MLP kernel): forked at the start of the iteration, beside the hash encode, in the slim shape that fits next to its CTAs — the
// optimizer sweep that follows the MLP kernel there fills every SM.  Measured both ways in both phases (profiles/r9c_timeline_*):
// beside the encode the scatter gets the whole chip (-2.3 us) but the encode loses 2.7 us.  MON_EARLY_FORK=0 / 1 forces one form.
static bool early_fork(const mon_object* o, bool fused) {
    static const int env = [] { const char* e = getenv("MON_EARLY_FORK"); return e ? atoi(e) : -1; }();
    if (o->occ_bits) return false;     // the occupancy mode's sample-points kernel writes single-buffered lists E and M read
    return env >= 0 ? env != 0 : fused;
}
static int launch_batch(mon_object* o, const MonBatch& b, cudaStream_t st, bool slim = false) {
    mon_launch_generate_batch(b, o->scene, st, MonLaunchOpt(), slim);
    return MON_OK;
}
static int launch_points(mon_object* o, const MonBatch& b, cudaStream_t st, bool pdl = true) {
    MonLaunchOpt lo; lo.pdl = pdl;
    mon_launch_sample_points(o->N, MON_S, b.rays, nullptr, b.inj_dt, o->seed, b.ctrl, 2, 0, o->scene.bmin, o->scene.bmax, const_cast<float*>(b.pts), st, lo, nullptr,
                             o->occ_bits ? &b.occ : nullptr);
    return MON_OK;
}
static int launch_encode(mon_object* o, const MonBatch& b, cudaStream_t st, bool pdl = true) {
    MonLaunchOpt lo; lo.pdl = pdl;
    cudaError_t e = mon_launch_encode_forward(o->grid, o->N, b.pts, o->ph_planar, o->enc, b.ctrl, (uint32_t)o->sm_count, st, 0, 0xffffffffu, lo,
                                              o->occ_bits ? o->occ_list : nullptr, o->occ_bits ? o->occ_count : nullptr);
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "hash encode launch: %s", cudaGetErrorString(e));
    return MON_OK;
}
static int launch_mlp(mon_object* o, const MonBatch& b, cudaStream_t st, bool fused = false) {
    cudaError_t e = mon_launch_mlp_train_tc(b, o->lc, o->cfg.n_hidden_layers, o->n_mlp, o->n_ctas, st, MonLaunchOpt(), fused ? &o->grid : nullptr, o->gh + o->n_mlp);
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "fused MLP launch: %s", cudaGetErrorString(e));
    return MON_OK;
}
// gradient scatter: one launch; the kernel takes the shared-memory resident path or the global reductions by the iteration's
// live-sample count.  Configurations the unified kernel does not cover use the stand-alone global-reduction kernel.
static int launch_scatter(mon_object* o, cudaStream_t st, bool leave_spare_sms = true) {
    if (o->scatter_unified) {
        // occupancy mode: the next iteration's batch cluster runs beside this kernel (capture_graph) and needs SMs of its own
        cudaError_t e = mon_launch_scatter(o->grid, o->N, o->resident_min_live, o->live_cnt, o->pts_c, o->genc, o->ctrl_late, o->gcls, o->gh + o->n_mlp,
                                           (uint32_t)o->sm_count, st, MonLaunchOpt(), leave_spare_sms);
        if (e != cudaSuccess) return fail(MON_ERR_CUDA, "gradient scatter launch: %s", cudaGetErrorString(e));
        return MON_OK;
    }
    mon_launch_encode_backward(o->grid, o->N, 0xffffffffu, o->live_cnt, o->pts_c, o->genc, o->ctrl_late, o->gh + o->n_mlp, st);
    return MON_OK;
}
static int scatter_launches(const mon_object*) { return 1; }
// optimizer sweep: MLP weights (fixed-order reduction of the per-CTA partials, Adam, EMA) + logged loss + the grid (Adam with
// per-parameter steps, EMA, gradient zeroing, planar weight copy)
static void launch_optimizer(mon_object* o, cudaStream_t st, bool pdl, bool fused = false) {
    MonLaunchOpt lo; lo.pdl = pdl;
    // fused: the gradients are in the entry-ordered table whatever the live count was
    mon_launch_optimizer(o->opt, o->ctrl_late, o->pf, o->ph, o->gh, o->partials, o->m, o->v, o->ps, o->ema, o->loss, o->R, o->grid,
                         o->ph_planar, st, MON_OPT_ALL, 0, 0xffffffffu, lo, o->gcls, o->live_cnt, fused ? 0xffffffffu : o->resident_min_live, (uint32_t)o->sm_count);
}
// which graph variant the next call takes: the scatter fused into the MLP kernel when the object's recent iterations had few live
// samples (the counts of the last two iterations of the previous call, if their read-back has landed)
static bool use_fused(const mon_object* o) {
    if (!o->fuse_supported || o->occ_bits || o->fuse_mode < 0) return false;
    if (o->fuse_mode > 0) return true;
    const uint32_t a = reinterpret_cast<volatile uint32_t*>(o->h_live)[0], b = reinterpret_cast<volatile uint32_t*>(o->h_live)[1];
    if (o->resident_min_live == 0u) return false;      // MON_SCATTER_RESIDENT_MIN=0 (tests): always the resident scatter kernel
    return std::max(a, b) < o->fuse_below;
}

// serial version (injected / profiled iterations).  ev (optional, MON_N_STAGES+1 events): recorded before each
// stage and after the last one.  Returns the number of kernels launched through n_launched.
static int enqueue_iteration(mon_object* o, const MonBatch& b, bool snapshot_grad, int* n_launched, cudaEvent_t* ev = nullptr, bool fused = false) {
    cudaStream_t st = o->stream;
    int n = 0, rc;
    if (ev) CK(cudaEventRecord(ev[0], st));
    launch_batch(o, b, st); ++n;
    if (ev) CK(cudaEventRecord(ev[1], st));
    launch_points(o, b, st); ++n;
    if (ev) CK(cudaEventRecord(ev[2], st));
    if ((rc = launch_encode(o, b, st)) != MON_OK) return rc;
    ++n;
    if (ev) CK(cudaEventRecord(ev[3], st));
    if ((rc = launch_mlp(o, b, st, fused)) != MON_OK) return rc;
    ++n;
    if (ev) CK(cudaEventRecord(ev[4], st));
    if (!fused) {
        if ((rc = launch_scatter(o, st)) != MON_OK) return rc;
        n += scatter_launches(o);
    }
    if (snapshot_grad) { mon_launch_snapshot_grad(o->P, o->n_mlp, o->opt.n_partials, o->gh, o->partials, o->grad_snap, st, o->grid, o->gcls); ++n; }
    if (ev) CK(cudaEventRecord(ev[5], st));
    launch_optimizer(o, st, !snapshot_grad && !o->scatter_unified, fused); ++n;
    if (ev) CK(cudaEventRecord(ev[6], st));
    CK(cudaGetLastError());
    if (n_launched) *n_launched = n;
    return MON_OK;
}

static int capture_graph(mon_object* o, int iters, bool fused, cudaGraphExec_t* out) {
    // two batch sets alternate inside a graph; every graph starts with set 0 (its first batch is generated in the open)
    const MonBatch bs[2] = {make_batch(o, false, false, 0), make_batch(o, false, false, 1)};
    cudaGraph_t g = nullptr;
    cudaStream_t st = o->stream, aux = o->aux;
    // Where the next iteration's batch + sample points run (early_fork above).  With a scatter kernel in the graph:
    //     main:  E(i) -> M(i) -> S(i) ---> O(i) --join--> E(i+1) ...            the branch forks behind M(i), the last reader of the
    //     aux :            \--> B(i+1) -> P(i+1) --/                            batch set it rewrites (set 0 throughout)
    // without one (steady state, scatter fused into M):
    //     main:  E(i) ----------> M(i) -> O(i) --join--> E(i+1) ...             forked at the START of iteration i, beside its hash
    //     aux :  B(i+1) -> P(i+1) ---------------/                              encode; the two batch sets alternate
    const bool early = early_fork(o, fused);
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = MON_OK;
    cudaError_t e = cudaSuccess;
    launch_batch(o, bs[0], st);
    launch_points(o, bs[0], st);
    auto fork_next = [&](const MonBatch& nb) {
        if ((e = cudaEventRecord(o->ev_fork_m, st)) != cudaSuccess) return;
        if ((e = cudaStreamWaitEvent(aux, o->ev_fork_m, 0)) != cudaSuccess) return;
        launch_batch(o, nb, aux, early);
        launch_points(o, nb, aux, false);
        e = cudaEventRecord(o->ev_join, aux);
    };
    for (int i = 0; i < iters && rc == MON_OK && e == cudaSuccess; ++i) {
        const MonBatch& b = bs[early ? (i & 1) : 0];
        const MonBatch& nb = bs[early ? ((i + 1) & 1) : 0];
        const bool fork = i + 1 < iters;
        if (fork && early) { fork_next(nb); if (e != cudaSuccess) break; }
        // after a fork or a join the encode kernel's predecessor in the stream is not a plain kernel node: no programmatic edge there
        if ((rc = launch_encode(o, b, st, i == 0 && !(fork && early))) != MON_OK) break;
        if ((rc = launch_mlp(o, b, st, fused)) != MON_OK) break;
        if (fork && !early) { fork_next(nb); if (e != cudaSuccess) break; }
        // fused variant (steady state): no scatter kernel.  The sweep follows the MLP kernel through a PLAIN edge as well: with a
        // programmatic one its CTAs land on SMs that still hold MLP CTAs and keep their maximum shared-memory carve-out, and its
        // streaming loads have too little L1 (19.4 instead of 13.5 us, profiles/r9b_timeline.txt)
        if (!fused && (rc = launch_scatter(o, st, !early)) != MON_OK) break;
        // no programmatic edge behind the unified scatter kernel: sweep CTAs that become resident while its 1024-thread CTAs still
        // run slowed the sweep by 3-5 us (profiles/r5i_timeline*.txt); a plain edge costs a 2 us gap (deferring the scatter's trigger
        // to its CTAs' exit changes nothing either way: profiles/r8b_timeline_late_edge.txt)
        launch_optimizer(o, st, !o->scatter_unified, fused);
        if (fork && (e = cudaStreamWaitEvent(st, o->ev_join, 0)) != cudaSuccess) break;
    }
    cudaError_t e2 = cudaStreamEndCapture(st, &g);
    if (rc != MON_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) { if (g) cudaGraphDestroy(g); return fail(MON_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e)); }
    if (e2 != cudaSuccess) return fail(MON_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e2));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    // the first launch of an executable graph also uploads it to the device (~10 us per iteration of the graph, measured as the
    // difference between a call's event time and its kernels' timeline): do that here, where the graph is prepared
    if ((e = cudaGraphUpload(*out, st)) != cudaSuccess) { cudaGraphExecDestroy(*out); *out = nullptr; return fail(MON_ERR_CUDA, "cudaGraphUpload: %s", cudaGetErrorString(e)); }
    return MON_OK;
}

// how a call of `iters` iterations is cut into graph replays: n_chunk replays of a `chunk`-iteration graph + one `rem`-iteration graph
struct GraphPlan { uint32_t chunk, n_chunk, rem; };
static GraphPlan graph_plan(uint32_t iters) {
    if (iters <= MON_GRAPH_CHUNK) return {0u, 0u, iters};
    for (uint32_t c = MON_GRAPH_CHUNK; c >= MON_GRAPH_CHUNK / 2; --c)
        if (iters % c == 0) return {c, iters / c, 0u};
    return {MON_GRAPH_CHUNK, iters / MON_GRAPH_CHUNK, iters % MON_GRAPH_CHUNK};
}

// the instantiated graph of exactly `iters` iterations (1 <= iters <= MON_GRAPH_CHUNK), captured on first use
static int graph_for(mon_object* o, uint32_t iters, bool fused, cudaGraphExec_t* out) {
    for (auto& g : o->graphs) if (g.iters == iters && g.fused == fused) { g.stamp = ++o->graph_clock; *out = g.exec; return MON_OK; }
    cudaGraphExec_t exec = nullptr;
    int rc = capture_graph(o, (int)iters, fused, &exec);
    if (rc != MON_OK) return rc;
    size_t n_rem = 0, oldest = SIZE_MAX;
    for (size_t k = 0; k < o->graphs.size(); ++k) {
        if (o->graphs[k].iters == MON_GRAPH_CHUNK) continue;     // the full-length chunk graph is never dropped
        ++n_rem;
        if (oldest == SIZE_MAX || o->graphs[k].stamp < o->graphs[oldest].stamp) oldest = k;
    }
    if (iters != MON_GRAPH_CHUNK && n_rem >= MON_GRAPH_CACHE && oldest != SIZE_MAX) {
        // replays of the dropped graph that are still in flight keep their resources until they finish
        cudaGraphExecDestroy(o->graphs[oldest].exec);
        o->graphs.erase(o->graphs.begin() + (long)oldest);
    }
    o->graphs.push_back({iters, fused, exec, ++o->graph_clock});
    *out = exec;
    return MON_OK;
}

// kernels launched by one replay of an `iters`-iteration graph
static uint64_t launches_in_graph(const mon_object* o, uint32_t iters, bool fused) { return (5ull + (fused ? 0ull : (uint64_t)scatter_launches(o))) * iters; }

int mon_object_create(mon_dataset* ds, const mon_config* cfg, uint32_t seed, uint8_t instance_id,
                      const float obj_Tow[16], const float bmin[3], const float bmax[3], mon_object** out) {
    if (!out) return fail(MON_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!ds || !cfg || !obj_Tow || !bmin || !bmax) return fail(MON_ERR_ARG, "NULL argument");
    std::string why;
    MonGrid grid;
    if (!validate_config(*cfg, why) || !make_grid(*cfg, grid, why)) return fail(MON_ERR_ARG, "unsupported config: %s", why.c_str());
    for (int k = 0; k < 3; ++k) if (!(bmax[k] > bmin[k])) return fail(MON_ERR_ARG, "empty bounding box");
    CK(cudaSetDevice(ds->gpu));
    mon_object* o = new mon_object();
    o->ds = ds; o->cfg = *cfg; o->grid = grid; o->seed = seed;
    o->n_mlp = n_mlp_params(*cfg);
    o->n_grid = grid.offset[cfg->n_levels] * 2;
    o->P = o->n_mlp + o->n_grid;
    o->R = cfg->rays_per_batch;
    o->N = o->R * MON_S;
    memcpy(o->scene.Tow, obj_Tow, 64);
    memcpy(o->scene.bmin, bmin, 12); memcpy(o->scene.bmax, bmax, 12);
    memcpy(o->scene.K, ds->K, 16);
    o->scene.H = ds->H; o->scene.W = ds->W; o->scene.instance_id = instance_id; o->scene.use_depth = ds->use_depth;
    o->lc.loss_scale = cfg->loss_scale; o->lc.depth_lambda = cfg->depth_lambda; o->lc.mask_lambda = cfg->mask_lambda;
    o->lc.bg_density_reg = cfg->bg_density_reg;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ds->gpu));
    o->sm_count = prop.multiProcessorCount;
    // fused MLP kernel: persistent grid of all co-resident CTAs (4 per SM with one hidden layer, 2 with two: TMEM
    // columns and shared memory, kernels_mlp_tc.cu), never more CTAs than 4-ray tiles
    o->n_ctas = std::min<uint32_t>((uint32_t)o->sm_count * (cfg->n_hidden_layers == 1 ? 4u : 2u), (o->R + 3) / 4);
    if (const char* env = getenv("MON_MLP_CTAS")) {   // A/B: persistent-grid size of the fused MLP kernel
        const int v = atoi(env);
        if (v > 0) o->n_ctas = std::min<uint32_t>((uint32_t)v, o->n_ctas);
    }
    if (o->n_ctas > MON_MAX_MLP_CTAS) o->n_ctas = MON_MAX_MLP_CTAS;
    o->opt.lr = cfg->learning_rate; o->opt.beta1 = cfg->beta1; o->opt.beta2 = cfg->beta2; o->opt.eps = cfg->epsilon;
    o->opt.l2_reg = cfg->l2_reg; o->opt.ema_decay = cfg->ema_decay; o->opt.loss_scale = cfg->loss_scale;
    o->opt.decay_start = cfg->decay_start; o->opt.decay_interval = cfg->decay_interval; o->opt.decay_base = cfg->decay_base;
    o->opt.n_mlp = o->n_mlp; o->opt.n_params = o->P; o->opt.n_partials = o->n_ctas;
    o->opt.log2_beta1 = std::log2(cfg->beta1); o->opt.log2_beta2 = std::log2(cfg->beta2);
    {
        int ex = 0;
        const float mant = std::frexp(cfg->loss_scale, &ex);
        o->opt.loss_scale_pow2 = (mant == 0.5f) ? 1u : 0u;
        o->opt.inv_loss_scale = 1.0f / cfg->loss_scale;
    }

    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        mon_object_destroy(o);
        return fail(MON_ERR_CUDA, "object setup: %s", cudaGetErrorString(e));
    }
#define OALLOC(ptr, bytes)                                                                                  \
    do {                                                                                                    \
        cudaError_t e_ = mon_dev_malloc(reinterpret_cast<void**>(&(ptr)), (bytes), o->stream);               \
        if (e_ == cudaSuccess) e_ = cudaMemsetAsync((ptr), 0, (bytes), o->stream);                          \
        if (e_ != cudaSuccess) { mon_object_destroy(o); return fail(MON_ERR_CUDA, "object allocation (%zu B): %s", (size_t)(bytes), cudaGetErrorString(e_)); } \
    } while (0)
    const size_t P = o->P, R = o->R, N = o->N;
    OALLOC(o->pf, P * 4); OALLOC(o->m, P * 4); OALLOC(o->v, P * 4); OALLOC(o->ps, P * 2 + 16);
    OALLOC(o->ph, P * 2 + 16); OALLOC(o->gh, P * 2 + 16); OALLOC(o->ema, P * 2 + 16);
    OALLOC(o->ctrl_state, sizeof(MonCtrl)); OALLOC(o->ctrl, 2 * sizeof(MonCtrl)); OALLOC(o->ctrl_late, sizeof(MonCtrl));
    // two batch sets (rays, targets, control block, sample positions): inside the iteration graphs the batch of iteration i + 1 is
    // generated into set (i + 1) & 1 while iteration i is encoded from set i & 1 (capture_graph); set 1 follows set 0 in each buffer
    OALLOC(o->rays, 2 * R * sizeof(MonRay)); OALLOC(o->ray_inst, 2 * R);
    OALLOC(o->target, 2 * R * 12); OALLOC(o->target_depth, 2 * R * 4); OALLOC(o->bg, 2 * R * 12);
    OALLOC(o->rgb_rays, R * 12); OALLOC(o->depth_rays, R * 4); OALLOC(o->mask_rays, R * 4); OALLOC(o->loss, R * 4);
    OALLOC(o->pts, 2 * N * 12); OALLOC(o->enc, N * MON_IN * 2);
    OALLOC(o->pts_c, N * 16); OALLOC(o->genc, N * (size_t)MON_MAX_LEVELS * 4); OALLOC(o->live_cnt, 8);
    OALLOC(o->ph_planar, (size_t)o->n_grid * 2 + 16);
    OALLOC(o->gcls, (size_t)o->n_grid * 2 + 16);
    OALLOC(o->partials, (size_t)o->n_ctas * o->n_mlp * 4);
    OALLOC(o->debias_lut, (size_t)MON_DEBIAS_LUT * 4);
    OALLOC(o->d_boxes, sizeof(mon_bbox2d) * MON_BOX_CAP0);
    o->box_cap = MON_BOX_CAP0;
    o->opt.debias_lut = o->debias_lut; o->opt.n_debias_lut = MON_DEBIAS_LUT;
#undef OALLOC
    // live-sample count from which an iteration's gradients are scattered through shared memory (MON_SCATTER_RESIDENT_MIN: A/B
    // and tests; 0 = always, -1 = never).  Configurations the resident kernel does not cover keep the global reductions.
    o->resident_min_live = MON_RESIDENT_MIN_LIVE;
    if (const char* env = getenv("MON_SCATTER_RESIDENT_MIN")) { const long v = atol(env); o->resident_min_live = v < 0 ? 0xffffffffu : (uint32_t)v; }
    if (!mon_scatter_resident_supported(grid)) { o->resident_min_live = 0xffffffffu; o->scatter_unified = false; }
    o->fuse_supported = o->scatter_unified;        // power-of-two tables (scatter_level_pow2)
    if (const char* env = getenv("MON_SCATTER_FUSED")) o->fuse_mode = atoi(env);
    if (const char* env = getenv("MON_SCATTER_FUSED_BELOW")) o->fuse_below = (uint32_t)atol(env);
    if ((e = cudaMallocHost(&o->h_live, 2 * sizeof(uint32_t))) == cudaSuccess) o->h_live[0] = o->h_live[1] = 0xffffffffu;
    if (e != cudaSuccess ||
        (e = cudaMallocHost(&o->h_ctrl, sizeof(MonCtrl))) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&o->aux, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&o->ev_fork_m, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&o->ev_join, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreate(&o->ev0)) != cudaSuccess || (e = cudaEventCreate(&o->ev1)) != cudaSuccess) {
        mon_object_destroy(o);
        return fail(MON_ERR_CUDA, "object setup: %s", cudaGetErrorString(e));
    }
    memset(o->h_ctrl, 0, sizeof(MonCtrl));

    // ---- A12 parameter initialisation (Trainer ctor, trainer.h:53-90)
    std::seed_seq seq{seed};
    std::vector<uint32_t> seeds(2);
    seq.generate(seeds.begin(), seeds.end());
    HostPcg32 rng(seeds.front());
    std::vector<float> mlp(o->n_mlp);
    {
        float* p = mlp.data();
        auto xavier = [&](uint32_t rows, uint32_t cols) {  // gpu_matrix.h:291-303
            float scale = 1.0f;
            scale *= std::sqrt(6.0f / (float)(rows + cols));
            for (uint32_t i = 0; i < rows * cols; ++i) p[i] = rng.next_float() * 2.0f * scale - scale;
            p += rows * cols;
        };
        xavier(MON_WIDTH, MON_IN);
        for (uint32_t l = 1; l < cfg->n_hidden_layers; ++l) xavier(MON_WIDTH, MON_WIDTH);
        xavier(MON_OUT, MON_WIDTH);
    }
    if ((e = cudaMemcpyAsync(o->pf, mlp.data(), o->n_mlp * 4, cudaMemcpyHostToDevice, o->stream)) != cudaSuccess) {
        mon_object_destroy(o);
        return fail(MON_ERR_CUDA, "param upload: %s", cudaGetErrorString(e));
    }
    mon_launch_init_grid(rng.state, rng.inc, o->n_grid, o->pf + o->n_mlp, o->stream);
    mon_launch_cast_params((uint32_t)P, o->pf, o->ph, o->stream);
    mon_launch_planarize(o->grid, o->ph + o->n_mlp, o->ph_planar, o->stream);
    o->launches += 3;
    {   // Adam bias correction by step count, the reference's expression in host float arithmetic (kernels_optim.cu)
        std::vector<float> lut(MON_DEBIAS_LUT);
        for (uint32_t s = 0; s < MON_DEBIAS_LUT; ++s)
            lut[s] = sqrtf(1.0f - powf(cfg->beta2, (float)s)) / (1.0f - powf(cfg->beta1, (float)s));
        if ((e = cudaMemcpy(o->debias_lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
            mon_object_destroy(o);
            return fail(MON_ERR_CUDA, "bias-correction table upload: %s", cudaGetErrorString(e));
        }
    }
    if ((e = cudaStreamSynchronize(o->stream)) != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        mon_object_destroy(o);
        return fail(MON_ERR_CUDA, "param init: %s", cudaGetErrorString(e));
    }
    *out = o;
    return MON_OK;
}

int mon_object_destroy(mon_object* o) {
    if (!o) return MON_OK;
    cudaSetDevice(o->ds->gpu);
    if (o->stream) cudaStreamSynchronize(o->stream);
    drop_graphs(o);
    void* ptrs[] = {o->pf, o->m, o->v, o->ps, o->ph, o->gh, o->ema, o->ctrl_state, o->ctrl, o->ctrl_late, o->d_boxes, o->rays, o->ray_inst, o->target,
                    o->target_depth, o->bg, o->rgb_rays, o->depth_rays, o->mask_rays, o->loss, o->pts, o->pts_c, o->genc, o->live_cnt, o->debias_lut, o->enc,
                    o->d_enc, o->ph_planar, o->gcls, o->partials, o->occ_bits, o->occ_density, o->occ_ray_mask, o->occ_list, o->occ_count,
                    o->dbg_out, o->dbg_dout, o->inj_xy, o->inj_col, o->inj_dt, o->grad_snap, o->r_rays, o->r_orig, o->r_nhit, o->r_enc,
                    o->r_jit, o->r_rgb, o->r_depth, o->r_mask, o->r_Twc, o->r_pts, o->r_planar};
    for (void* p : ptrs) mon_dev_free(p, o->stream);
    for (void* p : o->scr) mon_dev_free(p, o->stream);
    if (o->stream) cudaStreamSynchronize(o->stream);
    if (o->h_ctrl) cudaFreeHost(o->h_ctrl);
    if (o->h_live) cudaFreeHost(o->h_live);
    cudaEvent_t evs[] = {o->ev0, o->ev1, o->ev_fork_m, o->ev_join};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    cudaStream_t sts[] = {o->aux, o->stream};
    for (cudaStream_t st : sts) if (st) cudaStreamDestroy(st);
    delete o;
    return MON_OK;
}

static int upload_boxes(mon_object* o, uint32_t first) {
    CK(cudaSetDevice(o->ds->gpu));
    const uint32_t n = (uint32_t)o->h_boxes.size();
    std::unique_lock<std::mutex> ds_lock(o->ds->mu);   // the ingest thread appends frames concurrently
    for (uint32_t i = first; i < n; ++i) {
        const mon_bbox2d& b = o->h_boxes[i];
        if (b.FrameId >= o->ds->max_frames || !o->ds->h_frames[b.FrameId].rgb)
            return fail(MON_ERR_ARG, "box %u references frame %u which is not in the dataset", i, b.FrameId);
        // the reference samples x in [b.x, b.x+b.w] inclusive (curand's (0,1])
        if (b.w == 0 || b.h == 0 || (uint64_t)b.x + b.w > (uint64_t)o->ds->W || (uint64_t)b.y + b.h > (uint64_t)o->ds->H)
            return fail(MON_ERR_ARG, "box %u (x=%u y=%u h=%u w=%u) lies outside the %dx%d image", i, b.x, b.y, b.h, b.w, o->ds->W, o->ds->H);
    }
    ds_lock.unlock();
    CK(cudaStreamSynchronize(o->stream));
    if (n > o->box_cap) {
        const uint32_t cap = std::max<uint32_t>(256, n * 2);
        mon_bbox2d* p = nullptr;
        CK(mon_dev_malloc(reinterpret_cast<void**>(&p), sizeof(mon_bbox2d) * cap, o->stream));
        mon_dev_free(o->d_boxes, o->stream);
        o->d_boxes = p; o->box_cap = cap;
        first = 0;
        drop_graphs(o);  // the captured graphs hold the old pointer
    }
    if (n > first) CK(cudaMemcpyAsync(o->d_boxes + first, o->h_boxes.data() + first, sizeof(mon_bbox2d) * (n - first), cudaMemcpyHostToDevice, o->stream));
    CK(cudaMemcpyAsync(&o->ctrl_state->n_boxes, &n, 4, cudaMemcpyHostToDevice, o->stream));
    CK(cudaStreamSynchronize(o->stream));
    return MON_OK;
}

int mon_object_set_bboxes(mon_object* o, const mon_bbox2d* boxes, uint32_t n) {
    if (!o || (!boxes && n)) return fail(MON_ERR_ARG, "NULL argument");
    std::vector<mon_bbox2d> keep = o->h_boxes;
    o->h_boxes.assign(boxes, boxes + n);
    const int rc = upload_boxes(o, 0);
    if (rc != MON_OK) o->h_boxes = keep;
    return rc;
}

int mon_object_add_bboxes(mon_object* o, const mon_bbox2d* boxes, uint32_t n) {
    if (!o || (!boxes && n)) return fail(MON_ERR_ARG, "NULL argument");
    const uint32_t first = (uint32_t)o->h_boxes.size();
    o->h_boxes.insert(o->h_boxes.end(), boxes, boxes + n);
    const int rc = upload_boxes(o, first);
    if (rc != MON_OK) o->h_boxes.resize(first);
    return rc;
}

static int finish_timing(mon_object* o) {
    if (o->timing_pending) {
        CK(cudaEventElapsedTime(&o->last_ms, o->ev0, o->ev1));
        o->timing_pending = false;
    }
    return MON_OK;
}

int mon_object_prepare_train(mon_object* o, uint32_t iters) {
    if (!o) return fail(MON_ERR_ARG, "obj is NULL");
    // allowed before the first boxes arrive (online mode prepares at object creation): the graphs capture the address of the
    // box buffer, which exists from the start, and the live box count is read on the device
    CK(cudaSetDevice(o->ds->gpu));
    cudaGraphExec_t g = nullptr;
    const GraphPlan plan = graph_plan(iters);
    // both variants a call of this length may take (with and without the scatter kernel, use_fused)
    for (int fused = 0; fused < 2; ++fused) {
        if (fused ? (!o->fuse_supported || o->occ_bits || o->fuse_mode < 0) : o->fuse_mode > 0 && o->fuse_supported && !o->occ_bits) continue;
        if (plan.n_chunk) { int rc = graph_for(o, plan.chunk, fused != 0, &g); if (rc != MON_OK) return rc; }
        if (plan.rem) { int rc = graph_for(o, plan.rem, fused != 0, &g); if (rc != MON_OK) return rc; }
    }
    return MON_OK;
}

static int occ_update(mon_object* o);      // refresh of the opt-in occupancy grid (defined with the inference helpers below)
static int occ_maybe_update(mon_object* o) {
    if (!o->occ_res || o->iters_enqueued < o->occ_warmup) return MON_OK;
    if (o->occ_started && o->iters_enqueued - o->occ_last_update < o->occ_interval) return MON_OK;
    return occ_update(o);
}

int mon_object_train_async(mon_object* o, uint32_t iters) {
    if (!o) return fail(MON_ERR_ARG, "obj is NULL");
    if (o->h_boxes.empty()) return fail(MON_ERR_STATE, "no 2-D boxes: call mon_object_set_bboxes first");
    CK(cudaSetDevice(o->ds->gpu));
    // a call of n iterations = replays of one chunk graph (+ ONE graph of exactly the remainder, graph_plan): every
    // iteration but the first of each graph has its batch generation hidden behind the previous iteration's scatter
    cudaGraphExec_t g_chunk = nullptr, g_rem = nullptr;
    const GraphPlan plan = graph_plan(iters);
    const uint32_t rem = plan.rem;
    const bool fused = use_fused(o);
    if (plan.n_chunk) { int rc = graph_for(o, plan.chunk, fused, &g_chunk); if (rc != MON_OK) return rc; }
    if (rem) { int rc = graph_for(o, rem, fused, &g_rem); if (rc != MON_OK) return rc; }
    CK(cudaStreamWaitEvent(o->stream, o->ds->ev_uploaded, 0));   // frames uploaded asynchronously from pinned buffers
    CK(cudaEventRecord(o->ev0, o->stream));
    // opt-in occupancy mode: the grid is refreshed between graph replays (enqueued on the same stream, no host wait)
    for (uint32_t k = 0; k < plan.n_chunk; ++k) {
        if (o->occ_res) { int rc = occ_maybe_update(o); if (rc != MON_OK) return rc; }
        CK(cudaGraphLaunch(g_chunk, o->stream));
        o->iters_enqueued += plan.chunk;
    }
    if (rem) {
        if (o->occ_res) { int rc = occ_maybe_update(o); if (rc != MON_OK) return rc; }
        CK(cudaGraphLaunch(g_rem, o->stream));
        o->iters_enqueued += rem;
    }
    CK(cudaEventRecord(o->ev1, o->stream));
    // the live-sample counts of the call's last two iterations, for the next call's choice of graph variant (never waited for)
    CK(cudaMemcpyAsync(o->h_live, o->live_cnt, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, o->stream));
    o->timing_pending = true;
    o->launches += launches_in_graph(o, iters, fused);
    o->have_injected = false;
    return MON_OK;
}

int mon_object_sync(mon_object* o) {
    if (!o) return fail(MON_ERR_ARG, "obj is NULL");
    CK(cudaSetDevice(o->ds->gpu));
    CK(cudaStreamSynchronize(o->stream));
    return finish_timing(o);
}

static int read_ctrl(mon_object* o) {
    CK(cudaMemcpyAsync(o->h_ctrl, o->ctrl_late, sizeof(MonCtrl), cudaMemcpyDeviceToHost, o->stream));
    CK(cudaStreamSynchronize(o->stream));
    return MON_OK;
}

int mon_object_train(mon_object* o, uint32_t iters, float* loss) {
    int rc = mon_object_train_async(o, iters);
    if (rc != MON_OK) return rc;
    rc = read_ctrl(o);  // one 32-byte D2H read per call: the logged loss, like Train_Step (:1650-1658)
    if (rc != MON_OK) return rc;
    rc = finish_timing(o);
    if (rc != MON_OK) return rc;
    if (loss) *loss = o->h_ctrl->loss_mean;
    return MON_OK;
}

int mon_object_train_profiled(mon_object* o, uint32_t iters, float* stage_ms, uint32_t n_stages) {
    if (!o || !stage_ms) return fail(MON_ERR_ARG, "NULL argument");
    if (n_stages != MON_N_STAGES) return fail(MON_ERR_ARG, "n_stages must be %d", MON_N_STAGES);
    if (o->h_boxes.empty()) return fail(MON_ERR_STATE, "no 2-D boxes: call mon_object_set_bboxes first");
    CK(cudaSetDevice(o->ds->gpu));
    // all iterations are enqueued back to back (one event set per iteration, one synchronisation at the end), so the
    // stage times are not inflated by the host catching up after a per-iteration sync
    const uint32_t chunk = std::min<uint32_t>(iters, 64u);
    std::vector<cudaEvent_t> ev((size_t)chunk * (MON_N_STAGES + 1), nullptr);
    for (auto& x : ev) CK(cudaEventCreate(&x));
    double acc[MON_N_STAGES] = {0};
    const MonBatch b = make_batch(o, false, false);
    const bool fused = use_fused(o);      // the variant a graph call would take now; its scatter stage is then empty
    int rc = MON_OK;
    for (uint32_t done = 0; done < iters && rc == MON_OK; done += chunk) {
        const uint32_t n_it = std::min(chunk, iters - done);
        for (uint32_t it = 0; it < n_it && rc == MON_OK; ++it) {
            int n = 0;
            rc = enqueue_iteration(o, b, false, &n, &ev[(size_t)it * (MON_N_STAGES + 1)], fused);
            if (rc == MON_OK) o->launches += (uint64_t)n;
        }
        if (rc != MON_OK) break;
        cudaMemcpyAsync(o->h_live, o->live_cnt, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, o->stream);
        cudaError_t e = cudaStreamSynchronize(o->stream);
        if (e != cudaSuccess) { rc = fail(MON_ERR_CUDA, "profiled iterations: %s", cudaGetErrorString(e)); break; }
        for (uint32_t it = 0; it < n_it; ++it)
            for (int k = 0; k < MON_N_STAGES; ++k) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ev[(size_t)it * (MON_N_STAGES + 1) + k], ev[(size_t)it * (MON_N_STAGES + 1) + k + 1]);
                acc[k] += ms;
            }
    }
    for (auto& x : ev) cudaEventDestroy(x);
    if (rc != MON_OK) return rc;
    for (int k = 0; k < MON_N_STAGES; ++k) stage_ms[k] = iters ? (float)(acc[k] / iters) : 0.0f;
    o->have_injected = false;
    return MON_OK;
}

int mon_object_last_train_ms(mon_object* o, float* ms) {
    if (!o || !ms) return fail(MON_ERR_ARG, "NULL argument");
    if (o->timing_pending) return fail(MON_ERR_STATE, "training still in flight: call mon_object_sync first");
    *ms = o->last_ms;
    return MON_OK;
}

int mon_object_step_count(mon_object* o, uint32_t* step) {
    if (!o || !step) return fail(MON_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(o->ds->gpu));
    int rc = read_ctrl(o);
    if (rc != MON_OK) return rc;
    *step = o->h_ctrl->step;
    return MON_OK;
}

int mon_object_live_samples(mon_object* o, uint32_t* n_live, uint32_t* n_points) {
    if (!o || !n_live) return fail(MON_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(o->ds->gpu));
    int rc = read_ctrl(o);   // waits for the stream; the control block names the last iteration
    if (rc != MON_OK) return rc;
    uint32_t cnt[2] = {0, 0};
    CK(cudaMemcpy(cnt, o->live_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    *n_live = o->h_ctrl->iter ? cnt[(o->h_ctrl->iter - 1) & 1u] : 0u;
    if (n_points) *n_points = o->N;
    return MON_OK;
}

int mon_object_launch_count(mon_object* o, uint64_t* n) {
    if (!o || !n) return fail(MON_ERR_ARG, "NULL argument");
    *n = o->launches;
    return MON_OK;
}

// ------------------------------------------------------------------------------- parity hooks
static int ensure_hooks(mon_object* o) {
    if (o->dbg_out) return MON_OK;
    const size_t R = o->R, N = o->N;
    // all or nothing: the buffers are committed to the object only when every allocation succeeded
    void* tmp[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const size_t bytes[7] = {N * 16, N * 16, R * 8, R * 12, N * 4, (size_t)o->P * 4, N * MON_IN * 2};
    for (int k = 0; k < 7; ++k) {
        cudaError_t e = mon_dev_malloc(&tmp[k], bytes[k], o->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(tmp[k], 0, bytes[k], o->stream);
        if (e != cudaSuccess) {
            for (void* p : tmp) mon_dev_free(p, o->stream);
            return fail(MON_ERR_CUDA, "parity-hook buffers (%zu B): %s", bytes[k], cudaGetErrorString(e));
        }
    }
    o->dbg_dout = static_cast<float*>(tmp[1]); o->inj_xy = static_cast<float*>(tmp[2]); o->inj_col = static_cast<float*>(tmp[3]);
    o->inj_dt = static_cast<float*>(tmp[4]); o->grad_snap = static_cast<float*>(tmp[5]); o->d_enc = static_cast<__half*>(tmp[6]);
    o->dbg_out = static_cast<float*>(tmp[0]);
    return MON_OK;
}

int mon_object_train_injected(mon_object* o, const float* sample_xy, const float* rand_colors,
                              const float* rand_dt, float* loss, uint32_t* n_in_box) {
    if (!o || !sample_xy || !rand_colors || !rand_dt) return fail(MON_ERR_ARG, "NULL argument");
    if (o->h_boxes.empty()) return fail(MON_ERR_STATE, "no 2-D boxes: call mon_object_set_bboxes first");
    CK(cudaSetDevice(o->ds->gpu));
    int rc = ensure_hooks(o);
    if (rc != MON_OK) return rc;
    CK(cudaMemcpyAsync(o->inj_xy, sample_xy, (size_t)o->R * 8, cudaMemcpyHostToDevice, o->stream));
    CK(cudaMemcpyAsync(o->inj_col, rand_colors, (size_t)o->R * 12, cudaMemcpyHostToDevice, o->stream));
    CK(cudaMemcpyAsync(o->inj_dt, rand_dt, (size_t)o->N * 4, cudaMemcpyHostToDevice, o->stream));
    CK(cudaMemsetAsync(o->dbg_out, 0, (size_t)o->N * 16, o->stream));
    CK(cudaMemsetAsync(o->dbg_dout, 0, (size_t)o->N * 16, o->stream));
    CK(cudaMemsetAsync(o->d_enc, 0, (size_t)o->N * MON_IN * 2, o->stream));
    const MonBatch b = make_batch(o, true, true);
    int n = 0;
    CK(cudaEventRecord(o->ev0, o->stream));
    // the parity hooks run the chain with the scatter kernel; MON_SCATTER_FUSED=1 (tests) puts the fused form under them
    rc = enqueue_iteration(o, b, true, &n, nullptr, o->fuse_mode > 0 && o->fuse_supported && !o->occ_bits);
    if (rc != MON_OK) return rc;
    CK(cudaEventRecord(o->ev1, o->stream));
    o->timing_pending = true;
    o->launches += (uint64_t)n;
    rc = read_ctrl(o);
    if (rc != MON_OK) return rc;
    rc = finish_timing(o);
    if (rc != MON_OK) return rc;
    o->have_injected = true;
    if (loss) *loss = o->h_ctrl->loss_mean;
    if (n_in_box) *n_in_box = o->h_ctrl->n_in;
    return MON_OK;
}

namespace {
__global__ void k_half_to_float(size_t n, const __half* __restrict__ in, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __half2float(in[i]);
}
__global__ void k_u16_to_float(size_t n, const uint16_t* __restrict__ in, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void k_u32_to_float(size_t n, const uint32_t* __restrict__ in, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void k_u8_to_float(size_t n, const uint8_t* __restrict__ in, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void k_float_to_half(size_t n, const float* __restrict__ in, __half* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(in[i]);
}
__global__ void k_lattice_points(uint32_t rx, uint32_t ry, uint32_t rz, float* __restrict__ out) {
    // generate_grid_samples_nerf_uniform (nerf_model.cu:296-309): pos = idx / (res - 1), x fastest
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)rx * ry * rz;
    if (i >= n) return;
    const uint32_t x = (uint32_t)(i % rx), y = (uint32_t)((i / rx) % ry), z = (uint32_t)(i / ((size_t)rx * ry));
    out[3 * i + 0] = __fdiv_rn((float)x, (float)(rx - 1));
    out[3 * i + 1] = __fdiv_rn((float)y, (float)(ry - 1));
    out[3 * i + 2] = __fdiv_rn((float)z, (float)(rz - 1));
}
// level-major fp16 pairs [C/2][n][2] -> point-major [n][C] (parity hooks only)
__global__ void k_soa_to_rows_float(size_t n, uint32_t C, const __half* __restrict__ soa, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const size_t p = i / C, c = i - p * C;
    out[i] = __half2float(soa[((c >> 1) * n + p) * 2 + (c & 1)]);
}
__global__ void k_soa_to_rows_half(size_t n, uint32_t C, const __half* __restrict__ soa, __half* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const size_t p = i / C, c = i - p * C;
    out[i] = soa[((c >> 1) * n + p) * 2 + (c & 1)];
}
// occupancy grid update: cell (x, y, z) of a res^3 grid takes the maximum density exp(logit) over its 8 corners on the (res+1)^3
// lattice, the running grid is max(decay * old, new), and the cell's bit is set while that exceeds the threshold (32 consecutive
// cells = one word, written by ballot).  res^3 is a multiple of 32.
__global__ void k_occ_fold(uint32_t res, const float* __restrict__ out4, float* __restrict__ density, uint32_t* __restrict__ bits, float decay, float thresh) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = res * res * res;
    bool occ = false;
    if (i < n) {
        const uint32_t x = i % res, y = (i / res) % res, z = i / (res * res), r1 = res + 1;
        float m = -1e30f;
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const size_t li = (size_t)(x + (c & 1u)) + (size_t)r1 * ((y + ((c >> 1) & 1u)) + (size_t)r1 * (z + (c >> 2)));
            m = fmaxf(m, out4[4 * li + 3]);
        }
        const float d = fmaxf(density[i] * decay, __expf(fminf(m, 30.0f)));
        density[i] = d;
        occ = d > thresh;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, occ);
    if ((threadIdx.x & 31u) == 0u && i < n) bits[i >> 5] = word;
}
__global__ void k_extract_sigma(size_t n, const float* __restrict__ out4, float* __restrict__ sigma) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sigma[i] = out4[4 * i + 3];
}

// copies n values of a device array to the host as float, converting on the device
int fetch_as_float(mon_object* o, const void* src, int kind /*0 f32 1 f16 2 u32 3 u8 4 u16*/, size_t n, float* out) {
    if (n == 0) return MON_OK;
    float* tmp = nullptr;
    const float* from = nullptr;
    if (kind == 0) {
        from = static_cast<const float*>(src);
    } else {
        CK(mon_dev_malloc(reinterpret_cast<void**>(&tmp), n * 4, o->stream));
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (kind == 1) k_half_to_float<<<blocks, 256, 0, o->stream>>>(n, static_cast<const __half*>(src), tmp);
        else if (kind == 2) k_u32_to_float<<<blocks, 256, 0, o->stream>>>(n, static_cast<const uint32_t*>(src), tmp);
        else if (kind == 4) k_u16_to_float<<<blocks, 256, 0, o->stream>>>(n, static_cast<const uint16_t*>(src), tmp);
        else k_u8_to_float<<<blocks, 256, 0, o->stream>>>(n, static_cast<const uint8_t*>(src), tmp);
        o->launches += 1;
        from = tmp;
    }
    cudaError_t e = cudaMemcpyAsync(out, from, n * 4, cudaMemcpyDeviceToHost, o->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(o->stream);
    mon_dev_free(tmp, o->stream);
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "state read-back: %s", cudaGetErrorString(e));
    return MON_OK;
}
}  // namespace

int mon_object_get_state(mon_object* o, int which, float* out, size_t n) {
    if (!o || !out) return fail(MON_ERR_ARG, "NULL argument");
    if (n != o->P) return fail(MON_ERR_ARG, "n must equal the parameter count %u", o->P);
    CK(cudaSetDevice(o->ds->gpu));
    switch (which) {
        case 0: return fetch_as_float(o, o->pf, 0, n, out);
        case 1:
            // the interleaved grid part is rebuilt from the planar working copy (the sweep maintains only that one)
            mon_launch_deplanarize(o->grid, o->ph_planar, o->ph + o->n_mlp, o->stream);
            o->launches += 1;
            return fetch_as_float(o, o->ph, 1, n, out);
        case 2: return fetch_as_float(o, o->ema, 1, n, out);
        case 3:
            if (!o->have_injected) return fail(MON_ERR_STATE, "the gradient snapshot exists only after mon_object_train_injected");
            return fetch_as_float(o, o->grad_snap, 0, n, out);
        case 4: return fetch_as_float(o, o->m, 0, n, out);
        case 5: return fetch_as_float(o, o->v, 0, n, out);
        case 6: return fetch_as_float(o, o->ps, 4, n, out);
    }
    return fail(MON_ERR_ARG, "unknown state selector %d", which);
}

int mon_object_set_params(mon_object* o, const float* params, size_t n) {
    if (!o || !params) return fail(MON_ERR_ARG, "NULL argument");
    if (n != o->P) return fail(MON_ERR_ARG, "n must equal the parameter count %u", o->P);
    CK(cudaSetDevice(o->ds->gpu));
    CK(cudaMemcpyAsync(o->pf, params, n * 4, cudaMemcpyHostToDevice, o->stream));
    mon_launch_cast_params((uint32_t)n, o->pf, o->ph, o->stream);
    mon_launch_planarize(o->grid, o->ph + o->n_mlp, o->ph_planar, o->stream);
    o->launches += 2;
    CK(cudaStreamSynchronize(o->stream));
    return MON_OK;
}

int mon_object_last(mon_object* o, int which, float* out, size_t cap, size_t* n_out) {
    if (!o || !out) return fail(MON_ERR_ARG, "NULL argument");
    if (!o->have_injected) return fail(MON_ERR_STATE, "intermediates exist only after mon_object_train_injected");
    CK(cudaSetDevice(o->ds->gpu));
    const size_t R = o->R, N = o->N;
    const void* src = nullptr; int kind = 0; size_t n = 0;
    switch (which) {
        case 0: src = o->rays; n = R * 9; break;
        case 1: src = o->pts; n = N * 3; break;
        case 3: {   // stored feature-major on the device; handed out point-major [N][32] like the other hooks
            n = N * MON_IN;
            if (n_out) *n_out = n;
            if (cap < n) return fail(MON_ERR_ARG, "buffer too small: need %zu floats", n);
            float* tmp = nullptr;
            CK(mon_dev_malloc(reinterpret_cast<void**>(&tmp), n * 4, o->stream));
            k_soa_to_rows_float<<<(unsigned)((n + 255) / 256), 256, 0, o->stream>>>(N, MON_IN, o->enc, tmp);
            o->launches += 1;
            const int rc = fetch_as_float(o, tmp, 0, n, out);
            mon_dev_free(tmp, o->stream);
            return rc;
        }
        case 4: src = o->dbg_out; n = N * 4; break;
        case 5: src = o->rgb_rays; n = R * 3; break;
        case 6: src = o->depth_rays; n = R; break;
        case 7: src = o->mask_rays; n = R; break;
        case 8: src = o->dbg_dout; n = N * 4; break;
        case 9: src = o->d_enc; kind = 1; n = N * MON_IN; break;
        case 10: src = o->target; n = R * 3; break;
        case 11: src = o->target_depth; n = R; break;
        case 12: src = o->ray_inst; kind = 3; n = R; break;
        case 13: src = o->loss; n = R; break;
        default: return fail(MON_ERR_ARG, "unknown intermediate selector %d", which);
    }
    if (n_out) *n_out = n;
    if (cap < n) return fail(MON_ERR_ARG, "buffer too small: need %zu floats", n);
    return fetch_as_float(o, src, kind, n, out);
}

// ------------------------------------------------------------------------------- render
static int ensure_render_ws(mon_object* o, uint32_t n_rays, size_t jitter_floats) {
    const uint32_t S2 = o->cfg.render_samples_per_ray;
    const uint32_t tile = 16384;  // rays per pass: 1 Mi points, 64 MiB of fp16 features
    if (n_rays > o->r_cap_rays) {
        void* old[] = {o->r_rays, o->r_orig, o->r_rgb, o->r_depth, o->r_mask};
        for (void* p : old) mon_dev_free(p, o->stream);
        o->r_rays = nullptr; o->r_orig = nullptr; o->r_rgb = o->r_depth = o->r_mask = nullptr;
        o->r_cap_rays = 0;
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_rays), (size_t)n_rays * sizeof(MonRay), o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_orig), (size_t)n_rays * 4, o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_rgb), (size_t)n_rays * 12, o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_depth), (size_t)n_rays * 4, o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_mask), (size_t)n_rays * 4, o->stream));
        o->r_cap_rays = n_rays;
    }
    if (!o->r_enc) {
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_enc), (size_t)tile * S2 * MON_IN * 2, o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_pts), (size_t)tile * S2 * 12, o->stream));
        if (!o->r_planar) CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_planar), (size_t)o->n_grid * 2 + 16, o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_Twc), 64, o->stream));
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_nhit), 4, o->stream));
        o->r_tile = tile;
    }
    if (jitter_floats > o->r_jit_cap) {
        mon_dev_free(o->r_jit, o->stream);
        o->r_jit = nullptr; o->r_jit_cap = 0;
        CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_jit), jitter_floats * 4, o->stream));
        o->r_jit_cap = jitter_floats;
    }
    return MON_OK;
}

static int render_impl(mon_object* o, mon_bbox2d box, const float Twc[16], bool object_centric, int use_ema,
                       const float* rand_dt, float* rgb, float* depth, float* mask) {
    if (!o || !Twc || !rgb || !depth || !mask) return fail(MON_ERR_ARG, "NULL argument");
    if (box.w == 0 || box.h == 0) return fail(MON_ERR_ARG, "empty render box");
    if ((uint64_t)box.w * box.h > (1u << 26)) return fail(MON_ERR_ARG, "render box too large");
    CK(cudaSetDevice(o->ds->gpu));
    const uint32_t n_rays = box.w * box.h, S2 = o->cfg.render_samples_per_ray;
    int rc = ensure_render_ws(o, n_rays, rand_dt ? (size_t)n_rays * S2 : 0);
    if (rc != MON_OK) return rc;
    cudaStream_t st = o->stream;
    CK(cudaMemcpyAsync(o->r_Twc, Twc, 64, cudaMemcpyHostToDevice, st));
    if (rand_dt) CK(cudaMemcpyAsync(o->r_jit, rand_dt, (size_t)n_rays * S2 * 4, cudaMemcpyHostToDevice, st));
    const float bgc = 1.0f;   // Render / RenderVideo composite over white (nerf_model.cu:1787,1953)
    MonScene sc = o->scene;
    if (object_centric)       // GenerateRenderVideoRays (nerf_model.cu:495-533): the pose is camera -> OBJECT, no world hop
        for (int k = 0; k < 16; ++k) sc.Tow[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    // rays of the box pixels; misses get their final value here, hits are compacted (ray + pixel index)
    mon_launch_render_rays(n_rays, box, sc, o->r_Twc, bgc, o->r_rays, o->r_orig, o->r_nhit, o->r_rgb, o->r_depth, o->r_mask, st);
    o->launches += 1;
    uint32_t n_hit = 0;
    CK(cudaMemcpyAsync(&n_hit, o->r_nhit, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const __half* params = use_ema ? o->ema : o->ph;
    const __half* planar = o->ph_planar;
    if (use_ema) {   // the EMA weights have no resident planar copy: renders are rare, refresh one per call
        mon_launch_planarize(o->grid, o->ema + o->n_mlp, o->r_planar, st);
        o->launches += 1;
        planar = o->r_planar;
    }
    const uint32_t rc_id = o->render_count++;
    const float* jit = rand_dt ? o->r_jit : nullptr;       // indexed by the original pixel through r_orig
    const uint32_t iter_fixed = rc_id * 4099u;
    for (uint32_t r0 = 0; r0 < n_hit; r0 += o->r_tile) {
        const uint32_t nr = std::min(o->r_tile, n_hit - r0);
        mon_launch_sample_points(nr * S2, S2, o->r_rays + r0, nullptr, jit, o->seed, nullptr, 3, iter_fixed,
                                 o->scene.bmin, o->scene.bmax, o->r_pts, st, MonLaunchOpt(), o->r_orig + r0);
        cudaError_t e = mon_launch_encode_forward(o->grid, nr * S2, o->r_pts, planar, o->r_enc, nullptr, (uint32_t)o->sm_count, st);
        if (e == cudaSuccess)
            e = mon_launch_mlp_render_tc(nr, S2, o->cfg.n_hidden_layers, o->r_rays + r0, nullptr, jit, o->seed, iter_fixed, params,
                                         o->r_enc, bgc, o->r_rgb, o->r_depth, o->r_mask, st, o->r_orig + r0);
        if (e != cudaSuccess) return fail(MON_ERR_CUDA, "render launch: %s", cudaGetErrorString(e));
        o->launches += 3;
    }
    CK(cudaMemcpyAsync(rgb, o->r_rgb, (size_t)n_rays * 12, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(depth, o->r_depth, (size_t)n_rays * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(mask, o->r_mask, (size_t)n_rays * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return MON_OK;
}

int mon_object_render(mon_object* o, mon_bbox2d box, const float Twc[16], int use_ema,
                      const float* rand_dt, float* rgb, float* depth, float* mask) {
    return render_impl(o, box, Twc, false, use_ema, rand_dt, rgb, depth, mask);
}

int mon_object_render_object_centric(mon_object* o, mon_bbox2d box, const float Toc[16], int use_ema,
                                     const float* rand_dt, float* rgb, float* depth, float* mask) {
    return render_impl(o, box, Toc, true, use_ema, rand_dt, rgb, depth, mask);
}

// slot 0: encodings of infer_points_device, 1: positions, 2: logits, 3: sigma
static int scratch(mon_object* o, int slot, size_t bytes, void** out) {
    if (bytes > o->scr_cap[slot]) {
        CK(cudaStreamSynchronize(o->stream));
        mon_dev_free(o->scr[slot], o->stream);
        o->scr[slot] = nullptr; o->scr_cap[slot] = 0;
        CK(mon_dev_malloc(&o->scr[slot], bytes, o->stream));
        o->scr_cap[slot] = bytes;
    }
    *out = o->scr[slot];
    return MON_OK;
}

// network logits at arbitrary unit-cube positions (device pointer in, device out4 [n][4]); the inference half of
// DifferentiableObject::inference as GetDensityOnGrid / compute_mesh_vertex_colors use it (nerf_model.cu:2007-2067)
static int infer_points_device(mon_object* o, const float* d_pts, uint32_t n, int use_ema, float* d_out4, bool sync = true) {
    cudaStream_t st = o->stream;
    const __half* params = use_ema ? o->ema : o->ph;
    const __half* planar = o->ph_planar;
    if (use_ema) {
        if (!o->r_planar) CK(mon_dev_malloc(reinterpret_cast<void**>(&o->r_planar), (size_t)o->n_grid * 2 + 16, o->stream));
        mon_launch_planarize(o->grid, o->ema + o->n_mlp, o->r_planar, st);
        o->launches += 1;
        planar = o->r_planar;
    }
    __half* enc = nullptr;
    int rc = scratch(o, 0, (size_t)n * MON_IN * 2, reinterpret_cast<void**>(&enc));
    if (rc != MON_OK) return rc;
    cudaError_t e = mon_launch_encode_forward(o->grid, n, d_pts, planar, enc, nullptr, (uint32_t)o->sm_count, st);
    if (e == cudaSuccess) e = mon_launch_mlp_infer_tc(n, o->cfg.n_hidden_layers, params, enc, d_out4, st);
    o->launches += 2;
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "inference: %s", cudaGetErrorString(e));
    return MON_OK;
}

int mon_object_query_points(mon_object* o, const float* points_unit, uint32_t n, int use_ema, float* out4) {
    if (!o || !points_unit || !out4) return fail(MON_ERR_ARG, "NULL argument");
    if (n == 0) return MON_OK;
    CK(cudaSetDevice(o->ds->gpu));
    float *pts = nullptr, *res = nullptr;
    int rc = scratch(o, 1, (size_t)n * 12, reinterpret_cast<void**>(&pts));
    if (rc == MON_OK) rc = scratch(o, 2, (size_t)n * 16, reinterpret_cast<void**>(&res));
    if (rc != MON_OK) return rc;
    cudaError_t e = cudaMemcpyAsync(pts, points_unit, (size_t)n * 12, cudaMemcpyHostToDevice, o->stream);
    if (e == cudaSuccess) rc = infer_points_device(o, pts, n, use_ema, res);
    if (e == cudaSuccess && rc == MON_OK) {
        e = cudaMemcpyAsync(out4, res, (size_t)n * 16, cudaMemcpyDeviceToHost, o->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(o->stream);
    }
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "mon_object_query_points: %s", cudaGetErrorString(e));
    return rc;
}

int mon_object_density_grid(mon_object* o, const uint32_t res[3], float* out) {
    if (!o || !res || !out) return fail(MON_ERR_ARG, "NULL argument");
    if (res[0] < 2 || res[1] < 2 || res[2] < 2) return fail(MON_ERR_ARG, "resolution must be >= 2 per axis");
    const size_t n = (size_t)res[0] * res[1] * res[2];
    if (n > (1u << 27)) return fail(MON_ERR_ARG, "lattice too large");
    CK(cudaSetDevice(o->ds->gpu));
    cudaStream_t st = o->stream;
    float *pts = nullptr, *out4 = nullptr, *sigma = nullptr;
    int rc = scratch(o, 1, n * 12, reinterpret_cast<void**>(&pts));
    if (rc == MON_OK) rc = scratch(o, 2, n * 16, reinterpret_cast<void**>(&out4));
    if (rc == MON_OK) rc = scratch(o, 3, n * 4, reinterpret_cast<void**>(&sigma));
    if (rc != MON_OK) return rc;
    cudaError_t e = cudaSuccess;
    {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        k_lattice_points<<<blocks, 256, 0, st>>>(res[0], res[1], res[2], pts);
        // inference weights (EMA), like mpNetwork->inference_mixed_precision_impl(..., use_inference_params = true) (:2028)
        rc = infer_points_device(o, pts, (uint32_t)n, 1, out4);
        if (rc == MON_OK) {
            k_extract_sigma<<<blocks, 256, 0, st>>>(n, out4, sigma);
            o->launches += 2;
            e = cudaMemcpyAsync(out, sigma, n * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    if (rc != MON_OK) return rc;
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "density grid: %s", cudaGetErrorString(e));
    return MON_OK;
}

// ------------------------------------------------------------------------------- mesh extraction on the GPU (kernels_mesh.cu)
// GenerateMesh (nerf_model.cu:1993-2043) = GetDensityOnGrid + MarchingCubes + compute_mesh_1ring + compute_mesh_vertex_colors, all on
// the device: density lattice (EMA weights) -> count + exclusive scans -> vertices / indices / 1-ring normals -> the network's rgb at
// the vertices.  The result stays on the device until mon_mesh_read.
struct mon_mesh {
    int gpu = 0;
    cudaStream_t st = nullptr;     // the stream the mesh was built on: its object's, or one of its own (own_stream)
    bool own_stream = false;
    uint32_t n_verts = 0, n_surface = 0, n_indices = 0;
    float *verts = nullptr, *normals = nullptr;
    uint8_t* colors = nullptr;
    uint32_t* indices = nullptr;
};

int mon_mesh_destroy(mon_mesh* m) {
    if (!m) return MON_OK;
    cudaSetDevice(m->gpu);
    // pool memory, released stream-ordered (a plain cudaFree is a device-wide synchronisation and would stall every training
    // stream of the GPU and the frontend's keyframe upload behind them)
    mon_dev_free(m->verts, m->st); mon_dev_free(m->normals, m->st); mon_dev_free(m->colors, m->st); mon_dev_free(m->indices, m->st);
    if (m->own_stream) { cudaStreamSynchronize(m->st); cudaStreamDestroy(m->st); }
    delete m;
    return MON_OK;
}

// the iso-surface of a lattice that is already in device memory, on stream st (synchronises once, for the two counts)
static int mesh_from_device_lattice(int gpu, const float* d_sigma, uint32_t res, const float bmin[3], const float bmax[3], float thresh, cudaStream_t st,
                                    mon_mesh** out) {
    const size_t n = (size_t)res * res * res;
    mon_mesh* m = new mon_mesh();
    m->gpu = gpu;
    m->st = st;
    uint32_t *v_off = nullptr, *i_off = nullptr, *totals = nullptr, *sums = nullptr, *vid = nullptr;
    auto cleanup = [&](bool all) {
        mon_dev_free(v_off, st); mon_dev_free(i_off, st); mon_dev_free(totals, st); mon_dev_free(sums, st); mon_dev_free(vid, st);
        if (all) mon_mesh_destroy(m);
    };
    cudaError_t e;
    if ((e = mon_dev_malloc(reinterpret_cast<void**>(&v_off), n * 4, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&i_off), n * 4, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&totals), mon_mesh_scan_scratch_words(res) * 4, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&sums), 8, st)) != cudaSuccess ||
        (e = mon_launch_mc_count(d_sigma, res, thresh, v_off, i_off, totals, sums, st)) != cudaSuccess) {
        cleanup(true);
        return fail(MON_ERR_CUDA, "marching cubes (count): %s", cudaGetErrorString(e));
    }
    uint32_t h_sums[2] = {0, 0};
    if ((e = cudaMemcpyAsync(h_sums, sums, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess || (e = cudaStreamSynchronize(st)) != cudaSuccess) {
        cleanup(true);
        return fail(MON_ERR_CUDA, "marching cubes (count): %s", cudaGetErrorString(e));
    }
    m->n_surface = h_sums[0];
    m->n_indices = h_sums[1];
    m->n_verts = (m->n_surface + 127u) & ~127u;      // "round for later nn stuff" (marching_cubes.cu:499): zero vertices
    const size_t nv = std::max<size_t>(m->n_verts, 1), ni = std::max<size_t>(m->n_indices, 1);
    if ((e = mon_dev_malloc(reinterpret_cast<void**>(&m->verts), nv * 12, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&m->normals), nv * 12, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&m->colors), nv * 3, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&m->indices), ni * 4, st)) != cudaSuccess ||
        (e = cudaMemsetAsync(m->colors, 0, nv * 3, st)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&vid), n * 12, st)) != cudaSuccess ||
        (e = mon_launch_mc_build(d_sigma, res, thresh, bmin, bmax, v_off, i_off, m->n_surface, m->n_verts, m->n_indices, vid, m->verts, m->normals,
                                 m->indices, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess) {
        cleanup(true);
        return fail(MON_ERR_CUDA, "marching cubes (build): %s", cudaGetErrorString(e));
    }
    cleanup(false);
    *out = m;
    return MON_OK;
}

int mon_mesh_from_lattice(int gpu, const float* sigma, uint32_t res, const float bmin[3], const float bmax[3], float thresh, mon_mesh** out) {
    if (!sigma || !bmin || !bmax || !out) return fail(MON_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (res < 2 || res > 512) return fail(MON_ERR_ARG, "resolution must be 2..512");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(MON_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    if (gpu < 0 || gpu >= n_dev) return fail(MON_ERR_ARG, "gpu %d out of range (0..%d)", gpu, n_dev - 1);
    CK(cudaSetDevice(gpu));
    const size_t n = (size_t)res * res * res;
    cudaStream_t st = nullptr;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    float* d_sigma = nullptr;
    cudaError_t e = mon_dev_malloc(reinterpret_cast<void**>(&d_sigma), n * 4, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sigma, sigma, n * 4, cudaMemcpyHostToDevice, st);
    int rc = e == cudaSuccess ? mesh_from_device_lattice(gpu, d_sigma, res, bmin, bmax, thresh, st, out) : fail(MON_ERR_CUDA, "lattice upload: %s", cudaGetErrorString(e));
    mon_dev_free(d_sigma, st);
    if (rc == MON_OK) (*out)->own_stream = true;      // destroyed with the mesh
    else { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    return rc;
}

int mon_object_extract_mesh(mon_object* o, uint32_t res, float thresh, mon_mesh** out) {
    if (!o || !out) return fail(MON_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (res < 2 || res > 512) return fail(MON_ERR_ARG, "resolution must be 2..512");
    CK(cudaSetDevice(o->ds->gpu));
    cudaStream_t st = o->stream;
    const size_t n = (size_t)res * res * res;
    float *pts = nullptr, *out4 = nullptr, *sigma = nullptr;
    int rc = scratch(o, 1, n * 12, reinterpret_cast<void**>(&pts));
    if (rc == MON_OK) rc = scratch(o, 2, n * 16, reinterpret_cast<void**>(&out4));
    if (rc == MON_OK) rc = scratch(o, 3, n * 4, reinterpret_cast<void**>(&sigma));
    if (rc != MON_OK) return rc;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_lattice_points<<<blocks, 256, 0, st>>>(res, res, res, pts);
    rc = infer_points_device(o, pts, (uint32_t)n, 1, out4, false);      // inference (EMA) weights, like GetDensityOnGrid (:2028)
    if (rc != MON_OK) return rc;
    k_extract_sigma<<<blocks, 256, 0, st>>>(n, out4, sigma);
    o->launches += 2;
    mon_mesh* m = nullptr;
    rc = mesh_from_device_lattice(o->ds->gpu, sigma, res, o->scene.bmin, o->scene.bmax, thresh, st, &m);
    if (rc != MON_OK) return rc;
    o->launches += 11;
    if (m->n_verts) {
        // compute_mesh_vertex_colors (:2045-2067): the network at WarpPoint(vertex) — the padding vertices included, as there
        float *unit = nullptr, *rgb4 = nullptr;
        rc = scratch(o, 1, (size_t)m->n_verts * 12, reinterpret_cast<void**>(&unit));
        if (rc == MON_OK) rc = scratch(o, 2, (size_t)m->n_verts * 16, reinterpret_cast<void**>(&rgb4));
        if (rc == MON_OK) {
            mon_launch_mesh_unit_points(m->n_verts, m->verts, o->scene.bmin, o->scene.bmax, unit, st);
            rc = infer_points_device(o, unit, m->n_verts, 1, rgb4, false);
        }
        if (rc == MON_OK) {
            mon_launch_mesh_colors(m->n_verts, rgb4, m->colors, st);
            o->launches += 2;
            const cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = fail(MON_ERR_CUDA, "mesh colours: %s", cudaGetErrorString(e));
        }
        if (rc != MON_OK) { mon_mesh_destroy(m); return rc; }
    }
    *out = m;
    return MON_OK;
}

int mon_mesh_counts(const mon_mesh* m, uint32_t* n_verts, uint32_t* n_surface_verts, uint32_t* n_indices) {
    if (!m) return fail(MON_ERR_ARG, "mesh is NULL");
    if (n_verts) *n_verts = m->n_verts;
    if (n_surface_verts) *n_surface_verts = m->n_surface;
    if (n_indices) *n_indices = m->n_indices;
    return MON_OK;
}

int mon_mesh_read(const mon_mesh* m, float* verts, float* normals, uint8_t* colors, uint32_t* indices) {
    if (!m) return fail(MON_ERR_ARG, "mesh is NULL");
    CK(cudaSetDevice(m->gpu));
    // on the mesh's own (non-blocking) stream: the synchronous cudaMemcpy goes through the legacy default stream
    if (verts && m->n_verts) CK(cudaMemcpyAsync(verts, m->verts, (size_t)m->n_verts * 12, cudaMemcpyDeviceToHost, m->st));
    if (normals && m->n_verts) CK(cudaMemcpyAsync(normals, m->normals, (size_t)m->n_verts * 12, cudaMemcpyDeviceToHost, m->st));
    if (colors && m->n_verts) CK(cudaMemcpyAsync(colors, m->colors, (size_t)m->n_verts * 3, cudaMemcpyDeviceToHost, m->st));
    if (indices && m->n_indices) CK(cudaMemcpyAsync(indices, m->indices, (size_t)m->n_indices * 4, cudaMemcpyDeviceToHost, m->st));
    CK(cudaStreamSynchronize(m->st));
    return MON_OK;
}

// ------------------------------------------------------------------------------- opt-in occupancy grid
// The reference carries instant-ngp's accelerators as dead code (Step / VolumeRenderGradient with compaction,
// nerf_model.cu:957-1132,1504-1550; BASELINE north_star: "occupancy-grid ray marching with warp-ballot sample compaction").
// Here: a res^3 bit grid over the object's box, refreshed from the network's own density between graph replays; the
// sample-points kernel tests every stratified sample against it and compacts the occupied ones by warp ballot, the encode
// kernel walks that list only, the fused MLP kernel treats the others as empty space.  It changes which samples contribute,
// hence the results: OFF by default and in every parity run.
static int occ_update(mon_object* o) {
    cudaStream_t st = o->stream;
    const uint32_t r1 = o->occ_res + 1;
    const size_t n = (size_t)r1 * r1 * r1;
    float *pts = nullptr, *out4 = nullptr;
    int rc = scratch(o, 1, n * 12, reinterpret_cast<void**>(&pts));
    if (rc == MON_OK) rc = scratch(o, 2, n * 16, reinterpret_cast<void**>(&out4));
    if (rc != MON_OK) return rc;
    k_lattice_points<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r1, r1, r1, pts);
    rc = infer_points_device(o, pts, (uint32_t)n, 0, out4, false);        // training weights, asynchronous
    if (rc != MON_OK) return rc;
    const uint32_t cells = o->occ_res * o->occ_res * o->occ_res;
    k_occ_fold<<<(cells + 255) / 256, 256, 0, st>>>(o->occ_res, out4, o->occ_density, o->occ_bits, o->occ_started ? o->occ_decay : 0.0f, o->occ_sigma_thresh);
    o->launches += 2;
    CK(cudaGetLastError());
    o->occ_started = true;
    o->occ_last_update = o->iters_enqueued;
    return MON_OK;
}

int mon_object_set_occupancy(mon_object* o, uint32_t grid_res, uint32_t warmup_iters, uint32_t update_interval, float alpha_threshold) {
    if (!o) return fail(MON_ERR_ARG, "obj is NULL");
    CK(cudaSetDevice(o->ds->gpu));
    CK(cudaStreamSynchronize(o->stream));
    drop_graphs(o);                      // the iteration graphs capture the grid's pointers (or their absence)
    void* old[] = {o->occ_bits, o->occ_density, o->occ_ray_mask, o->occ_list, o->occ_count};
    for (void* p : old) mon_dev_free(p, o->stream);
    o->occ_bits = nullptr; o->occ_density = nullptr; o->occ_ray_mask = o->occ_list = o->occ_count = nullptr;
    o->occ_res = 0; o->occ_started = false;
    if (grid_res == 0) return MON_OK;
    if (grid_res < 8 || grid_res > 256 || (grid_res & 3u)) return fail(MON_ERR_ARG, "grid_res must be 0 (off) or a multiple of 4 in 8..256");
    if (!(alpha_threshold > 0.0f && alpha_threshold < 1.0f)) return fail(MON_ERR_ARG, "alpha_threshold must be in (0, 1)");
    const size_t cells = (size_t)grid_res * grid_res * grid_res;
    cudaError_t e;
    if ((e = mon_dev_malloc(reinterpret_cast<void**>(&o->occ_bits), cells / 8, o->stream)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&o->occ_density), cells * 4, o->stream)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&o->occ_ray_mask), (size_t)o->R * 4, o->stream)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&o->occ_list), (size_t)o->N * 4, o->stream)) != cudaSuccess ||
        (e = mon_dev_malloc(reinterpret_cast<void**>(&o->occ_count), 4, o->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(o->occ_bits, 0xff, cells / 8, o->stream)) != cudaSuccess ||          // everything occupied until the first refresh
        (e = cudaMemsetAsync(o->occ_density, 0, cells * 4, o->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(o->occ_ray_mask, 0xff, (size_t)o->R * 4, o->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(o->occ_count, 0, 4, o->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(o->stream)) != cudaSuccess)
        return fail(MON_ERR_CUDA, "occupancy grid allocation: %s", cudaGetErrorString(e));
    o->occ_res = grid_res;
    o->occ_warmup = warmup_iters;
    o->occ_interval = update_interval ? update_interval : 16u;
    // a cell counts as empty while the opacity of one average sample interval inside it stays below alpha_threshold:
    // 1 - exp(-sigma * dt) < alpha  <=>  sigma < -ln(1 - alpha) / dt,  dt = mean box edge / samples per ray
    const float edge = ((o->scene.bmax[0] - o->scene.bmin[0]) + (o->scene.bmax[1] - o->scene.bmin[1]) + (o->scene.bmax[2] - o->scene.bmin[2])) / 3.0f;
    o->occ_sigma_thresh = -logf(1.0f - alpha_threshold) / (edge / (float)MON_S);
    o->occ_last_update = o->iters_enqueued;
    return MON_OK;
}

int mon_object_occupancy_stats(mon_object* o, float* occupied_cell_fraction, float* occupied_sample_fraction) {
    if (!o) return fail(MON_ERR_ARG, "obj is NULL");
    if (!o->occ_res) return fail(MON_ERR_STATE, "the occupancy grid is off (mon_object_set_occupancy)");
    CK(cudaSetDevice(o->ds->gpu));
    const size_t words = (size_t)o->occ_res * o->occ_res * o->occ_res / 32;
    std::vector<uint32_t> bits(words);
    uint32_t cnt = 0;
    CK(cudaMemcpyAsync(bits.data(), o->occ_bits, words * 4, cudaMemcpyDeviceToHost, o->stream));
    CK(cudaMemcpyAsync(&cnt, o->occ_count, 4, cudaMemcpyDeviceToHost, o->stream));
    CK(cudaStreamSynchronize(o->stream));
    size_t set = 0;
    for (uint32_t w : bits) set += (size_t)__builtin_popcount(w);
    if (occupied_cell_fraction) *occupied_cell_fraction = (float)set / (float)(words * 32);
    if (occupied_sample_fraction) *occupied_sample_fraction = (float)cnt / (float)o->N;      // of the last iteration
    return MON_OK;
}

// ------------------------------------------------------------------------------- stage hooks
int mon_debug_encode_pieces(const mon_config* cfg, uint32_t n_points, uint32_t n_ctas, uint32_t level_begin, uint32_t level_end, uint32_t* out4) {
    if (!cfg || !out4 || n_ctas == 0) return fail(MON_ERR_ARG, "NULL argument");
    std::string why;
    MonGrid grid;
    if (!validate_config(*cfg, why) || !make_grid(*cfg, grid, why)) return fail(MON_ERR_ARG, "unsupported config: %s", why.c_str());
    if (level_end > grid.n_levels) level_end = grid.n_levels;
    if (level_begin >= level_end) return fail(MON_ERR_ARG, "empty level range");
    mon_encode_pieces_host(grid, n_points, n_ctas, level_begin, level_end, out4);
    return MON_OK;
}

int mon_debug_scatter_pieces(const mon_config* cfg, uint32_t n_live, uint32_t n_ctas, uint32_t* out4) {
    if (!cfg || !out4 || n_ctas == 0) return fail(MON_ERR_ARG, "NULL argument");
    std::string why;
    MonGrid grid;
    if (!validate_config(*cfg, why) || !make_grid(*cfg, grid, why)) return fail(MON_ERR_ARG, "unsupported config: %s", why.c_str());
    if (!mon_scatter_resident_supported(grid)) return fail(MON_ERR_ARG, "the shared-memory resident scatter does not cover this configuration");
    mon_scatter_resident_pieces_host(grid, n_live, n_ctas, out4);
    return MON_OK;
}

int mon_stage_encode(const mon_config* cfg, const uint16_t* grid_fp16, size_t n_grid_params,
                     const float* points_unit, uint32_t n_points, uint16_t* enc_out) {
    if (!cfg || !grid_fp16 || !points_unit || !enc_out) return fail(MON_ERR_ARG, "NULL argument");
    MonGrid g; std::string why;
    if (!make_grid(*cfg, g, why)) return fail(MON_ERR_ARG, "%s", why.c_str());
    if (n_grid_params != (size_t)g.offset[cfg->n_levels] * 2) return fail(MON_ERR_ARG, "n_grid_params must be %u", g.offset[cfg->n_levels] * 2);
    if (n_points == 0) return MON_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(MON_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    __half *d_grid = nullptr, *d_planar = nullptr, *d_soa = nullptr, *d_enc = nullptr; float* d_pts = nullptr;
    const size_t enc_bytes = (size_t)n_points * 2 * cfg->n_levels * 2;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, 0);
    if (e == cudaSuccess) e = cudaSetDevice(0);
    if (e == cudaSuccess) e = cudaMalloc(&d_grid, n_grid_params * 2);
    if (e == cudaSuccess) e = cudaMalloc(&d_planar, n_grid_params * 2);
    if (e == cudaSuccess) e = cudaMalloc(&d_pts, (size_t)n_points * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_soa, enc_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_enc, enc_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d_grid, grid_fp16, n_grid_params * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_pts, points_unit, (size_t)n_points * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        mon_launch_planarize(g, d_grid, d_planar, nullptr);
        e = mon_launch_encode_forward(g, n_points, d_pts, d_planar, d_soa, nullptr, (uint32_t)prop.multiProcessorCount, nullptr);
        const size_t total = (size_t)n_points * 2 * cfg->n_levels;
        if (e == cudaSuccess) k_soa_to_rows_half<<<(unsigned)((total + 255) / 256), 256>>>(n_points, 2 * cfg->n_levels, d_soa, d_enc);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(enc_out, d_enc, enc_bytes, cudaMemcpyDeviceToHost);
    if (d_grid) cudaFree(d_grid);
    if (d_planar) cudaFree(d_planar);
    if (d_pts) cudaFree(d_pts);
    if (d_soa) cudaFree(d_soa);
    if (d_enc) cudaFree(d_enc);
    if (e != cudaSuccess) return fail(MON_ERR_CUDA, "mon_stage_encode: %s", cudaGetErrorString(e));
    return MON_OK;
}

}  // extern "C"
