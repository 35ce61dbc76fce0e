// ref_harness.cu — headless driver of the REFERENCE's own neural runtime (vendored tiny-cuda-nn, compiled
// unmodified from /root/reference by oracle/ref/Makefile into oracle/_ref/libtcnn_ref.a).
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under ro_map_b200/ links or loads this.
//
// What is the reference and what is ours in this file:
//   * hash-grid encoding, FullyFusedMLP forward/backward, the CUTLASS weight-gradient GEMMs, Adam /
//     ExponentialDecay / EMA and the parameter initialisation are tiny-cuda-nn's code, driven through
//     exactly the objects NeRF_Model::ResetNetwork builds (MON/Core/src/nerf_model.cu:1286-1342) and the
//     calls Step_No_Compacted / Train_Step make (:1552-1607, :1630-1648).
//   * RO-MAP's own glue kernels (GenerateRays, fill_rollover_rays, GenerateInputPoints, VolumeRender,
//     VolumeRenderGradient_No_Compacted, SumLoss, GenerateRenderRays, GenerateRenderInputPoints,
//     VolumeRender_Render; nerf_model.cu:280-294,369-493,536-626,735-954,1134-1253): with -DROMAP_GENUINE (what
//     oracle/ref/Makefile builds) this file #includes the reference's nerf_model.cu FROM WHERE IT LIES, unmodified,
//     and launches those very kernels.  Every Core translation unit includes Eigen, OpenCV 3 and GLEW, none of which
//     is installed here, so oracle/ref/shim/ provides stand-ins for exactly the types and calls that file names
//     (fixed-size vectors, a 4x4 matrix, cv::Mat as a buffer, GL handle types); the arithmetic of the fixed-size
//     expressions is written the way Eigen evaluates them (see shim/Eigen/Core).
//     Without the flag the kernels RESTATED below in the reference's launch shape are used (kept for comparison).
#include <tiny-cuda-nn/common_device.h>
#include <tiny-cuda-nn/encodings/grid.h>
#include <tiny-cuda-nn/loss.h>
#include <tiny-cuda-nn/network.h>
#include <tiny-cuda-nn/network_with_input_encoding.h>
#include <tiny-cuda-nn/optimizer.h>
#include <tiny-cuda-nn/trainer.h>

#ifdef ROMAP_GENUINE
// the reference's own translation units, in place (NeRF_Model's host methods compile too but are never called here)
#include ROMAP_NERF_MODEL_CU
#include ROMAP_MARCHING_CUBES_CU
#endif

#include <curand.h>

#include <cfloat>
#include <chrono>
#include <memory>
#include <string>
#include <vector>

using namespace tcnn;
using precision_t = network_precision_t;
using json = nlohmann::json;

namespace {

struct RRay { float o[3], d[3], d_norm, tmin, tmax; };          // nerf::Ray, 36 B
struct RBox { uint32_t FrameId, x, y, h, w; };                   // nerf::FrameIdAndBbox
struct RMeta { const float* pixels; const float* depth; const uint8_t* instance; const float* pose; };  // nerf::MetaData

struct Ref {
    json config;
    std::shared_ptr<Loss<precision_t>> loss;
    std::shared_ptr<Optimizer<precision_t>> optimizer;
    std::shared_ptr<NetworkWithInputEncoding<precision_t>> network;
    std::shared_ptr<Trainer<float, precision_t, precision_t>> trainer;
    cudaStream_t stream = nullptr;
    std::unique_ptr<Context> ctx;
    GPUMatrixDynamic<float> last_input;
    GPUMatrixDynamic<precision_t> last_output;
    uint32_t n_params = 0;
    std::string err;
    // scene for the whole-iteration loop
    std::vector<GPUMemory<float>> px, dp, poses;
    std::vector<GPUMemory<uint8_t>> inst;
    GPUMemory<RMeta> meta;
    GPUMemory<RBox> boxes;
    GPUMemory<float> K;
    uint32_t n_boxes = 0;
    int H = 0, W = 0, use_depth = 0;
    uint8_t instance_id = 0;
    float Tow[16], bmin[3], bmax[3];
    curandGenerator_t gen = nullptr;
    uint32_t R = 0;
    GPUMemory<float> sxy, rcol, rdt, target, target_depth, points, dist, rgb_rays, depth_rays, mask_rays, lossbuf, losssum;
    GPUMemory<RRay> rays;
    GPUMemory<uint8_t> rinst;
    GPUMemory<uint32_t> counter;
    GPUMemory<precision_t> out, dout;
    int step = 0;
};

__device__ void h_rot(const float* M, const float* v, float* r) {
    for (int i = 0; i < 3; ++i) r[i] = M[i] * v[0] + M[4 + i] * v[1] + M[8 + i] * v[2];
}

__device__ bool h_slab(const float* lo, const float* hi, const float* o, const float* d, float& t0, float& t1) {
    float a = (lo[0] - o[0]) / d[0], b = (hi[0] - o[0]) / d[0];
    if (a > b) { float s = a; a = b; b = s; }
    for (int k = 1; k < 3; ++k) {
        float c = (lo[k] - o[k]) / d[k], e = (hi[k] - o[k]) / d[k];
        if (c > e) { float s = c; c = e; e = s; }
        if (a > e || c > b) return false;
        if (c > a) a = c;
        if (e < b) b = e;
    }
    t0 = a; t1 = b;
    return a != FLT_MAX;
}

struct Scene { float Tow[16], lo[3], hi[3]; };

#ifdef ROMAP_GENUINE
static_assert(sizeof(nerf::Ray) == sizeof(RRay) && sizeof(nerf::FrameIdAndBbox) == sizeof(RBox) && sizeof(nerf::MetaData) == sizeof(RMeta) &&
              sizeof(Eigen::Matrix4f) == 64, "harness buffers must have the reference's POD layouts");
nerf::BoundingBox genuine_box(const Ref* r) {
    nerf::BoundingBox b;
    b.min = Eigen::Vector3f(r->bmin[0], r->bmin[1], r->bmin[2]);
    b.max = Eigen::Vector3f(r->bmax[0], r->bmax[1], r->bmax[2]);
    return b;
}
Eigen::Matrix4f genuine_Tow(const Ref* r) {
    Eigen::Matrix4f T;
    for (int i = 0; i < 16; ++i) T.data()[i] = r->Tow[i];
    return T;
}
#endif

__global__ void g_rays(uint32_t R, uint32_t n_boxes, Scene sc, uint32_t* counter, const RBox* boxes, const RMeta* meta,
                       const float* sxy, const float* rcol, RRay* rays, uint8_t* rinst, float* tgt, float* tgtd,
                       const float* K, int H, int W, uint8_t obj, bool use_depth) {
    const uint32_t i = threadIdx.x + blockDim.x * blockIdx.x;
    if (i >= R) return;
    const RBox b = boxes[i % n_boxes];
    const RMeta m = meta[b.FrameId];
    const uint32_t x = b.x + (uint32_t)(sxy[2 * i] * (int)b.w), y = b.y + (uint32_t)(sxy[2 * i + 1] * (int)b.h);
    const uint8_t id = m.instance[y * W + x];
    if (id != 0 && id != obj) return;
    float dir[3] = {((float)x - K[2]) / K[0], ((float)y - K[3]) / K[1], 1.0f};
    const float nrm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float dn[3] = {dir[0] / nrm, dir[1] / nrm, dir[2] / nrm}, dw[3], d[3], ow[3] = {m.pose[12], m.pose[13], m.pose[14]}, o[3];
    h_rot(m.pose, dn, dw);
    h_rot(sc.Tow, dw, d);
    h_rot(sc.Tow, ow, o);
    for (int k = 0; k < 3; ++k) o[k] += sc.Tow[12 + k];
    float t0, t1;
    if (!h_slab(sc.lo, sc.hi, o, d, t0, t1)) return;
    const uint32_t s = atomicAdd(counter, 1);
    RRay r;
    for (int k = 0; k < 3; ++k) { r.o[k] = o[k]; r.d[k] = d[k]; }
    r.d_norm = nrm; r.tmin = fmaxf(t0, 0.0f); r.tmax = t1;
    rays[s] = r;
    if (id != 0) {
        for (int k = 0; k < 3; ++k) tgt[s * 3 + k] = m.pixels[(y * W + x) * 3 + k];
        tgtd[s] = use_depth ? m.depth[y * W + x] * nrm : 0.0f;
        rinst[s] = 1;
    } else {
        for (int k = 0; k < 3; ++k) tgt[s * 3 + k] = rcol[s * 3 + k];
        tgtd[s] = 0.0f;
        rinst[s] = 0;
    }
}

__global__ void g_rollover(uint32_t R, const uint32_t* counter, RRay* rays, uint8_t* rinst, float* tgt, float* tgtd) {
    const uint32_t i = threadIdx.x + blockDim.x * blockIdx.x, n = *counter;
    if (i < n || i >= R || n == 0) return;
    const uint32_t s = i % n;
    rays[i] = rays[s]; rinst[i] = rinst[s]; tgtd[i] = tgtd[s];
    for (int k = 0; k < 3; ++k) tgt[i * 3 + k] = tgt[s * 3 + k];
}

__global__ void g_points(uint32_t R, uint32_t S, Scene sc, const RRay* rays, float* pts, float* dist, const float* rdt) {
    const uint32_t i = threadIdx.x + blockDim.x * blockIdx.x;
    if (i >= R) return;
    const RRay r = rays[i];
    const float dt = (r.tmax - r.tmin) / (float)S;
    for (uint32_t n = 0; n < S; ++n) {
        const float t = r.tmin + dt * ((float)n + rdt[i * S + n]);
        for (int k = 0; k < 3; ++k) pts[(i * S + n) * 3 + k] = ((r.o[k] + t * r.d[k]) - sc.lo[k]) / (sc.hi[k] - sc.lo[k]);
        dist[i * S + n] = t;
    }
}

__device__ float h_sig(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void g_render(uint32_t R, uint32_t S, uint32_t ow, const precision_t* out, const float* dist, const float* rcol,
                         const uint32_t* counter, float* rgb_rays, float* depth_rays, float* mask_rays) {
    const uint32_t i = threadIdx.x + blockDim.x * blockIdx.x;
    if (i >= R) return;
    out += (size_t)i * S * ow; dist += i * S;
    const float* bg = rcol + (i % *counter) * 3;
    float T = 1.0f, c[3] = {0, 0, 0}, dep = 0.0f, last = 0.0f;
    for (uint32_t n = 0; n < S; ++n) {
        if (T < 1e-4f) break;
        const float cur = dist[n], dt = cur - last;
        const float a = 1.0f - __expf(-__expf((float)out[3]) * dt), w = a * T;
        for (int k = 0; k < 3; ++k) c[k] += w * h_sig((float)out[k]);
        dep += w * cur;
        T *= 1.0f - a;
        out += ow; last = cur;
    }
    for (int k = 0; k < 3; ++k) rgb_rays[i * 3 + k] = c[k] + T * bg[k];
    depth_rays[i] = dep; mask_rays[i] = 1.0f - T;
}

__global__ void g_grad(uint32_t R, uint32_t S, uint32_t ow, float scale, const precision_t* out, const float* dist,
                       const uint8_t* rinst, const float* tgt, const float* tgtd, const float* rgb_rays, const float* depth_rays,
                       const float* mask_rays, precision_t* dout, float* loss) {
    const uint32_t i = threadIdx.x + blockDim.x * blockIdx.x;
    if (i >= R) return;
    out += (size_t)i * S * ow; dout += (size_t)i * S * ow; dist += i * S;
    float g[3], C[3], ml = 0.0f;
    for (int k = 0; k < 3; ++k) { C[k] = rgb_rays[i * 3 + k]; const float df = C[k] - tgt[i * 3 + k]; g[k] = 2.0f * df; ml += df * df; }
    ml /= 3.0f;
    const float Dt = tgtd[i], D = depth_rays[i], mask = mask_rays[i];
    float dd = 0.0f;
    if (Dt > 0.0f) dd = 0.5f * (D - Dt >= 0.0f ? 1.0f : -1.0f);
    const bool objray = rinst[i] == 1;
    loss[i] = objray ? ml + dd * (D - Dt) + (1 - mask) : ml + mask;
    scale /= R;
    float T = 1.0f, c2[3] = {0, 0, 0}, d2 = 0.0f, last = 0.0f;
    for (uint32_t n = 0; n < S; ++n) {
        if (T < 1e-4f) break;
        const float cur = dist[n], dt = cur - last;
        float rgb[3];
        for (int k = 0; k < 3; ++k) rgb[k] = h_sig((float)out[k]);
        const float a = 1.0f - __expf(-__expf((float)out[3]) * dt), w = a * T;
        for (int k = 0; k < 3; ++k) c2[k] += w * rgb[k];
        d2 += w * cur;
        T *= 1.0f - a;
        for (int k = 0; k < 3; ++k) { const float s = h_sig((float)out[k]); dout[k] = (precision_t)(scale * ((w * g[k]) * (s * (1 - s)))); }
        const float ds = __expf(fminf(fmaxf((float)out[3], -15.0f), 15.0f));
        float v;
        if (objray) {
            float dot = 0.0f;
            for (int k = 0; k < 3; ++k) dot += g[k] * (T * rgb[k] - (C[k] - c2[k]));
            v = ds * dt * (dot + dd * (T * cur - (D - d2)) + 0.5f * (mask >= 1 ? 1.0f : -1.0f) * (1 - mask));
        } else {
            v = ds * dt * (0.5f * (mask >= 0 ? 1.0f : -1.0f)) * (1 - mask) + ds * 0.01f;
        }
        dout[3] = (precision_t)(scale * v);
        out += ow; dout += ow; last = cur;
    }
}

__global__ void g_sumloss(uint32_t R, const float* loss, float* part) {
    __shared__ float s[256];
    const int t = threadIdx.x, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)R) return;
    s[t] = loss[i];
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) { if (t < k) s[t] += s[t + k]; __syncthreads(); }
    if (t == 0) part[blockIdx.x] = s[0];
}

template <typename T>
void to_float(cudaStream_t st, const T* src, float* host, size_t n) {
    GPUMemory<float> tmp(n);
    float* d = tmp.data();
    parallel_for_gpu(st, n, [src, d] __device__(size_t i) { d[i] = (float)src[i]; });
    CUDA_CHECK_THROW(cudaStreamSynchronize(st));
    tmp.copy_to_host(host, n);
}

Scene scene_of(const Ref* r) {
    Scene s;
    memcpy(s.Tow, r->Tow, 64); memcpy(s.lo, r->bmin, 12); memcpy(s.hi, r->bmax, 12);
    return s;
}

#define GUARD(r, ...) try { __VA_ARGS__; return 0; } catch (const std::exception& e) { (r)->err = e.what(); return -1; }

}  // namespace

extern "C" {

const char* ref_last_error(void* h) { return static_cast<Ref*>(h)->err.c_str(); }

// NeRF_Model::ResetNetwork (nerf_model.cu:1286-1342): loss otype forced to L2, Trainer seed 1337
void* ref_create(const char* json_text, uint32_t seed) {
    Ref* r = new Ref();
    try {
        r->config = json::parse(json_text, nullptr, true, true);
        r->config["loss"]["otype"] = "L2";
        r->loss.reset(create_loss<precision_t>(r->config["loss"]));
        r->optimizer.reset(create_optimizer<precision_t>(r->config["optimizer"]));
        r->network = std::make_shared<NetworkWithInputEncoding<precision_t>>(3, 4, r->config["encoding"], r->config["network"]);
        r->trainer = std::make_shared<Trainer<float, precision_t, precision_t>>(r->network, r->optimizer, r->loss, seed);
        r->n_params = (uint32_t)r->network->n_params();
        CUDA_CHECK_THROW(cudaStreamCreate(&r->stream));
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_create: %s\n", e.what());
        delete r;
        return nullptr;
    }
    return r;
}

void ref_destroy(void* h) {
    Ref* r = static_cast<Ref*>(h);
    if (!r) return;
    cudaDeviceSynchronize();
    r->ctx.reset();
    if (r->gen) curandDestroyGenerator(r->gen);
    delete r;
}

uint32_t ref_n_params(void* h) { return static_cast<Ref*>(h)->n_params; }
uint32_t ref_padded_output_width(void* h) { return static_cast<Ref*>(h)->network->padded_output_width(); }

// which: 0 fp32 master, 1 fp16 params, 2 inference (EMA) params, 3 fp16 gradients
int ref_get(void* h, int which, float* out) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const size_t n = r->n_params;
        if (which == 0) CUDA_CHECK_THROW(cudaMemcpy(out, r->trainer->params_full_precision(), n * 4, cudaMemcpyDeviceToHost));
        else if (which == 1) to_float(r->stream, r->trainer->params(), out, n);
        else if (which == 2) to_float(r->stream, r->trainer->params_inference(), out, n);
        else to_float(r->stream, r->trainer->param_gradients(), out, n);
    })
}

// overwrite fp32 master + fp16 working weights (EMA and Adam state untouched)
int ref_set_params(void* h, const float* master) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const size_t n = r->n_params;
        float* pf = r->trainer->params_full_precision();
        precision_t* p = r->trainer->params();
        CUDA_CHECK_THROW(cudaMemcpy(pf, master, n * 4, cudaMemcpyHostToDevice));
        parallel_for_gpu(r->stream, n, [pf, p] __device__(size_t i) { p[i] = (precision_t)pf[i]; });
        CUDA_CHECK_THROW(cudaStreamSynchronize(r->stream));
    })
}

// GridEncoding forward alone; enc_out is point-major [N][32] fp16 bit patterns. N % 128 == 0.
int ref_encode(void* h, const float* pts, uint32_t N, uint16_t* enc_out) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        auto enc = r->network->encoding();
        GPUMatrixDynamic<float> in(3, N, r->stream, CM);
        CUDA_CHECK_THROW(cudaMemcpyAsync(in.data(), pts, (size_t)N * 12, cudaMemcpyHostToDevice, r->stream));
        GPUMatrixDynamic<precision_t> out(enc->padded_output_width(), N, r->stream, enc->preferred_output_layout());
        enc->inference_mixed_precision(r->stream, in, out, false);
        const uint32_t Wd = enc->padded_output_width();
        std::vector<uint16_t> raw((size_t)Wd * N);
        CUDA_CHECK_THROW(cudaMemcpyAsync(raw.data(), out.data(), raw.size() * 2, cudaMemcpyDeviceToHost, r->stream));
        CUDA_CHECK_THROW(cudaStreamSynchronize(r->stream));
        const bool soa = out.layout() == RM;  // RM == SoA: feature-major
        for (uint32_t i = 0; i < N; ++i)
            for (uint32_t c = 0; c < Wd; ++c) enc_out[(size_t)i * Wd + c] = soa ? raw[(size_t)c * N + i] : raw[(size_t)i * Wd + c];
    })
}

// NetworkWithInputEncoding::forward with the training weights; out is [N][16] fp16 bit patterns; keeps the context
int ref_forward(void* h, const float* pts, uint32_t N, uint16_t* out) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        r->ctx.reset();
        r->last_input = GPUMatrixDynamic<float>(3, N, r->stream, CM);
        r->last_output = GPUMatrixDynamic<precision_t>(r->network->padded_output_width(), N, r->stream, CM);
        CUDA_CHECK_THROW(cudaMemcpyAsync(r->last_input.data(), pts, (size_t)N * 12, cudaMemcpyHostToDevice, r->stream));
        r->ctx = r->network->forward(r->stream, r->last_input, &r->last_output, false, false);
        CUDA_CHECK_THROW(cudaMemcpyAsync(out, r->last_output.data(), (size_t)N * r->network->padded_output_width() * 2, cudaMemcpyDeviceToHost, r->stream));
        CUDA_CHECK_THROW(cudaStreamSynchronize(r->stream));
    })
}

// NetworkWithInputEncoding::backward(..., EGradientMode::Overwrite) on the kept context; dout [N][16] fp16 bits
int ref_backward(void* h, const uint16_t* dout, uint32_t N) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        if (!r->ctx || r->last_input.n() != N) throw std::runtime_error("ref_backward without matching ref_forward");
        GPUMatrixDynamic<precision_t> d(r->network->padded_output_width(), N, r->stream, CM);
        CUDA_CHECK_THROW(cudaMemcpyAsync(d.data(), dout, (size_t)N * r->network->padded_output_width() * 2, cudaMemcpyHostToDevice, r->stream));
        r->network->backward(r->stream, *r->ctx, r->last_input, r->last_output, d, nullptr, false, EGradientMode::Overwrite);
        CUDA_CHECK_THROW(cudaStreamSynchronize(r->stream));
    })
}

int ref_optimizer_step(void* h, float loss_scale) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        r->trainer->optimizer_step(r->stream, loss_scale);
        CUDA_CHECK_THROW(cudaStreamSynchronize(r->stream));
    })
}

// DifferentiableObject::inference (EMA weights, fp32 [N][4]) as Render uses it (nerf_model.cu:1795)
int ref_inference(void* h, const float* pts, uint32_t N, float* out4) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        GPUMatrixDynamic<float> in(3, N, r->stream, CM);
        GPUMatrixDynamic<float> out(4, N, r->stream, CM);
        CUDA_CHECK_THROW(cudaMemcpyAsync(in.data(), pts, (size_t)N * 12, cudaMemcpyHostToDevice, r->stream));
        r->network->inference(r->stream, in, out);
        CUDA_CHECK_THROW(cudaMemcpyAsync(out4, out.data(), (size_t)N * 16, cudaMemcpyDeviceToHost, r->stream));
        CUDA_CHECK_THROW(cudaStreamSynchronize(r->stream));
    })
}

// ---- whole-iteration loop in the reference's shape -------------------------------------------------
// frames: rgb u8 RGB HxWx3 (stored as float = u8/255 like DataToGPU, nerf_data.cu:157-168), instance u8, depth f32
int ref_scene(void* h, uint32_t n_frames, const uint8_t* const* rgb, const uint8_t* const* inst, const float* const* depth,
              const float* poses16, int H, int W, const float* K4, const void* boxes_v /* FrameIdAndBbox[n_boxes] */, uint32_t n_boxes, const float* Tow16,
              const float* bmin, const float* bmax, uint8_t instance_id, int use_depth, uint32_t rays_per_batch) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const size_t px = (size_t)H * W;
        r->px.resize(n_frames); r->dp.resize(n_frames); r->inst.resize(n_frames); r->poses.resize(n_frames);
        std::vector<RMeta> meta(n_frames);
        std::vector<float> tmp(px * 3);
        for (uint32_t f = 0; f < n_frames; ++f) {
            for (size_t i = 0; i < px * 3; ++i) tmp[i] = (float)rgb[f][i] * (float)(1.0 / 255.0);
            r->px[f].resize_and_copy_from_host(tmp);
            r->inst[f].resize(px); r->inst[f].copy_from_host(inst[f], px);
            if (use_depth) { r->dp[f].resize(px); r->dp[f].copy_from_host(depth[f], px); }
            r->poses[f].resize(16); r->poses[f].copy_from_host(poses16 + (size_t)f * 16, 16);
            meta[f] = RMeta{r->px[f].data(), use_depth ? r->dp[f].data() : nullptr, r->inst[f].data(), r->poses[f].data()};
        }
        r->meta.resize_and_copy_from_host(meta);
        const RBox* boxes = static_cast<const RBox*>(boxes_v);
        std::vector<RBox> b(boxes, boxes + n_boxes);
        r->boxes.resize_and_copy_from_host(b);
        r->n_boxes = n_boxes;
        std::vector<float> k(K4, K4 + 4);
        r->K.resize_and_copy_from_host(k);
        r->H = H; r->W = W; r->use_depth = use_depth; r->instance_id = instance_id;
        memcpy(r->Tow, Tow16, 64); memcpy(r->bmin, bmin, 12); memcpy(r->bmax, bmax, 12);
        // AllocateBatchWorkspace (nerf_model.cu:1344-1427)
        const uint32_t R = rays_per_batch, S = 32, N = R * S, ow = r->network->padded_output_width();
        r->R = R;
        if (!r->gen) {
            if (curandCreateGenerator(&r->gen, CURAND_RNG_PSEUDO_XORWOW) != CURAND_STATUS_SUCCESS) throw std::runtime_error("curandCreateGenerator");
            curandSetStream(r->gen, r->stream);
        }
        r->sxy.resize(R * 2); r->rcol.resize(R * 3); r->rdt.resize(N); r->rays.resize(R); r->rinst.resize(R);
        r->target.resize(R * 3); r->target_depth.resize(R); r->points.resize((size_t)N * 3); r->dist.resize(N);
        r->out.resize((size_t)N * ow); r->dout.resize((size_t)N * ow);
        r->rgb_rays.resize(R * 3); r->depth_rays.resize(R); r->mask_rays.resize(R); r->lossbuf.resize(R); r->losssum.resize(16);
        r->counter.resize(1);
        r->losssum.memset(0);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
    })
}

// Train_Step's loop body `iters` times (nerf_model.cu:1635-1648), with the reference's syncs.
// inject != 0: use host-provided randoms (sample_xy 2R, colors 3R, dt R*32) instead of cuRAND (single iteration use).
int ref_train(void* h, uint32_t iters, const float* inj_xy, const float* inj_col, const float* inj_dt,
              float* device_ms, float* wall_ms, float* loss_out, uint32_t* n_in_out) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const uint32_t R = r->R, S = 32, N = R * S, ow = r->network->padded_output_width();
        const Scene sc = scene_of(r);
        cudaStream_t st = r->stream;
        cudaEvent_t e0, e1;
        CUDA_CHECK_THROW(cudaEventCreate(&e0)); CUDA_CHECK_THROW(cudaEventCreate(&e1));
        GPUMatrixDynamic<float> pts(r->points.data(), 3, N, CM);
        GPUMatrixDynamic<precision_t> out(r->out.data(), ow, N, CM), dout(r->dout.data(), ow, N, CM);
        CUDA_CHECK_THROW(cudaStreamSynchronize(st));
        const auto w0 = std::chrono::steady_clock::now();
        CUDA_CHECK_THROW(cudaEventRecord(e0, st));
        for (uint32_t it = 0; it < iters; ++it) {
            // GenerateBatch (:1429-1479)
            if (inj_xy) {
                CUDA_CHECK_THROW(cudaMemcpyAsync(r->sxy.data(), inj_xy, R * 8, cudaMemcpyHostToDevice, st));
                CUDA_CHECK_THROW(cudaMemcpyAsync(r->rcol.data(), inj_col, R * 12, cudaMemcpyHostToDevice, st));
            } else {
                curandGenerateUniform(r->gen, r->sxy.data(), R * 2);
                curandGenerateUniform(r->gen, r->rcol.data(), R * 3);
            }
            CUDA_CHECK_THROW(cudaMemsetAsync(r->counter.data(), 0, 4, st));
#ifdef ROMAP_GENUINE
            linear_kernel(nerf::GenerateRays, 0, st, R, (size_t)r->n_boxes, genuine_box(r), r->counter.data(), reinterpret_cast<nerf::FrameIdAndBbox*>(r->boxes.data()),
                          genuine_Tow(r), reinterpret_cast<nerf::MetaData*>(r->meta.data()), r->sxy.data(), r->rcol.data(),
                          reinterpret_cast<nerf::Ray*>(r->rays.data()), r->rinst.data(), r->target.data(), r->target_depth.data(), r->K.data(), r->H, r->W,
                          r->instance_id, r->use_depth != 0);
            CUDA_CHECK_THROW(cudaStreamSynchronize(st));
            linear_kernel(nerf::fill_rollover_rays, 0, st, R, r->counter.data(), reinterpret_cast<nerf::Ray*>(r->rays.data()), r->rinst.data(), r->target.data(), r->target_depth.data());
#else
            linear_kernel(g_rays, 0, st, R, r->n_boxes, sc, r->counter.data(), r->boxes.data(), r->meta.data(), r->sxy.data(), r->rcol.data(),
                          r->rays.data(), r->rinst.data(), r->target.data(), r->target_depth.data(), r->K.data(), r->H, r->W,
                          r->instance_id, r->use_depth != 0);
            CUDA_CHECK_THROW(cudaStreamSynchronize(st));
            g_rollover<<<n_blocks_linear(R), n_threads_linear, 0, st>>>(R, r->counter.data(), r->rays.data(), r->rinst.data(), r->target.data(), r->target_depth.data());
#endif
            if (inj_dt) CUDA_CHECK_THROW(cudaMemcpyAsync(r->rdt.data(), inj_dt, (size_t)N * 4, cudaMemcpyHostToDevice, st));
            else curandGenerateUniform(r->gen, r->rdt.data(), N);
            CUDA_CHECK_THROW(cudaStreamSynchronize(st));
#ifdef ROMAP_GENUINE
            linear_kernel(nerf::GenerateInputPoints, 0, st, R, S, genuine_box(r), reinterpret_cast<nerf::Ray*>(r->rays.data()), r->points.data(), r->dist.data(), r->rdt.data());
#else
            linear_kernel(g_points, 0, st, R, S, sc, r->rays.data(), r->points.data(), r->dist.data(), r->rdt.data());
#endif
            // Step_No_Compacted (:1552-1607)
            {
                auto ctx = r->network->forward(st, pts, &out, false, false);
#ifdef ROMAP_GENUINE
                linear_kernel(nerf::VolumeRender, 0, st, R, S, ow, genuine_box(r), nerf::ENerfActivation::Logistic, nerf::ENerfActivation::Exponential,
                              r->out.data(), r->points.data(), r->dist.data(), reinterpret_cast<nerf::Ray*>(r->rays.data()), r->rcol.data(), r->counter.data(),
                              r->rgb_rays.data(), r->depth_rays.data(), r->mask_rays.data());
                CUDA_CHECK_THROW(cudaMemsetAsync(r->dout.data(), 0, (size_t)N * ow * sizeof(precision_t), st));
                linear_kernel(nerf::VolumeRenderGradient_No_Compacted, 0, st, R, S, ow, genuine_box(r), nerf::ENerfActivation::Logistic, nerf::ENerfActivation::Exponential,
                              128.0f, r->out.data(), r->points.data(), r->dist.data(), reinterpret_cast<nerf::Ray*>(r->rays.data()), r->rinst.data(), r->target.data(),
                              r->target_depth.data(), r->rgb_rays.data(), r->depth_rays.data(), r->mask_rays.data(), r->dout.data(), r->lossbuf.data());
                nerf::SumLoss<<<16, 256, 0, st>>>(R, r->lossbuf.data(), r->losssum.data());
#else
                linear_kernel(g_render, 0, st, R, S, ow, r->out.data(), r->dist.data(), r->rcol.data(), r->counter.data(),
                              r->rgb_rays.data(), r->depth_rays.data(), r->mask_rays.data());
                CUDA_CHECK_THROW(cudaMemsetAsync(r->dout.data(), 0, (size_t)N * ow * sizeof(precision_t), st));
                linear_kernel(g_grad, 0, st, R, S, ow, 128.0f, r->out.data(), r->dist.data(), r->rinst.data(), r->target.data(), r->target_depth.data(),
                              r->rgb_rays.data(), r->depth_rays.data(), r->mask_rays.data(), r->dout.data(), r->lossbuf.data());
                g_sumloss<<<16, 256, 0, st>>>(R, r->lossbuf.data(), r->losssum.data());
#endif
                r->network->backward(st, *ctx, pts, out, dout, nullptr, false, EGradientMode::Overwrite);
            }
            r->trainer->optimizer_step(st, 128.0f);
            CUDA_CHECK_THROW(cudaStreamSynchronize(st));
            r->step++;
        }
        CUDA_CHECK_THROW(cudaEventRecord(e1, st));
        CUDA_CHECK_THROW(cudaStreamSynchronize(st));
        const auto w1 = std::chrono::steady_clock::now();
        float ms = 0;
        CUDA_CHECK_THROW(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (device_ms) *device_ms = ms;
        if (wall_ms) *wall_ms = std::chrono::duration<float, std::milli>(w1 - w0).count();
        float part[16];
        r->losssum.copy_to_host(part, 16);
        float s = 0;
        for (int i = 0; i < 16; ++i) s += part[i];
        if (loss_out) *loss_out = s / R;
        if (n_in_out) r->counter.copy_to_host(n_in_out, 1);
    })
}

// per-ray / per-point buffers of the last iteration: 0 rays(9) 1 points(3) 2 dist 4 out(ow, as float) 5 rgb 6 depth 7 mask 8 dout(ow) 10 target 11 target depth 12 flag 13 loss
int ref_last(void* h, int which, float* out) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const uint32_t R = r->R, N = R * 32, ow = r->network->padded_output_width();
        switch (which) {
            case 0: CUDA_CHECK_THROW(cudaMemcpy(out, r->rays.data(), (size_t)R * 36, cudaMemcpyDeviceToHost)); break;
            case 1: r->points.copy_to_host(out, (size_t)N * 3); break;
            case 2: r->dist.copy_to_host(out, N); break;
            case 4: to_float(r->stream, r->out.data(), out, (size_t)N * ow); break;
            case 5: r->rgb_rays.copy_to_host(out, R * 3); break;
            case 6: r->depth_rays.copy_to_host(out, R); break;
            case 7: r->mask_rays.copy_to_host(out, R); break;
            case 8: to_float(r->stream, r->dout.data(), out, (size_t)N * ow); break;
            case 10: r->target.copy_to_host(out, R * 3); break;
            case 11: r->target_depth.copy_to_host(out, R); break;
            case 12: to_float(r->stream, r->rinst.data(), out, R); break;
            case 13: r->lossbuf.copy_to_host(out, R); break;
            default: throw std::runtime_error("unknown selector");
        }
    })
}

// every later call of this host thread drives GPU `dev` (NeRF::mGPUid / cudaSetDevice in the reference's per-object threads, nerf.cu:123,192)
int ref_use_device(int dev) { return cudaSetDevice(dev) == cudaSuccess ? 0 : -1; }

// 1: the RO-MAP kernels in this library are the reference's own (nerf_model.cu compiled in place), 0: the restatement
int ref_is_genuine(void) {
#ifdef ROMAP_GENUINE
    return 1;
#else
    return 0;
#endif
}

#ifdef ROMAP_GENUINE
// NeRF_Model::Render's device work (nerf_model.cu:1702-1830) with the reference's own kernels and tiny-cuda-nn's inference
// on the EMA weights: box = {FrameId, x, y, h, w}; rand_dt: h*w*64 floats in (0,1] (injected instead of cuRAND);
// outputs rgb[h*w*3], depth[h*w], mask[h*w], plus the rays (9 floats each) and the in-box flags for stage-level checks.
int ref_render(void* h, const void* box_v, const float* Twc16, const float* rand_dt, float* rgb, float* depth, float* mask, float* rays_out, int* inbox_out,
               float* points_out /* [h*w*64][3] */, float* dist_out /* [h*w*64] */, float* out4_out /* [h*w*64][4] */) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const nerf::FrameIdAndBbox box = *static_cast<const nerf::FrameIdAndBbox*>(box_v);
        const uint32_t S2 = 64, n_rays = box.h * box.w;
        const uint32_t per = 128 / S2, n128 = (n_rays + per - 1) / per * per, batch = n128 * S2;   // RenderRaysPerBatch128 (:1747-1749)
        cudaStream_t st = r->stream;
        GPUMemory<nerf::Ray> rays(n_rays);
        GPUMemory<int> inbox(n_rays);
        GPUMemory<float> pts((size_t)3 * batch), dist((size_t)S2 * n_rays), out4((size_t)4 * batch), d_rgb(3 * n_rays), d_depth(n_rays), d_mask(n_rays), rdt((size_t)S2 * n_rays);
        pts.memset(0); out4.memset(0); dist.memset(0); rays.memset(0);
        rdt.copy_from_host(rand_dt, (size_t)S2 * n_rays);
        Eigen::Matrix4f Twc;
        for (int i = 0; i < 16; ++i) Twc.data()[i] = Twc16[i];
        linear_kernel(nerf::GenerateRenderRays, 0, st, n_rays, box, genuine_box(r), Twc, genuine_Tow(r), rays.data(), inbox.data(), r->K.data());
        linear_kernel(nerf::GenerateRenderInputPoints, 0, st, n_rays, S2, genuine_box(r), rays.data(), inbox.data(), pts.data(), dist.data(), rdt.data());
        GPUMatrixDynamic<float> in(pts.data(), 3, batch, CM), o4(out4.data(), 4, batch, CM);
        r->network->inference(st, in, o4);
        linear_kernel(nerf::VolumeRender_Render, 0, st, n_rays, S2, 4u, genuine_box(r), nerf::ENerfActivation::Logistic, nerf::ENerfActivation::Exponential, 1.0f,
                      out4.data(), pts.data(), dist.data(), rays.data(), inbox.data(), d_rgb.data(), d_depth.data(), d_mask.data());
        CUDA_CHECK_THROW(cudaStreamSynchronize(st));
        d_rgb.copy_to_host(rgb, 3 * n_rays); d_depth.copy_to_host(depth, n_rays); d_mask.copy_to_host(mask, n_rays);
        if (rays_out) CUDA_CHECK_THROW(cudaMemcpy(rays_out, rays.data(), (size_t)n_rays * 36, cudaMemcpyDeviceToHost));
        if (inbox_out) inbox.copy_to_host(inbox_out, n_rays);
        if (points_out) pts.copy_to_host(points_out, (size_t)3 * S2 * n_rays);
        if (dist_out) dist.copy_to_host(dist_out, (size_t)S2 * n_rays);
        if (out4_out) out4.copy_to_host(out4_out, (size_t)4 * S2 * n_rays);
    })
}


// NeRF_Model::Render (object_centric == 0, nerf_model.cu:1702-1830) or one view of NeRF_Model::RenderVideo (object_centric == 1:
// GenerateRenderVideoRays with a camera -> object pose, :495-534,1915-1968), device work + the three D2H copies, without the
// stage-level dumps of ref_render: what bench.py --impl reference times as "rendered rays/s".  rand_dt == NULL: cuRAND on the
// device like the reference (:1781,1929).  device_ms: CUDA events around the device work.
int ref_render2(void* h, const void* box_v, const float* T16, int object_centric, const float* rand_dt, float* rgb, float* depth, float* mask,
                float* rays_out, int* inbox_out, float* device_ms) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        const nerf::FrameIdAndBbox box = *static_cast<const nerf::FrameIdAndBbox*>(box_v);
        const uint32_t S2 = 64, n_rays = box.h * box.w;
        const uint32_t per = 128 / S2, n128 = (n_rays + per - 1) / per * per, batch = n128 * S2;
        cudaStream_t st = r->stream;
        if (!r->gen) {
            if (curandCreateGenerator(&r->gen, CURAND_RNG_PSEUDO_XORWOW) != CURAND_STATUS_SUCCESS) throw std::runtime_error("curandCreateGenerator");
            curandSetStream(r->gen, st);
        }
        cudaEvent_t e0, e1;
        CUDA_CHECK_THROW(cudaEventCreate(&e0)); CUDA_CHECK_THROW(cudaEventCreate(&e1));
        // the reference allocates its render workspace per call (allocate_workspace_and_distribute, :1751-1779)
        GPUMemory<nerf::Ray> rays(n_rays);
        GPUMemory<int> inbox(n_rays);
        GPUMemory<float> pts((size_t)3 * batch), dist((size_t)S2 * n_rays), out4((size_t)4 * batch), d_rgb(3 * n_rays), d_depth(n_rays), d_mask(n_rays), rdt((size_t)S2 * n_rays);
        Eigen::Matrix4f T;
        for (int i = 0; i < 16; ++i) T.data()[i] = T16[i];
        CUDA_CHECK_THROW(cudaEventRecord(e0, st));
        if (object_centric) linear_kernel(nerf::GenerateRenderVideoRays, 0, st, n_rays, box, genuine_box(r), T, rays.data(), inbox.data(), r->K.data());
        else linear_kernel(nerf::GenerateRenderRays, 0, st, n_rays, box, genuine_box(r), T, genuine_Tow(r), rays.data(), inbox.data(), r->K.data());
        if (rand_dt) CUDA_CHECK_THROW(cudaMemcpyAsync(rdt.data(), rand_dt, (size_t)S2 * n_rays * 4, cudaMemcpyHostToDevice, st));
        else curandGenerateUniform(r->gen, rdt.data(), (size_t)S2 * n_rays);
        linear_kernel(nerf::GenerateRenderInputPoints, 0, st, n_rays, S2, genuine_box(r), rays.data(), inbox.data(), pts.data(), dist.data(), rdt.data());
        CUDA_CHECK_THROW(cudaStreamSynchronize(st));      // :1793 / :1941
        GPUMatrixDynamic<float> in(pts.data(), 3, batch, CM), o4(out4.data(), 4, batch, CM);
        r->network->inference(st, in, o4);
        linear_kernel(nerf::VolumeRender_Render, 0, st, n_rays, S2, 4u, genuine_box(r), nerf::ENerfActivation::Logistic, nerf::ENerfActivation::Exponential, 1.0f,
                      out4.data(), pts.data(), dist.data(), rays.data(), inbox.data(), d_rgb.data(), d_depth.data(), d_mask.data());
        CUDA_CHECK_THROW(cudaEventRecord(e1, st));
        CUDA_CHECK_THROW(cudaStreamSynchronize(st));
        d_rgb.copy_to_host(rgb, 3 * n_rays); d_depth.copy_to_host(depth, n_rays); d_mask.copy_to_host(mask, n_rays);
        if (rays_out) CUDA_CHECK_THROW(cudaMemcpy(rays_out, rays.data(), (size_t)n_rays * 36, cudaMemcpyDeviceToHost));
        if (inbox_out) inbox.copy_to_host(inbox_out, n_rays);
        float ms = 0;
        CUDA_CHECK_THROW(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (device_ms) *device_ms = ms;
    })
}

// NeRF_Model::GenerateToc (nerf_model.cu:2186-2205), the reference's own host code: it reads no member, so it is called on
// raw zeroed storage of the class (no constructor: that one creates CUDA streams).  Toc16: column-major 4x4.
int ref_generate_toc(float theta, float phi, float radius, float* Toc16) {
    static std::vector<unsigned char> storage(sizeof(nerf::NeRF_Model) + 64, 0);
    nerf::NeRF_Model* m = reinterpret_cast<nerf::NeRF_Model*>((reinterpret_cast<uintptr_t>(storage.data()) + 63) & ~(uintptr_t)63);
    const Eigen::Matrix4f T = m->GenerateToc(theta, phi, radius);
    for (int i = 0; i < 16; ++i) Toc16[i] = T.data()[i];
    return 0;
}

// MarchingCubes + compute_mesh_1ring of the reference (marching_cubes.cu:474-510,655-663) on a host density lattice [res^3], x fastest.
// verts/normals: room for verts_cap vertices (xyz); n_verts receives the padded vertex count the reference allocates, n_idx 3 * triangles.
int ref_marching_cubes(void* h, const float* density, uint32_t res, float thresh, float* verts_out, float* normals_out, uint32_t verts_cap, uint32_t* n_verts,
                       uint32_t* idx_out, uint32_t idx_cap, uint32_t* n_idx) {
    Ref* r = static_cast<Ref*>(h);
    GUARD(r, {
        GPUMemory<float> d((size_t)res * res * res);
        d.copy_from_host(density, (size_t)res * res * res);
        GPUMemory<Eigen::Vector3f> verts, normals;
        GPUMemory<Eigen::Vector4f> smoothed;
        GPUMemory<uint32_t> indices;
        nerf::MarchingCubes(genuine_box(r), Eigen::Vector3i((int)res, (int)res, (int)res), thresh, d, verts, indices, r->stream);
        nerf::compute_mesh_1ring(verts, indices, smoothed, normals, r->stream);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        *n_verts = (uint32_t)verts.size(); *n_idx = (uint32_t)indices.size();
        if (verts.size() > verts_cap || indices.size() > idx_cap) throw std::runtime_error("ref_marching_cubes: output buffers too small");
        if (verts.size()) {
            CUDA_CHECK_THROW(cudaMemcpy(verts_out, verts.data(), verts.size() * 12, cudaMemcpyDeviceToHost));
            CUDA_CHECK_THROW(cudaMemcpy(normals_out, normals.data(), normals.size() * 12, cudaMemcpyDeviceToHost));
        }
        if (indices.size()) indices.copy_to_host(idx_out, indices.size());
    })
}
#endif

}  // extern "C"

#ifdef ROMAP_GENUINE
// ---- link-time stand-ins: nerf_model.cu's host methods (never called by this harness) name four OpenGL buffer calls; the
// library has to resolve them to load. They abort if reached.
namespace {
[[noreturn]] void ref_unreachable(const char* what) {
    fprintf(stderr, "oracle/ref harness: %s is a link-time stand-in and must never run\n", what);
    abort();
}
}  // namespace
extern "C" {
void glGenBuffers(GLsizei, GLuint*) { ref_unreachable("glGenBuffers"); }
void glBindBuffer(GLenum, GLuint) { ref_unreachable("glBindBuffer"); }
void glBufferData(GLenum, GLsizeiptr, const void*, GLenum) { ref_unreachable("glBufferData"); }
void glDeleteBuffers(GLsizei, const GLuint*) { ref_unreachable("glDeleteBuffers"); }
}
#endif
