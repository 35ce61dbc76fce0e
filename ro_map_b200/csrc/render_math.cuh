// render_math.cuh — warp-parallel volume rendering, loss and loss-gradient for one ray.
//
// One warp = one ray, one lane = one of the 32 stratified samples.  Replaces the reference's
// one-thread-per-ray serial loops VolumeRender (MON/Core/src/nerf_model.cu:735-815) and
// VolumeRenderGradient_No_Compacted (:817-954): transmittance is an exclusive warp prefix
// product, the accumulated colour/depth are warp prefix sums, the early stop `if (T < 1e-4)
// break` becomes a per-lane predicate (T is monotone non-increasing, so "visited" is a prefix).
// The summation order therefore differs from the serial loop; parity is a stated fp32
// tolerance (tests/test_gpu_parity.py), not bit-exactness.
#pragma once
#include "mon_device.cuh"

#define MON_FULL 0xffffffffu
#define MON_T_EPS 1e-4f

// tcnn::logistic = 1/(1+expf(-x)).  Evaluated with the SFU exponential and reciprocal (relative error ~2^-21, far below
// the fp16 resolution of the logits it is applied to); the density keeps the reference's own __expf.
MON_DEV float mon_logistic(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

MON_DEV float warp_incl_prod(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(MON_FULL, v, o); if (lane >= (uint32_t)o) v *= u; }
    return v;
}
MON_DEV float warp_incl_sum(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(MON_FULL, v, o); if (lane >= (uint32_t)o) v += u; }
    return v;
}

struct RayTargets { float tgt[3]; float tgt_depth; float bg[3]; bool is_obj; };
struct RayResult { float rgb[3]; float depth, mask, loss; };

// logits o[4] (r,g,b,sigma; already rounded through fp16 like the reference's network output),
// t = this lane's sample distance.  Returns dL/dlogits for this lane in dout[4] (fp32, caller
// rounds to fp16) and the per-ray results (identical on every lane).
MON_DEV RayResult warp_render_loss_grad(const float o[4], float t, uint32_t lane, const RayTargets& rt,
                                        float k /* loss_scale / R */, const MonLossCfg& lc, float dout[4]) {
    float t_prev = __shfl_up_sync(MON_FULL, t, 1);
    if (lane == 0) t_prev = 0.0f;  // first interval is measured from the ray origin (nerf_model.cu:769-784)
    const float dt = t - t_prev;
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = mon_logistic(o[c]);
    const float density = __expf(o[3]);                    // network_to_density: unclamped __expf (:49)
    const float alpha = 1.0f - __expf(-density * dt);
    const float om = 1.0f - alpha;
    const float Tincl = warp_incl_prod(om, lane);          // T_{n+1}
    float Texcl = __shfl_up_sync(MON_FULL, Tincl, 1);      // T_n
    if (lane == 0) Texcl = 1.0f;
    const bool visited = Texcl >= MON_T_EPS;
    const float w = visited ? alpha * Texcl : 0.0f;
    float C2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) C2[c] = warp_incl_sum(w * rgb[c], lane);
    const float D2 = warp_incl_sum(w * t, lane);
    const uint32_t vis = __ballot_sync(MON_FULL, visited);
    const int last = 31 - __clz((int)vis);                 // lane 0 is always visited (T_0 = 1)
    const float Tfinal = __shfl_sync(MON_FULL, Tincl, last);
    RayResult rr;
#pragma unroll
    for (int c = 0; c < 3; ++c) rr.rgb[c] = __shfl_sync(MON_FULL, C2[c], 31) + Tfinal * rt.bg[c];
    rr.depth = __shfl_sync(MON_FULL, D2, 31);
    rr.mask = 1.0f - Tfinal;

    // loss (:858-880)
    float g[3], mean_loss = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float diff = rr.rgb[c] - rt.tgt[c]; g[c] = 2.0f * diff; mean_loss += diff * diff; }
    mean_loss = mean_loss / 3.0f;
    float dd = 0.0f;
    if (rt.tgt_depth > 0.0f) dd = lc.depth_lambda * (rr.depth - rt.tgt_depth >= 0.0f ? 1.0f : -1.0f);
    rr.loss = rt.is_obj ? mean_loss + dd * (rr.depth - rt.tgt_depth) + (1.0f - rr.mask) : mean_loss + rr.mask;

    // gradient w.r.t. the 4 logits of this sample (:914-945)
    if (!visited) { dout[0] = dout[1] = dout[2] = dout[3] = 0.0f; return rr; }
#pragma unroll
    for (int c = 0; c < 3; ++c) dout[c] = k * ((w * g[c]) * (rgb[c] * (1.0f - rgb[c])));
    const float dsig = __expf(fminf(fmaxf(o[3], -15.0f), 15.0f));
    const float depth_sup = dd * (Tincl * t - (rr.depth - D2));
    const float dmask_dsig = 1.0f - rr.mask;
    float dmlp;
    if (rt.is_obj) {
        const float dmask = lc.mask_lambda * (rr.mask >= 1.0f ? 1.0f : -1.0f);
        float dot = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) dot += g[c] * (Tincl * rgb[c] - (rr.rgb[c] - C2[c]));
        dmlp = dsig * dt * (dot + depth_sup + dmask * dmask_dsig);
    } else {
        const float dmask = lc.mask_lambda * (rr.mask >= 0.0f ? 1.0f : -1.0f);
        dmlp = dsig * dt * dmask * dmask_dsig + dsig * lc.bg_density_reg;
    }
    dout[3] = k * dmlp;
    return rr;
}

// Inference compositing with carry across 32-sample chunks (VolumeRender_Render, :1134-1229).
struct RenderCarry { float T, C[3], D, last_t; };
MON_DEV void warp_render_chunk(const float o[4], float t, uint32_t lane, RenderCarry& cr) {
    float t_prev = __shfl_up_sync(MON_FULL, t, 1);
    if (lane == 0) t_prev = cr.last_t;
    const float dt = t - t_prev;
    const float alpha = 1.0f - __expf(-__expf(o[3]) * dt);
    const float om = 1.0f - alpha;
    const float Tincl = cr.T * warp_incl_prod(om, lane);
    float Texcl = __shfl_up_sync(MON_FULL, Tincl, 1);
    if (lane == 0) Texcl = cr.T;
    const bool visited = Texcl >= MON_T_EPS;
    const float w = visited ? alpha * Texcl : 0.0f;
    float c3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) c3[c] = __shfl_sync(MON_FULL, warp_incl_sum(w * mon_logistic(o[c]), lane), 31);
    const float d = __shfl_sync(MON_FULL, warp_incl_sum(w * t, lane), 31);
    const uint32_t vis = __ballot_sync(MON_FULL, visited);
    if (vis) {
        const int last = 31 - __clz((int)vis);
        cr.T = __shfl_sync(MON_FULL, Tincl, last);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) cr.C[c] += c3[c];
    cr.D += d;
    cr.last_t = __shfl_sync(MON_FULL, t, 31);
}
