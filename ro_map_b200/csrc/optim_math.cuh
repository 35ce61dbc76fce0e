// optim_math.cuh — the per-parameter optimizer arithmetic of the optimizer sweep (kernels_optim.cu).
//   adam_step<__half>                         TCNN optimizers/adam.h:48-118
//   ema_step_half_precision<__half>           TCNN optimizers/ema.h:62-76,102-136
#pragma once
#include "mon_device.cuh"

// Adam's bias correction sqrt(1 - beta2^s) / (1 - beta1^s) depends on the parameter's own step count s only
// (adam.h:103-104): the sweep reads it from a per-object table indexed by s that mon_core.cu fills on the HOST with the
// reference's expression, sqrtf(1 - powf(beta2, s)) / (1 - powf(beta1, s)).  That replaces ~45 instructions per touched
// parameter by one cached load, and it makes the value identical to the CPU restatement's: 1 - beta^s cancels, so one
// ulp of difference between two pow implementations is 3e-6 of the learning rate.  Steps beyond the table (32768
// updates of one parameter) evaluate beta^s as exp2f(s * log2 beta) on the device.
__device__ __forceinline__ float adam_debias(const MonOpt& o, uint32_t cs) {
    const float b1s = exp2f((float)cs * o.log2_beta1), b2s = exp2f((float)cs * o.log2_beta2);
    return __fdiv_rn(__fsqrt_rn(1.0f - b2s), 1.0f - b1s);
}

// one Adam update (adam.h:65-118); returns the new weight
__device__ __forceinline__ float adam_one(const MonOpt& o, float lr_base, float gradient, bool is_mlp, float w, float& m, float& v, uint32_t& cs) {
    if (is_mlp) gradient = __fmaf_rn(o.l2_reg, w, gradient);
    const float gsq = __fmul_rn(gradient, gradient);
    m = __fmaf_rn(o.beta1, m, __fmul_rn(1.0f - o.beta1, gradient));
    v = __fmaf_rn(o.beta2, v, __fmul_rn(1.0f - o.beta2, gsq));
    cs = min(cs + 1u, 65535u);      // 16-bit counter in memory (see OptimPtrs)
    const float lr = __fmul_rn(lr_base, cs < o.n_debias_lut ? __ldg(o.debias_lut + cs) : adam_debias(o, cs));
    // IEEE sqrt and division like the reference's sqrtf and '/' (adam.h:107): with identical gradients the weights stay
    // bit-identical to the CPU restatement (the SFU approximations would save ~12 instructions and cost that property)
    const float eff = fminf(fmaxf(__fdiv_rn(lr, __fadd_rn(__fsqrt_rn(v), o.eps)), 0.0f), FLT_MAX);
    return __fmaf_rn(-eff, m, w);
}

// ps: per-parameter step counters as SATURATING 16-bit integers (the reference keeps uint32, adam.h:58): 4 bytes less read + written
// per touched parameter of a sweep that is bound by the bytes it moves.  The count only enters through the bias correction
// sqrt(1 - beta2^s) / (1 - beta1^s), which is exactly 1.0f in fp32 long before 65535 updates of one parameter for beta <= 0.9997
// (beta^s < 2^-25; base.json: 0.9 / 0.99 -> from s = 1656 on), so the saturated counter gives bit-identical weights.
struct OptimPtrs { float* pf; __half* ph; float* m; float* v; uint16_t* ps; __half* ema; };

// Adam + EMA for 4 consecutive parameters starting at i4 (= 2 table entries).  g: their loss-scaled gradients (fp16 values
// widened to float); wraw / eraw: their fp16 weights and EMA weights (always needed, so the caller fetches them beside the
// gradient).  Grid parameters with a zero gradient are skipped by Adam (adam.h:75-79) but still EMA-filtered.  planar_f0 (grid
// only): the first entry's slot in the feature-0 array of the planar weight copy, feature 1 lies planar_stride halves further.
__device__ __forceinline__ void optim_quad(const MonOpt& o, float lr_base, float old_db, float new_db, bool is_mlp, uint32_t i4, float (&g)[4],
                                           uint2 wraw, uint2 eraw, const OptimPtrs& p, __half* planar_f0, uint32_t planar_stride) {
    bool touched[4];
    bool any = is_mlp;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        g[k] = o.loss_scale_pow2 ? __fmul_rn(g[k], o.inv_loss_scale) : __fdiv_rn(g[k], o.loss_scale);
        touched[k] = is_mlp || g[k] != 0.0f;       // grid: zero gradient => Adam skips the parameter (adam.h:75-79)
        any |= touched[k];
    }

    // ---- fp16 weights of the 4 parameters (needed by the EMA in any case)
    __half wh[4] = {__ushort_as_half((unsigned short)(wraw.x & 0xffffu)), __ushort_as_half((unsigned short)(wraw.x >> 16)),
                    __ushort_as_half((unsigned short)(wraw.y & 0xffffu)), __ushort_as_half((unsigned short)(wraw.y >> 16))};
    if (any) {
        float4 w4 = *reinterpret_cast<const float4*>(p.pf + i4);
        float4 m4 = *reinterpret_cast<const float4*>(p.m + i4);
        float4 v4 = *reinterpret_cast<const float4*>(p.v + i4);
        const uint2 sraw = *reinterpret_cast<const uint2*>(p.ps + i4);
        uint32_t sp[4] = {sraw.x & 0xffffu, sraw.x >> 16, sraw.y & 0xffffu, sraw.y >> 16};
        float* wp = &w4.x; float* mp = &m4.x; float* vp = &v4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (touched[k]) {
                wp[k] = adam_one(o, lr_base, g[k], is_mlp, wp[k], mp[k], vp[k], sp[k]);
                wh[k] = __float2half_rn(wp[k]);
            }
        }
        *reinterpret_cast<float4*>(p.pf + i4) = w4;
        *reinterpret_cast<float4*>(p.m + i4) = m4;
        *reinterpret_cast<float4*>(p.v + i4) = v4;
        *reinterpret_cast<uint2*>(p.ps + i4) = make_uint2(sp[0] | (sp[1] << 16), sp[2] | (sp[3] << 16));
        if (planar_f0) {
            // grid: the planar copy is the fp16 working copy (2 bytes less written per touched parameter than keeping the interleaved one too)
            *reinterpret_cast<__half2*>(planar_f0) = __halves2half2(wh[0], wh[2]);
            *reinterpret_cast<__half2*>(planar_f0 + planar_stride) = __halves2half2(wh[1], wh[3]);
        } else {
            const __half2 a = __halves2half2(wh[0], wh[1]), b = __halves2half2(wh[2], wh[3]);
            *reinterpret_cast<uint2*>(p.ph + i4) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        }
    }

    // ---- EMA over all params with the global step (ema.h:62-76)
    const __half2 e01 = *reinterpret_cast<const __half2*>(&eraw.x), e23 = *reinterpret_cast<const __half2*>(&eraw.y);
    const float ev[4] = {__low2float(e01), __high2float(e01), __low2float(e23), __high2float(e23)};
    float nf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        nf[k] = __fmul_rn(__fmaf_rn(__half2float(wh[k]), 1.0f - o.ema_decay, __fmul_rn(__fmul_rn(ev[k], o.ema_decay), old_db)), new_db);
    const __half2 n01 = __floats2half2_rn(nf[0], nf[1]), n23 = __floats2half2_rn(nf[2], nf[3]);
    *reinterpret_cast<uint2*>(p.ema + i4) = make_uint2(*reinterpret_cast<const uint32_t*>(&n01), *reinterpret_cast<const uint32_t*>(&n23));
}
