// kernels_mlp_wmma.cu — LEGACY-TENSOR-PATH validation kernel (mma.sync via nvcuda::wmma).
//
// Same fusion as the tcgen05 product kernel (kernels_mlp_tc.cu): per ray, MLP forward ->
// volume render -> loss gradient -> MLP backward (dgrad + wgrad) without any activation leaving
// the SM; it exists so the tcgen05 kernel can be cross-checked on the device and so the first
// end-to-end parity run does not depend on hand-built UMMA descriptors.  Selected only through
// mon_object_set_mlp_impl(obj, 1).  fp16 operands, fp32 accumulation (the reference accumulates
// in fp16, TCNN/src/fully_fused_mlp.cu:68,198; see DESIGN.md "precision").
//
// Reference counterparts: kernel_mlp_fused (fully_fused_mlp.cu:499-557), VolumeRender /
// VolumeRenderGradient_No_Compacted (MON/Core/src/nerf_model.cu:735-954),
// kernel_mlp_fused_backward (:150-259), the three CUTLASS GEMMs of backward_impl (:785-834).
#include <mma.h>

#include "mon_kernels.h"
#include "render_math.cuh"

using namespace nvcuda;

#define WM_WARPS 4
#define WM_THREADS (WM_WARPS * 32)

struct alignas(32) WarpTile {
    __half enc[32][40];
    __half hid[32][72];
    __half dhid[32][72];
    __half dout[32][24];
    float stage[32][68];
};
struct alignas(32) CtaSmem {
    WarpTile w[WM_WARPS];
    __half Win[64 * 32];
    __half Wout[16 * 64];
    float dW[64 * 32 + 16 * 64];
};

typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> FragA;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::col_major> FragAT;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> FragB;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> FragBT;
typedef wmma::fragment<wmma::accumulator, 16, 16, 16, float> FragC;

__device__ __forceinline__ void warp_mlp_forward(WarpTile& ws, const __half* sWin, const __half* sWout, uint32_t lane) {
    // hidden = relu(enc[32x32] * W_in^T), W_in is [64][32] row-major == col-major B (K x N) with ld 32
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            FragC acc; wmma::fill_fragment(acc, 0.0f);
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                FragA a; FragBT b;
                wmma::load_matrix_sync(a, &ws.enc[mt * 16][kt * 16], 40);
                wmma::load_matrix_sync(b, sWin + nt * 16 * 32 + kt * 16, 32);
                wmma::mma_sync(acc, a, b, acc);
            }
            wmma::store_matrix_sync(&ws.stage[mt * 16][nt * 16], acc, 68, wmma::mem_row_major);
        }
    __syncwarp();
    for (uint32_t idx = lane; idx < 32 * 64; idx += 32) {
        const uint32_t r = idx >> 6, c = idx & 63;
        ws.hid[r][c] = __float2half_rn(fmaxf(ws.stage[r][c], 0.0f));
    }
    __syncwarp();
    // out = hid[32x64] * W_out^T, W_out is [16][64] row-major == col-major B with ld 64
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        FragC acc; wmma::fill_fragment(acc, 0.0f);
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            FragA a; FragBT b;
            wmma::load_matrix_sync(a, &ws.hid[mt * 16][kt * 16], 72);
            wmma::load_matrix_sync(b, sWout + kt * 16, 64);
            wmma::mma_sync(acc, a, b, acc);
        }
        wmma::store_matrix_sync(&ws.stage[mt * 16][0], acc, 68, wmma::mem_row_major);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WM_THREADS)
k_mlp_train_wmma(MonBatch b, MonLossCfg lc, uint32_t n_mlp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* partial = b.mlp_partials + (size_t)blockIdx.x * n_mlp;
    if (b.ctrl->skip) return;

    for (uint32_t i = tid; i < 64 * 32; i += WM_THREADS) sm.Win[i] = b.params[i];
    for (uint32_t i = tid; i < 16 * 64; i += WM_THREADS) sm.Wout[i] = b.params[64 * 32 + i];
    for (uint32_t i = tid; i < 64 * 32 + 16 * 64; i += WM_THREADS) sm.dW[i] = 0.0f;
    WarpTile& ws = sm.w[warp];
    for (uint32_t i = lane; i < 32 * 24; i += 32) (&ws.dout[0][0])[i] = __float2half_rn(0.0f);
    __syncthreads();

    FragC accWout[4], accWin[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { wmma::fill_fragment(accWout[i], 0.0f); wmma::fill_fragment(accWin[i][0], 0.0f); wmma::fill_fragment(accWin[i][1], 0.0f); }

    const float k = lc.loss_scale / (float)b.R;
    const uint32_t iter = b.ctrl->iter - 1;

    for (uint32_t ray = blockIdx.x * WM_WARPS + warp; ray < b.R; ray += gridDim.x * WM_WARPS) {
        // 1. stage this ray's 32x32 fp16 encoding tile (2 KB contiguous)
        const uint4* src = reinterpret_cast<const uint4*>(b.enc + (size_t)ray * 32 * MON_IN);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = lane + 32 * j;
            *reinterpret_cast<uint4*>(&ws.enc[idx >> 2][(idx & 3) * 8]) = src[idx];
        }
        __syncwarp();
        // 2. forward
        warp_mlp_forward(ws, sm.Win, sm.Wout, lane);
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = __half2float(__float2half_rn(ws.stage[lane][c]));  // network output is fp16
        // 3. render + loss + dL/dout, lane = sample
        const MonRay r = b.rays[ray];
        const uint32_t pt = ray * 32 + lane;
        const float xi = mon_rand(b.inj_dt, b.seed, iter, 2, pt);
        const float t = mon_sample_t(r, lane, xi, 32.0f);
        RayTargets rt;
#pragma unroll
        for (int c = 0; c < 3; ++c) { rt.tgt[c] = b.target[ray * 3 + c]; rt.bg[c] = b.bg[ray * 3 + c]; }
        rt.tgt_depth = b.target_depth[ray];
        rt.is_obj = b.ray_inst[ray] == 1;
        float go[4];
        const RayResult rr = warp_render_loss_grad(o, t, lane, rt, k, lc, go);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) b.rgb_rays[ray * 3 + c] = rr.rgb[c];
            b.depth_rays[ray] = rr.depth; b.mask_rays[ray] = rr.mask; b.loss[ray] = rr.loss;
        }
        __half gh[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { gh[c] = __float2half_rn(go[c]); ws.dout[lane][c] = gh[c]; }
        if (b.dbg_out) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { b.dbg_out[(size_t)pt * 4 + c] = o[c]; b.dbg_dout[(size_t)pt * 4 + c] = __half2float(gh[c]); }
        }
        __syncwarp();
        // 4. d_hid = (dout[32x16] * W_out[16x64]) masked by relu
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                FragC acc; wmma::fill_fragment(acc, 0.0f);
                FragA a; FragB bb;
                wmma::load_matrix_sync(a, &ws.dout[mt * 16][0], 24);
                wmma::load_matrix_sync(bb, sm.Wout + nt * 16, 64);
                wmma::mma_sync(acc, a, bb, acc);
                wmma::store_matrix_sync(&ws.stage[mt * 16][nt * 16], acc, 68, wmma::mem_row_major);
            }
        __syncwarp();
        for (uint32_t idx = lane; idx < 32 * 64; idx += 32) {
            const uint32_t rr2 = idx >> 6, c = idx & 63;
            ws.dhid[rr2][c] = __half2float(ws.hid[rr2][c]) > 0.0f ? __float2half_rn(ws.stage[rr2][c]) : __float2half_rn(0.0f);
        }
        __syncwarp();
        // 5. weight gradients, accumulated across this warp's rays
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                FragAT a; FragB bb;
                wmma::load_matrix_sync(a, &ws.dout[kt * 16][0], 24);        // dout^T
                wmma::load_matrix_sync(bb, &ws.hid[kt * 16][nt * 16], 72);
                wmma::mma_sync(accWout[nt], a, bb, accWout[nt]);
            }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int kt = 0; kt < 2; ++kt) {
                    FragAT a; FragB bb;
                    wmma::load_matrix_sync(a, &ws.dhid[kt * 16][mt * 16], 72);  // dhid^T
                    wmma::load_matrix_sync(bb, &ws.enc[kt * 16][nt * 16], 40);
                    wmma::mma_sync(accWin[mt][nt], a, bb, accWin[mt][nt]);
                }
        // 6. d_enc = dhid[32x64] * W_in[64x32]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                FragC acc; wmma::fill_fragment(acc, 0.0f);
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) {
                    FragA a; FragB bb;
                    wmma::load_matrix_sync(a, &ws.dhid[mt * 16][kt * 16], 72);
                    wmma::load_matrix_sync(bb, sm.Win + kt * 16 * 32 + nt * 16, 32);
                    wmma::mma_sync(acc, a, bb, acc);
                }
                wmma::store_matrix_sync(&ws.stage[mt * 16][nt * 16], acc, 68, wmma::mem_row_major);
            }
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(b.d_enc + (size_t)ray * 32 * MON_IN);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = lane + 32 * j, row = idx >> 2, c0 = (idx & 3) * 8;
            __half2 h[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(ws.stage[row][c0 + 2 * q], ws.stage[row][c0 + 2 * q + 1]);
            dst[idx] = *reinterpret_cast<uint4*>(h);
        }
        __syncwarp();
    }

    // 7. reduce the 4 warps' weight-gradient fragments in shared memory, then one partial row per CTA
    for (int wsel = 0; wsel < WM_WARPS; ++wsel) {
        if ((int)warp == wsel) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    wmma::store_matrix_sync(&ws.stage[0][0], accWin[mt][nt], 16, wmma::mem_row_major);
                    __syncwarp();
                    for (uint32_t idx = lane; idx < 256; idx += 32) {
                        const uint32_t rr2 = idx >> 4, c = idx & 15;
                        sm.dW[(mt * 16 + rr2) * 32 + nt * 16 + c] += (&ws.stage[0][0])[idx];
                    }
                    __syncwarp();
                }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                wmma::store_matrix_sync(&ws.stage[0][0], accWout[nt], 16, wmma::mem_row_major);
                __syncwarp();
                for (uint32_t idx = lane; idx < 256; idx += 32) {
                    const uint32_t rr2 = idx >> 4, c = idx & 15;
                    sm.dW[64 * 32 + rr2 * 64 + nt * 16 + c] += (&ws.stage[0][0])[idx];
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    for (uint32_t i = tid; i < n_mlp; i += WM_THREADS) partial[i] = sm.dW[i];
}

// ---- inference: MLP forward + compositing for test renders (VolumeRender_Render, :1134-1229)
__global__ void __launch_bounds__(WM_THREADS)
k_mlp_render_wmma(uint32_t n_rays, uint32_t S2, const MonRay* __restrict__ rays, const int* __restrict__ in_box,
                  const float* __restrict__ jitter, uint32_t seed, uint32_t iter, const __half* __restrict__ params,
                  const __half* __restrict__ enc, float bgc, float* __restrict__ rgb, float* __restrict__ depth,
                  float* __restrict__ mask) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 64 * 32; i += WM_THREADS) sm.Win[i] = params[i];
    for (uint32_t i = tid; i < 16 * 64; i += WM_THREADS) sm.Wout[i] = params[64 * 32 + i];
    __syncthreads();
    WarpTile& ws = sm.w[warp];
    for (uint32_t ray = blockIdx.x * WM_WARPS + warp; ray < n_rays; ray += gridDim.x * WM_WARPS) {
        if (!in_box[ray]) {
            if (lane == 0) { rgb[ray * 3] = rgb[ray * 3 + 1] = rgb[ray * 3 + 2] = bgc; depth[ray] = 0.0f; mask[ray] = 0.0f; }
            continue;
        }
        const MonRay r = rays[ray];
        RenderCarry cr; cr.T = 1.0f; cr.C[0] = cr.C[1] = cr.C[2] = 0.0f; cr.D = 0.0f; cr.last_t = 0.0f;
        for (uint32_t chunk = 0; chunk < S2 / 32; ++chunk) {
            const uint32_t pt0 = ray * S2 + chunk * 32;
            const uint4* src = reinterpret_cast<const uint4*>(enc + (size_t)pt0 * MON_IN);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t idx = lane + 32 * j;
                *reinterpret_cast<uint4*>(&ws.enc[idx >> 2][(idx & 3) * 8]) = src[idx];
            }
            __syncwarp();
            warp_mlp_forward(ws, sm.Win, sm.Wout, lane);
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) o[c] = __half2float(__float2half_rn(ws.stage[lane][c]));
            const uint32_t n = chunk * 32 + lane;
            const float xi = mon_rand(jitter, seed, iter, 3, pt0 + lane);
            const float t = mon_sample_t(r, n, xi, (float)S2);
            warp_render_chunk(o, t, lane, cr);
            __syncwarp();
        }
        if (lane == 0) {
            if (1.0f - cr.T > 0.5f) {
#pragma unroll
                for (int c = 0; c < 3; ++c) rgb[ray * 3 + c] = cr.C[c] + cr.T * bgc;
                depth[ray] = __fdiv_rn(cr.D, r.d_norm);
                mask[ray] = 1.0f;
            } else {
                rgb[ray * 3] = rgb[ray * 3 + 1] = rgb[ray * 3 + 2] = bgc; depth[ray] = 0.0f; mask[ray] = 0.0f;
            }
        }
    }
}

// raw network output (4 logits per point) for the density grid / parity hooks
__global__ void __launch_bounds__(WM_THREADS)
k_mlp_infer_wmma(uint32_t n_points, const __half* __restrict__ params, const __half* __restrict__ enc, float* __restrict__ out4) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 64 * 32; i += WM_THREADS) sm.Win[i] = params[i];
    for (uint32_t i = tid; i < 16 * 64; i += WM_THREADS) sm.Wout[i] = params[64 * 32 + i];
    __syncthreads();
    WarpTile& ws = sm.w[warp];
    const uint32_t n_tiles = (n_points + 31) / 32;
    for (uint32_t tile = blockIdx.x * WM_WARPS + warp; tile < n_tiles; tile += gridDim.x * WM_WARPS) {
        const uint32_t pt0 = tile * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = lane + 32 * j, row = idx >> 2;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (pt0 + row < n_points) v = reinterpret_cast<const uint4*>(enc + (size_t)pt0 * MON_IN)[idx];
            *reinterpret_cast<uint4*>(&ws.enc[row][(idx & 3) * 8]) = v;
        }
        __syncwarp();
        warp_mlp_forward(ws, sm.Win, sm.Wout, lane);
        if (pt0 + lane < n_points) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out4[(size_t)(pt0 + lane) * 4 + c] = __half2float(__float2half_rn(ws.stage[lane][c]));
        }
        __syncwarp();
    }
}

static size_t wmma_smem_bytes() { return sizeof(CtaSmem); }

cudaError_t mon_launch_mlp_train_wmma(const MonBatch& b, const MonLossCfg& lc, uint32_t n_mlp, uint32_t n_ctas, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mlp_train_wmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wmma_smem_bytes());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k_mlp_train_wmma<<<n_ctas, WM_THREADS, wmma_smem_bytes(), st>>>(b, lc, n_mlp);
    return cudaGetLastError();
}
cudaError_t mon_launch_mlp_render_wmma(uint32_t n_rays, uint32_t S2, const MonRay* rays, const int* in_box, const float* jitter,
                                       uint32_t seed, uint32_t iter, const __half* params, const __half* enc, float bgc,
                                       float* rgb, float* depth, float* mask, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mlp_render_wmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wmma_smem_bytes());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    uint32_t ctas = (n_rays + WM_WARPS - 1) / WM_WARPS;
    if (ctas > 148 * 2 * 4) ctas = 148 * 2 * 4;
    if (ctas == 0) ctas = 1;
    k_mlp_render_wmma<<<ctas, WM_THREADS, wmma_smem_bytes(), st>>>(n_rays, S2, rays, in_box, jitter, seed, iter, params, enc, bgc, rgb, depth, mask);
    return cudaGetLastError();
}
cudaError_t mon_launch_mlp_infer_wmma(uint32_t n_points, const __half* params, const __half* enc, float* out4, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mlp_infer_wmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wmma_smem_bytes());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    uint32_t ctas = ((n_points + 31) / 32 + WM_WARPS - 1) / WM_WARPS;
    if (ctas > 148 * 8) ctas = 148 * 8;
    if (ctas == 0) ctas = 1;
    k_mlp_infer_wmma<<<ctas, WM_THREADS, wmma_smem_bytes(), st>>>(n_points, params, enc, out4);
    return cudaGetLastError();
}
