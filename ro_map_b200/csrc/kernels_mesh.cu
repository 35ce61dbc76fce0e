// kernels_mesh.cu — marching cubes on the GPU: the iso-surface of a density lattice as vertices, 1-ring normals and triangle
// indices.  Replaces gen_vertices / gen_faces / accumulate_1ring + the host glue of MarchingCubes and compute_mesh_1ring
// (MON/Core/src/marching_cubes.cu:41-91,93-435,437-472,474-510), with the reference's conventions: one vertex per sign-changing
// lattice edge, owned by the edge's lower lattice point, at (x + (thresh - f0) / (f1 - f0)) * scale + min; the reference's corner /
// edge numbering of a cell (:391-421) and its 256-row triangle table; un-normalised (area weighted) 1-ring normals
// (pb - pa) x (pa - pc) summed per vertex, then normalised; the vertex count padded to a multiple of 128 with zero vertices (:499).
//
// The reference hands out vertex and triangle slots with atomicAdd (its output order is a race).  Here both are EXCLUSIVE SCANS over the
// lattice — count, scan, write — so the order is the lattice order (x fastest, per point the +x, +y, +z edge; per cell the table's
// own triangle order): reproducible, and identical to the CPU statement of the same algorithm (tests/host/mesh_cpu.h, the host
// implementation these kernels replaced, held against the reference's output in tests/test_golden_romap.py), vertex for vertex
// and index for index (tests/test_gpu_mesh.py).  Only the normals depend on an order (float atomics, as in the reference).
#include "mon_kernels.h"
#include "../host/mc_table.h"

namespace {

__constant__ int8_t c_tri[256][16];
// cell corners / edges in the reference's numbering (marching_cubes.cu:393-421); per edge: the corner that owns it and its axis
__constant__ uint8_t c_corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
__constant__ uint8_t c_edge_owner[12] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3};
__constant__ uint8_t c_edge_axis[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};

#define MESH_THREADS 256
#define SCAN_BLOCK 1024     // elements per CTA of the scan (256 threads x 4)

struct Lattice {
    const float* sigma;
    uint32_t res;
    float thresh;
};

__device__ __forceinline__ uint32_t cell_mask(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) {
    const size_t r1 = L.res, r2 = (size_t)L.res * L.res;
    uint32_t mask = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (L.sigma[(x + c_corner[c][0]) + (y + c_corner[c][1]) * r1 + (z + c_corner[c][2]) * r2] > L.thresh) mask |= 1u << c;
    return mask;
}

// per lattice point: number of vertices it owns (sign changes on its +x, +y, +z edges) and number of indices its cell emits
__global__ void k_mc_count(Lattice L, uint32_t* __restrict__ n_vert, uint32_t* __restrict__ n_idx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t r1 = L.res, r2 = r1 * r1, n = r2 * r1;
    if (i >= n) return;
    const uint32_t x = (uint32_t)(i % r1), y = (uint32_t)((i / r1) % r1), z = (uint32_t)(i / r2);
    const bool inside = L.sigma[i] > L.thresh;
    uint32_t nv = 0;
    if (x + 1 < L.res && inside != (L.sigma[i + 1] > L.thresh)) ++nv;
    if (y + 1 < L.res && inside != (L.sigma[i + r1] > L.thresh)) ++nv;
    if (z + 1 < L.res && inside != (L.sigma[i + r2] > L.thresh)) ++nv;
    n_vert[i] = nv;
    uint32_t ni = 0;
    if (x + 1 < L.res && y + 1 < L.res && z + 1 < L.res) {
        const int8_t* tri = c_tri[cell_mask(L, x, y, z)];
        while (ni < 15 && tri[ni] >= 0) ++ni;
    }
    n_idx[i] = ni;
}

// ---- exclusive scan of n uint32 (in place), three small kernels: per-CTA scan + CTA totals, scan of the totals, add-back
__global__ void k_scan_blocks(uint32_t n, uint32_t* __restrict__ data, uint32_t* __restrict__ totals) {
    __shared__ uint32_t warp_sum[MESH_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + k < n ? data[base + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += u; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < MESH_THREADS / 32 ? warp_sum[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += u; }
        if (lane < MESH_THREADS / 32) warp_sum[lane] = wi - w;
        if (lane == MESH_THREADS / 32 - 1) totals[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sum[warp] + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
}
// one CTA: exclusive scan of the CTA totals (any count, in chunks of the CTA's width); the grand total goes to *sum
__global__ void k_scan_totals(uint32_t n_blocks, uint32_t* __restrict__ totals, uint32_t* __restrict__ sum) {
    __shared__ uint32_t warp_sum[MESH_THREADS / 32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += MESH_THREADS) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < n_blocks ? totals[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += u; }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < MESH_THREADS / 32 ? warp_sum[lane] : 0u, wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += u; }
            if (lane < MESH_THREADS / 32) warp_sum[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sum[warp] + incl - v;
        if (i < n_blocks) totals[i] = excl;
        __syncthreads();
        if (threadIdx.x == MESH_THREADS - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *sum = carry;
}
__global__ void k_scan_add(uint32_t n, uint32_t* __restrict__ data, const uint32_t* __restrict__ totals) {
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    const uint32_t add = totals[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + k < n) data[i + k] += add;
}

// gen_vertices (:41-91): the vertices of a lattice point's own edges, in the order +x, +y, +z, at the point's scan offset;
// vid[a][i] = id + 1 of the vertex on edge a of point i (0 = none)
__global__ void k_mc_vertices(Lattice L, const uint32_t* __restrict__ v_off, float3 bmin, float3 scale, float* __restrict__ verts,
                              uint32_t* __restrict__ vid) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t r1 = L.res, r2 = r1 * r1, n = r2 * r1;
    if (i >= n) return;
    const uint32_t c[3] = {(uint32_t)(i % r1), (uint32_t)((i / r1) % r1), (uint32_t)(i / r2)};
    const size_t step[3] = {1, r1, r2};
    const float mn[3] = {bmin.x, bmin.y, bmin.z}, sc[3] = {scale.x, scale.y, scale.z};
    const float f0 = L.sigma[i];
    const bool inside = f0 > L.thresh;
    uint32_t id = v_off[i];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        uint32_t mine = 0u;
        if (c[a] + 1 < L.res) {
            const float f1 = L.sigma[i + step[a]];
            if (inside != (f1 > L.thresh)) {
                const float dt = __fdiv_rn(__fsub_rn(L.thresh, f0), __fsub_rn(f1, f0));
#pragma unroll
                for (int k = 0; k < 3; ++k) verts[3 * (size_t)id + k] = __fmaf_rn(__fadd_rn((float)c[k], k == a ? dt : 0.0f), sc[k], mn[k]);
                mine = ++id;
            }
        }
        vid[i + n * a] = mine;
    }
}

// gen_faces (:93-435): the cell's triangles in the table's order at the cell's scan offset
__global__ void k_mc_faces(Lattice L, const uint32_t* __restrict__ i_off, const uint32_t* __restrict__ vid, uint32_t* __restrict__ indices) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t r1 = L.res, r2 = r1 * r1, n = r2 * r1;
    if (i >= n) return;
    const uint32_t x = (uint32_t)(i % r1), y = (uint32_t)((i / r1) % r1), z = (uint32_t)(i / r2);
    if (x + 1 >= L.res || y + 1 >= L.res || z + 1 >= L.res) return;
    const uint32_t mask = cell_mask(L, x, y, z);
    if (mask == 0u || mask == 255u) return;
    const int8_t* tri = c_tri[mask];
    uint32_t out = i_off[i];
    for (int k = 0; k < 15 && tri[k] >= 0; ++k) {
        const int e = tri[k], lo = c_edge_owner[e], axis = c_edge_axis[e];
        const size_t li = i + c_corner[lo][0] + c_corner[lo][1] * r1 + c_corner[lo][2] * r2;
        indices[out++] = vid[li + n * axis] - 1u;
    }
}

// accumulate_1ring (:437-472): n = (pb - pa) x (pa - pc), un-normalised, added to the triangle's three vertices
__global__ void k_mc_normals(uint32_t n_tri, const uint32_t* __restrict__ indices, const float* __restrict__ verts, float* __restrict__ normals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    const uint32_t a = indices[3 * t], b = indices[3 * t + 1], c = indices[3 * t + 2];
    const float* pa = verts + 3 * (size_t)a; const float* pb = verts + 3 * (size_t)b; const float* pc = verts + 3 * (size_t)c;
    const float u[3] = {__fsub_rn(pb[0], pa[0]), __fsub_rn(pb[1], pa[1]), __fsub_rn(pb[2], pa[2])};
    const float w[3] = {__fsub_rn(pa[0], pc[0]), __fsub_rn(pa[1], pc[1]), __fsub_rn(pa[2], pc[2])};
    const float nrm[3] = {__fsub_rn(__fmul_rn(u[1], w[2]), __fmul_rn(u[2], w[1])), __fsub_rn(__fmul_rn(u[2], w[0]), __fmul_rn(u[0], w[2])),
                          __fsub_rn(__fmul_rn(u[0], w[1]), __fmul_rn(u[1], w[0]))};
    const uint32_t vs[3] = {a, b, c};
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(normals + 3 * (size_t)vs[v] + k, nrm[k]);
}
__global__ void k_mc_normalize(uint32_t n_verts, float* __restrict__ normals) {   // Eigen's normalized() leaves a zero vector alone
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_verts) return;
    float* nv = normals + 3 * (size_t)v;
    const float z2 = __fadd_rn(__fadd_rn(__fmul_rn(nv[0], nv[0]), __fmul_rn(nv[1], nv[1])), __fmul_rn(nv[2], nv[2]));
    if (z2 > 0.0f) {
        const float len = __fsqrt_rn(z2);
        nv[0] = __fdiv_rn(nv[0], len); nv[1] = __fdiv_rn(nv[1], len); nv[2] = __fdiv_rn(nv[2], len);
    }
}

// compute_mesh_vertex_colors (nerf_model.cu:2045-2067): vertices -> unit cube (WarpPoint), and the network's rgb logits -> u8
__global__ void k_mesh_unit_points(uint32_t n_verts, const float* __restrict__ verts, float3 bmin, float3 bmax, float* __restrict__ unit) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_verts) return;
    unit[3 * (size_t)v + 0] = __fdiv_rn(__fsub_rn(verts[3 * (size_t)v + 0], bmin.x), __fsub_rn(bmax.x, bmin.x));
    unit[3 * (size_t)v + 1] = __fdiv_rn(__fsub_rn(verts[3 * (size_t)v + 1], bmin.y), __fsub_rn(bmax.y, bmin.y));
    unit[3 * (size_t)v + 2] = __fdiv_rn(__fsub_rn(verts[3 * (size_t)v + 2], bmin.z), __fsub_rn(bmax.z, bmin.z));
}
__global__ void k_mesh_colors(uint32_t n_verts, const float* __restrict__ out4, uint8_t* __restrict__ colors) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_verts) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float c = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-out4[4 * (size_t)v + k])));
        colors[3 * (size_t)v + k] = (uint8_t)fminf(fmaxf(__fmul_rn(c, 255.0f), 0.0f), 255.0f);
    }
}

cudaError_t scan_exclusive(uint32_t n, uint32_t* data, uint32_t* totals, uint32_t* sum, cudaStream_t st) {
    const uint32_t blocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    k_scan_blocks<<<blocks, MESH_THREADS, 0, st>>>(n, data, totals);
    k_scan_totals<<<1, MESH_THREADS, 0, st>>>(blocks, totals, sum);
    k_scan_add<<<blocks, MESH_THREADS, 0, st>>>(n, data, totals);
    return cudaGetLastError();
}

}  // namespace

size_t mon_mesh_scan_scratch_words(uint32_t res) {
    const size_t n = (size_t)res * res * res;
    return 2 * ((n + SCAN_BLOCK - 1) / SCAN_BLOCK) + 2;
}

// Phase 1: counts + scans.  v_off / i_off: [res^3] each (become the exclusive offsets), totals: mon_mesh_scan_scratch_words(res)
// words, sums: [2] device words = number of surface vertices, number of indices.
cudaError_t mon_launch_mc_count(const float* sigma, uint32_t res, float thresh, uint32_t* v_off, uint32_t* i_off, uint32_t* totals, uint32_t* sums,
                                cudaStream_t st) {
    // the 4 KB table goes up with every extraction, on the caller's stream: a one-time synchronous cudaMemcpyToSymbol would go through
    // the legacy default stream (concurrent calls write the same bytes)
    const cudaError_t prep = cudaMemcpyToSymbolAsync(c_tri, mesh::mc::TRIANGLES, sizeof(c_tri), 0, cudaMemcpyHostToDevice, st);
    if (prep != cudaSuccess) return prep;
    const size_t n = (size_t)res * res * res;
    const Lattice L = {sigma, res, thresh};
    k_mc_count<<<(unsigned)((n + MESH_THREADS - 1) / MESH_THREADS), MESH_THREADS, 0, st>>>(L, v_off, i_off);
    const size_t half = (n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1;
    cudaError_t e = scan_exclusive((uint32_t)n, v_off, totals, sums, st);
    if (e != cudaSuccess) return e;
    return scan_exclusive((uint32_t)n, i_off, totals + half, sums + 1, st);
}

// Phase 2: vertices (n_verts_padded x 3, the padding zeroed here), vid scratch [3 * res^3], indices, normals
cudaError_t mon_launch_mc_build(const float* sigma, uint32_t res, float thresh, const float bmin[3], const float bmax[3], const uint32_t* v_off,
                                const uint32_t* i_off, uint32_t n_surface, uint32_t n_verts_padded, uint32_t n_indices, uint32_t* vid, float* verts,
                                float* normals, uint32_t* indices, cudaStream_t st) {
    const size_t n = (size_t)res * res * res;
    const Lattice L = {sigma, res, thresh};
    const float3 mn = make_float3(bmin[0], bmin[1], bmin[2]);
    const float3 sc = make_float3((bmax[0] - bmin[0]) / (float)(res - 1), (bmax[1] - bmin[1]) / (float)(res - 1), (bmax[2] - bmin[2]) / (float)(res - 1));
    cudaError_t e;
    if (n_verts_padded) {
        if ((e = cudaMemsetAsync(verts + 3 * (size_t)n_surface, 0, (size_t)(n_verts_padded - n_surface) * 12, st)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(normals, 0, (size_t)n_verts_padded * 12, st)) != cudaSuccess) return e;
    }
    const unsigned blocks = (unsigned)((n + MESH_THREADS - 1) / MESH_THREADS);
    k_mc_vertices<<<blocks, MESH_THREADS, 0, st>>>(L, v_off, mn, sc, verts, vid);
    k_mc_faces<<<blocks, MESH_THREADS, 0, st>>>(L, i_off, vid, indices);
    if (n_indices) {
        k_mc_normals<<<(n_indices / 3 + MESH_THREADS - 1) / MESH_THREADS, MESH_THREADS, 0, st>>>(n_indices / 3, indices, verts, normals);
        k_mc_normalize<<<(n_verts_padded + MESH_THREADS - 1) / MESH_THREADS, MESH_THREADS, 0, st>>>(n_verts_padded, normals);
    }
    return cudaGetLastError();
}

void mon_launch_mesh_unit_points(uint32_t n_verts, const float* verts, const float bmin[3], const float bmax[3], float* unit, cudaStream_t st) {
    if (!n_verts) return;
    k_mesh_unit_points<<<(n_verts + MESH_THREADS - 1) / MESH_THREADS, MESH_THREADS, 0, st>>>(n_verts, verts, make_float3(bmin[0], bmin[1], bmin[2]),
                                                                                              make_float3(bmax[0], bmax[1], bmax[2]), unit);
}
void mon_launch_mesh_colors(uint32_t n_verts, const float* out4, uint8_t* colors, cudaStream_t st) {
    if (!n_verts) return;
    k_mesh_colors<<<(n_verts + MESH_THREADS - 1) / MESH_THREADS, MESH_THREADS, 0, st>>>(n_verts, out4, colors);
}
