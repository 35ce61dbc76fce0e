// mesh_cpu.h — TEST INFRASTRUCTURE: the CPU statement of the mesh extraction the B200 core runs on the GPU (csrc/kernels_mesh.cu),
// i.e. gen_vertices / gen_faces / accumulate_1ring + the host glue of the reference's MarchingCubes and compute_mesh_1ring
// (MON/Core/src/marching_cubes.cu:41-91,93-435,437-472,474-510) with the scan order the kernels use (lattice order).  Held against the
// reference's own output by tests/test_golden_romap.py (mesh_golden.cpp); the GPU kernels are held against it element for element
// by tests/test_gpu_mesh.py; mesh_check.cpp runs it on analytic fields.  Nothing in ro_map_b200/ includes this file.
#pragma once
#include <cmath>

#include "mesh.h"

namespace mesh_cpu {
using mesh::Extracted;
namespace mc = mesh::mc;

// the iso-surface of a lattice alone (no colours)
inline void marching_cubes(const float* sigma, uint32_t res, const float bmin[3], const float bmax[3], float thresh, Extracted& out) {
    out = Extracted();
    const size_t res1 = res, res2 = (size_t)res * res, res3 = res2 * res;
    float scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = (bmax[k] - bmin[k]) / (float)(res - 1);
    // gen_vertices: lattice point (x,y,z) owns its +x, +y, +z edges; vertex ids are stored +1 (0 = none)
    std::vector<uint32_t> vid(res3 * 3, 0);
    for (uint32_t z = 0; z < res; ++z)
        for (uint32_t y = 0; y < res; ++y)
            for (uint32_t x = 0; x < res; ++x) {
                const size_t idx = x + y * res1 + z * res2;
                const float f0 = sigma[idx];
                const bool inside = f0 > thresh;
                const uint32_t c[3] = {x, y, z};
                const size_t step[3] = {1, res1, res2};
                for (int a = 0; a < 3; ++a) {
                    if (c[a] + 1 >= res) continue;
                    const float f1 = sigma[idx + step[a]];
                    if (inside == (f1 > thresh)) continue;
                    const float dt = (thresh - f0) / (f1 - f0);
                    vid[idx + res3 * a] = (uint32_t)(out.verts.size() / 3) + 1;
                    for (int k = 0; k < 3; ++k) out.verts.push_back(std::fmaf((float)c[k] + (k == a ? dt : 0.0f), scale[k], bmin[k]));
                }
            }
    // gen_faces
    for (uint32_t z = 0; z + 1 < res; ++z)
        for (uint32_t y = 0; y + 1 < res; ++y)
            for (uint32_t x = 0; x + 1 < res; ++x) {
                const size_t idx = x + y * res1 + z * res2;
                int mask = 0;
                for (int c = 0; c < 8; ++c)
                    if (sigma[idx + mc::CORNER[c][0] + mc::CORNER[c][1] * res1 + mc::CORNER[c][2] * res2] > thresh) mask |= 1 << c;
                if (mask == 0 || mask == 255) continue;
                const int8_t* tri = mc::TRIANGLES[mask];
                for (int i = 0; i < 15 && tri[i] >= 0; ++i) {
                    const int e = tri[i], a = mc::EDGE[e][0], b = mc::EDGE[e][1];
                    const int axis = mc::CORNER[a][0] != mc::CORNER[b][0] ? 0 : (mc::CORNER[a][1] != mc::CORNER[b][1] ? 1 : 2);
                    const int lo = (mc::CORNER[a][axis] == 0) ? a : b;   // the lattice point that owns the edge
                    const size_t li = idx + mc::CORNER[lo][0] + mc::CORNER[lo][1] * res1 + mc::CORNER[lo][2] * res2;
                    out.indices.push_back(vid[li + res3 * axis] - 1);
                }
            }
    out.n_surface_verts = (uint32_t)(out.verts.size() / 3);
    const size_t nv = ((size_t)out.n_surface_verts + 127) & ~(size_t)127;   // "round for later nn stuff" (marching_cubes.cu:499): zero vertices
    out.verts.resize(nv * 3, 0.0f);
    // accumulate_1ring: n = (pb - pa) x (pa - pc), un-normalised (area weighted), summed into the three corners; then normalised
    out.normals.assign(nv * 3, 0.0f);
    for (size_t i = 0; i + 2 < out.indices.size(); i += 3) {
        const uint32_t a = out.indices[i], b = out.indices[i + 1], c = out.indices[i + 2];
        const float* pa = &out.verts[3 * a]; const float* pb = &out.verts[3 * b]; const float* pc = &out.verts[3 * c];
        const float u[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, w[3] = {pa[0] - pc[0], pa[1] - pc[1], pa[2] - pc[2]};
        const float n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
        for (uint32_t v : {a, b, c}) for (int k = 0; k < 3; ++k) out.normals[3 * v + k] += n[k];
    }
    for (size_t v = 0; v < nv; ++v) {   // trans_mesh_data: Eigen's normalized() leaves a zero vector alone
        float* n = &out.normals[3 * v];
        const float z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        if (z2 > 0.0f) { const float len = std::sqrt(z2); n[0] /= len; n[1] /= len; n[2] /= len; }
    }
}


// density lattice and vertex colours through two C-ABI calls (stubbed with analytic fields by mesh_check.cpp)
inline bool extract(mon_object* obj, const float bmin[3], const float bmax[3], uint32_t res, float thresh, Extracted& out, std::string& err) {
    const uint32_t r3[3] = {res, res, res};
    std::vector<float> sigma((size_t)res * res * res);
    if (mon_object_density_grid(obj, r3, sigma.data()) != MON_OK) { err = mon_last_error(); return false; }
    marching_cubes(sigma.data(), res, bmin, bmax, thresh, out);
    const size_t nv = out.verts.size() / 3;
    // compute_mesh_vertex_colors: network (EMA weights) at WarpPoint(vertex), logistic on the rgb logits, *255 truncated to u8
    out.colors.assign(nv * 3, 0);
    if (nv) {
        std::vector<float> unit(nv * 3), out4(nv * 4);
        for (size_t v = 0; v < nv; ++v)
            for (int k = 0; k < 3; ++k) unit[3 * v + k] = (out.verts[3 * v + k] - bmin[k]) / (bmax[k] - bmin[k]);
        if (mon_object_query_points(obj, unit.data(), (uint32_t)nv, 1, out4.data()) != MON_OK) { err = mon_last_error(); return false; }
        for (size_t v = 0; v < nv; ++v)
            for (int k = 0; k < 3; ++k) {
                const float c = 1.0f / (1.0f + std::exp(-out4[4 * v + k]));
                out.colors[3 * v + k] = (uint8_t)std::fmin(std::fmax(c * 255.0f, 0.0f), 255.0f);
            }
    }
    return true;
}
}  // namespace mesh_cpu
