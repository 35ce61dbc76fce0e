// mesh.h — surface extraction for the viewer / PLY output: GenerateMesh + TransCPUMesh + SaveMesh of the reference
// (MON/Core/src/nerf_model.cu:1993-2105, marching_cubes.cu:567-605), host side.
//
// The density lattice (raw sigma logit on res^3 points of the object box, EMA weights) and the vertex colours come
// from the B200 core through the C ABI (mon_object_density_grid / mon_object_query_points).  The iso-surface at
// `thresh` (reference: 2.0 on the raw logit, marching_cubes.h:30-31) is extracted on the CPU by marching TETRAHEDRA
// (each lattice cell split into six tetrahedra around its main diagonal): the same surface as the reference's marching
// cubes up to the triangulation inside a cell, with shared vertices on lattice edges, area-weighted vertex normals
// (the reference's 1-ring normals) and the reference's ASCII PLY layout (reversed winding, u8 colours).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "mon_c.h"

namespace mesh {

struct Extracted {
    std::vector<float> verts, normals;   // object space, xyz per vertex
    std::vector<uint8_t> colors;         // rgb per vertex
    std::vector<uint32_t> indices;       // 3 per triangle
};

inline bool extract(mon_object* obj, const float bmin[3], const float bmax[3], uint32_t res, float thresh, Extracted& out, std::string& err) {
    const uint32_t r3[3] = {res, res, res};
    std::vector<float> sigma((size_t)res * res * res);
    if (mon_object_density_grid(obj, r3, sigma.data()) != MON_OK) { err = mon_last_error(); return false; }
    out = Extracted();
    std::vector<float> unit;   // unit-cube coordinates of the vertices, for the colour query
    std::unordered_map<uint64_t, uint32_t> edge_vertex;
    auto lattice = [&](uint32_t x, uint32_t y, uint32_t z) { return ((size_t)z * res + y) * res + x; };
    auto vertex_on_edge = [&](size_t a, size_t b) -> uint32_t {
        const uint64_t key = a < b ? ((uint64_t)a << 32 | b) : ((uint64_t)b << 32 | a);
        auto it = edge_vertex.find(key);
        if (it != edge_vertex.end()) return it->second;
        const float va = sigma[a], vb = sigma[b];
        float t = (thresh - va) / (vb - va);
        t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
        const uint32_t ca[3] = {(uint32_t)(a % res), (uint32_t)((a / res) % res), (uint32_t)(a / ((size_t)res * res))};
        const uint32_t cb[3] = {(uint32_t)(b % res), (uint32_t)((b / res) % res), (uint32_t)(b / ((size_t)res * res))};
        const uint32_t id = (uint32_t)(out.verts.size() / 3);
        for (int k = 0; k < 3; ++k) {
            const float u = ((float)ca[k] + t * ((float)cb[k] - (float)ca[k])) / (float)(res - 1);
            unit.push_back(u);
            out.verts.push_back(bmin[k] + u * (bmax[k] - bmin[k]));   // UnWarpPoint (nerf_model.cu:146-150)
        }
        edge_vertex.emplace(key, id);
        return id;
    };
    auto emit = [&](uint32_t i0, uint32_t i1, uint32_t i2, const float inside[3]) {
        if (i0 == i1 || i1 == i2 || i0 == i2) return;
        const float* p0 = &out.verts[3 * i0]; const float* p1 = &out.verts[3 * i1]; const float* p2 = &out.verts[3 * i2];
        const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const float d = n[0] * (p0[0] - inside[0]) + n[1] * (p0[1] - inside[1]) + n[2] * (p0[2] - inside[2]);
        if (d >= 0.0f) { out.indices.push_back(i0); out.indices.push_back(i1); out.indices.push_back(i2); }   // normal points away from the dense side
        else { out.indices.push_back(i0); out.indices.push_back(i2); out.indices.push_back(i1); }
    };
    static const int tets[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};
    for (uint32_t z = 0; z + 1 < res; ++z)
        for (uint32_t y = 0; y + 1 < res; ++y)
            for (uint32_t x = 0; x + 1 < res; ++x) {
                size_t corner[8];
                bool any_in = false, any_out = false;
                for (int c = 0; c < 8; ++c) {
                    corner[c] = lattice(x + (c & 1), y + ((c >> 1) & 1), z + ((c >> 2) & 1));
                    const bool in = sigma[corner[c]] > thresh;
                    any_in |= in; any_out |= !in;
                }
                if (!any_in || !any_out) continue;
                for (const auto& t : tets) {
                    int in_idx[4], out_idx[4], n_in = 0, n_out = 0;
                    for (int k = 0; k < 4; ++k) {
                        if (sigma[corner[t[k]]] > thresh) in_idx[n_in++] = t[k]; else out_idx[n_out++] = t[k];
                    }
                    if (n_in == 0 || n_in == 4) continue;
                    // a point on the dense side of the surface, to orient the triangles
                    float inside[3] = {0, 0, 0};
                    for (int k = 0; k < n_in; ++k) {
                        const int c = in_idx[k];
                        const float u[3] = {(float)(x + (c & 1)) / (res - 1), (float)(y + ((c >> 1) & 1)) / (res - 1), (float)(z + ((c >> 2) & 1)) / (res - 1)};
                        for (int d = 0; d < 3; ++d) inside[d] += (bmin[d] + u[d] * (bmax[d] - bmin[d])) / (float)n_in;
                    }
                    if (n_in == 1 || n_in == 3) {
                        const int apex = n_in == 1 ? in_idx[0] : out_idx[0];
                        const int* others = n_in == 1 ? out_idx : in_idx;
                        const uint32_t a = vertex_on_edge(corner[apex], corner[others[0]]);
                        const uint32_t b = vertex_on_edge(corner[apex], corner[others[1]]);
                        const uint32_t c2 = vertex_on_edge(corner[apex], corner[others[2]]);
                        emit(a, b, c2, inside);
                    } else {   // two inside, two outside: a quad
                        const uint32_t a = vertex_on_edge(corner[in_idx[0]], corner[out_idx[0]]);
                        const uint32_t b = vertex_on_edge(corner[in_idx[0]], corner[out_idx[1]]);
                        const uint32_t c2 = vertex_on_edge(corner[in_idx[1]], corner[out_idx[1]]);
                        const uint32_t d2 = vertex_on_edge(corner[in_idx[1]], corner[out_idx[0]]);
                        emit(a, b, c2, inside);
                        emit(a, c2, d2, inside);
                    }
                }
            }
    const size_t nv = out.verts.size() / 3;
    // area-weighted vertex normals over the 1-ring
    out.normals.assign(nv * 3, 0.0f);
    for (size_t i = 0; i + 2 < out.indices.size(); i += 3) {
        const uint32_t a = out.indices[i], b = out.indices[i + 1], c = out.indices[i + 2];
        const float* p0 = &out.verts[3 * a]; const float* p1 = &out.verts[3 * b]; const float* p2 = &out.verts[3 * c];
        const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        for (uint32_t v : {a, b, c}) for (int k = 0; k < 3; ++k) out.normals[3 * v + k] += n[k];
    }
    for (size_t v = 0; v < nv; ++v) {
        float* n = &out.normals[3 * v];
        const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (len > 0.0f) { n[0] /= len; n[1] /= len; n[2] /= len; }
    }
    // vertex colours: network at the vertex positions with the EMA weights, logistic on the rgb logits
    out.colors.assign(nv * 3, 0);
    if (nv) {
        std::vector<float> out4(nv * 4);
        if (mon_object_query_points(obj, unit.data(), (uint32_t)nv, 1, out4.data()) != MON_OK) { err = mon_last_error(); return false; }
        for (size_t v = 0; v < nv; ++v)
            for (int k = 0; k < 3; ++k) {
                const float c = 1.0f / (1.0f + std::exp(-out4[4 * v + k]));
                out.colors[3 * v + k] = (uint8_t)std::fmin(std::fmax(c * 255.0f, 0.0f), 255.0f);
            }
    }
    return true;
}

// the reference's ASCII PLY (marching_cubes.cu:567-605): positions (v - offset) / scale with offset 0, scale 1
inline bool save_ply(const std::string& path, const std::vector<float>& verts, const std::vector<float>& normals,
                     const std::vector<uint8_t>& colors, const std::vector<uint32_t>& indices) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const unsigned nv = (unsigned)(verts.size() / 3), nf = (unsigned)(indices.size() / 3);
    fprintf(f,
            "ply\nformat ascii 1.0\ncomment multi-object NeRF mesh (reference layout: instant-ngp PLY)\nelement vertex %u\n"
            "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
            "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %u\nproperty list uchar int vertex_index\nend_header\n",
            nv, nf);
    for (unsigned i = 0; i < nv; ++i)
        fprintf(f, "%0.5f %0.5f %0.5f %0.3f %0.3f %0.3f %d %d %d\n", verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], normals[3 * i], normals[3 * i + 1],
                normals[3 * i + 2], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
    for (unsigned i = 0; i < nf; ++i) fprintf(f, "3 %d %d %d\n", indices[3 * i + 2], indices[3 * i + 1], indices[3 * i]);
    fclose(f);
    return true;
}

}  // namespace mesh
