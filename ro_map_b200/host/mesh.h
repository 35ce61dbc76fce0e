// mesh.h — surface extraction for the viewer / PLY output: GenerateMesh + TransCPUMesh + SaveMesh of the reference
// (MON/Core/src/nerf_model.cu:1993-2105, marching_cubes.cu:41-91,93-435,437-472,474-510,567-605), host side.
//
// The density lattice (raw sigma logit on res^3 points of the object box, EMA weights) and the vertex colours come
// from the B200 core through the C ABI (mon_object_density_grid / mon_object_query_points).  The iso-surface at
// `thresh` (reference: 2.0 on the raw logit, marching_cubes.h:30-31) is extracted on the CPU by MARCHING CUBES with
// the reference's conventions: one vertex per sign-changing lattice edge at `(x + (thresh-f0)/(f1-f0)) * scale + min`
// (gen_vertices), the reference's corner / edge numbering of a cell (gen_faces :393-421), triangles wound so that their
// geometric normal points into the dense side, un-normalised 1-ring normals `(pb-pa) x (pa-pc)` summed per vertex
// (accumulate_1ring), the vertex count padded to a multiple of 128 with zero vertices (MarchingCubes :499), colours =
// logistic(rgb logits) at the warped vertex positions, and the reference's ASCII PLY layout (reversed winding, u8 colours).
//
// The per-configuration triangle lists are not a copied table: they are DERIVED at first use (mc::table) — on every
// cell face the sign-changing edges are joined (a face whose corners alternate joins the two edges around each dense
// corner, a rule that depends on the face alone, so neighbouring cells always agree and the surface is closed), the
// resulting closed loops are oriented and triangulated (avoiding diagonals that would lie in a cell face).  Against the
// reference run on a B200 (tests/golden/romap_mesh_golden.npz) this gives the bit-identical vertex set and the same triangle
// count — on a sphere and on white noise that contains all 256 configurations — with about half of the triangles identical;
// the others differ by the diagonal chosen inside a polygon.  No cracks by construction.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "common.h"
#include "mon_c.h"

namespace mesh {

namespace mc {
// cell corners and edges in the reference's numbering (marching_cubes.cu:393-421)
static const int CORNER[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int EDGE[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int FACE[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};   // corners in cyclic order

struct Case { int n = 0; int8_t tri[24]; };   // n triangle corners (3 per triangle), each an edge number

inline int edge_between(int a, int b) {
    for (int e = 0; e < 12; ++e)
        if ((EDGE[e][0] == a && EDGE[e][1] == b) || (EDGE[e][0] == b && EDGE[e][1] == a)) return e;
    return -1;
}

inline std::array<Case, 256> derive() {
    std::array<Case, 256> T{};
    for (int mask = 1; mask < 255; ++mask) {
        auto in = [&](int c) { return (mask >> c) & 1; };
        int link[12][2], deg[12];
        for (int e = 0; e < 12; ++e) { deg[e] = 0; link[e][0] = link[e][1] = -1; }
        auto join = [&](int a, int b) { link[a][deg[a]++] = b; link[b][deg[b]++] = a; };
        for (const auto& f : FACE) {
            int cross[4], nc = 0, fe[4];
            for (int k = 0; k < 4; ++k) {
                fe[k] = edge_between(f[k], f[(k + 1) & 3]);
                if (in(f[k]) != in(f[(k + 1) & 3])) cross[nc++] = k;
            }
            if (nc == 2) join(fe[cross[0]], fe[cross[1]]);
            else if (nc == 4)   // corners alternate: cut off each dense corner k (edges k-1 and k of the cycle)
                for (int k = 0; k < 4; ++k) if (in(f[k])) join(fe[(k + 3) & 3], fe[k]);
        }
        bool used[12] = {};
        Case& cs = T[mask];
        for (int e0 = 0; e0 < 12; ++e0) {
            if (deg[e0] != 2 || used[e0]) continue;
            int loop[12], n = 0, prev = -1, cur = e0;
            do {
                loop[n++] = cur; used[cur] = true;
                const int nxt = link[cur][0] != prev ? link[cur][0] : link[cur][1];
                prev = cur; cur = nxt;
            } while (cur != e0 && n < 12);
            // orientation: Newell normal of the loop (edge midpoints) against the dense -> empty direction of its edges
            float P[12][3], nrm[3] = {0, 0, 0}, score = 0;
            for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) P[i][k] = 0.5f * (CORNER[EDGE[loop[i]][0]][k] + CORNER[EDGE[loop[i]][1]][k]);
            for (int i = 0; i < n; ++i) {
                const float* a = P[i]; const float* b = P[(i + 1) % n];
                nrm[0] += (a[1] - b[1]) * (a[2] + b[2]); nrm[1] += (a[2] - b[2]) * (a[0] + b[0]); nrm[2] += (a[0] - b[0]) * (a[1] + b[1]);
            }
            for (int i = 0; i < n; ++i) {
                const int a = EDGE[loop[i]][0], b = EDGE[loop[i]][1], dense = in(a) ? a : b, empty = in(a) ? b : a;
                for (int k = 0; k < 3; ++k) score += nrm[k] * (float)(CORNER[empty][k] - CORNER[dense][k]);
            }
            if (score > 0) for (int i = 0; i < n / 2; ++i) { const int t = loop[i]; loop[i] = loop[n - 1 - i]; loop[n - 1 - i] = t; }   // normal must point INTO the dense side
            // triangulate the loop; a diagonal that joins two edges of one cell face would lie IN that face (and could coincide with
            // a diagonal of the neighbouring cell): minimise their number over all triangulations (interval DP, n <= 12)
            auto in_face = [&](int ea, int eb) {
                for (const auto& f : FACE) {
                    bool ha = false, hb = false;
                    for (int k = 0; k < 4; ++k) { const int fe = edge_between(f[k], f[(k + 1) & 3]); ha |= fe == ea; hb |= fe == eb; }
                    if (ha && hb) return 1;
                }
                return 0;
            };
            int cost[12][12] = {}, split[12][12] = {};
            auto diag = [&](int i, int j) { return (j - i == 1 || (i == 0 && j == n - 1)) ? 0 : in_face(loop[i], loop[j]); };
            for (int len = 2; len < n; ++len)
                for (int i = 0; i + len < n; ++i) {
                    const int j = i + len;
                    cost[i][j] = 1 << 20;
                    for (int k = i + 1; k < j; ++k) {
                        const int c = cost[i][k] + cost[k][j] + diag(i, k) + diag(k, j);
                        if (c < cost[i][j]) { cost[i][j] = c; split[i][j] = k; }
                    }
                }
            int stack[24][2], sp = 0;
            stack[sp][0] = 0; stack[sp][1] = n - 1; ++sp;
            while (sp) {
                --sp;
                const int i = stack[sp][0], j = stack[sp][1];
                if (j - i < 2) continue;
                const int k = split[i][j];
                cs.tri[cs.n++] = (int8_t)loop[i]; cs.tri[cs.n++] = (int8_t)loop[k]; cs.tri[cs.n++] = (int8_t)loop[j];
                stack[sp][0] = k; stack[sp][1] = j; ++sp;
                stack[sp][0] = i; stack[sp][1] = k; ++sp;
            }
        }
    }
    return T;
}

inline const std::array<Case, 256>& table() {
    static const std::array<Case, 256> T = derive();
    return T;
}
}  // namespace mc

struct Extracted {
    std::vector<float> verts, normals;   // object space, xyz per vertex; count padded to a multiple of 128 like the reference's
    std::vector<uint8_t> colors;         // rgb per vertex
    std::vector<uint32_t> indices;       // 3 per triangle, the reference's internal winding (save_ply reverses it)
    uint32_t n_surface_verts = 0;        // vertices that lie on the surface (the rest is padding at the origin)
};

// the iso-surface of a lattice alone (no colours): used by extract() and by the parity tests
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
    const auto& T = mc::table();
    for (uint32_t z = 0; z + 1 < res; ++z)
        for (uint32_t y = 0; y + 1 < res; ++y)
            for (uint32_t x = 0; x + 1 < res; ++x) {
                const size_t idx = x + y * res1 + z * res2;
                int mask = 0;
                for (int c = 0; c < 8; ++c)
                    if (sigma[idx + mc::CORNER[c][0] + mc::CORNER[c][1] * res1 + mc::CORNER[c][2] * res2] > thresh) mask |= 1 << c;
                if (mask == 0 || mask == 255) continue;
                const mc::Case& cs = T[mask];
                for (int i = 0; i < cs.n; ++i) {
                    const int e = cs.tri[i], a = mc::EDGE[e][0], b = mc::EDGE[e][1];
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
