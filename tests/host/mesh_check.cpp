// mesh_check.cpp — CPU unit check of the mesh extraction's CPU statement (tests/host/mesh_cpu.h) and of the PLY writer of
// ro_map_b200/host/mesh.h, with the two C-ABI calls the CPU statement makes replaced by an
// analytic field (sigma = 2 + 40 * (0.3 - |p - c|), i.e. the threshold-2.0 surface is a sphere of radius 0.3 in
// unit-cube coordinates; colour logits encode the position).  Prints facts the pytest asserts on.
#include <cstdio>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "mesh_cpu.h"

static const float C0[3] = {0.5f, 0.45f, 0.55f};
static bool g_random = false;   // "random": white noise with an empty border — every one of the 256 cell configurations occurs
static float field(float x, float y, float z) {
    if (g_random) {
        if (x <= 0.0f || y <= 0.0f || z <= 0.0f || x >= 1.0f || y >= 1.0f || z >= 1.0f) return 0.0f;
        uint32_t h = (uint32_t)(x * 4096.0f) * 73856093u ^ (uint32_t)(y * 4096.0f) * 19349663u ^ (uint32_t)(z * 4096.0f) * 83492791u;
        h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
        return 4.0f * (float)(h >> 8) / 16777216.0f;   // [0,4): half of the lattice points are above the threshold 2.0
    }
    const float dx = x - C0[0], dy = y - C0[1], dz = z - C0[2];
    return 2.0f + 40.0f * (0.3f - std::sqrt(dx * dx + dy * dy + dz * dz));
}
extern "C" {
const char* mon_last_error(void) { return "stub"; }
int mon_object_extract_mesh(mon_object*, uint32_t, float, mon_mesh**) { return 1; }
int mon_mesh_counts(const mon_mesh*, uint32_t*, uint32_t*, uint32_t*) { return 1; }
int mon_mesh_read(const mon_mesh*, float*, float*, uint8_t*, uint32_t*) { return 1; }
int mon_mesh_destroy(mon_mesh*) { return 1; }
int mon_object_density_grid(mon_object*, const uint32_t res[3], float* out) {
    for (uint32_t z = 0; z < res[2]; ++z)
        for (uint32_t y = 0; y < res[1]; ++y)
            for (uint32_t x = 0; x < res[0]; ++x)
                out[((size_t)z * res[1] + y) * res[0] + x] = field((float)x / (res[0] - 1), (float)y / (res[1] - 1), (float)z / (res[2] - 1));
    return 0;
}
int mon_object_query_points(mon_object*, const float* p, uint32_t n, int, float* out4) {
    for (uint32_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) out4[4 * i + k] = 8.0f * (p[3 * i + k] - 0.5f);
        out4[4 * i + 3] = field(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    }
    return 0;
}
}

int main(int argc, char** argv) {
    const uint32_t res = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    g_random = argc > 3 && std::string(argv[3]) == "random";
    const float bmin[3] = {-1.0f, -2.0f, -0.5f}, bmax[3] = {1.0f, 2.0f, 0.5f};
    mesh::Extracted m;
    std::string err;
    if (!mesh_cpu::extract(nullptr, bmin, bmax, res, 2.0f, m, err)) { printf("error %s\n", err.c_str()); return 1; }
    const size_t nv = m.n_surface_verts, nf = m.indices.size() / 3, n_padded = m.verts.size() / 3;
    size_t bad_padding = (n_padded % 128 != 0) + (n_padded < nv) + (n_padded >= nv + 128);   // padded to the next multiple of 128 ...
    for (size_t v = nv; v < n_padded; ++v)                                                       // ... with zero vertices and zero normals
        for (int k = 0; k < 3; ++k) bad_padding += (m.verts[3 * v + k] != 0.0f) + (m.normals[3 * v + k] != 0.0f);
    size_t cases_seen = 0, unreferenced = 0;
    {
        std::vector<char> ref(nv, 0);
        for (uint32_t i : m.indices) { if (i >= nv) ++bad_padding; else ref[i] = 1; }
        for (char r : ref) unreferenced += !r;
        for (int mask = 0; mask < 256; ++mask) cases_seen += mesh::mc::TRIANGLES[mask][0] >= 0;
    }
    // every undirected edge shared by exactly two triangles, and traversed once in each direction (consistent winding)
    std::map<std::pair<uint32_t, uint32_t>, int> directed;
    std::set<std::pair<uint32_t, uint32_t>> undirected;
    for (size_t f = 0; f < nf; ++f)
        for (int k = 0; k < 3; ++k) {
            const uint32_t a = m.indices[3 * f + k], b = m.indices[3 * f + (k + 1) % 3];
            directed[{a, b}]++;
            undirected.insert({a < b ? a : b, a < b ? b : a});
        }
    size_t bad_edges = 0;
    for (auto& e : undirected)
        if (directed[{e.first, e.second}] != 1 || directed[{e.second, e.first}] != 1) ++bad_edges;
    // vertices on the sphere (unit-cube metric), normals outward, colours = logistic of the position code
    double max_r_err = 0.0, min_dot = 1.0;
    size_t bad_col = 0;
    for (size_t v = 0; v < nv; ++v) {
        float u[3], n_unit[3];
        for (int k = 0; k < 3; ++k) u[k] = (m.verts[3 * v + k] - bmin[k]) / (bmax[k] - bmin[k]);
        const float d[3] = {u[0] - C0[0], u[1] - C0[1], u[2] - C0[2]};
        const double r = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        max_r_err = std::fmax(max_r_err, std::fabs(r - 0.3));
        // object-space normal -> unit-cube normal direction (covariant: multiply by the box extent)
        for (int k = 0; k < 3; ++k) n_unit[k] = m.normals[3 * v + k] * (bmax[k] - bmin[k]);
        const double nl = std::sqrt(n_unit[0] * n_unit[0] + n_unit[1] * n_unit[1] + n_unit[2] * n_unit[2]);
        min_dot = std::fmin(min_dot, (n_unit[0] * d[0] + n_unit[1] * d[1] + n_unit[2] * d[2]) / (nl * r));
        for (int k = 0; k < 3; ++k) {
            const int want = (int)(255.0f / (1.0f + std::exp(-8.0f * (u[k] - 0.5f))));
            if (std::abs((int)m.colors[3 * v + k] - want) > 1) ++bad_col;
        }
    }
    // signed volume enclosed by the surface in unit-cube coordinates (divergence theorem); the reference's winding makes the
    // geometric normals point INTO the dense side, so the enclosed volume comes out negative
    double vol6 = 0.0;
    for (size_t f = 0; f < nf; ++f) {
        double p[3][3];
        for (int k = 0; k < 3; ++k)
            for (int d = 0; d < 3; ++d) p[k][d] = (m.verts[3 * m.indices[3 * f + k] + d] - bmin[d]) / (bmax[d] - bmin[d]);
        vol6 += p[0][0] * (p[1][1] * p[2][2] - p[1][2] * p[2][1]) - p[0][1] * (p[1][0] * p[2][2] - p[1][2] * p[2][0]) + p[0][2] * (p[1][0] * p[2][1] - p[1][1] * p[2][0]);
    }
    // the table respects the cube's symmetries: a configuration and its image under a rotation about the z axis or the x axis
    // (together they generate all 24 rotations) have the same number of triangles
    size_t asym = 0;
    {
        const int rz[8] = {1, 2, 3, 0, 5, 6, 7, 4};   // corner c -> its image under a quarter turn about z
        const int rx[8] = {3, 2, 6, 7, 0, 1, 5, 4};   // ... about x: (x,y,z) -> (x, 1-z, y)
        for (int mask = 0; mask < 256; ++mask)
            for (const int* rot : {rz, rx}) {
                int img = 0;
                for (int c = 0; c < 8; ++c) if (mask >> c & 1) img |= 1 << rot[c];
                auto count = [](int m) { int n = 0; while (n < 15 && mesh::mc::TRIANGLES[m][n] >= 0) ++n; return n; };
                asym += count(mask) != count(img);
            }
    }
    printf("volume %.6f asymmetric_cases %zu ", vol6 / 6.0, asym);
    printf("verts %zu faces %zu edges %zu bad_edges %zu euler %ld max_r_err %.6f min_normal_dot %.4f bad_colors %zu padded_verts %zu bad_padding %zu "
           "unreferenced %zu cases %zu\n", nv, nf, undirected.size(), bad_edges, (long)nv - (long)undirected.size() + (long)nf, max_r_err, min_dot, bad_col,
           n_padded, bad_padding, unreferenced, cases_seen);
    if (argc > 2) return mesh::save_ply(argv[2], m.verts, m.normals, m.colors, m.indices) ? 0 : 2;
    return 0;
}
