// mesh_golden.cpp — runs the CPU statement of the core's marching cubes (tests/host/mesh_cpu.h) on a density lattice read from a raw float32 file
// ([z][y][x], res^3 values) and writes vertices / normals / indices as raw arrays for tests/test_golden_romap.py to hold
// against the reference's own MarchingCubes output (tests/golden/romap_mesh_golden.npz).  CPU only, no C ABI involved.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mesh_cpu.h"

extern "C" {   // the headers name them in their extract() functions; not used here
const char* mon_last_error(void) { return "unused"; }
int mon_object_density_grid(mon_object*, const uint32_t*, float*) { return 1; }
int mon_object_query_points(mon_object*, const float*, uint32_t, int, float*) { return 1; }
int mon_object_extract_mesh(mon_object*, uint32_t, float, mon_mesh**) { return 1; }
int mon_mesh_counts(const mon_mesh*, uint32_t*, uint32_t*, uint32_t*) { return 1; }
int mon_mesh_read(const mon_mesh*, float*, float*, uint8_t*, uint32_t*) { return 1; }
int mon_mesh_destroy(mon_mesh*) { return 1; }
}

static bool dump(const std::string& path, const void* p, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = bytes == 0 || fwrite(p, 1, bytes, f) == bytes;
    fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc < 11) { fprintf(stderr, "usage: mesh_golden lattice.f32 res thresh minx miny minz maxx maxy maxz out_prefix\n"); return 2; }
    const uint32_t res = (uint32_t)atoi(argv[2]);
    const float thresh = (float)atof(argv[3]);
    const float bmin[3] = {(float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6])}, bmax[3] = {(float)atof(argv[7]), (float)atof(argv[8]), (float)atof(argv[9])};
    std::vector<float> sigma((size_t)res * res * res);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(sigma.data(), 4, sigma.size(), f) != sigma.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    mesh::Extracted m;
    mesh_cpu::marching_cubes(sigma.data(), res, bmin, bmax, thresh, m);
    const std::string out = argv[10];
    if (!dump(out + ".verts", m.verts.data(), m.verts.size() * 4) || !dump(out + ".normals", m.normals.data(), m.normals.size() * 4) ||
        !dump(out + ".indices", m.indices.data(), m.indices.size() * 4)) return 3;
    printf("%u %zu %zu\n", m.n_surface_verts, m.verts.size() / 3, m.indices.size() / 3);
    return 0;
}
