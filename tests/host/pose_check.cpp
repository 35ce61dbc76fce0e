// pose_check.cpp — prints facts about ro_map_b200/host/pose_math.h for tests/test_host_facade.py:
//   pose_check toc <theta> <phi> <r>        -> 16 floats (column-major Toc)
//   pose_check quat <16 floats col-major>   -> qx qy qz qw
//   pose_check mul <16> <16>                -> 16 floats
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pose_math.h"

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    float a[16], b[16], c[16], q[4];
    if (!strcmp(argv[1], "toc") && argc == 5) {
        pose_math::turntable_toc((float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4]), c);
        for (float v : c) printf("%.9g ", v);
    } else if (!strcmp(argv[1], "quat") && argc == 18) {
        for (int i = 0; i < 16; ++i) a[i] = (float)atof(argv[2 + i]);
        pose_math::rot_to_quat(a, q);
        for (float v : q) printf("%.9g ", v);
    } else if (!strcmp(argv[1], "mul") && argc == 34) {
        for (int i = 0; i < 16; ++i) { a[i] = (float)atof(argv[2 + i]); b[i] = (float)atof(argv[18 + i]); }
        pose_math::mul44(a, b, c);
        for (float v : c) printf("%.9g ", v);
    } else {
        return 1;
    }
    printf("\n");
    return 0;
}
