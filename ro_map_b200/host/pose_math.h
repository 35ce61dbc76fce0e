// pose_math.h — the little rigid-transform arithmetic the on-disk outputs need, on plain column-major float[16]
// (the memory layout of Eigen::Matrix4f), so it compiles with Eigen present or with the mon_compat.h stand-ins.
//   * object-centric camera pose Toc = ObjTow * Twc and its (t, q) line in test.txt / train.txt (nerf.cu:331-336,388-393)
//   * the rotation -> quaternion conversion Eigen performs in `Eigen::Quaternionf q(R)` (trace / largest-diagonal branches)
//   * the turn-table poses of the 360-degree video (NeRF_Model::GenerateToc, nerf_model.cu:2186-2205; 60 views, phi = 30)
#pragma once
#include <cmath>

namespace pose_math {

inline float at(const float* m, int r, int c) { return m[c * 4 + r]; }

// C = A * B, all column-major 4x4
inline void mul44(const float* A, const float* B, float* C) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += at(A, r, k) * at(B, k, c);
            C[c * 4 + r] = s;
        }
}

// q = (x, y, z, w) of the rotation block of T; same branch structure as Eigen's matrix -> quaternion assignment
inline void rot_to_quat(const float* T, float q[4]) {
    const float m00 = at(T, 0, 0), m11 = at(T, 1, 1), m22 = at(T, 2, 2);
    float t = m00 + m11 + m22;
    if (t > 0.0f) {
        t = std::sqrt(t + 1.0f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (at(T, 2, 1) - at(T, 1, 2)) * t;
        q[1] = (at(T, 0, 2) - at(T, 2, 0)) * t;
        q[2] = (at(T, 1, 0) - at(T, 0, 1)) * t;
    } else {
        int i = 0;
        if (m11 > m00) i = 1;
        if (m22 > at(T, i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(at(T, i, i) - at(T, j, j) - at(T, k, k) + 1.0f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (at(T, k, j) - at(T, j, k)) * t;
        q[j] = (at(T, j, i) + at(T, i, j)) * t;
        q[k] = (at(T, k, i) + at(T, i, k)) * t;
    }
}

// camera -> object pose on the turn-table: camera at spherical (r, theta, phi) looking at the object origin,
// x axis tangent to the circle of latitude (GenerateToc).  Angles in degrees.
inline void turntable_toc(float theta_deg, float phi_deg, float r, float* Toc) {
    const float d2r = float(M_PI) / 180.0f;
    const float z = r * std::sin(phi_deg * d2r);
    const float x = r * std::cos(phi_deg * d2r) * std::cos(theta_deg * d2r);
    const float y = r * std::cos(phi_deg * d2r) * std::sin(theta_deg * d2r);
    const float tn = std::sqrt(x * x + y * y + z * z);
    const float zax[3] = {-x / tn, -y / tn, -z / tn};
    const float rv = (theta_deg + 90.0f) * d2r;
    float xax[3] = {std::cos(rv), std::sin(rv), 0.0f};
    const float xn = std::sqrt(xax[0] * xax[0] + xax[1] * xax[1]);
    xax[0] /= xn; xax[1] /= xn;
    float yax[3] = {zax[1] * xax[2] - zax[2] * xax[1], zax[2] * xax[0] - zax[0] * xax[2], zax[0] * xax[1] - zax[1] * xax[0]};
    const float yn = std::sqrt(yax[0] * yax[0] + yax[1] * yax[1] + yax[2] * yax[2]);
    for (float& v : yax) v /= yn;
    for (int i = 0; i < 16; ++i) Toc[i] = 0.0f;
    for (int i = 0; i < 3; ++i) { Toc[0 + i] = xax[i]; Toc[4 + i] = yax[i]; Toc[8 + i] = zax[i]; }
    Toc[12] = x; Toc[13] = y; Toc[14] = z; Toc[15] = 1.0f;
}

}  // namespace pose_math
