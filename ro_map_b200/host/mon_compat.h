// mon_compat.h — the value types that cross the libMON.so C++ boundary (MON/Core/include/*.h pull them from
// Eigen and OpenCV).  With the real libraries installed (the RO-MAP build tree) this header just includes them;
// in a tree without them (this repository's CI image) it provides layout-compatible minimal stand-ins so the
// facade, the headless OfflineNeRF driver and the tests still compile.  Only what the boundary touches is covered:
// Eigen::Matrix4f (column-major float[16]), Vector3f, Quaternionf::toRotationMatrix, cv::Mat as a typed 2-D buffer.
#pragma once

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(MON_FORCE_SHIMS)
#define MON_HAVE_EIGEN 1
#endif
#if __has_include(<opencv2/core.hpp>) && !defined(MON_FORCE_SHIMS)
#define MON_HAVE_OPENCV 1
#endif
#endif

#ifdef MON_HAVE_EIGEN
#include <Eigen/Core>
#include <Eigen/Dense>
#else
#include <cmath>
#include <cstring>
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace Eigen {
struct Vector3f {
    float v[3] = {0, 0, 0};
    Vector3f() = default;
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    static Vector3f Zero() { return Vector3f(); }
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
    const float* data() const { return v; }
    float* data() { return v; }
};
inline Vector3f operator*(float s, const Vector3f& a) { return Vector3f(s * a.v[0], s * a.v[1], s * a.v[2]); }
inline Vector3f operator*(const Vector3f& a, float s) { return s * a; }

struct Matrix4f {   // column-major, like Eigen's default
    float m[16] = {0};
    static Matrix4f Zero() { return Matrix4f(); }
    static Matrix4f Identity() { Matrix4f r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
    float& operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    const float* data() const { return m; }
    float* data() { return m; }
    // inverse of a rigid transform [R t; 0 1] (all poses on this boundary are rigid)
    Matrix4f inverse() const {
        Matrix4f r = Identity();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i);
        for (int i = 0; i < 3; ++i) r(i, 3) = -(r(i, 0) * (*this)(0, 3) + r(i, 1) * (*this)(1, 3) + r(i, 2) * (*this)(2, 3));
        return r;
    }
};

struct Quaternionf {
    float w_, x_, y_, z_;
    Quaternionf(float w, float x, float y, float z) : w_(w), x_(x), y_(y), z_(z) {}
    // same formula as Eigen::QuaternionBase::toRotationMatrix (no normalisation)
    void toRotationMatrix(float R[9]) const {   // row-major 3x3
        const float tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
        const float twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
        R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
        R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
    }
};
}  // namespace Eigen
#endif

#ifdef MON_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#include <cstdint>
#include <memory>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_32FC3 21
namespace cv {
// A typed 2-D buffer with shared ownership, the subset of cv::Mat this boundary uses.
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elemSize());
        data = buf_->data();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { const int d = type_ & 7; return d == 0 ? 1 : d == 2 ? 2 : 4; }
    size_t elemSize() const { return elemSize1() * channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T* ptr(int r = 0, int c = 0) { return reinterpret_cast<T*>(data + ((size_t)r * cols + c) * elemSize()); }
    template <typename T> const T* ptr(int r = 0, int c = 0) const { return reinterpret_cast<const T*>(data + ((size_t)r * cols + c) * elemSize()); }
private:
    int type_ = 0;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
}  // namespace cv
#endif

namespace mon_compat {
// column-major float[16] view of a pose, whichever Matrix4f is in use
inline const float* mat16(const Eigen::Matrix4f& m) { return m.data(); }
inline Eigen::Matrix4f pose_from_tq(float tx, float ty, float tz, float qx, float qy, float qz, float qw) {
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
#ifdef MON_HAVE_EIGEN
    Eigen::Quaternionf q(qw, qx, qy, qz);
    T.topLeftCorner(3, 3) = q.toRotationMatrix();
#else
    float R[9];
    Eigen::Quaternionf(qw, qx, qy, qz).toRotationMatrix(R);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T(i, j) = R[i * 3 + j];
#endif
    T(0, 3) = tx; T(1, 3) = ty; T(2, 3) = tz;
    return T;
}
}  // namespace mon_compat
