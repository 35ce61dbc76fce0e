// Stand-in for the handful of OpenCV 3 names RO-MAP's Core headers and nerf_model.cu mention (cv::Mat as a typed 2-D
// buffer, imwrite/cvtColor/normalize/convertTo as no-ops).  TEST INFRASTRUCTURE ONLY: it exists so that the reference's
// nerf_model.cu compiles unmodified; none of these host paths is executed by the golden-vector harness.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_32FC3 21
#define CV_MINMAX 32
namespace cv {
enum { COLOR_RGB2BGR = 4, COLOR_BGR2RGB = 4, NORM_MINMAX = 32 };
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(std::make_shared<std::vector<unsigned char>>((size_t)r * c * ((type >> 3) + 1) * ((type & 7) == 0 ? 1 : (type & 7) == 2 ? 2 : 4))) { data = buf_->data(); }
    int channels() const { return (type_ >> 3) + 1; }
    bool empty() const { return data == nullptr; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == 0 ? 1 : (type_ & 7) == 2 ? 2 : 4); }
    template <typename T> T* ptr(int r = 0, int c = 0) { return reinterpret_cast<T*>(data + ((size_t)r * cols + c) * elemSize()); }
    void convertTo(Mat& dst, int, double = 1.0, double = 0.0) const { dst = *this; }
private:
    int type_ = 0;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src; }
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void normalize(const Mat& src, Mat& dst, double = 1.0, double = 0.0, int = 0) { dst = src; }
}  // namespace cv
