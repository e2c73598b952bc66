// oracle/ref_shim/cvshim.h -- TEST INFRASTRUCTURE ONLY.
//
// The few cv:: names the reference's hand-written pair blend ([BLEND]:141-717) touches, so that THAT block -- taken
// from /root/reference at build time by line range, never stored in this repository -- compiles and runs here without
// OpenCV.  Everything arithmetic in the block is the reference's own code; what this header supplies is storage
// (a continuous float matrix), cvtColor(CV_RGB2GRAY) on CV_32FC3 (OpenCV's scalar formula; the reference's
// OpenCV 3.4.2 may associate it differently in its SIMD body), and no-ops for imwrite / cout / tick counters.
//
// Buffers are continuous like cv::Mat's and carry a zeroed guard band on both ends: the block reads a few floats
// past row ends ([BLEND]:276-278, :501), which in OpenCV lands in the neighbouring row or just outside the buffer.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <vector>

namespace cv {

enum { CV_32FC1 = 5, CV_32FC3 = 21, CV_RGB2GRAY = 7 };

struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    void create(int r, int c, int type) {
        rows = r; cols = c; ch_ = (type == CV_32FC3) ? 3 : 1;
        store_.assign((size_t)r * c * ch_ + 2 * GUARD, 0.f);
    }
    void setTo(float v) { std::fill(store_.begin() + GUARD, store_.end() - GUARD, v); }
    void copyTo(Mat& dst) const { dst.rows = rows; dst.cols = cols; dst.ch_ = ch_; dst.store_ = store_; }
    template <typename T> T* ptr(int y) { return reinterpret_cast<T*>(store_.data() + GUARD + (size_t)y * cols * ch_); }
    template <typename T> const T* ptr(int y) const { return reinterpret_cast<const T*>(store_.data() + GUARD + (size_t)y * cols * ch_); }
    int channels() const { return ch_; }
    float* data() { return store_.data() + GUARD; }
    const float* data() const { return store_.data() + GUARD; }

protected:
    static const size_t GUARD = 64;
    int ch_ = 1;
    std::vector<float> store_;
};

typedef Mat UMat;

template <typename T> class Mat_ : public Mat {
public:
    void create(int r, int c) { Mat::create(r, c, CV_32FC1); }
};

inline void cvtColor(const Mat& src, Mat& dst, int /*code: CV_RGB2GRAY*/) {
    dst.create(src.rows, src.cols, CV_32FC1);
    for (int y = 0; y < src.rows; ++y) {
        const float* s = src.ptr<float>(y);
        float* d = dst.ptr<float>(y);
        for (int x = 0; x < src.cols; ++x) d[x] = s[3 * x] * 0.299f + s[3 * x + 1] * 0.587f + s[3 * x + 2] * 0.114f;
    }
}

inline bool imwrite(const char*, const Mat&) { return true; }
inline double getTickCount() { return 0.0; }
inline double getTickFrequency() { return 1.0; }

namespace detail {
template <typename T> static inline T sqr(T x) { return x * x; }   // opencv2/stitching/detail/util_inl.hpp
}

}  // namespace cv

// `cout << anything << endl` in the block goes nowhere
struct RefNullStream {};
template <typename T> inline RefNullStream& operator<<(RefNullStream& s, const T&) { return s; }
static RefNullStream ref_null_stream;
#define cout ref_null_stream
#define endl 0
