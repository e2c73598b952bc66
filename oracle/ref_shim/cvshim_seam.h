// oracle/ref_shim/cvshim_seam.h -- TEST INFRASTRUCTURE ONLY.
//
// The cv:: names the reference's refactored DP seam finder ([SEAM]:29-1093, a free-function copy of
// cv::detail::DpSeamFinder) touches, so that THOSE functions -- taken from /root/reference at build time by line range,
// never stored in this repository -- compile and run here without OpenCV.  The seam finder's logic and arithmetic
// (components, contours, edges, seam tips, costs, the DP, the label update) is the reference's own text; this header
// supplies containers (Mat with shared storage and sub-rectangle views, Mat_<T>, Point, Rect, Size) and the four
// OpenCV routines the text calls: floodFill (4-connected, exact match: the defaults), cv::partition (classes numbered
// in order of first appearance), cvRound, normL2 (squared norm, no root -- detail/util_inl.hpp), plus cvtColor + Sobel
// for COLOR_GRAD (float association of OpenCV's vector body, see oracle/seam.cpp).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <set>
#include <stdexcept>
#include <utility>
#include <vector>

namespace cv {

typedef unsigned char uchar;
typedef int64_t int64;

enum { CV_8U = 0, CV_32S = 4, CV_32F = 5 };
#define REF_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
enum {
    CV_8UC1 = REF_MAKETYPE(CV_8U, 1), CV_8UC3 = REF_MAKETYPE(CV_8U, 3), CV_8UC4 = REF_MAKETYPE(CV_8U, 4),
    CV_32SC1 = REF_MAKETYPE(CV_32S, 1),
    CV_32FC1 = REF_MAKETYPE(CV_32F, 1), CV_32FC3 = REF_MAKETYPE(CV_32F, 3), CV_32FC4 = REF_MAKETYPE(CV_32F, 4)
};
enum { ACCESS_READ = 1 << 24, ACCESS_WRITE = 1 << 25, ACCESS_RW = 3 << 24 };
enum { COLOR_BGR2GRAY = 6, COLOR_BGRA2GRAY = 10 };
namespace Error { enum { StsBadArg = -5, StsAssert = -215 }; }

struct RefError : std::runtime_error {
    int code;
    RefError(int c, const char* what) : std::runtime_error(what), code(c) {}
};
#define CV_Assert(expr) do { if (!(expr)) throw cv::RefError(cv::Error::StsAssert, #expr); } while (0)
#define CV_Error(code, msg) throw cv::RefError((code), (msg))

struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
inline Point operator+(const Point& a, const Point& b) { return Point(a.x + b.x, a.y + b.y); }
inline Point operator-(const Point& a, const Point& b) { return Point(a.x - b.x, a.y - b.y); }
inline bool operator==(const Point& a, const Point& b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const Point& a, const Point& b) { return !(a == b); }
inline Point& operator+=(Point& a, const Point& b) { a.x += b.x; a.y += b.y; return a; }

struct Point3f {
    float x = 0, y = 0, z = 0;
    Point3f() {}
    Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
    Size(const Point& p) : width(p.x), height(p.y) {}
};
inline bool operator==(const Size& a, const Size& b) { return a.width == b.width && a.height == b.height; }
inline bool operator!=(const Size& a, const Size& b) { return !(a == b); }

struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
    Rect(const Point& tl, const Point& br) : x(std::min(tl.x, br.x)), y(std::min(tl.y, br.y)), width(std::max(tl.x, br.x) - x), height(std::max(tl.y, br.y) - y) {}
    Rect(const Point& tl, const Size& sz) : x(tl.x), y(tl.y), width(sz.width), height(sz.height) {}
    Point tl() const { return Point(x, y); }
    Point br() const { return Point(x + width, y + height); }
    Size size() const { return Size(width, height); }
};

inline int cvRound(double v) { return (int)lrint(v); }

template <typename T> struct RefDepth;
template <> struct RefDepth<uchar> { enum { value = CV_8U }; };
template <> struct RefDepth<int> { enum { value = CV_32S }; };
template <> struct RefDepth<float> { enum { value = CV_32F }; };

class Mat {
public:
    int rows = 0, cols = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* external, size_t step) : rows(r), cols(c), data_((uchar*)external), step_(step), type_(type) {}

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }          // create() zero-fills
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

    void create(int r, int c, int type) {
        if (data_ && r == rows && c == cols && type == type_) return;                      // cv::Mat::create: nothing to do
        rows = r; cols = c; type_ = type;
        step_ = (size_t)c * elem_size();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step_ + 64, (uchar)0);
        data_ = buf_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }

    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elem_size() const { return (size_t)channels() * (depth() == CV_8U ? 1 : 4); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }

    template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data_ + (size_t)y * step_); }
    template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data_ + (size_t)y * step_); }
    template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }

    Mat operator()(const Rect& r) const {                                                  // view onto the same storage
        Mat m;
        m.rows = r.height; m.cols = r.width; m.type_ = type_; m.step_ = step_; m.buf_ = buf_;
        m.data_ = data_ + (size_t)r.y * step_ + (size_t)r.x * elem_size();
        return m;
    }
    void copyTo(Mat& dst) const {                                                          // cv::Mat::copyTo: create() then copy
        dst.create(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data_ + (size_t)y * dst.step_, data_ + (size_t)y * step_, (size_t)cols * elem_size());
    }
    Mat& setTo(double v) {
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols * channels(); ++x) {
                if (depth() == CV_8U) ptr<uchar>(y)[x] = (uchar)v;
                else if (depth() == CV_32S) ptr<int>(y)[x] = (int)v;
                else ptr<float>(y)[x] = (float)v;
            }
        return *this;
    }
    Mat getMat(int /*access*/) const { return *this; }                                    // UMat::getMat

protected:
    std::shared_ptr<std::vector<uchar>> buf_;
    uchar* data_ = nullptr;
    size_t step_ = 0;
    int type_ = 0;
};

typedef Mat UMat;

template <typename T> class Mat_ : public Mat {
public:
    Mat_() { type_ = RefDepth<T>::value; }
    Mat_(const Mat& m) : Mat(m) { check(); }
    Mat_& operator=(const Mat& m) { Mat::operator=(m); check(); return *this; }
    void create(int r, int c) { Mat::create(r, c, RefDepth<T>::value); }
    void create(Size s) { Mat::create(s.height, s.width, RefDepth<T>::value); }
    T& operator()(int y, int x) { return this->template ptr<T>(y)[x]; }
    const T& operator()(int y, int x) const { return this->template ptr<T>(y)[x]; }
    T& operator()(Point p) { return this->template ptr<T>(p.y)[p.x]; }
    const T& operator()(Point p) const { return this->template ptr<T>(p.y)[p.x]; }
    Mat_ operator()(const Rect& r) const { return Mat_(Mat::operator()(r)); }

private:
    void check() const {
        if (!empty() && type() != RefDepth<T>::value) throw RefError(Error::StsAssert, "Mat_<T>: element type mismatch (no conversion in the shim)");
    }
};

// cv::floodFill(image, seed, newVal) with its defaults: loDiff = upDiff = 0, 4-connectivity; returns the filled area
inline int floodFill(Mat& image, Point seed, int newVal) {
    if (image.type() != CV_32SC1) throw RefError(Error::StsBadArg, "floodFill shim: CV_32SC1 only");
    const int old = image.at<int>(seed.y, seed.x);
    if (old == newVal) return 0;
    std::vector<Point> stack(1, seed);
    image.at<int>(seed.y, seed.x) = newVal;
    int area = 0;
    while (!stack.empty()) {
        const Point p = stack.back();
        stack.pop_back();
        ++area;
        const Point nb[4] = {Point(p.x - 1, p.y), Point(p.x + 1, p.y), Point(p.x, p.y - 1), Point(p.x, p.y + 1)};
        for (const Point& q : nb)
            if (q.x >= 0 && q.y >= 0 && q.x < image.cols && q.y < image.rows && image.at<int>(q.y, q.x) == old) {
                image.at<int>(q.y, q.x) = newVal;
                stack.push_back(q);
            }
    }
    return area;
}

// cv::partition (core/operations.hpp): equivalence classes of the transitive closure of `predicate`; class ids in order
// of first appearance
template <typename Tp, class Pred> int partition(const std::vector<Tp>& vec, std::vector<int>& labels, Pred predicate = Pred()) {
    const int N = (int)vec.size();
    std::vector<int> parent(N);
    for (int i = 0; i < N; ++i) parent[i] = i;
    auto root = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            if (i != j && predicate(vec[i], vec[j])) {
                const int a = root(i), b = root(j);
                if (a != b) parent[b] = a;
            }
    labels.assign(N, -1);
    std::vector<int> cls(N, -1);
    int n = 0;
    for (int i = 0; i < N; ++i) {
        const int r = root(i);
        if (cls[r] < 0) cls[r] = n++;
        labels[i] = cls[r];
    }
    return n;
}

// COLOR_GRAD only.  CV_32F: OpenCV's vector-body association; CV_8U: its 14-bit fixed point (B 1868, G 9617, R 4899).
inline void cvtColor(const Mat& src, Mat& dst, int code) {
    (void)code;
    const int cn = src.channels();   // cv::cvtColor takes the channel count from the source (BGR2GRAY accepts 3 or 4, alpha ignored)
    if (src.depth() == CV_32F) {
        dst.create(src.rows, src.cols, CV_32FC1);
        for (int y = 0; y < src.rows; ++y)
            for (int x = 0; x < src.cols; ++x) {
                const float* p = src.ptr<float>(y) + cn * x;
                dst.at<float>(y, x) = std::fmaf(p[2], 0.299f, std::fmaf(p[0], 0.114f, p[1] * 0.587f));
            }
    } else {
        dst.create(src.rows, src.cols, CV_8UC1);
        for (int y = 0; y < src.rows; ++y)
            for (int x = 0; x < src.cols; ++x) {
                const uchar* p = src.ptr<uchar>(y) + cn * x;
                dst.at<uchar>(y, x) = (uchar)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + (1 << 13)) >> 14);
            }
    }
}

inline void Sobel(const Mat& src, Mat& dst, int /*ddepth = CV_32F*/, int dx, int /*dy*/) {
    const int H = src.rows, W = src.cols;
    auto refl = [](int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); };
    auto g = [&](int y, int x) { return src.depth() == CV_32F ? src.at<float>(y, x) : (float)src.at<uchar>(y, x); };
    Mat rowf(H, W, CV_32FC1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float a = g(y, refl(x - 1, W)), b = g(y, x), c = g(y, refl(x + 1, W));
            rowf.at<float>(y, x) = dx ? (c - a) : ((a + c) + b * 2.f);
        }
    Mat out(H, W, CV_32FC1);
    for (int y = 0; y < H; ++y) {
        const int ya = refl(y - 1, H), yc = refl(y + 1, H);
        for (int x = 0; x < W; ++x)
            out.at<float>(y, x) = dx ? ((rowf.at<float>(ya, x) + rowf.at<float>(yc, x)) + rowf.at<float>(y, x) * 2.f)
                                     : (rowf.at<float>(yc, x) - rowf.at<float>(ya, x));
    }
    dst = out;
}

inline int64 getTickCount() { return 0; }
inline double getTickFrequency() { return 1.0; }

namespace detail {
template <typename T> static inline T sqr(T x) { return x * x; }                          // detail/util_inl.hpp
static inline float normL2(const Point3f& a, const Point3f& b) { return sqr(a.x - b.x) + sqr(a.y - b.y) + sqr(a.z - b.z); }
}

}  // namespace cv
