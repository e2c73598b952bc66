// oracle/orb.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the reference's ORB features finder:
//   find                  [FEAT]:948-1021   (gray conversion, 3 x 1 grid of cells, one detectAndCompute per cell)
//   detectAndCompute      [FEAT]:727-946    (scale pyramid by INTER_LINEAR_EXACT resize, key points, blur, descriptors)
//   computeKeyPoints      [FEAT]:56-191     (per level: FAST(20, nonmax) -> border filter -> retainBest(2 n) -> Harris ->
//                                            retainBest(n); then IC angles, scaling of the points)
//   HarrisResponses       [FEAT]:205-248
//   ICAngles              [FEAT]:250-283
//   computeOrbDescriptors [FEAT]:288-418    (wta_k = 2, the configuration of the reference)
// and of the OpenCV calls it delegates to (un-vendored; pinned to cv2 4.13 in tests/test_oracle_orb.py): cvtColor(BGR2GRAY) 8-bit,
// resize(INTER_LINEAR_EXACT), FAST-9/16 with non-maximum suppression, KeyPointsFilter::runByImageBorder / retainBest, fastAtan2,
// GaussianBlur(7 x 7, sigma 2) on an 8-bit sub-matrix (sepFilter2D with the float kernel).
//
// Borders: every pixel a retained key point reads -- FAST radius 3, Harris 4, orientation 15, descriptor ceil(15 sqrt 2) + blur 3 =
// 25 -- lies inside its level because runByImageBorder keeps points at least edgeThreshold = 31 pixels from the edge, so the
// copyMakeBorder frames of [FEAT]:776-840 never reach the output; this restatement works on the bare levels (reads outside
// clamp, for the blur's own three edge pixels, which nothing reads).
//
// Compile with -ffp-contract=off.
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

static const int kBitPattern31[256 * 4] = {
#include "orb_pattern.inc"
};

struct Img {
    int rows = 0, cols = 0;
    std::vector<uint8_t> d;
    Img() {}
    Img(int r, int c) : rows(r), cols(c), d((size_t)r * c) {}
    uint8_t* row(int y) { return d.data() + (size_t)y * cols; }
    const uint8_t* row(int y) const { return d.data() + (size_t)y * cols; }
    uint8_t at(int y, int x) const { return d[(size_t)std::min(std::max(y, 0), rows - 1) * cols + std::min(std::max(x, 0), cols - 1)]; }
};

struct KeyPt { float x, y, size, angle, response; int octave; };

inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvRoundF(float v) { return (int)lrintf(v); }

// cvtColor(COLOR_BGR2GRAY / BGRA2GRAY), 8 bit: 15-bit coefficients (OpenCV 4: BY15 3735, GY15 19235, RY15 9798)
void bgr2gray(const uint8_t* src, int rows, int cols, int ch, size_t step, Img& g) {
    g = Img(rows, cols);
    for (int y = 0; y < rows; ++y) {
        const uint8_t* s = src + (size_t)y * step;
        uint8_t* d = g.row(y);
        for (int x = 0; x < cols; ++x, s += ch) d[x] = (uint8_t)((s[0] * 3735 + s[1] * 19235 + s[2] * 9798 + (1 << 14)) >> 15);
    }
}

// resize(src, dst, dsize, 0, 0, INTER_LINEAR_EXACT) for 8UC1: resize_bitExact with interpolationLinear<ufixedpoint16>.  The source
// coordinate is evaluated in (soft)double -- IEEE double here --, the fraction rounded to 8 bits, rows and columns interpolated in
// 8.8 / 16.16 fixed point with one rounding at the end.
void linear_coeffs(int src, int dst, std::vector<int>& ofs, std::vector<int>& c1) {
    ofs.resize(dst); c1.resize(dst);
    const double inv_scale = (double)dst / src;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        const double f = scale * ((double)d + 0.5) - 0.5;
        const int i = (int)std::floor(f);
        if (i >= 0 && src > 1) {
            if (i < src - 1) { ofs[d] = i; c1[d] = cvRoundD((f - (double)i) * 256.0); }
            else { ofs[d] = src - 1; c1[d] = 0; }
        } else { ofs[d] = 0; c1[d] = 0; }
    }
}

void resize_linear_exact(const Img& s, Img& d, int drows, int dcols) {
    d = Img(drows, dcols);
    std::vector<int> xo, xc, yo, yc;
    linear_coeffs(s.cols, dcols, xo, xc);
    linear_coeffs(s.rows, drows, yo, yc);
    std::vector<uint32_t> h0(dcols), h1(dcols);
    for (int y = 0; y < drows; ++y) {
        const uint8_t* r0 = s.row(yo[y]);
        const uint8_t* r1 = s.row(std::min(yo[y] + 1, s.rows - 1));
        for (int x = 0; x < dcols; ++x) {
            const int a = xo[x], b = std::min(a + 1, s.cols - 1);
            h0[x] = (uint32_t)(256 - xc[x]) * r0[a] + (uint32_t)xc[x] * r0[b];      // ufixedpoint16, 8.8
            h1[x] = (uint32_t)(256 - xc[x]) * r1[a] + (uint32_t)xc[x] * r1[b];
        }
        uint8_t* o = d.row(y);
        for (int x = 0; x < dcols; ++x) {
            const uint32_t v = (uint32_t)(256 - yc[y]) * h0[x] + (uint32_t)yc[y] * h1[x];   // ufixedpoint32, 16.16
            o[x] = (uint8_t)std::min<uint32_t>(255u, (v + (1u << 15)) >> 16);
        }
    }
}

// cv::FAST(img, keypoints, threshold, true) = FAST_t<16>: 9 contiguous pixels of the 16-pixel circle all brighter than v + t or
// all darker than v - t; score = the largest threshold for which the pixel stays a corner (cornerScore<16>); a corner is kept
// when its score is strictly greater than those of its eight neighbours; key points come out in raster order.
const int kCircle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

int fast_score(const Img& g, int x, int y, int threshold) {      // 0: not a corner
    const int v = g.row(y)[x];
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - g.row(y + kCircle[k][1])[x + kCircle[k][0]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    bool corner = false;
    int best = 0;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        if (mn > threshold || mx < -threshold) corner = true;
        best = std::max(best, std::max(mn, -mx));
    }
    return corner ? best - 1 : 0;
}

void fast_detect(const Img& g, int threshold, std::vector<KeyPt>& out) {
    out.clear();
    if (g.rows < 7 || g.cols < 7) return;
    std::vector<int> score((size_t)g.rows * g.cols, 0);
    for (int y = 3; y < g.rows - 3; ++y)
        for (int x = 3; x < g.cols - 3; ++x) score[(size_t)y * g.cols + x] = fast_score(g, x, y, threshold);
    for (int y = 3; y < g.rows - 3; ++y)
        for (int x = 3; x < g.cols - 3; ++x) {
            const int s = score[(size_t)y * g.cols + x];
            if (s == 0) continue;
            bool keep = true;
            for (int dy = -1; dy <= 1 && keep; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    if ((dx || dy) && score[(size_t)(y + dy) * g.cols + x + dx] >= s) { keep = false; break; }
            if (keep) out.push_back(KeyPt{(float)x, (float)y, 7.f, -1.f, (float)s, 0});
        }
}

// KeyPointsFilter::runByImageBorder: keep points inside Rect(Point(b, b), Point(cols - b, rows - b))
void run_by_image_border(std::vector<KeyPt>& k, int cols, int rows, int b) {
    if (b <= 0) return;
    if (rows <= b * 2 || cols <= b * 2) { k.clear(); return; }
    std::vector<KeyPt> o;
    for (const KeyPt& p : k)
        if (p.x >= (float)b && p.x < (float)(cols - b) && p.y >= (float)b && p.y < (float)(rows - b)) o.push_back(p);
    k.swap(o);
}

// KeyPointsFilter::retainBest: the n strongest, plus everything that ties with the n-th
void retain_best(std::vector<KeyPt>& k, int n) {
    if (n < 0 || k.size() <= (size_t)n) return;
    if (n == 0) { k.clear(); return; }
    std::nth_element(k.begin(), k.begin() + n - 1, k.end(), [](const KeyPt& a, const KeyPt& b) { return a.response > b.response; });
    const float amb = k[(size_t)n - 1].response;
    auto e = std::partition(k.begin() + n, k.end(), [amb](const KeyPt& a) { return a.response >= amb; });
    k.resize((size_t)(e - k.begin()));
}

void harris_responses(const Img& g, std::vector<KeyPt>& pts, int blockSize, float harris_k) {   // [FEAT]:205-248
    const int r = blockSize / 2;
    float scale = 1.f / ((1 << 2) * blockSize * 255.f);
    float scale_sq_sq = scale * scale * scale * scale;
    for (KeyPt& p : pts) {
        const int x0 = cvRoundF(p.x), y0 = cvRoundF(p.y);
        int a = 0, b = 0, c = 0;
        for (int i = 0; i < blockSize; ++i)
            for (int j = 0; j < blockSize; ++j) {
                const int x = x0 - r + j, y = y0 - r + i;
                const int Ix = (g.at(y, x + 1) - g.at(y, x - 1)) * 2 + (g.at(y - 1, x + 1) - g.at(y - 1, x - 1)) + (g.at(y + 1, x + 1) - g.at(y + 1, x - 1));
                const int Iy = (g.at(y + 1, x) - g.at(y - 1, x)) * 2 + (g.at(y + 1, x - 1) - g.at(y - 1, x - 1)) + (g.at(y + 1, x + 1) - g.at(y - 1, x + 1));
                a += Ix * Ix;
                b += Iy * Iy;
                c += Ix * Iy;
            }
        p.response = ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
    }
}

// cv::fastAtan2 (degrees, the scalar polynomial of mathfuncs_core)
float fast_atan2(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / 3.141592653589793238462643383279502884197169399375);
    static const float p3 = -0.3258083974640975f * (float)(180 / 3.141592653589793238462643383279502884197169399375);
    static const float p5 = 0.1555786518463281f * (float)(180 / 3.141592653589793238462643383279502884197169399375);
    static const float p7 = -0.04432655554792128f * (float)(180 / 3.141592653589793238462643383279502884197169399375);
    float ax = std::abs(x), ay = std::abs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

void umax_table(int half, std::vector<int>& umax) {                 // [FEAT]:85-100
    umax.assign((size_t)half + 2, 0);
    int v, v0, vmax = (int)std::floor(half * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(half * std::sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(std::sqrt((double)half * half - v * v));
    for (v = half, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

void ic_angles(const Img& g, std::vector<KeyPt>& pts, const std::vector<int>& u_max, int half_k) {   // [FEAT]:250-283
    for (KeyPt& p : pts) {
        const int cx = cvRoundF(p.x), cy = cvRoundF(p.y);
        int m_01 = 0, m_10 = 0;
        for (int u = -half_k; u <= half_k; ++u) m_10 += u * g.at(cy, cx + u);
        for (int v = 1; v <= half_k; ++v) {
            int v_sum = 0;
            int d = u_max[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = g.at(cy + v, cx + u), val_minus = g.at(cy - v, cx + u);
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        p.angle = fast_atan2((float)m_01, (float)m_10);
    }
}

// GaussianBlur(m, m, Size(7, 7), 2, 2, BORDER_REFLECT_101) on an 8-bit SUB-matrix ([FEAT]:921-926).  OpenCV's bit-exact fixed-point
// Gaussian needs an isolated matrix, so this call falls through to sepFilter2D with the CV_32F kernel getGaussianKernel(7, 2):
// rows as float sums left to right (RowFilter<uchar, float>), columns symmetric from the centre outwards
// (SymmColumnFilter<Cast<float, uchar>>), saturate_cast<uchar> (round half to even) at the end.  Pinned to cv2.sepFilter2D and,
// through the descriptors, to cv2.ORB.  Every accumulation step is a fused multiply-add: that is what OpenCV's filter code
// compiles to in the AVX2 / FMA3 dispatch every current x86 host selects (measured: 0 of 4.5 M pixels differ from cv2 with fused
// steps; with separate multiplies and adds -- cv2 after setUseOptimized(false), the SSE baseline -- about one rounded byte in
// 10^5 differs).
void gaussian_kernel7(float k[7]) {                                // getGaussianKernel(7, 2, CV_32F): double arithmetic, float result
    const double sigma = 2.0, scale2X = -0.5 / (sigma * sigma);
    double v[7], sum = 0;
    for (int i = 0; i < 7; ++i) { const double x = i - 3; v[i] = std::exp(scale2X * x * x); sum += v[i]; }
    sum = 1. / sum;
    for (int i = 0; i < 7; ++i) k[i] = (float)(v[i] * sum);
}

void gaussian7(const Img& s, Img& d) {
    float k[7];
    gaussian_kernel7(k);
    d = Img(s.rows, s.cols);
    std::vector<float> h((size_t)s.rows * s.cols);
    for (int y = 0; y < s.rows; ++y)
        for (int x = 0; x < s.cols; ++x) {
            float acc = k[0] * (float)s.at(y, x - 3);
            for (int t = 1; t < 7; ++t) acc = std::fmaf(k[t], (float)s.at(y, x - 3 + t), acc);
            h[(size_t)y * s.cols + x] = acc;
        }
    auto H = [&](int y, int x) { return h[(size_t)std::min(std::max(y, 0), s.rows - 1) * s.cols + x]; };
    for (int y = 0; y < s.rows; ++y)
        for (int x = 0; x < s.cols; ++x) {
            float acc = std::fmaf(k[3], H(y, x), 0.f);
            for (int t = 1; t <= 3; ++t) acc = std::fmaf(k[3 + t], H(y + t, x) + H(y - t, x), acc);
            const int r = cvRoundF(acc);
            d.row(y)[x] = (uint8_t)std::min(255, std::max(0, r));
        }
}

void orb_descriptors(const std::vector<Img>& blurred, const std::vector<float>& layerScale, const std::vector<KeyPt>& kps, uint8_t* desc) {   // [FEAT]:288-418, wta_k = 2
    for (size_t j = 0; j < kps.size(); ++j) {
        const KeyPt& kpt = kps[j];
        const Img& g = blurred[(size_t)kpt.octave];
        float scale = 1.f / layerScale[(size_t)kpt.octave];
        float angle = kpt.angle;
        angle *= (float)(3.1415926535897932384626433832795 / 180.f);
        float a = (float)cosf(angle), b = (float)sinf(angle);
        const int cy = cvRoundF(kpt.y * scale), cx = cvRoundF(kpt.x * scale);
        const int* pattern = kBitPattern31;
        auto value = [&](int idx) {
            float x = pattern[2 * idx] * a - pattern[2 * idx + 1] * b;
            float y = pattern[2 * idx] * b + pattern[2 * idx + 1] * a;
            return (int)g.at(cy + cvRoundF(y), cx + cvRoundF(x));
        };
        for (int i = 0; i < 32; ++i, pattern += 32) {
            int val = 0;
            for (int t = 0; t < 8; ++t) val |= (value(2 * t) < value(2 * t + 1)) << t;
            desc[j * 32 + (size_t)i] = (uint8_t)val;
        }
    }
}

// detectAndCompute [FEAT]:727-946 on one gray cell
void detect_and_compute(const Img& image, int nfeatures, double scaleFactor, int nlevels, int edgeThreshold, int patchSize, int fastThreshold,
                        std::vector<KeyPt>& all, std::vector<uint8_t>& desc) {
    std::vector<Img> levels((size_t)nlevels);
    std::vector<float> layerScale((size_t)nlevels);
    for (int level = 0; level < nlevels; ++level) {
        float scale = (float)std::pow(scaleFactor, (double)level);
        layerScale[(size_t)level] = scale;
        const int w = cvRoundF(image.cols / scale), h = cvRoundF(image.rows / scale);
        if (level == 0) levels[0] = image;
        else resize_linear_exact(levels[(size_t)level - 1], levels[(size_t)level], h, w);
    }
    // computeKeyPoints [FEAT]:56-191
    std::vector<int> nfeaturesPerLevel((size_t)nlevels);
    float factor = (float)(1.0 / scaleFactor);
    float ndesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sumFeatures = 0;
    for (int level = 0; level < nlevels - 1; ++level) {
        nfeaturesPerLevel[(size_t)level] = cvRoundF(ndesired);
        sumFeatures += nfeaturesPerLevel[(size_t)level];
        ndesired *= factor;
    }
    nfeaturesPerLevel[(size_t)nlevels - 1] = std::max(nfeatures - sumFeatures, 0);
    const int half = patchSize / 2;
    std::vector<int> umax;
    umax_table(half, umax);
    all.clear();
    std::vector<int> counters((size_t)nlevels);
    std::vector<KeyPt> kp;
    for (int level = 0; level < nlevels; ++level) {
        const int featuresNum = nfeaturesPerLevel[(size_t)level];
        const Img& img = levels[(size_t)level];
        fast_detect(img, fastThreshold, kp);
        run_by_image_border(kp, img.cols, img.rows, edgeThreshold);
        retain_best(kp, 2 * featuresNum);
        counters[(size_t)level] = (int)kp.size();
        for (KeyPt& p : kp) { p.octave = level; p.size = patchSize * layerScale[(size_t)level]; }
        all.insert(all.end(), kp.begin(), kp.end());
    }
    if (all.empty()) { desc.clear(); return; }
    {
        std::vector<KeyPt> next;
        size_t offset = 0;
        for (int level = 0; level < nlevels; ++level) {
            kp.assign(all.begin() + (long)offset, all.begin() + (long)offset + counters[(size_t)level]);
            offset += (size_t)counters[(size_t)level];
            harris_responses(levels[(size_t)level], kp, 7, 0.04f);
            retain_best(kp, nfeaturesPerLevel[(size_t)level]);
            ic_angles(levels[(size_t)level], kp, umax, half);
            next.insert(next.end(), kp.begin(), kp.end());
        }
        all.swap(next);
    }
    for (KeyPt& p : all) { float s = layerScale[(size_t)p.octave]; p.x *= s; p.y *= s; }
    std::vector<Img> blurred((size_t)nlevels);
    for (int level = 0; level < nlevels; ++level) gaussian7(levels[(size_t)level], blurred[(size_t)level]);
    desc.resize(all.size() * 32);
    orb_descriptors(blurred, layerScale, all, desc.data());
}

}  // namespace

extern "C" {

void orc_bgr2gray(const uint8_t* src, int rows, int cols, int ch, size_t step, uint8_t* dst) {
    Img g;
    bgr2gray(src, rows, cols, ch, step, g);
    std::memcpy(dst, g.d.data(), g.d.size());
}

void orc_resize_linear_exact_u8(const uint8_t* src, int rows, int cols, uint8_t* dst, int drows, int dcols) {
    Img s(rows, cols), d;
    std::memcpy(s.d.data(), src, s.d.size());
    resize_linear_exact(s, d, drows, dcols);
    std::memcpy(dst, d.d.data(), d.d.size());
}

/* cv::FAST(img, kps, threshold, true): out = x, y, score per key point, raster order; returns the count (capped) */
int orc_fast(const uint8_t* src, int rows, int cols, int threshold, int* out, int cap) {
    Img s(rows, cols);
    std::memcpy(s.d.data(), src, s.d.size());
    std::vector<KeyPt> k;
    fast_detect(s, threshold, k);
    int n = 0;
    for (const KeyPt& p : k) {
        if (n >= cap) break;
        out[3 * n] = (int)p.x; out[3 * n + 1] = (int)p.y; out[3 * n + 2] = (int)p.response;
        ++n;
    }
    return (int)k.size();
}

void orc_gaussian7_u8(const uint8_t* src, int rows, int cols, uint8_t* dst) {
    Img s(rows, cols), d;
    std::memcpy(s.d.data(), src, s.d.size());
    gaussian7(s, d);
    std::memcpy(dst, d.d.data(), d.d.size());
}

float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }

/* find() [FEAT]:948-1021: image 8UC1 / 8UC3 / 8UC4, grid_w x grid_h cells.  kps: 6 floats per key point (x, y, size, angle,
 * response, octave); desc: 32 bytes per key point.  Returns the number of key points (outputs filled up to cap). */
int orc_orb_find(const uint8_t* img, int rows, int cols, int channels, size_t step, int grid_w, int grid_h, int nfeatures, float scale_factor,
                 int nlevels, float* kps, uint8_t* desc, int cap) {
    Img gray;
    if (channels == 1) {
        gray = Img(rows, cols);
        for (int y = 0; y < rows; ++y) std::memcpy(gray.row(y), img + (size_t)y * step, (size_t)cols);
    } else {
        bgr2gray(img, rows, cols, channels, step, gray);
    }
    int n = 0;
    for (int r = 0; r < grid_h; ++r)
        for (int c = 0; c < grid_w; ++c) {
            const int xl = c * cols / grid_w, yl = r * rows / grid_h, xr = (c + 1) * cols / grid_w, yr = (r + 1) * rows / grid_h;
            Img part(yr - yl, xr - xl);
            for (int y = yl; y < yr; ++y) std::memcpy(part.row(y - yl), gray.row(y) + xl, (size_t)(xr - xl));
            std::vector<KeyPt> k;
            std::vector<uint8_t> d;
            detect_and_compute(part, nfeatures, (double)scale_factor, nlevels, 31, 31, 20, k, d);
            for (size_t i = 0; i < k.size(); ++i, ++n) {
                if (n >= cap) continue;
                float* o = kps + 6 * (size_t)n;
                o[0] = k[i].x + xl; o[1] = k[i].y + yl; o[2] = k[i].size; o[3] = k[i].angle; o[4] = k[i].response; o[5] = (float)k[i].octave;
                std::memcpy(desc + 32 * (size_t)n, d.data() + 32 * i, 32);
            }
        }
    return n;
}

}  // extern "C"
