// oracle/linblend.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the reference's hand-written pair blend, [BLEND]:141-717: geometry, overlap cost
// map costV, greedy per-row seam, overlap classification masks, per-row left/right scan, seam-guided
// linear weights and the three-region composite.
//
// Parity pin: the reference's OWN block -- [BLEND]:141-717 cut out of its main() at build time and compiled against
// oracle/ref_shim/cvshim.h (`make -C oracle ref` -> oracle/_ref/libref_linblend.so).  This restatement reproduces it bit
// for bit (panorama incl. its NaNs, greedy seam, cost map): tests/test_oracle_reference_build.py, live where
// /root/reference exists and against tests/golden/linblend_ref_cases.npz everywhere.  Not covered by the pin: OpenCV's
// own cvtColor (the shim uses the scalar formula below; a gray value exactly at a threshold may round differently in
// OpenCV 3.4.2's SIMD body) and the out-of-buffer reads listed next.
//
// Defined behaviour where the reference reads out of bounds:
//   * cv::Mat_ buffers are continuous, so row-relative out-of-row reads of costV / mask_r2
//     ([BLEND]:276-278, :501) land in the neighbouring row; that is restated with flat indexing,
//     and anything outside the whole buffer reads as 0.
//   * image rows past the end of an image ([BLEND]:216-219 when panoHe_-dy2 > rows) are skipped.
// Precondition of the reference's index math: image 1 is the left image (tl2.x >= tl1.x).
//
// Compile with -ffp-contract=off.
#include "oracle.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

inline float sqr(float v) { return v * v; }

struct Geo {
    int panoBr, panoHe, dx2, dy, dy1, dy2, height, width, interSectBr;
    bool overlap;
};

Geo geometry(int rows1, int cols1, int rows2, int cols2, int tl1x, int tl1y, int tl2x, int tl2y) {
    Geo g;
    g.panoBr = tl2x - tl1x + cols2;                                                   // :152
    g.panoHe = std::max(tl1y + rows1, tl2y + rows2) - std::min(tl1y, tl2y);           // :153
    g.dx2 = tl2x - tl1x;                                                              // :158
    g.dy = tl2y - tl1y;                                                               // :162
    g.dy1 = g.dy < 0 ? -g.dy : 0;
    g.dy2 = g.dy > 0 ? g.dy : 0;
    int itx = std::max(tl1x, tl2x), ity = std::max(tl1y, tl2y);                       // :175-178
    int ibx = std::min(tl1x + cols1, tl2x + cols2), iby = std::min(tl1y + rows1, tl2y + rows2);
    g.overlap = !(itx >= ibx || ity >= iby);
    g.height = iby - ity;
    g.width = ibx - itx;
    g.interSectBr = cols1 - g.dx2;                                                    // :191
    return g;
}

struct FlatF {   // continuous float matrix with "flat" out-of-row semantics
    int rows, cols;
    std::vector<float> v;
    FlatF(int r, int c) : rows(r), cols(c), v((size_t)r * c, 0.f) {}
    inline float get(int y, long x) const {
        long i = (long)y * cols + x;
        return (i < 0 || i >= (long)v.size()) ? 0.f : v[(size_t)i];
    }
    inline float& at(int y, int x) { return v[(size_t)y * cols + x]; }
};

}  // namespace

extern "C" {

void orc_lin_geometry(int rows1, int cols1, int rows2, int cols2, int tl1x, int tl1y, int tl2x, int tl2y,
                      int* panoHe, int* panoBr) {
    Geo g = geometry(rows1, cols1, rows2, cols2, tl1x, tl1y, tl2x, tl2y);
    *panoHe = g.panoHe;
    *panoBr = g.panoBr;
}

int orc_lin_blend(const float* img1, int rows1, int cols1, const float* img2, int rows2, int cols2,
                  int tl1x, int tl1y, int tl2x, int tl2y, float* pano, int* seam_x, float* costV_out) {
    const Geo g = geometry(rows1, cols1, rows2, cols2, tl1x, tl1y, tl2x, tl2y);
    if (!g.overlap) return 1;                                                         // :182-183
    const int dx2 = g.dx2, dy = g.dy, dy1 = g.dy1, dy2 = g.dy2, height = g.height, width = g.width;
    const int IB = g.interSectBr, He = g.panoHe;
    auto row1 = [&](int y) { return img1 + (size_t)y * cols1 * 3; };
    auto row2 = [&](int y) { return img2 + (size_t)y * cols2 * 3; };

    // ---- costV  :206-261
    FlatF costV(He, IB + 2);
    int y0, y1, off2;
    if (dy > 0) { y0 = dy2; y1 = He - dy2; off2 = dy2; }
    else if (dy < 0) { y0 = dy1; y1 = He - dy1; off2 = dy1; }
    else { y0 = 0; y1 = std::min(rows1, rows2); off2 = 0; }
    for (int y = y0; y < y1; ++y) {
        if (y >= rows1 || y - off2 < 0 || y - off2 >= rows2) continue;   // reference: undefined
        const float* p1 = row1(y);
        const float* p2 = row2(y - off2);
        for (int x = 1; x < IB - 1; ++x) {
            if (x + dx2 + 1 >= cols1 || x >= cols2) continue;            // reference: undefined
            costV.at(y, x) =
                ((sqr(p1[(x + dx2) * 3] - p2[x * 3]) + sqr(p1[(x + dx2) * 3 + 1] - p2[x * 3 + 1]) + sqr(p1[(x + dx2) * 3 + 2] - p2[x * 3 + 2])) +
                 (sqr(p1[(x + dx2 + 1) * 3] - p2[(x - 1) * 3]) + sqr(p1[(x + dx2 + 1) * 3 + 1] - p2[(x - 1) * 3 + 1]) +
                  sqr(p1[(x + dx2 + 1) * 3 + 2] - p2[(x - 1) * 3 + 2]))) / 2;
        }
    }
    if (costV_out) std::memcpy(costV_out, costV.v.data(), costV.v.size() * sizeof(float));

    // ---- greedy seam  :268-307
    std::vector<int> seam(He);
    {
        int px = IB / 2, py = 0;
        seam[0] = px;
        while (py < He - 1) {
            float a = costV.get(py + 1, (long)px - 1);
            float b = costV.get(py + 1, (long)px);
            float c = costV.get(py + 1, (long)px + 1);
            if (a == b && a == c) { }
            else if (a <= b && a <= c) px -= 1;
            else if (b <= a && b <= c) { }
            else if (c <= a && c <= b) px += 1;
            else { /* NaN: the reference would spin forever; keep column */ }
            py += 1;
            seam[py] = px;
        }
    }
    for (int i = 0; i < He; ++i) seam_x[i] = seam[i];

    // ---- gray + classification  :311-470   (cvtColor RGB2GRAY on float: c0*0.299 + c1*0.587 + c2*0.114)
    auto gray = [](const float* p) { return p[0] * 0.299f + p[1] * 0.587f + p[2] * 0.114f; };
    FlatF m1(height, width + 2), m2(height, width + 2);
    const float thr = dy == 0 ? 10.f : 20.f;
    for (int y = 0; y < height; ++y) {
        m1.at(y, 0) = 128; m2.at(y, 0) = 128;
        m1.at(y, width + 1) = 128; m2.at(y, width + 1) = 128;
        const float* p1 = row1(dy > 0 ? y + dy2 : y);
        const float* p2 = row2(dy < 0 ? y + dy1 : y);
        for (int x = 1; x < width + 1; ++x) {
            float g1 = gray(p1 + (size_t)(x + dx2 - 1) * 3), g2 = gray(p2 + (size_t)(x - 1) * 3);
            if (g1 >= thr && g2 >= thr) { m1.at(y, x) = 255; m2.at(y, x) = 255; }
            if (g1 >= thr && g2 < thr) { m1.at(y, x) = 1; m2.at(y, x) = 0; }
            if (g1 < thr && g2 >= thr) { m1.at(y, x) = 0; m2.at(y, x) = 1; }
            if (g1 < thr && g2 < thr) { m1.at(y, x) = 1; m2.at(y, x) = 1; }
        }
    }

    // ---- left/right scan + weights  :475-558
    for (int y = 0; y < height; ++y) {
        int left = 0, right = 0;
        for (int x = 1; x < width + 1; ++x) {                                         // :494-506
            float c0 = m2.get(y, x - 1), c1 = m2.get(y, x), c2 = m2.get(y, x + 1);
            if (c1 == 255 && c0 == 0 && c2 == 1) left = x;
            if ((c1 == 255 && c0 == 0 && c2 == 255) ||
                (c0 == 128 && c1 == 255 && c2 == 255 && m2.get(y, x + 2) == 255 && m2.get(y, x + 3) == 255))
                left = x;
        }
        for (int x = 1; x < width + 1; ++x) {                                         // :511-523
            float c0 = m2.get(y, x - 1), c1 = m2.get(y, x), c2 = m2.get(y, x + 1);
            if (c0 == 0 && c1 == 255 && c2 == 1) right = x;
            if (c0 == 255 && c1 == 255 && (c2 == 1 || c2 == 128)) right = x;
        }
        const int sx = seam[y + dy2 + dy1];
        for (int x = 1; x < width + 1; ++x) {                                         // :531-552
            if (m2.at(y, x) == 255) {
                if (left && left == right) {
                    m1.at(y, x) = 1;
                    m2.at(y, x) = 0;
                } else if (x <= (sx + 1)) {
                    m1.at(y, x) = (float)(1 - 0.5 * (x - left) / (sx + 1 - left));
                    m2.at(y, x) = 1 - m1.at(y, x);
                } else if (x > (sx + 1) && x <= right) {
                    m1.at(y, x) = (float)(0.5 * (right - x) / (right - sx - 1));
                    m2.at(y, x) = 1 - m1.at(y, x);
                }
            }
        }
    }
    for (int y = 0; y < height; ++y)                                                  // :560-572
        for (int x = 0; x < width + 1; ++x)
            if (m1.at(y, x) == 255) { m1.at(y, x) = 1; m2.at(y, x) = 0; }

    // ---- composite  :579-711
    const int Br = g.panoBr;
    std::memset(pano, 0, sizeof(float) * (size_t)He * Br * 3);
    auto prow = [&](int y) { return pano + (size_t)y * Br * 3; };
    for (int y = 0; y < rows1; ++y) {                       // image 1, columns [0, dx2)
        const float* p1 = row1(y);
        float* p2 = prow(dy < 0 ? y + dy1 : y);
        for (int x = 0; x < dx2; ++x) { p2[3 * x] = p1[3 * x]; p2[3 * x + 1] = p1[3 * x + 1]; p2[3 * x + 2] = p1[3 * x + 2]; }
    }
    for (int y = (dy > 0 ? dy2 : 0); y < rows2; ++y) {      // image 2, columns [cols1, panoBr)
        const float* p1 = row2(dy > 0 ? y - dy2 : y);
        float* p2 = prow(y);
        for (int x = cols1; x < Br; ++x) {
            p2[3 * x] = p1[3 * (x - dx2)]; p2[3 * x + 1] = p1[3 * (x - dx2) + 1]; p2[3 * x + 2] = p1[3 * (x - dx2) + 2];
        }
    }
    for (int y = 0; y < height; ++y) {                      // overlap
        const float* p1 = row1(dy > 0 ? y + dy2 : y);
        const float* p2 = row2(dy < 0 ? y + dy1 : y);
        float* p3 = prow(dy > 0 ? y + dy2 : (dy < 0 ? y + dy1 : y));
        for (int x = dx2; x < dx2 + width; ++x) {
            float w1 = m1.at(y, x - dx2 + 1), w2 = m2.at(y, x - dx2 + 1);
            p3[3 * x] = p1[3 * x] * w1 + p2[3 * (x - dx2)] * w2;
            p3[3 * x + 1] = p1[3 * x + 1] * w1 + p2[3 * (x - dx2) + 1] * w2;
            p3[3 * x + 2] = p1[3 * x + 2] * w1 + p2[3 * (x - dx2) + 2] * w2;
        }
    }
    return 0;
}

}  // extern "C"
