// oracle/feather.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the blend path the reference's mains actually execute ([SEAM]:1249-1280, [WARP]:276-313):
//   Mat element = getStructuringElement(MORPH_RECT, Size(20, 20));
//   dilate(masks_seam[k], masks_seam[k], element);  masks_seam[k] = masks_seam[k] & masks_warped[k];      [SEAM]:1257-1270
//   blender = Blender::createDefault(Blender::FEATHER); fb->setSharpness(0.1); prepare / feed / blend       [SEAM]:1249-1252,1271,1280
// The arithmetic lives in un-vendored OpenCV 3.4.2 (imgproc dilate / distanceTransform, stitching blenders.cpp);
// restated from its published algorithm and pinned against OpenCV 4.13 (tests/test_oracle_cv2.py):
//   dilate, rectangular element kw x kh, anchor at the centre (kw/2, kh/2): dst(x,y) = max of src over
//     x - kw/2 .. x - kw/2 + kw - 1 (same in y); outside the image does not count.
//   createWeightMap: w = min(distanceTransform(mask, DIST_L1, 3) * sharpness, 1)  -- the 3x3 L1 chamfer is the exact
//     city-block distance to the nearest zero pixel (FLT_MAX when the mask has none), product in float.
//   feed: dst[tl + p] += (short)(src * w) per channel (float product, truncation), wsum += w (float, feed order).
//   blend: dst = (short)(dst / (wsum + 1e-5f)), dst_mask = wsum > 1e-5f, dst = 0 where the mask is 0.
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstring>
#include <vector>

extern "C" {

void orc_dilate_rect(const uint8_t* src, int rows, int cols, int kw, int kh, uint8_t* dst) {
    std::vector<uint8_t> tmp((size_t)rows * cols);
    const int ax = kw / 2, ay = kh / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int m = 0;
            for (int k = std::max(0, x - ax); k <= std::min(cols - 1, x - ax + kw - 1); ++k) m = std::max(m, (int)src[(size_t)y * cols + k]);
            tmp[(size_t)y * cols + x] = (uint8_t)m;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int m = 0;
            for (int k = std::max(0, y - ay); k <= std::min(rows - 1, y - ay + kh - 1); ++k) m = std::max(m, (int)tmp[(size_t)k * cols + x]);
            dst[(size_t)y * cols + x] = (uint8_t)m;
        }
}

// distanceTransform(mask, DIST_L1, 3) as CV_32F: exact city-block distance to the nearest zero pixel
void orc_distance_l1(const uint8_t* mask, int rows, int cols, float* dist) {
    const int INF = INT_MAX / 4;
    std::vector<int> d((size_t)rows * cols);
    for (int y = 0; y < rows; ++y) {                    // along the rows
        int* r = d.data() + (size_t)y * cols;
        const uint8_t* m = mask + (size_t)y * cols;
        int run = INF;
        for (int x = 0; x < cols; ++x) { run = m[x] ? (run >= INF ? INF : run + 1) : 0; r[x] = run; }
        run = INF;
        for (int x = cols - 1; x >= 0; --x) { run = m[x] ? (run >= INF ? INF : run + 1) : 0; r[x] = std::min(r[x], run); }
    }
    for (int x = 0; x < cols; ++x) {                    // along the columns: min-plus with |dy|
        for (int y = 1; y < rows; ++y) { int& v = d[(size_t)y * cols + x]; const int u = d[(size_t)(y - 1) * cols + x]; if (u < INF && u + 1 < v) v = u + 1; }
        for (int y = rows - 2; y >= 0; --y) { int& v = d[(size_t)y * cols + x]; const int u = d[(size_t)(y + 1) * cols + x]; if (u < INF && u + 1 < v) v = u + 1; }
    }
    for (size_t i = 0; i < (size_t)rows * cols; ++i) dist[i] = d[i] >= INF ? FLT_MAX : (float)d[i];
}

void orc_feather_weight(const uint8_t* mask, int rows, int cols, float sharpness, float* weight) {
    orc_distance_l1(mask, rows, cols, weight);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
        const float v = weight[i] * sharpness;           // multiply(weight, sharpness, tmp)
        weight[i] = v > 1.f ? 1.f : v;                   // threshold(tmp, weight, 1.f, 1.f, THRESH_TRUNC)
    }
}

struct orc_fb {
    float sharpness;
    int roi[4];
    std::vector<int16_t> dst;
    std::vector<float> wsum;
};

orc_fb* orc_fb_create(float sharpness) { orc_fb* f = new orc_fb(); f->sharpness = sharpness; return f; }
void orc_fb_destroy(orc_fb* f) { delete f; }

void orc_fb_prepare(orc_fb* f, const int* roi_xywh) {
    std::memcpy(f->roi, roi_xywh, sizeof(int) * 4);
    f->dst.assign((size_t)roi_xywh[2] * roi_xywh[3] * 3, 0);
    f->wsum.assign((size_t)roi_xywh[2] * roi_xywh[3], 0.f);
}

void orc_fb_feed(orc_fb* f, const int16_t* img, const uint8_t* mask, int rows, int cols, int tlx, int tly) {
    std::vector<float> w((size_t)rows * cols);
    orc_feather_weight(mask, rows, cols, f->sharpness, w.data());
    const int dx = tlx - f->roi[0], dy = tly - f->roi[1], W = f->roi[2];
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const float wv = w[(size_t)y * cols + x];
            const size_t o = (size_t)(dy + y) * W + dx + x;
            for (int c = 0; c < 3; ++c)
                f->dst[3 * o + c] = (int16_t)(f->dst[3 * o + c] + (int16_t)((float)img[((size_t)y * cols + x) * 3 + c] * wv));
            f->wsum[o] += wv;
        }
}

void orc_fb_blend(orc_fb* f, int16_t* pano, uint8_t* pano_mask) {
    const size_t n = (size_t)f->roi[2] * f->roi[3];
    for (size_t o = 0; o < n; ++o) {
        const float ws = f->wsum[o];
        const bool on = ws > 1e-5f;
        for (int c = 0; c < 3; ++c) pano[3 * o + c] = on ? (int16_t)((float)f->dst[3 * o + c] / (ws + 1e-5f)) : (int16_t)0;
        pano_mask[o] = on ? 255 : 0;
    }
}

}  // extern "C"
