// oracle/ref_shim/seam_ref.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own refactored DP seam finder: its forward declarations, globals and functions from find() to
// updateLabelsUsingSeam ([SEAM]:29-1093) are included below from a file that oracle/Makefile extracts from
// /root/reference at build time (deleted again after compiling), against oracle/ref_shim/cvshim_seam.h.
#include "cvshim_seam.h"

#include <iostream>
#include <limits>
#include <map>

using namespace cv;
using namespace std;
using namespace detail;

// Seam capture without touching the reference's text: its resolveConflicts() calls
//     updateLabelsUsingSeam(c1, c2, seam, isHorizontalSeam)          [SEAM]:470
// with a NON-const std::vector<Point> lvalue, while its own function takes a const reference.  The overload declared
// here takes a non-const reference, so overload resolution picks it at that call site; it records the seam (the
// vector estimateSeam() produced) and forwards to the reference's function.
static std::vector<int32_t> g_seam_trace;   // per seam: comp, horizontal, npts, then npts x (x, y) in panorama coordinates
void updateLabelsUsingSeam(int comp1, int comp2, std::vector<Point>& seam, bool isHorizontalSeam);

#include "seam_block.inc"

void updateLabelsUsingSeam(int comp1, int comp2, std::vector<Point>& seam, bool isHorizontalSeam) {
    g_seam_trace.push_back(comp1);
    g_seam_trace.push_back(isHorizontalSeam ? 1 : 0);
    g_seam_trace.push_back((int32_t)seam.size());
    for (const Point& p : seam) {
        g_seam_trace.push_back(p.x + unionTl_.x);
        g_seam_trace.push_back(p.y + unionTl_.y);
    }
    updateLabelsUsingSeam(comp1, comp2, static_cast<const std::vector<Point>&>(seam), isHorizontalSeam);
}

// images: n pointers to tightly packed rows x cols x cn (uint8 or float32); masks: n pointers, modified in place.
// is_u8: bit 0 = 8-bit images, bit 1 = four channels (CV_8UC4 / CV_32FC4, [SEAM]:745-748) instead of three.
// returns 0, or the cv::Error code the reference's CV_Assert / CV_Error raised.
extern "C" int ref_dp_seam_find(int n, const void* const* images, int is_u8, const int* rows, const int* cols, const int* corners_xy,
                                uint8_t* const* masks, int cost_fn) {
    try {
        std::vector<UMat> src(n), msk(n);
        std::vector<Point> corners(n);
        for (int i = 0; i < n; ++i) {
            const bool u8 = (is_u8 & 1) != 0;
            const int cn = (is_u8 & 2) ? 4 : 3;
            const int type = u8 ? (cn == 4 ? CV_8UC4 : CV_8UC3) : (cn == 4 ? CV_32FC4 : CV_32FC3);
            src[i] = Mat(rows[i], cols[i], type, const_cast<void*>(images[i]), (size_t)cols[i] * cn * (u8 ? 1 : 4));
            msk[i] = Mat(rows[i], cols[i], CV_8UC1, masks[i], (size_t)cols[i]);
            corners[i] = Point(corners_xy[2 * i], corners_xy[2 * i + 1]);
        }
        costFunc_ = cost_fn ? COLOR_GRAD : COLOR;
        g_seam_trace.clear();
        find(src, corners, msk);
        return 0;
    } catch (const RefError& e) {
        return e.code;
    }
}

// the seams of the last ref_dp_seam_find call, in the order they were estimated: returns the number of int32 values and
// copies at most cap of them
extern "C" size_t ref_last_seam_trace(int32_t* out, size_t cap) {
    const size_t n = g_seam_trace.size();
    if (out) std::memcpy(out, g_seam_trace.data(), sizeof(int32_t) * std::min(n, cap));
    return n;
}

// computeCosts ([SEAM]:733-803) alone: the globals it reads are set from the arguments (labels: H x W int32 in the union
// frame, component label l with bounding box roi = x, y, w, h).  costV: h x (w+1), costH: (h+1) x w.
extern "C" int ref_seam_costs(const void* img1, const void* img2, int is_u8, int rows1, int cols1, int rows2, int cols2, int tl1x, int tl1y,
                              int tl2x, int tl2y, const int32_t* labels, int H, int W, int union_tlx, int union_tly, int l, const int roi[4],
                              int cost_fn, float* costV_out, float* costH_out) {
    try {
        const bool u8 = (is_u8 & 1) != 0;
        const int cn = (is_u8 & 2) ? 4 : 3;
        const int type = u8 ? (cn == 4 ? CV_8UC4 : CV_8UC3) : (cn == 4 ? CV_32FC4 : CV_32FC3);
        const size_t es = u8 ? 1 : 4;
        Mat image1(rows1, cols1, type, const_cast<void*>(img1), (size_t)cols1 * cn * es);
        Mat image2(rows2, cols2, type, const_cast<void*>(img2), (size_t)cols2 * cn * es);
        unionTl_ = Point(union_tlx, union_tly);
        unionBr_ = Point(union_tlx + W, union_tly + H);
        unionSize_ = Size(W, H);
        labels_.create(unionSize_);
        for (int y = 0; y < H; ++y) std::memcpy(labels_.ptr<int>(y), labels + (size_t)y * W, sizeof(int) * (size_t)W);
        const int comp = l - 1;
        states_.assign(comp + 1, INTERS);
        tls_.assign(comp + 1, Point(0, 0));
        brs_.assign(comp + 1, Point(0, 0));
        tls_[comp] = Point(roi[0], roi[1]);
        brs_[comp] = Point(roi[0] + roi[2], roi[1] + roi[3]);
        costFunc_ = cost_fn ? COLOR_GRAD : COLOR;
        if (costFunc_ == COLOR_GRAD) computeGradients(image1, image2);
        Mat_<float> costV, costH;
        computeCosts(image1, image2, Point(tl1x, tl1y), Point(tl2x, tl2y), comp, costV, costH);
        for (int y = 0; y < costV.rows; ++y) std::memcpy(costV_out + (size_t)y * costV.cols, costV.ptr<float>(y), sizeof(float) * (size_t)costV.cols);
        for (int y = 0; y < costH.rows; ++y) std::memcpy(costH_out + (size_t)y * costH.cols, costH.ptr<float>(y), sizeof(float) * (size_t)costH.cols);
        return 0;
    } catch (const RefError& e) {
        return e.code;
    }
}

