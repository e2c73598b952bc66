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

#include "seam_block.inc"

// images: n pointers to tightly packed rows x cols x 3 (uint8 or float32); masks: n pointers, modified in place.
// returns 0, or the cv::Error code the reference's CV_Assert / CV_Error raised.
extern "C" int ref_dp_seam_find(int n, const void* const* images, int is_u8, const int* rows, const int* cols, const int* corners_xy,
                                uint8_t* const* masks, int cost_fn) {
    try {
        std::vector<UMat> src(n), msk(n);
        std::vector<Point> corners(n);
        for (int i = 0; i < n; ++i) {
            const int type = is_u8 ? CV_8UC3 : CV_32FC3;
            src[i] = Mat(rows[i], cols[i], type, const_cast<void*>(images[i]), (size_t)cols[i] * 3 * (is_u8 ? 1 : 4));
            msk[i] = Mat(rows[i], cols[i], CV_8UC1, masks[i], (size_t)cols[i]);
            corners[i] = Point(corners_xy[2 * i], corners_xy[2 * i + 1]);
        }
        costFunc_ = cost_fn ? COLOR_GRAD : COLOR;
        find(src, corners, msk);
        return 0;
    } catch (const RefError& e) {
        return e.code;
    }
}
