// oracle/ref_shim/linblend_ref.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own hand-written pair blend: the statements of its main() from [BLEND]:141 to :717 are
// included below from a file that oracle/Makefile extracts from /root/reference at build time into oracle/_ref/
// (git-ignored; transcoded GBK -> UTF-8 so that no multi-byte character ends in a backslash).  Used by
// tests/test_oracle_reference_build.py to pin oracle/linblend.cpp; absent when /root/reference is absent.
#include "cvshim.h"

#include <cstdint>
#include <cstring>

using namespace cv;
using namespace std;
using namespace detail;

// returns 0 = blended, 1 = the block's own early "no conflicts" return ([BLEND]:182-183)
static int run_block(vector<Mat>& images_warped, vector<UMat>& images_warped_f, vector<Point>& corners, Mat& pano_out, vector<Point>& seam_out,
                     Mat& cost_out) {
#include "blend_block.inc"
    pano.copyTo(pano_out);
    seam_out = seam;
    costV.copyTo(cost_out);
    return 1000;
}

extern "C" int ref_lin_blend(const float* img1, int rows1, int cols1, const float* img2, int rows2, int cols2, int tl1x, int tl1y, int tl2x, int tl2y,
                             float* pano, int pano_rows, int pano_cols, int32_t* seam_x, float* costV, int cost_cols) {
    vector<Mat> images_warped(2);
    vector<UMat> images_warped_f(2);
    const float* src[2] = {img1, img2};
    const int rows[2] = {rows1, rows2}, cols[2] = {cols1, cols2};
    for (int i = 0; i < 2; ++i) {
        images_warped_f[i].create(rows[i], cols[i], CV_32FC3);
        std::memcpy(images_warped_f[i].data(), src[i], sizeof(float) * (size_t)rows[i] * cols[i] * 3);
        images_warped[i].create(rows[i], cols[i], CV_32FC3);           // only .rows / .cols are read from it
    }
    vector<Point> corners = {Point(tl1x, tl1y), Point(tl2x, tl2y)};
    Mat p, c;
    vector<Point> seam;
    const int rc = run_block(images_warped, images_warped_f, corners, p, seam, c);
    if (rc != 1000) return 1;
    if (p.rows != pano_rows || p.cols != pano_cols || (int)seam.size() != pano_rows || c.cols != cost_cols) return -1;
    std::memcpy(pano, p.data(), sizeof(float) * (size_t)p.rows * p.cols * 3);
    for (int y = 0; y < pano_rows; ++y) seam_x[y] = seam[y].x;
    if (costV) std::memcpy(costV, c.data(), sizeof(float) * (size_t)c.rows * c.cols);
    return 0;
}
