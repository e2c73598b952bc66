// oracle/ref_shim/warp_ref.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own cylindrical projector arithmetic: its globals, mapForward, mapBackward and detectResultRoi
// ([WARP]:30-89) are included below from a file that oracle/Makefile extracts from /root/reference at build time
// (deleted again after compiling).  The per-pixel loop of buildMaps ([WARP]:133-143) is repeated here around the
// reference's mapBackward because the original writes through cv::OutputArray.  The camera products k_rinv / r_kinv
// come from the caller: setCameraParams ([WARP]:90-120) forms them with OpenCV matrix operators.
#include "cvshim.h"

#include <cmath>
#include <limits>

using namespace cv;
using namespace std;

#include "warp_block.inc"

extern "C" void ref_warp_set(const float* k_rinv_in, const float* r_kinv_in, float scale_in) {
    for (int i = 0; i < 9; ++i) { k_rinv[i] = k_rinv_in[i]; r_kinv[i] = r_kinv_in[i]; }
    scale = scale_in;
}

extern "C" void ref_detect_roi(int width, int height, int tlbr[4]) {
    Point tl, br;
    detectResultRoi(Size(width, height), tl, br);
    tlbr[0] = tl.x; tlbr[1] = tl.y; tlbr[2] = br.x; tlbr[3] = br.y;
}

extern "C" void ref_build_maps(int tlx, int tly, int brx, int bry, float* xmap, float* ymap) {
    const int w = brx - tlx + 1;
    float x, y;
    for (int v = tly; v <= bry; ++v)
        for (int u = tlx; u <= brx; ++u) {
            mapBackward(static_cast<float>(u), static_cast<float>(v), x, y);
            xmap[(size_t)(v - tly) * w + (u - tlx)] = x;
            ymap[(size_t)(v - tly) * w + (u - tlx)] = y;
        }
}
