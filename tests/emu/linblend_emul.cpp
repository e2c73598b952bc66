// tests/emu/linblend_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked region of imagestitch_b200/csrc/linblend.cu (geometry + the six kernels of the pair blend), launched in
// the order and with the shapes is_linear_blend_pair uses.  k_lin_seam (shared memory + __syncthreads) and k_lin_rowscan (warp
// shuffles) run on the multi-threaded block emulator, the per-pixel kernels on the sequential one.
#include "cuda_host_emul_mt.h"

#include <vector>

#include "../../include/imagestitch.h"

namespace is {
#include "linblend_region.inc"
}
using namespace is;

static inline unsigned div_up(int a, int b) { return (unsigned)((a + b - 1) / b); }

// returns 1 for the early "no overlap" return, 0 otherwise; pano: panoHe x panoBr x 3 floats, seam_x: panoHe ints
extern "C" int emu_linear_blend_pair(const float* img1, int rows1, int cols1, const float* img2, int rows2, int cols2, int tl1x, int tl1y, int tl2x,
                                     int tl2y, float* pano, int* seam_x) {
    is_point tl1{tl1x, tl1y}, tl2{tl2x, tl2y};
    LinGeo g = lin_geometry(rows1, cols1, rows2, cols2, tl1, tl2);
    if (!g.overlap) return 1;
    FImg fa{img1, sizeof(float) * 3 * (size_t)cols1}, fb{img2, sizeof(float) * 3 * (size_t)cols2};
    const int CW = g.IB + 2, MW = g.width + 2;
    std::vector<float> costV((size_t)g.panoHe * CW), m1((size_t)g.height * MW), m2((size_t)g.height * MW);
    std::vector<int> seam(g.panoHe), left(g.height), right(g.height);
    dim3 block(32, 8);
    emu_launch(dim3(div_up(CW, 32), div_up(g.panoHe, 8)), block, [&] { k_lin_cost(fa, fb, g, costV.data()); });
    emu_launch_mt(1, 256, [&] { k_lin_seam(costV.data(), g.panoHe, CW, g.IB / 2, seam.data()); });
    emu_launch(dim3(div_up(MW, 32), div_up(g.height, 8)), block, [&] { k_lin_classify(fa, fb, g, m1.data(), m2.data()); });
    emu_launch_mt(div_up(g.height, 8), 256, [&] { k_lin_rowscan(m2.data(), g.height, g.width, left.data(), right.data()); });
    emu_launch(dim3(div_up(g.width + 1, 32), div_up(g.height, 8)), block, [&] { k_lin_weights(m1.data(), m2.data(), g, left.data(), right.data(), seam.data()); });
    emu_launch(dim3(div_up(g.panoBr, 32), div_up(g.panoHe, 8)), block,
               [&] { k_lin_composite(fa, fb, g, m1.data(), m2.data(), pano, sizeof(float) * 3 * (size_t)g.panoBr); });
    std::memcpy(seam_x, seam.data(), sizeof(int) * (size_t)g.panoHe);
    return 0;
}
