// tests/emu/feather_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked regions of imagestitch_b200/csrc/feather.cu (dilate, L1 distance transform -> weight map, the gather
// blend), launched with the shapes of mask_dilate_and_device / feather_weight_device / feather_blend_device.
#include "cuda_host_emul_mt.h"

#include <cfloat>
#include <climits>
#include <vector>

namespace is {
#include "feather_regions.inc"
}
using namespace is;

static inline unsigned div_up(int a, int b) { return (unsigned)((a + b - 1) / b); }

extern "C" void emu_mask_dilate_and(uint8_t* mask, int rows, int cols, int kw, int kh, const uint8_t* andm) {
    std::vector<uint8_t> tmp((size_t)rows * cols);
    for (int y = 0; y < rows; ++y)                                           // grid (div_up(cols, DIL_T), rows): one row of blocks at a time
        for (unsigned bx = 0; bx < div_up(cols, DIL_T); ++bx) {
            emu_launch_mt(1, DIL_T, [&] {
                blockIdx = EmuDim3(bx, (unsigned)y);                          // emu_launch_mt numbers blocks along x only
                k_dilate_rows(mask, (size_t)cols, tmp.data(), (size_t)cols, rows, cols, kw);
            });
        }
    emu_launch(dim3(div_up(cols, 64), div_up(rows, 4)), dim3(64, 4),
               [&] { k_dilate_cols_and(tmp.data(), (size_t)cols, mask, (size_t)cols, andm, (size_t)cols, rows, cols, kh); });
}

static void weight_map(const uint8_t* mask, int rows, int cols, float sharpness, float* w) {
    std::vector<int> dh((size_t)rows * cols);
    emu_launch_mt(div_up(rows, 8), 256, [&] { k_dt_rows(mask, (size_t)cols, rows, cols, dh.data()); });
    emu_launch(dim3(div_up(cols, 128)), dim3(128), [&] { k_dt_cols_down(dh.data(), rows, cols); });
    emu_launch(dim3(div_up(cols, 128)), dim3(128), [&] { k_dt_cols_up_weight(dh.data(), rows, cols, sharpness, w, (size_t)cols); });
}

extern "C" void emu_feather_weight(const uint8_t* mask, int rows, int cols, float sharpness, float* w) { weight_map(mask, rows, cols, sharpness, w); }

// imgs: n tightly packed images (u8 x 3 or s16 x 3), masks u8; corners relative to the ROI origin; dst: H x W x 3 s16, dmask: H x W u8
extern "C" void emu_feather_blend(int n, const void* const* imgs, int is_u8, const uint8_t* const* masks, const int* rows, const int* cols, const int* x0,
                                  const int* y0, float sharpness, int W, int H, int16_t* dst, uint8_t* dmask) {
    std::vector<std::vector<float>> w((size_t)n);
    std::vector<FeatherImg> table((size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i) {
        w[i].resize((size_t)rows[i] * cols[i]);
        weight_map(masks[i], rows[i], cols[i], sharpness, w[i].data());
        table[i] = FeatherImg{imgs[i], (size_t)cols[i] * 3 * (is_u8 ? 1 : 2), is_u8, w[i].data(), x0[i], y0[i], rows[i], cols[i]};
    }
    emu_launch(dim3(div_up(W, 64), div_up(H, 4)), dim3(64, 4),
               [&] { k_feather_blend(table.data(), n, W, H, dst, sizeof(int16_t) * 3 * (size_t)W, dmask, (size_t)W); });
}
