// tests/emu/seam_cost_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked (@emu-begin / @emu-end) regions of imagestitch_b200/csrc/seam.cu: MaskView, Frame, ImgView, the
// COLOR / COLOR_GRAD cost functions, k_sobel_window, k_cost_maps and k_cost_pq.  tests/test_kernel_host_emulation.py extracts
// the regions into seam_regions.inc next to the build output and drives the entry points below.
#include "cuda_host_emul.h"

namespace is {
#include "seam_regions.inc"
}
using namespace is;

static inline unsigned div_up(int a, int b) { return (unsigned)((a + b - 1) / b); }

// gradients of one image over the window (ox, oy, ww, wh) of the union frame; image coords = union coords + (dx, dy)
extern "C" void emu_sobel_window(const void* img, int is_u8, int rows, int cols, int dx, int dy, int ox, int oy, int ww, int wh, float* gx, float* gy,
                                 int pitch) {
    dim3 block(64, 4), grid(div_up(ww, 64), div_up(wh, 4));
    if (is_u8) {
        ImgView<uint8_t> v{(const uint8_t*)img, (size_t)cols * 3, rows, cols, dx, dy};
        emu_launch(grid, block, [&] { k_sobel_window<uint8_t>(v, ox, oy, ww, wh, gx, gy, pitch); });
    } else {
        ImgView<float> v{(const float*)img, (size_t)cols * 12, rows, cols, dx, dy};
        emu_launch(grid, block, [&] { k_sobel_window<float>(v, ox, oy, ww, wh, gx, gy, pitch); });
    }
}

// k_cost_pq over the component bounding box (rx, ry, rw, rh) of label l; labels: full union frame H x W (window = frame).
// grad = 4 planes (gx1, gy1, gx2, gy2) of gpitch x wh floats over the window at (gox, goy), or NULL for COLOR.
extern "C" void emu_cost_pq(const void* img1, const void* img2, int is_u8, int rows1, int cols1, int rows2, int cols2, int dx1, int dy1, int dx2,
                            int dy2, const int* labels, int H, int W, int l, int rx, int ry, int rw, int rh, int horizontal, const float* grad,
                            int gpitch, int gox, int goy, int gwh, float* P, float* Q, int pitch) {
    Frame f{W, H, 0, 0, W, H, MaskView{nullptr, 0, 0, 0, 0, 0}, MaskView{nullptr, 0, 0, 0, 0, 0}};
    const size_t plane = (size_t)gpitch * gwh;
    GradView g{grad, grad ? grad + plane : nullptr, grad ? grad + 2 * plane : nullptr, grad ? grad + 3 * plane : nullptr, gpitch, gox, goy};
    const int steps = horizontal ? rw : rh;
    dim3 block(64, 4), grid(div_up(pitch, 64), div_up(steps, 4));
    const size_t es = is_u8 ? 1 : 4;
    if (is_u8) {
        ImgView<uint8_t> a{(const uint8_t*)img1, (size_t)cols1 * 3 * es, rows1, cols1, dx1, dy1}, b{(const uint8_t*)img2, (size_t)cols2 * 3 * es, rows2, cols2, dx2, dy2};
        if (grad) emu_launch(grid, block, [&] { k_cost_pq<uint8_t, true>(a, b, labels, f, l, rx, ry, rw, rh, horizontal, P, Q, pitch, g); });
        else emu_launch(grid, block, [&] { k_cost_pq<uint8_t, false>(a, b, labels, f, l, rx, ry, rw, rh, horizontal, P, Q, pitch, g); });
    } else {
        ImgView<float> a{(const float*)img1, (size_t)cols1 * 3 * es, rows1, cols1, dx1, dy1}, b{(const float*)img2, (size_t)cols2 * 3 * es, rows2, cols2, dx2, dy2};
        if (grad) emu_launch(grid, block, [&] { k_cost_pq<float, true>(a, b, labels, f, l, rx, ry, rw, rh, horizontal, P, Q, pitch, g); });
        else emu_launch(grid, block, [&] { k_cost_pq<float, false>(a, b, labels, f, l, rx, ry, rw, rh, horizontal, P, Q, pitch, g); });
    }
}
