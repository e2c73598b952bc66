// tests/emu/imgio_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked region of imagestitch_b200/csrc/imgio.cu (bitmap header parse / write, k_bmp_unpack, k_bmp_pack), driven
// the way is_imread_bmp / is_imwrite_bmp drive it.
#include "cuda_host_emul.h"

#include <vector>

#include "../../include/imagestitch.h"

namespace is {
#include "imgio_region.inc"
}
using namespace is;

static inline unsigned div_up(int a, int b) { return (unsigned)((a + b - 1) / b); }

extern "C" int emu_bmp_info(const uint8_t* file, size_t size, int* rows, int* cols, int* bpp) {
    BmpInfo info;
    const int rc = bmp_parse(file, size, &info);
    if (rc != IS_OK) return rc;
    *rows = info.height; *cols = info.width; *bpp = info.bpp;
    return IS_OK;
}

extern "C" int emu_bmp_read(const uint8_t* file, size_t size, uint8_t* dst, size_t dstep) {
    BmpInfo info;
    const int rc = bmp_parse(file, size, &info);
    if (rc != IS_OK) return rc;
    emu_launch(dim3(div_up(info.width, 256), (unsigned)info.height), dim3(256), [&] { k_bmp_unpack(file, info, dst, dstep); });
    return IS_OK;
}

// out: 54 (+ 1024) + file_step * rows bytes; returns the file size
extern "C" size_t emu_bmp_write(const void* src, int depth, int rows, int cols, int channels, size_t sstep, uint8_t* out) {
    const size_t head = bmp_header(cols, rows, channels, out);
    const int file_step = (cols * channels + 3) & -4;
    dim3 grid(div_up(file_step, 256), (unsigned)rows);
    if (depth == IS_8U) emu_launch(grid, dim3(256), [&] { k_bmp_pack<uint8_t>((const uint8_t*)src, sstep, rows, cols, channels, out + head, file_step); });
    else if (depth == IS_16S) emu_launch(grid, dim3(256), [&] { k_bmp_pack<int16_t>((const int16_t*)src, sstep, rows, cols, channels, out + head, file_step); });
    else emu_launch(grid, dim3(256), [&] { k_bmp_pack<float>((const float*)src, sstep, rows, cols, channels, out + head, file_step); });
    return head + (size_t)file_step * rows;
}
