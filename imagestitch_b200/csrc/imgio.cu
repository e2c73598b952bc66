// imgio.cu -- the on-disk format either side of the path: Windows bitmaps, decoded into and encoded out of HBM.
//
// Replaces cv::imread(".bmp") [BLEND]:31-34, [SEAM]:1098-1100 and cv::imwrite(".bmp", mat) [BLEND]:717, [SEAM]:1195-1206 (OpenCV
// imgcodecs, un-vendored: BmpDecoder / BmpEncoder of grfmt_bmp.cpp, and imwrite's convertTo(CV_8U) for the CV_32F / CV_16S mats
// the mains hand it).  The file goes to the device as it is; one kernel turns it into the cv::Mat layout the path works on
// (rows top-down, BGR interleaved, palette looked up, padding and alpha dropped) -- and the other way round for writing, with
// the saturating conversion to 8 bit fused into the packing.  Byte-for-byte pinned against cv2.imread / cv2.imwrite
// (tests/test_imgio.py), also on the bitmaps the reference itself checked in.
#include "internal.cuh"

#include <cstdio>
#include <memory>
#include <new>
#include <stdexcept>

namespace is {

// @emu-begin (tests/test_imgio.py compiles the marked region for the host)
struct BmpInfo {
    int width, height;        // height > 0 always; top_down tells the row order of the file
    int bpp;                  // 1, 4, 8, 24, 32
    int top_down;
    int data_offset;          // first pixel byte
    int palette_offset;       // BGRA entries (BGR triples for the 12-byte core header: palette_entry 3)
    int palette_entry;
    int palette_count;        // entries present in the file (biClrUsed); OpenCV leaves the others black
    int file_step;            // bytes per row in the file
};

static inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// BmpDecoder::readHeader for the uncompressed layouts (BI_RGB; 32 bit also as BI_BITFIELDS, which OpenCV reads as BGRA)
static int bmp_parse(const uint8_t* f, size_t size, BmpInfo* o) {
    if (size < 26 || f[0] != 'B' || f[1] != 'M') return IS_ERR_BAD_ARG;
    const uint32_t off = rd32(f + 10), hs = rd32(f + 14);
    int w, h, bpp;
    uint32_t compression = 0, clr_used = 0;
    o->palette_entry = 4;
    if (hs == 12) {
        w = (int)rd16(f + 18); h = (int)rd16(f + 20); bpp = (int)rd16(f + 24);
        o->palette_entry = 3;
    } else if (hs >= 36 && size >= 14 + (size_t)hs) {
        w = (int)rd32(f + 18); h = (int)rd32(f + 22); bpp = (int)rd16(f + 28);
        compression = rd32(f + 30);
        clr_used = rd32(f + 46);
        if (compression == 3 && bpp == 32) {                 // BI_BITFIELDS: only the layout that is BGRA anyway
            if (hs < 52 && size < 66) return IS_ERR_UNSUPPORTED;
            if (rd32(f + 54) != 0x00ff0000u || rd32(f + 58) != 0x0000ff00u || rd32(f + 62) != 0x000000ffu) return IS_ERR_UNSUPPORTED;
        }
    } else {
        return IS_ERR_UNSUPPORTED;
    }
    o->top_down = h < 0;
    if (h < 0) h = -h;
    if (w <= 0 || h <= 0) return IS_ERR_BAD_ARG;
    if (!(bpp == 1 || bpp == 4 || bpp == 8 || bpp == 24 || bpp == 32)) return IS_ERR_UNSUPPORTED;     // 16-bit 555 / 565: not here
    if (!(compression == 0 || (compression == 3 && bpp == 32))) return IS_ERR_UNSUPPORTED;             // RLE4 / RLE8: not here
    o->width = w; o->height = h; o->bpp = bpp;
    o->palette_offset = 14 + (int)hs;
    o->file_step = (int)((((long long)w * bpp + 7) / 8 + 3) & ~3LL);
    o->data_offset = (int)off;
    o->palette_count = 0;
    if (bpp <= 8) {
        if (clr_used > 256u) return IS_ERR_BAD_ARG;
        o->palette_count = clr_used == 0 ? (1 << bpp) : (int)clr_used;
        if ((size_t)o->palette_offset + (size_t)o->palette_entry * (size_t)o->palette_count > size) return IS_ERR_BAD_ARG;
    }
    if ((size_t)off + (size_t)o->file_step * (size_t)h > size) return IS_ERR_BAD_ARG;
    return IS_OK;
}

// BmpEncoder::write's header: 14 + 40 bytes, and the gray palette for one channel (FillGrayPalette)
static size_t bmp_header(int width, int height, int channels, uint8_t* out) {
    const int file_step = (width * channels + 3) & -4;
    const int header = 54, palette = channels > 1 ? 0 : 1024;
    const uint32_t file_size = (uint32_t)((size_t)file_step * height + header + palette);
    auto w32 = [&](int at, uint32_t v) { out[at] = (uint8_t)v; out[at + 1] = (uint8_t)(v >> 8); out[at + 2] = (uint8_t)(v >> 16); out[at + 3] = (uint8_t)(v >> 24); };
    std::memset(out, 0, (size_t)header + palette);
    out[0] = 'B'; out[1] = 'M';
    w32(2, file_size); w32(10, (uint32_t)(header + palette));
    w32(14, 40); w32(18, (uint32_t)width); w32(22, (uint32_t)height);
    out[26] = 1; out[28] = (uint8_t)(channels << 3);
    for (int i = 0; palette && i < 256; ++i) { out[54 + 4 * i] = out[55 + 4 * i] = out[56 + 4 * i] = (uint8_t)i; }
    return (size_t)header + palette;
}

// file -> cv::Mat rows (8UC3, what imread's default IMREAD_COLOR returns for every bitmap): one thread per pixel
__global__ void __launch_bounds__(256) k_bmp_unpack(const uint8_t* __restrict__ file, BmpInfo B, uint8_t* __restrict__ dst, size_t dstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= B.width || y >= B.height) return;
    const uint8_t* row = file + (size_t)B.data_offset + (size_t)(B.top_down ? y : B.height - 1 - y) * (size_t)B.file_step;
    uint8_t* d = dst + (size_t)y * dstep + 3 * (size_t)x;
    if (B.bpp >= 24) {
        const uint8_t* s = row + (size_t)x * (size_t)(B.bpp >> 3);
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        return;
    }
    int idx;
    if (B.bpp == 8) idx = row[x];
    else if (B.bpp == 4) idx = (row[x >> 1] >> ((x & 1) ? 0 : 4)) & 15;          // high nibble first (FillColorRow4)
    else idx = (row[x >> 3] >> (7 - (x & 7))) & 1;                               // most significant bit first (FillColorRow1)
    if (idx >= B.palette_count) { d[0] = 0; d[1] = 0; d[2] = 0; return; }
    const uint8_t* pal = file + (size_t)B.palette_offset + (size_t)idx * (size_t)B.palette_entry;
    d[0] = pal[0]; d[1] = pal[1]; d[2] = pal[2];
}

// saturate_cast<uchar>: the convertTo(CV_8U) imwrite applies to anything that is not 8 bit.  Floats round half to even; NaN and
// values beyond the int range become cvRound's INT_MIN on the reference's CPUs, i.e. 0 after the saturation.
__device__ __forceinline__ uint8_t sat_u8(uint8_t v) { return v; }
__device__ __forceinline__ uint8_t sat_u8(int16_t v) { return (uint8_t)max(0, min(255, (int)v)); }
__device__ __forceinline__ uint8_t sat_u8(float v) {
    if (!(v >= -2147483648.f && v < 2147483648.f)) return 0;
    return (uint8_t)max(0, min(255, __float2int_rn(v)));
}

// cv::Mat rows -> the file's pixel array: bottom-up, rows padded to four bytes with zeros; one thread per file byte
template <typename T>
__global__ void __launch_bounds__(256) k_bmp_pack(const T* __restrict__ src, size_t sstep, int rows, int cols, int channels, uint8_t* __restrict__ out, int file_step) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, fy = blockIdx.y;
    if (b >= file_step || fy >= rows) return;
    uint8_t v = 0;
    if (b < cols * channels) {
        const T* row = reinterpret_cast<const T*>(reinterpret_cast<const char*>(src) + (size_t)(rows - 1 - fy) * sstep);
        v = sat_u8(row[b]);
    }
    out[(size_t)fy * (size_t)file_step + b] = v;
}
// @emu-end

static int read_file(is_ctx* ctx, const char* path, std::vector<uint8_t>* out) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(ctx, IS_ERR_BAD_ARG, "cannot open %s", path);
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? std::fread(out->data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != out->size() || out->empty()) return fail(ctx, IS_ERR_BAD_ARG, "cannot read %s", path);
    return IS_OK;
}

static int parse_or_fail(is_ctx* ctx, const char* path, const std::vector<uint8_t>& file, BmpInfo* info) {
    const int rc = bmp_parse(file.data(), file.size(), info);
    if (rc == IS_ERR_UNSUPPORTED) return fail(ctx, rc, "%s: only uncompressed 1 / 4 / 8 / 24 / 32-bit bitmaps are decoded", path);
    if (rc != IS_OK) return fail(ctx, rc, "%s: not a bitmap or truncated", path);
    return IS_OK;
}

}  // namespace is

using namespace is;

namespace {

int bmp_info_entry(is_ctx* ctx, const char* path, is_size* size, int* bits_per_pixel) {
    if (!ctx || !path) return IS_ERR_BAD_ARG;
    // the header decides (14 + 40 bytes, the three masks of BI_BITFIELDS behind it); the length checks see the real file size
    uint8_t head[80] = {0};
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(ctx, IS_ERR_BAD_ARG, "cannot open %s", path);
    const size_t got = std::fread(head, 1, sizeof(head), f);
    std::fseek(f, 0, SEEK_END);
    const long total = std::ftell(f);
    std::fclose(f);
    if (got < 26 || total < (long)got) return fail(ctx, IS_ERR_BAD_ARG, "%s: not a bitmap or truncated", path);
    BmpInfo info;
    const int rc = bmp_parse(head, (size_t)total, &info);
    if (rc == IS_ERR_UNSUPPORTED) return fail(ctx, rc, "%s: only uncompressed 1 / 4 / 8 / 24 / 32-bit bitmaps are decoded", path);
    if (rc != IS_OK) return fail(ctx, rc, "%s: not a bitmap or truncated", path);
    if (size) { size->width = info.width; size->height = info.height; }
    if (bits_per_pixel) *bits_per_pixel = info.bpp;
    return IS_OK;
}

int imread_bmp_entry(is_ctx* ctx, const char* path, is_mat* dst) {
    if (!ctx || !path) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, dst, "dst"));
    std::vector<uint8_t> file;
    IS_TRY(read_file(ctx, path, &file));
    BmpInfo info;
    IS_TRY(parse_or_fail(ctx, path, file, &info));
    IS_REQUIRE(ctx, dst->depth == IS_8U && dst->channels == 3 && dst->rows == info.height && dst->cols == info.width, IS_ERR_BAD_ARG,
               "dst must be 8UC3 of the size is_bmp_info reports (imread's IMREAD_COLOR)");
    IS_REQUIRE(ctx, info.height <= 65535, IS_ERR_UNSUPPORTED, "imread: at most 65535 rows (one grid row per image row)");
    DevBuf raw;
    IS_TRY(raw.alloc(ctx, file.size() + 16));
    IS_CUDA(ctx, cudaMemcpyAsync(raw.p, file.data(), file.size(), cudaMemcpyHostToDevice, ctx->stream));
    DevMat d;
    IS_TRY(stage_out(ctx, dst, &d, false));
    dim3 grid(div_up(info.width, 256), info.height);
    ctx->next_bytes = (double)info.file_step * info.height + 3. * info.width * info.height;
    IS_LAUNCH(ctx, k_bmp_unpack, grid, 256, 0, raw.as<uint8_t>(), info, d.ptr<uint8_t>(), d.step);
    IS_TRY(commit(ctx, &d));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));        // `file` (pageable) and `raw` leave scope
    return IS_OK;
}

int imwrite_bmp_entry(is_ctx* ctx, const char* path, const is_mat* src) {
    if (!ctx || !path) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, src, "src"));
    IS_REQUIRE(ctx, (src->channels == 1 || src->channels == 3) && (src->depth == IS_8U || src->depth == IS_16S || src->depth == IS_32F), IS_ERR_UNSUPPORTED,
               "imwrite: 1 or 3 channels of IS_8U / IS_16S / IS_32F");
    IS_REQUIRE(ctx, (long long)src->cols * src->channels < (1LL << 30) && src->rows <= 65535, IS_ERR_UNSUPPORTED, "imwrite: at most 65535 rows (one grid row per image row)");
    const int file_step = (src->cols * src->channels + 3) & -4;
    const size_t pixels = (size_t)file_step * (size_t)src->rows;
    IS_REQUIRE(ctx, pixels + 1078 < ((size_t)1 << 32), IS_ERR_UNSUPPORTED, "imwrite: a bitmap holds less than 4 GB");
    DevMat s;
    IS_TRY(stage_in(ctx, src, &s));
    DevBuf packed;
    IS_TRY(packed.alloc(ctx, pixels));
    dim3 grid(div_up(file_step, 256), src->rows);
    ctx->next_bytes = (double)s.row_bytes() * s.rows + (double)pixels;
    if (src->depth == IS_8U) IS_LAUNCH(ctx, k_bmp_pack<uint8_t>, grid, 256, 0, s.ptr<uint8_t>(), s.step, s.rows, s.cols, s.channels, packed.as<uint8_t>(), file_step);
    else if (src->depth == IS_16S) IS_LAUNCH(ctx, k_bmp_pack<int16_t>, grid, 256, 0, s.ptr<int16_t>(), s.step, s.rows, s.cols, s.channels, packed.as<uint8_t>(), file_step);
    else IS_LAUNCH(ctx, k_bmp_pack<float>, grid, 256, 0, s.ptr<float>(), s.step, s.rows, s.cols, s.channels, packed.as<uint8_t>(), file_step);
    uint8_t head[54 + 1024];
    const size_t head_bytes = bmp_header(src->cols, src->rows, src->channels, head);
    std::unique_ptr<uint8_t[]> host(new (std::nothrow) uint8_t[pixels]);
    IS_REQUIRE(ctx, host != nullptr, IS_ERR_NO_MEM, "imwrite: host buffer");
    IS_CUDA(ctx, cudaMemcpyAsync(host.get(), packed.p, pixels, cudaMemcpyDeviceToHost, ctx->stream));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(ctx, IS_ERR_BAD_ARG, "cannot create %s", path);
    const bool ok = std::fwrite(head, 1, head_bytes, f) == head_bytes && std::fwrite(host.get(), 1, pixels, f) == pixels;
    if (std::fclose(f) != 0 || !ok) return fail(ctx, IS_ERR_INTERNAL, "short write to %s", path);
    return IS_OK;
}

template <typename F>
int guarded(is_ctx* ctx, F&& f) {                             // no exception crosses the boundary (the file is read into a vector)
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return ctx ? fail(ctx, IS_ERR_NO_MEM, "image file: host memory") : IS_ERR_NO_MEM;
    } catch (const std::exception& e) {
        return ctx ? fail(ctx, IS_ERR_INTERNAL, "image file: %s", e.what()) : IS_ERR_INTERNAL;
    }
}

}  // namespace

extern "C" {

int is_bmp_info(is_ctx* ctx, const char* path, is_size* size, int* bits_per_pixel) {
    return guarded(ctx, [&] { return bmp_info_entry(ctx, path, size, bits_per_pixel); });
}
int is_imread_bmp(is_ctx* ctx, const char* path, is_mat* dst) {
    return guarded(ctx, [&] { return imread_bmp_entry(ctx, path, dst); });
}
int is_imwrite_bmp(is_ctx* ctx, const char* path, const is_mat* src) {
    return guarded(ctx, [&] { return imwrite_bmp_entry(ctx, path, src); });
}

}  // extern "C"
