// feather.cu -- the blend path the reference's mains execute: mask preparation + feather blend.
//
//   Mat element = getStructuringElement(MORPH_RECT, Size(20, 20));
//   dilate(masks_seam[k], masks_seam[k], element);  masks_seam[k] &= masks_warped[k];                    [SEAM]:1257-1270
//   blender = Blender::createDefault(Blender::FEATHER); fb->setSharpness(0.1);                            [SEAM]:1249-1251
//   blender->prepare(corners, sizes); blender->feed(img_s, mask, corner); blender->blend(result, mask);   [SEAM]:1252,1271,1280
// (cv::dilate, cv::distanceTransform and cv::detail::FeatherBlender of un-vendored OpenCV 3.4.2; restated in oracle/feather.cpp.)
//
// B200 formulation.
//   dilate        separable maximum (rows through shared memory, columns with coalesced row reads), the `&` fused in.
//   weight map    min(distanceTransform(mask, DIST_L1, 3) * sharpness, 1).  The 3x3 L1 chamfer is the exact city-block
//                 distance, which is separable: distance along the row (one warp per row, last-zero position by warp
//                 prefix maximum in both directions), then the min-plus pass with |dy| down and up every column (one
//                 thread per column, coalesced rows, loads independent of the running minimum).
//   feed / blend  OpenCV read-modify-writes the panorama-sized accumulators once per image; here blend() is one gather
//                 kernel: per panorama pixel the images covering it in feed order, dst += short(src * w) (float product,
//                 truncation, int16 wrap), wsum += w (float, feed order), then short(dst / (wsum + 1e-5f)) and the mask.
#include "internal.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>

namespace is {

// @emu-begin (tests/test_kernel_host_emulation.py compiles the marked regions for the host)
constexpr int DT_INF = INT_MAX / 4;

// ---- dilate (rectangular element kw x kh, anchor (kw/2, kh/2)): dst(x,y) = max src over x - kw/2 .. x - kw/2 + kw - 1 ----------
constexpr int DIL_T = 256, DIL_MAXK = 64;

__global__ void __launch_bounds__(DIL_T) k_dilate_rows(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst, size_t dstep, int rows, int cols,
                                                       int kw) {
    __shared__ uint8_t s[DIL_T + DIL_MAXK];
    const int y = blockIdx.y, x0 = blockIdx.x * DIL_T, ax = kw / 2;
    for (int e = threadIdx.x; e < DIL_T + kw - 1; e += DIL_T) {
        const int x = x0 - ax + e;
        s[e] = (x >= 0 && x < cols) ? src[(size_t)y * sstep + x] : 0;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= cols) return;
    int m = 0;
    for (int k = 0; k < kw; ++k) m = max(m, (int)s[threadIdx.x + k]);
    dst[(size_t)y * dstep + x] = (uint8_t)m;
}

__global__ void k_dilate_cols_and(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst, size_t dstep, const uint8_t* __restrict__ andm,
                                  size_t astep, int rows, int cols, int kh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const int ay = kh / 2;
    int m = 0;
    for (int k = max(0, y - ay); k <= min(rows - 1, y - ay + kh - 1); ++k) m = max(m, (int)src[(size_t)k * sstep + x]);
    if (andm) m &= andm[(size_t)y * astep + x];
    dst[(size_t)y * dstep + x] = (uint8_t)m;
}

// ---- distance along the row to the nearest zero pixel (DT_INF when the row has none) ---------------------------------------
__global__ void k_dt_rows(const uint8_t* __restrict__ mask, size_t mstep, int rows, int cols, int* __restrict__ dh) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* m = mask + (size_t)y * mstep;
    int* d = dh + (size_t)y * cols;
    // left to right: position of the last zero at or before x
    int last = -DT_INF;
    for (int base = 0; base < cols; base += 32) {
        const int x = base + lane;
        int z = (x < cols && m[x] == 0) ? x : -DT_INF;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z = max(z, t); }
        z = max(z, last);
        if (x < cols) d[x] = z <= -DT_INF ? DT_INF : x - z;
        last = __shfl_sync(0xffffffffu, z, 31);
    }
    // right to left: position of the first zero at or after x
    int next = DT_INF;
    for (int base = ((cols - 1) / 32) * 32; base >= 0; base -= 32) {
        const int x = base + lane;
        int z = (x < cols && m[x] == 0) ? x : DT_INF;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, z, o); if (lane + o < 32) z = min(z, t); }
        z = min(z, next);
        if (x < cols) { const int r = z >= DT_INF ? DT_INF : z - x; d[x] = min(d[x], r); }
        next = __shfl_sync(0xffffffffu, z, 0);
    }
}

// ---- min-plus with |dy| along the columns, then the weight min(d * sharpness, 1) -------------------------------------------
__global__ void k_dt_cols_down(int* __restrict__ dh, int rows, int cols) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cols) return;
    int g = DT_INF;
#pragma unroll 8
    for (int y = 0; y < rows; ++y) {
        const int v = dh[(size_t)y * cols + x];
        g = min(g >= DT_INF ? DT_INF : g + 1, v);
        dh[(size_t)y * cols + x] = g;
    }
}

__global__ void k_dt_cols_up_weight(const int* __restrict__ dh, int rows, int cols, float sharpness, float* __restrict__ w, size_t wstep_f) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cols) return;
    int g = DT_INF;
#pragma unroll 8
    for (int y = rows - 1; y >= 0; --y) {
        const int v = dh[(size_t)y * cols + x];
        g = min(g >= DT_INF ? DT_INF : g + 1, v);
        const float d = g >= DT_INF ? FLT_MAX : (float)g;
        w[(size_t)y * wstep_f + x] = fminf(__fmul_rn(d, sharpness), 1.f);   // multiply(weight, sharpness); threshold(TRUNC, 1)
    }
}

// @emu-end
int feather_weight_device(is_ctx* ctx, const DevMat& mask, float sharpness, float* w, size_t wstep_f) {
    DevBuf dh;
    IS_TRY(dh.alloc(ctx, sizeof(int) * (size_t)mask.rows * mask.cols));
    ctx->next_bytes = (double)mask.rows * mask.cols * 5;
    IS_LAUNCH(ctx, k_dt_rows, div_up(mask.rows, 8), 256, 0, mask.ptr<uint8_t>(), mask.step, mask.rows, mask.cols, dh.as<int>());
    ctx->next_bytes = (double)mask.rows * mask.cols * 8;
    IS_LAUNCH(ctx, k_dt_cols_down, div_up(mask.cols, 128), 128, 0, dh.as<int>(), mask.rows, mask.cols);
    ctx->next_bytes = (double)mask.rows * mask.cols * 8;
    IS_LAUNCH(ctx, k_dt_cols_up_weight, div_up(mask.cols, 128), 128, 0, dh.as<int>(), mask.rows, mask.cols, sharpness, w, wstep_f);
    return IS_OK;
}

// ---- blend(): gather over the fed images -------------------------------------------------------------------------------
// @emu-begin
struct FeatherImg {
    const void* img; size_t istep; int is_u8;
    const float* w;          // dense rows x cols
    int x0, y0, rows, cols;  // position inside the destination ROI
};

__global__ void k_feather_blend(const FeatherImg* __restrict__ imgs, int n, int W, int H, int16_t* __restrict__ dst, size_t dstep, uint8_t* __restrict__ dmask,
                                size_t mstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    int acc[3] = {0, 0, 0};
    float wsum = 0.f;
    for (int i = 0; i < n; ++i) {
        const FeatherImg& I = imgs[i];
        const int lx = x - I.x0, ly = y - I.y0;
        if ((unsigned)lx >= (unsigned)I.cols || (unsigned)ly >= (unsigned)I.rows) continue;
        const float w = I.w[(size_t)ly * I.cols + lx];
        int v[3];
        if (I.is_u8) {
            const uint8_t* p = reinterpret_cast<const uint8_t*>(I.img) + (size_t)ly * I.istep + 3 * lx;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
        } else {
            const int16_t* p = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(I.img) + (size_t)ly * I.istep) + 3 * lx;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
        }
        // dst += static_cast<short>(src * w): truncation toward zero, int16 wrap-around add; the weight sum in feed order
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = (int)(int16_t)(acc[c] + (int)(int16_t)__float2int_rz(__fmul_rn((float)v[c], w)));
        wsum = __fadd_rn(wsum, w);
    }
    const bool on = wsum > 1e-5f;
    int16_t* o = reinterpret_cast<int16_t*>(reinterpret_cast<char*>(dst) + (size_t)y * dstep) + 3 * x;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = on ? (int16_t)__float2int_rz(__fdiv_rn((float)acc[c], __fadd_rn(wsum, 1e-5f))) : (int16_t)0;
    dmask[(size_t)y * mstep + x] = on ? 255 : 0;
}

// @emu-end
// dilate (kw x kh rectangle, anchor at the centre) in place, optionally followed by `& andm`
int mask_dilate_and_device(is_ctx* ctx, const DevMat& m, int kw, int kh, const DevMat* andm) {
    IS_REQUIRE(ctx, kw >= 1 && kh >= 1 && kw <= DIL_MAXK && kh <= DIL_MAXK, IS_ERR_BAD_ARG, "structuring element must be 1..64 wide / high");
    DevMat tmp;
    IS_TRY(alloc_mat(ctx, m.rows, m.cols, 1, IS_8U, &tmp));
    ctx->next_bytes = 2.0 * m.rows * m.cols;
    IS_LAUNCH(ctx, k_dilate_rows, dim3(div_up(m.cols, DIL_T), m.rows), DIL_T, 0, m.ptr<uint8_t>(), m.step, tmp.ptr<uint8_t>(), tmp.step, m.rows, m.cols, kw);
    ctx->next_bytes = (andm ? 3.0 : 2.0) * m.rows * m.cols;
    dim3 block(64, 4), grid(div_up(m.cols, 64), div_up(m.rows, 4));
    IS_LAUNCH(ctx, k_dilate_cols_and, grid, block, 0, tmp.ptr<uint8_t>(), tmp.step, m.ptr<uint8_t>(), m.step, andm ? andm->ptr<uint8_t>() : nullptr,
              andm ? andm->step : 0, m.rows, m.cols, kh);
    return IS_OK;
}

// FeatherBlender prepare / feed x n / blend on device-resident images and masks (used in place)
int feather_blend_device(is_ctx* ctx, float sharpness, is_rect roi, int n, const DevMat* imgs, const DevMat* masks, const is_point* corners, const DevMat& d,
                         const DevMat& m) {
    std::vector<DevBuf> w((size_t)n);
    std::vector<FeatherImg> host((size_t)std::max(n, 1));
    double bytes = (double)roi.width * roi.height * 7;
    for (int i = 0; i < n; ++i) {
        IS_REQUIRE(ctx, corners[i].x >= roi.x && corners[i].y >= roi.y && corners[i].x + imgs[i].cols <= roi.x + roi.width &&
                            corners[i].y + imgs[i].rows <= roi.y + roi.height, IS_ERR_BAD_ARG, "image lies outside the prepared ROI");
        IS_TRY(w[i].alloc(ctx, sizeof(float) * (size_t)imgs[i].rows * imgs[i].cols));
        IS_TRY(feather_weight_device(ctx, masks[i], sharpness, w[i].as<float>(), (size_t)imgs[i].cols));
        host[i] = FeatherImg{imgs[i].data, imgs[i].step, imgs[i].depth == IS_8U ? 1 : 0, w[i].as<float>(), corners[i].x - roi.x, corners[i].y - roi.y,
                             imgs[i].rows, imgs[i].cols};
        bytes += (double)imgs[i].rows * imgs[i].cols * ((imgs[i].depth == IS_8U ? 3 : 6) + 4);
    }
    DevBuf table;
    IS_TRY(table.alloc(ctx, sizeof(FeatherImg) * host.size()));
    IS_TRY(upload(ctx, table.p, host.data(), sizeof(FeatherImg) * host.size()));
    dim3 block(64, 4), grid(div_up(roi.width, 64), div_up(roi.height, 4));
    ctx->next_bytes = bytes;
    IS_LAUNCH(ctx, k_feather_blend, grid, block, 0, table.as<FeatherImg>(), n, roi.width, roi.height, d.ptr<int16_t>(), d.step, m.ptr<uint8_t>(), m.step);
    return IS_OK;
}

}  // namespace is

using namespace is;

struct FeatherFed {
    DevMat img;
    DevBuf w;
    int tl_x, tl_y;
};

struct is_feather_blender {
    is_ctx* ctx = nullptr;
    float sharpness = 0.02f;
    bool prepared = false;
    is_rect roi{};
    std::vector<FeatherFed> fed;
};

extern "C" {

int is_mask_dilate_and(is_ctx* ctx, is_mat* mask, int kw, int kh, const is_mat* and_mask) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, mask, "mask"));
    IS_REQUIRE(ctx, mask->depth == IS_8U && mask->channels == 1, IS_ERR_ASSERT, "mask.type() == CV_8U");
    IS_REQUIRE(ctx, kw >= 1 && kh >= 1 && kw <= DIL_MAXK && kh <= DIL_MAXK, IS_ERR_BAD_ARG, "structuring element must be 1..64 wide / high");
    DevMat m, a;
    IS_TRY(stage_out(ctx, mask, &m, true));
    if (and_mask) {
        IS_TRY(check_mat(ctx, and_mask, "and_mask"));
        IS_REQUIRE(ctx, and_mask->depth == IS_8U && and_mask->channels == 1 && and_mask->rows == mask->rows && and_mask->cols == mask->cols, IS_ERR_ASSERT,
                   "and_mask must be CV_8U of the mask's size");
        IS_TRY(stage_in(ctx, and_mask, &a));
    }
    IS_TRY(mask_dilate_and_device(ctx, m, kw, kh, and_mask ? &a : nullptr));
    IS_TRY(commit(ctx, &m));
    return IS_OK;
}

int is_feather_weight_map(is_ctx* ctx, const is_mat* mask, float sharpness, is_mat* weight) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, mask, "mask"));
    IS_TRY(check_mat(ctx, weight, "weight"));
    IS_REQUIRE(ctx, mask->depth == IS_8U && mask->channels == 1, IS_ERR_ASSERT, "mask.type() == CV_8U");
    IS_REQUIRE(ctx, weight->depth == IS_32F && weight->channels == 1 && weight->rows == mask->rows && weight->cols == mask->cols, IS_ERR_BAD_ARG,
               "weight must be CV_32F of the mask's size");
    DevMat m, w;
    IS_TRY(stage_in(ctx, mask, &m));
    IS_TRY(stage_out(ctx, weight, &w, false));
    IS_REQUIRE(ctx, w.step % sizeof(float) == 0, IS_ERR_BAD_ARG, "weight rows must be float aligned");
    IS_TRY(feather_weight_device(ctx, m, sharpness, w.ptr<float>(), w.step / sizeof(float)));
    IS_TRY(commit(ctx, &w));
    return IS_OK;
}

int is_feather_create(is_ctx* ctx, float sharpness, is_feather_blender** out) {
    if (!ctx || !out) return IS_ERR_BAD_ARG;
    is_feather_blender* b = new is_feather_blender();
    b->ctx = ctx;
    b->sharpness = sharpness;
    *out = b;
    return IS_OK;
}

int is_feather_destroy(is_feather_blender* b) {
    if (b) {
        cudaSetDevice(b->ctx->device);
        delete b;
    }
    return IS_OK;
}

int is_feather_prepare_roi(is_feather_blender* b, is_rect dst_roi) {
    if (!b) return IS_ERR_BAD_ARG;
    IS_REQUIRE(b->ctx, dst_roi.width > 0 && dst_roi.height > 0, IS_ERR_BAD_ARG, "empty destination ROI");
    b->fed.clear();
    b->roi = dst_roi;
    b->prepared = true;
    return IS_OK;
}

int is_feather_prepare(is_feather_blender* b, int n, const is_point* corners, const is_size* sizes) {   // Blender::prepare(corners, sizes) -> resultRoi
    if (!b) return IS_ERR_BAD_ARG;
    IS_REQUIRE(b->ctx, n > 0 && corners && sizes, IS_ERR_BAD_ARG, "prepare needs at least one image");
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        tlx = std::min(tlx, corners[i].x); tly = std::min(tly, corners[i].y);
        brx = std::max(brx, corners[i].x + sizes[i].width); bry = std::max(bry, corners[i].y + sizes[i].height);
    }
    return is_feather_prepare_roi(b, is_rect{tlx, tly, brx - tlx, bry - tly});
}

int is_feather_dst_size(const is_feather_blender* b, is_size* size) {
    if (!b || !size || !b->prepared) return IS_ERR_BAD_ARG;
    size->width = b->roi.width;
    size->height = b->roi.height;
    return IS_OK;
}

int is_feather_feed(is_feather_blender* b, const is_mat* img, const is_mat* mask, is_point tl) {
    if (!b) return IS_ERR_BAD_ARG;
    is_ctx* ctx = b->ctx;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, b->prepared, IS_ERR_ASSERT, "feed() before prepare()");
    IS_TRY(check_mat(ctx, img, "img"));
    IS_TRY(check_mat(ctx, mask, "mask"));
    IS_REQUIRE(ctx, img->channels == 3 && (img->depth == IS_16S || img->depth == IS_8U), IS_ERR_ASSERT, "img.type() == CV_16SC3 || img.type() == CV_8UC3");
    IS_REQUIRE(ctx, mask->depth == IS_8U && mask->channels == 1 && mask->rows == img->rows && mask->cols == img->cols, IS_ERR_ASSERT,
               "mask.type() == CV_8U && mask.size() == img.size()");
    IS_REQUIRE(ctx, tl.x >= b->roi.x && tl.y >= b->roi.y && tl.x + img->cols <= b->roi.x + b->roi.width && tl.y + img->rows <= b->roi.y + b->roi.height,
               IS_ERR_BAD_ARG, "image lies outside the prepared ROI");
    FeatherFed f;
    f.tl_x = tl.x; f.tl_y = tl.y;
    // OpenCV's feed() consumes the image immediately: keep a private copy (the weight map is computed now)
    IS_TRY(alloc_mat(ctx, img->rows, img->cols, 3, img->depth, &f.img));
    IS_CUDA(ctx, cudaMemcpy2DAsync(f.img.data, f.img.step, img->data, img->step, f.img.row_bytes(), img->rows,
                                   img->device >= 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    DevMat m;
    IS_TRY(stage_in(ctx, mask, &m));
    IS_TRY(f.w.alloc(ctx, sizeof(float) * (size_t)img->rows * img->cols));
    IS_TRY(feather_weight_device(ctx, m, b->sharpness, f.w.as<float>(), (size_t)img->cols));
    if (mask->device < 0) IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the staged mask copy goes away with `m`
    b->fed.push_back(std::move(f));
    return IS_OK;
}

int is_feather_blend(is_feather_blender* b, is_mat* dst, is_mat* dst_mask) {
    if (!b) return IS_ERR_BAD_ARG;
    is_ctx* ctx = b->ctx;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, b->prepared, IS_ERR_ASSERT, "blend() before prepare()");
    IS_TRY(check_mat(ctx, dst, "dst"));
    IS_TRY(check_mat(ctx, dst_mask, "dst_mask"));
    IS_REQUIRE(ctx, dst->depth == IS_16S && dst->channels == 3 && dst->rows == b->roi.height && dst->cols == b->roi.width, IS_ERR_BAD_ARG,
               "dst must be CV_16SC3 of is_feather_dst_size");
    IS_REQUIRE(ctx, dst_mask->depth == IS_8U && dst_mask->channels == 1 && dst_mask->rows == dst->rows && dst_mask->cols == dst->cols, IS_ERR_BAD_ARG,
               "dst_mask must be CV_8U of is_feather_dst_size");
    DevMat d, m;
    IS_TRY(stage_out(ctx, dst, &d, false));
    IS_TRY(stage_out(ctx, dst_mask, &m, false));
    const int n = (int)b->fed.size();
    std::vector<FeatherImg> host((size_t)std::max(n, 1));
    double bytes = (double)b->roi.width * b->roi.height * 7;
    for (int i = 0; i < n; ++i) {
        const FeatherFed& f = b->fed[i];
        host[i] = FeatherImg{f.img.data, f.img.step, f.img.depth == IS_8U ? 1 : 0, f.w.as<float>(), f.tl_x - b->roi.x, f.tl_y - b->roi.y, f.img.rows, f.img.cols};
        bytes += (double)f.img.rows * f.img.cols * ((f.img.depth == IS_8U ? 3 : 6) + 4);
    }
    DevBuf table;
    IS_TRY(table.alloc(ctx, sizeof(FeatherImg) * host.size()));
    IS_TRY(upload(ctx, table.p, host.data(), sizeof(FeatherImg) * host.size()));
    dim3 block(64, 4), grid(div_up(b->roi.width, 64), div_up(b->roi.height, 4));
    ctx->next_bytes = bytes;
    IS_LAUNCH(ctx, k_feather_blend, grid, block, 0, table.as<FeatherImg>(), n, b->roi.width, b->roi.height, d.ptr<int16_t>(), d.step, m.ptr<uint8_t>(), m.step);
    IS_TRY(commit(ctx, &d));
    IS_TRY(commit(ctx, &m));
    if (dst->device >= 0 && dst_mask->device >= 0) IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the fed copies are released below
    b->fed.clear();
    b->prepared = false;
    return IS_OK;
}

}  // extern "C"
