// exposure.cu -- gain exposure compensation.
//
// Replaces cv::detail::GainCompensator as the reference's mains drive it between the warp loop and the seam finder:
//   compensator->feed(corners, images_warped, masks_warped)                [BLEND]:117-123, [SEAM]:1165-1171
//   compensator->apply(img_idx, corners[img_idx], img_warped, mask_warped) (compositing loop, before Blender::feed)
// (OpenCV 3.4.2 modules/stitching/src/exposure_compensate.cpp, un-vendored; restated in oracle/exposure.cpp.)
//
// feed(): for every pair i <= j of overlapping images, N = max(1, #pixels with both masks == 255) and the mean of
// sqrt(b^2 + g^2 + r^2) over those pixels for either image; then an n x n linear solve on the host.  The per-pixel part
// is one reduction kernel per pair over the overlap rectangle: per-thread double sums, block tree in a fixed order, one
// partial per block, the partials added on the host in block order -- deterministic, and within a few ulp of the
// reference's raster-order sum (the gains agree to ~1e-15 relative; they are only used as multipliers of 8-bit values).
// apply(): saturate_cast<uchar>(cvRound(double(pixel) * gain)), four bytes per thread.
#include "internal.cuh"

#include <algorithm>
#include <cmath>

namespace is {

struct GainPartial { double s1, s2; long long cnt; long long pad; };
constexpr int GAIN_BLOCKS = 296, GAIN_THREADS = 256;

__global__ void __launch_bounds__(GAIN_THREADS) k_gain_overlap(const uint8_t* __restrict__ img1, size_t step1, const uint8_t* __restrict__ mask1, size_t mstep1,
                                                               const uint8_t* __restrict__ img2, size_t step2, const uint8_t* __restrict__ mask2, size_t mstep2,
                                                               int w, int h, GainPartial* __restrict__ out) {
    double s1 = 0., s2 = 0.;
    long long cnt = 0;
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const uint8_t *m1 = mask1 + (size_t)y * mstep1, *m2 = mask2 + (size_t)y * mstep2;
        const uint8_t *p1 = img1 + (size_t)y * step1, *p2 = img2 + (size_t)y * step2;
        for (int x = threadIdx.x; x < w; x += GAIN_THREADS) {
            if (m1[x] != 255 || m2[x] != 255) continue;
            ++cnt;
            const int a0 = p1[3 * x], a1 = p1[3 * x + 1], a2 = p1[3 * x + 2];
            const int b0 = p2[3 * x], b1 = p2[3 * x + 1], b2 = p2[3 * x + 2];
            s1 += sqrt((double)(a0 * a0 + a1 * a1 + a2 * a2));
            s2 += sqrt((double)(b0 * b0 + b1 * b1 + b2 * b2));
        }
    }
    // block reduction in a fixed order
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    __shared__ double w1[GAIN_THREADS / 32], w2[GAIN_THREADS / 32];
    __shared__ long long wc[GAIN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { w1[wid] = s1; w2[wid] = s2; wc[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        GainPartial p{0., 0., 0, 0};
        for (int k = 0; k < GAIN_THREADS / 32; ++k) { p.s1 += w1[k]; p.s2 += w2[k]; p.cnt += wc[k]; }
        out[blockIdx.x] = p;
    }
}

__global__ void k_gain_apply(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst, size_t dstep, int row_bytes, int rows, double gain) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= row_bytes || y >= rows) return;
    const uint8_t* s = src + (size_t)y * sstep + x;
    uint8_t* d = dst + (size_t)y * dstep + x;
    auto f = [gain](unsigned v) { return (unsigned)min(255, max(0, __double2int_rn((double)v * gain))); };   // cvRound, saturate
    if (x + 4 <= row_bytes && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 3) == 0) {
        const unsigned v = *reinterpret_cast<const unsigned*>(s);
        *reinterpret_cast<unsigned*>(d) = f(v & 255u) | (f((v >> 8) & 255u) << 8) | (f((v >> 16) & 255u) << 16) | (f(v >> 24) << 24);
    } else {
        for (int k = 0; k < 4 && x + k < row_bytes; ++k) d[k] = (uint8_t)f(s[k]);
    }
}

// device-resident images / masks -> gains[n]
int gain_feed_device(is_ctx* ctx, int n, const DevMat* images, const DevMat* masks, const is_point* corners, double* gains) {
    struct Pair { int i, j, x0, y0, w, h; };
    std::vector<Pair> pairs;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            const int x0 = std::max(corners[i].x, corners[j].x), y0 = std::max(corners[i].y, corners[j].y);
            const int x1 = std::min(corners[i].x + images[i].cols, corners[j].x + images[j].cols);
            const int y1 = std::min(corners[i].y + images[i].rows, corners[j].y + images[j].rows);
            if (x0 < x1 && y0 < y1) pairs.push_back(Pair{i, j, x0, y0, x1 - x0, y1 - y0});   // overlapRoi
        }
    const size_t np = pairs.size();
    DevBuf part;
    IS_TRY(part.alloc(ctx, sizeof(GainPartial) * GAIN_BLOCKS * std::max<size_t>(np, 1)));
    for (size_t k = 0; k < np; ++k) {
        const Pair& p = pairs[k];
        const DevMat &a = images[p.i], &b = images[p.j], &ma = masks[p.i], &mb = masks[p.j];
        const int ax = p.x0 - corners[p.i].x, ay = p.y0 - corners[p.i].y, bx = p.x0 - corners[p.j].x, by = p.y0 - corners[p.j].y;
        ctx->next_bytes = (double)p.w * p.h * 8;
        IS_LAUNCH(ctx, k_gain_overlap, GAIN_BLOCKS, GAIN_THREADS, 0, a.ptr<uint8_t>() + (size_t)ay * a.step + 3 * (size_t)ax, a.step,
                  ma.ptr<uint8_t>() + (size_t)ay * ma.step + ax, ma.step, b.ptr<uint8_t>() + (size_t)by * b.step + 3 * (size_t)bx, b.step,
                  mb.ptr<uint8_t>() + (size_t)by * mb.step + bx, mb.step, p.w, p.h, part.as<GainPartial>() + k * GAIN_BLOCKS);
    }
    std::vector<GainPartial> host(GAIN_BLOCKS * std::max<size_t>(np, 1));
    if (np) IS_TRY(download(ctx, host.data(), part.p, sizeof(GainPartial) * GAIN_BLOCKS * np));
    std::vector<double> N((size_t)n * n, 0.), I((size_t)n * n, 0.);
    for (size_t k = 0; k < np; ++k) {
        double s1 = 0., s2 = 0.;
        long long cnt = 0;
        for (int bl = 0; bl < GAIN_BLOCKS; ++bl) { const GainPartial& q = host[k * GAIN_BLOCKS + bl]; s1 += q.s1; s2 += q.s2; cnt += q.cnt; }
        const int i = pairs[k].i, j = pairs[k].j;
        const double nn = (double)std::max<long long>(1, cnt);
        N[(size_t)i * n + j] = N[(size_t)j * n + i] = nn;
        I[(size_t)i * n + j] = s1 / nn;
        I[(size_t)j * n + i] = s2 / nn;
    }
    // normal equations of GainCompensator::feed (alpha = 0.01, beta = 100), cv::solve(DECOMP_LU)
    const double alpha = 0.01, beta = 100.;
    std::vector<double> A((size_t)n * n, 0.), b(n, 0.);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double nij = N[(size_t)i * n + j];
            b[i] += beta * nij;
            A[(size_t)i * n + i] += beta * nij;
            if (j == i) continue;
            A[(size_t)i * n + i] += 2 * alpha * I[(size_t)i * n + j] * I[(size_t)i * n + j] * nij;
            A[(size_t)i * n + j] -= 2 * alpha * I[(size_t)i * n + j] * I[(size_t)j * n + i] * nij;
        }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r) if (std::fabs(A[(size_t)r * n + k]) > std::fabs(A[(size_t)piv * n + k])) piv = r;
        IS_REQUIRE(ctx, std::fabs(A[(size_t)piv * n + k]) >= 1e-300, IS_ERR_ASSERT, "gain compensation: singular system");
        if (piv != k) { for (int c = 0; c < n; ++c) std::swap(A[(size_t)k * n + c], A[(size_t)piv * n + c]); std::swap(b[k], b[piv]); }
        const double d = -1. / A[(size_t)k * n + k];
        for (int r = k + 1; r < n; ++r) {
            const double f = A[(size_t)r * n + k] * d;
            for (int c = k + 1; c < n; ++c) A[(size_t)r * n + c] += f * A[(size_t)k * n + c];
            b[r] += f * b[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < n; ++c) s -= A[(size_t)k * n + c] * gains[c];
        gains[k] = s / A[(size_t)k * n + k];
    }
    return IS_OK;
}

int gain_apply_device(is_ctx* ctx, const DevMat& src, const DevMat& dst, double gain) {
    const int row_bytes = src.cols * src.channels;
    dim3 block(64, 4), grid(div_up(div_up(row_bytes, 4), 64), div_up(src.rows, 4));
    ctx->next_bytes = 2.0 * row_bytes * src.rows;
    IS_LAUNCH(ctx, k_gain_apply, grid, block, 0, src.ptr<uint8_t>(), src.step, dst.ptr<uint8_t>(), dst.step, row_bytes, src.rows, gain);
    return IS_OK;
}

}  // namespace is

using namespace is;

extern "C" {

int is_gain_feed(is_ctx* ctx, int n, const is_point* corners, const is_mat* images, const is_mat* masks, double* gains) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, n > 0 && corners && images && masks && gains, IS_ERR_BAD_ARG, "null argument");
    std::vector<DevMat> im(n), mk(n);
    for (int i = 0; i < n; ++i) {
        IS_TRY(check_mat(ctx, &images[i], "image"));
        IS_TRY(check_mat(ctx, &masks[i], "mask"));
        IS_REQUIRE(ctx, images[i].depth == IS_8U && images[i].channels == 3, IS_ERR_ASSERT, "images[i].type() == CV_8UC3");
        IS_REQUIRE(ctx, masks[i].depth == IS_8U && masks[i].channels == 1 && masks[i].rows == images[i].rows && masks[i].cols == images[i].cols,
                   IS_ERR_ASSERT, "masks[i].type() == CV_8U && masks[i].size() == images[i].size()");
        IS_TRY(stage_in(ctx, &images[i], &im[i]));
        IS_TRY(stage_in(ctx, &masks[i], &mk[i]));
    }
    return gain_feed_device(ctx, n, im.data(), mk.data(), corners, gains);
}

int is_gain_apply(is_ctx* ctx, is_mat* image, double gain) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, image, "image"));
    IS_REQUIRE(ctx, image->depth == IS_8U, IS_ERR_ASSERT, "image.depth() == CV_8U");
    DevMat d;
    IS_TRY(stage_out(ctx, image, &d, true));
    IS_TRY(gain_apply_device(ctx, d, d, gain));
    IS_TRY(commit(ctx, &d));
    return IS_OK;
}

}  // extern "C"
