// blend.cu -- Gaussian/Laplacian-pyramid multi-band blend.
//
// Replaces cv::detail::MultiBandBlender as the reference's mains drive it (prepare / feed / blend,
// [SEAM]:1244-1252,1271,1280) together with the OpenCV routines underneath it: copyMakeBorder, pyrDown,
// pyrUp, createLaplacePyr, normalizeUsingWeightMap, restoreImageFromLaplacePyr.
//
// B200 formulation.  OpenCV scatters: every feed() read-modify-writes the panorama-sized accumulators
// dst_pyr_laplace_[k] / dst_band_weights_[k], then blend() normalises and collapses them in further passes.
// Here the accumulation is a gather: feed() only builds the image's Gaussian pyramid (levels 1..nb) and its
// weight pyramid; blend() runs ONE kernel per level, coarse to fine, which for each panorama pixel
//   - sums over the images covering it  short(laplacian_k * weight_k)  and  weight_k   (in feed order),
//     forming the Laplacian  g_k - pyrUp(g_{k+1})  on the fly,
//   - normalises, adds pyrUp of the already collapsed level k+1, and stores the collapsed level k.
// The panorama accumulators are never read back and level 0 is written exactly once (cropped, masked).
// int16 accumulation wraps modulo 2^16 exactly like OpenCV's `short +=`, so it is order independent; the
// float weight sum keeps OpenCV's feed order.  Results equal the scatter formulation bit for bit.
#include "internal.cuh"
#include "tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace is {

constexpr int MAX_LEVELS = 16;
#define IS_WEIGHT_EPS 1e-5f

__host__ __device__ __forceinline__ int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

__host__ __device__ __forceinline__ int reflect_b(int p, int len) {   // BORDER_REFLECT
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

// Level-0 view of a fed image: the image with its BORDER_REFLECT frame (copyMakeBorder), never materialised.
struct Level0 {
    const void* img; size_t istep; int is_u8;
    const uint8_t* mask; size_t mstep;
    int rows, cols;        // fed image
    int top, left;         // border offsets
    int height, width;     // padded size
    const uint8_t* summary; int sum_w;   // k_mask_summary of the mask (may be null)
};

__device__ __forceinline__ void l0_pixel(const Level0& L, int y, int x, int* v) {
    const int sy = reflect_b(y - L.top, L.rows), sx = reflect_b(x - L.left, L.cols);
    if (L.is_u8) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(L.img) + (size_t)sy * L.istep + 3 * sx;
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    } else {
        const int16_t* p = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(L.img) + (size_t)sy * L.istep) + 3 * sx;
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    }
}

// level-0 weight: copyMakeBorder(BORDER_CONSTANT 0) of  mask * (1/255.f)   (CV_32F)  or  mask ? mask + 1 : 0  (CV_16S)
__device__ __forceinline__ int l0_mask(const Level0& L, int y, int x) {
    const int sy = y - L.top, sx = x - L.left;
    if ((unsigned)sy >= (unsigned)L.rows || (unsigned)sx >= (unsigned)L.cols) return 0;
    return L.mask[(size_t)sy * L.mstep + sx];
}
__device__ __forceinline__ float l0_weight_f(const Level0& L, int y, int x) { return __fmul_rn((float)l0_mask(L, y, x), (float)(1. / 255.)); }
__device__ __forceinline__ int l0_weight_s(const Level0& L, int y, int x) { int m = l0_mask(L, y, x); return m ? m + 1 : 0; }

// ---- pyrDown ------------------------------------------------------------------------------------------------
// 16S: exact integer 5x5 [1 4 6 4 1]^2 with BORDER_REFLECT_101, (s + 128) >> 8.
// 32F: scalar order of OpenCV's pyramids.cpp: row = s2*6 + (s1+s3)*4 + s0 + s4 ; out = (r2*6 + (r1+r3)*4 + r0 + r4) * (1/256)

// PART selects what a launch produces: the image level, the weight level, or both.  The image pyramid does not depend on
// the mask, so the pipeline builds it while the seam stage is still running and adds the weights afterwards.
enum { PD_BOTH = 0, PD_IMAGE = 1, PD_WEIGHT = 2 };

template <bool WF, int PART>   // WF: float weights, else int16 weights
__global__ void k_pyrdown(const int16_t* __restrict__ g, const void* __restrict__ w, int sh, int sw, int16_t* __restrict__ gd,
                          void* __restrict__ wd, int dh, int dw) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int xs[5], ys[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) { xs[a] = reflect101(2 * x + a - 2, sw); ys[a] = reflect101(2 * y + a - 2, sh); }
    const int kk[5] = {1, 4, 6, 4, 1};
    int acc[3] = {0, 0, 0};
    float rowf[5];
    int wacc = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        int r[3] = {0, 0, 0};
        float wf[5];
        int ws = 0;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const size_t i = (size_t)ys[a] * sw + xs[b];
            if (PART != PD_WEIGHT) { r[0] += kk[b] * g[3 * i]; r[1] += kk[b] * g[3 * i + 1]; r[2] += kk[b] * g[3 * i + 2]; }
            if (PART != PD_IMAGE) {
                if (WF) wf[b] = reinterpret_cast<const float*>(w)[i];
                else ws += kk[b] * reinterpret_cast<const int16_t*>(w)[i];
            }
        }
        acc[0] += kk[a] * r[0]; acc[1] += kk[a] * r[1]; acc[2] += kk[a] * r[2];
        if (PART != PD_IMAGE) {
            if (WF) rowf[a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wf[2], 6.f), __fmul_rn(__fadd_rn(wf[1], wf[3]), 4.f)), wf[0]), wf[4]);
            else wacc += kk[a] * ws;
        }
    }
    const size_t o = (size_t)y * dw + x;
    if (PART != PD_WEIGHT) {
        gd[3 * o] = (int16_t)sat16((acc[0] + 128) >> 8);
        gd[3 * o + 1] = (int16_t)sat16((acc[1] + 128) >> 8);
        gd[3 * o + 2] = (int16_t)sat16((acc[2] + 128) >> 8);
    }
    if (PART != PD_IMAGE) {
        if (WF) {
            float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(rowf[2], 6.f), __fmul_rn(__fadd_rn(rowf[1], rowf[3]), 4.f)), rowf[0]), rowf[4]);
            reinterpret_cast<float*>(wd)[o] = __fmul_rn(v, 1.f / 256.f);
        } else {
            reinterpret_cast<int16_t*>(wd)[o] = (int16_t)sat16((wacc + 128) >> 8);
        }
    }
}

// Weight levels >= 2 of all fed images in one launch per level (blockIdx.z = image): the levels are small, and thirty
// dependent launches cost more than their arithmetic.  Same float association order as k_pyrdown.
struct WeightLevel { const void* w_in; void* w_out; int sh, sw, dh, dw; };

template <bool WF>
__global__ void k_pyrdown_weights_batch(const WeightLevel* __restrict__ table) {
    const WeightLevel T = table[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= T.dw || y >= T.dh) return;
    int xs[5], ys[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) { xs[a] = reflect101(2 * x + a - 2, T.sw); ys[a] = reflect101(2 * y + a - 2, T.sh); }
    const int kk[5] = {1, 4, 6, 4, 1};
    float rowf[5];
    int wacc = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        float wf[5];
        int ws = 0;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const size_t i = (size_t)ys[a] * T.sw + xs[b];
            if (WF) wf[b] = reinterpret_cast<const float*>(T.w_in)[i];
            else ws += kk[b] * reinterpret_cast<const int16_t*>(T.w_in)[i];
        }
        if (WF) rowf[a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wf[2], 6.f), __fmul_rn(__fadd_rn(wf[1], wf[3]), 4.f)), wf[0]), wf[4]);
        else wacc += kk[a] * ws;
    }
    const size_t o = (size_t)y * T.dw + x;
    if (WF) {
        const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(rowf[2], 6.f), __fmul_rn(__fadd_rn(rowf[1], rowf[3]), 4.f)), rowf[0]), rowf[4]);
        reinterpret_cast<float*>(T.w_out)[o] = __fmul_rn(v, 1.f / 256.f);
    } else {
        reinterpret_cast<int16_t*>(T.w_out)[o] = (int16_t)sat16((wacc + 128) >> 8);
    }
}

// Image levels >= 2 of all fed images in one launch per level (blockIdx.z = image), through shared memory: a block computes a
// 32 x 16 tile of the coarser level from the (2 * 32 + 3) x (2 * 16 + 3) tile under it -- loaded once (32-bit words in the
// interior, element-wise through BORDER_REFLECT_101 on the rim), horizontal 5-tap sums to shared memory, vertical pass.
struct ImageLevel { const int16_t* g_in; int16_t* g_out; int sh, sw, dh, dw; };
constexpr int PU_TX = 32, PU_TY = 16, PU_IW = 2 * PU_TX + 3, PU_IH = 2 * PU_TY + 3;
constexpr int PU_ROW = 3 * PU_IW + 3;                         // int16 per shared-memory row: 201 values + padding to an even count

__global__ void __launch_bounds__(256) k_pyrdown_images_batch(const ImageLevel* __restrict__ table) {
    const ImageLevel T = table[blockIdx.z];
    const int ox0 = blockIdx.x * PU_TX, oy0 = blockIdx.y * PU_TY;
    if (ox0 >= T.dw || oy0 >= T.dh) return;                   // uniform over the block
    __shared__ __align__(16) int16_t tile[PU_IH][PU_ROW];
    __shared__ int hsum[PU_IH][PU_TX][3];
    const int tid = threadIdx.x;
    const int x_lo = 2 * ox0 - 2, y_lo = 2 * oy0 - 2;
    const bool interior = x_lo >= 0 && y_lo >= 0 && x_lo + PU_IW <= T.sw && y_lo + PU_IH <= T.sh;
    if (interior) {
        // a tile row is 67 pixels = 402 bytes starting at a multiple of 12 bytes: 100 words + one int16
        constexpr int WORDS = (3 * PU_IW) / 2;                // 100
        for (int e = tid; e < PU_IH * (WORDS + 1); e += 256) {
            const int r = e / (WORDS + 1), wd = e % (WORDS + 1);
            const int16_t* src = T.g_in + ((size_t)(y_lo + r) * T.sw + x_lo) * 3;
            if (wd < WORDS) reinterpret_cast<uint32_t*>(&tile[r][0])[wd] = reinterpret_cast<const uint32_t*>(src)[wd];
            else tile[r][3 * PU_IW - 1] = src[3 * PU_IW - 1];
        }
    } else {
        for (int e = tid; e < PU_IH * PU_IW; e += 256) {
            const int r = e / PU_IW, c = e % PU_IW;
            const int y = reflect101(y_lo + r, T.sh), x = reflect101(x_lo + c, T.sw);
            const int16_t* src = T.g_in + ((size_t)y * T.sw + x) * 3;
            tile[r][3 * c] = src[0]; tile[r][3 * c + 1] = src[1]; tile[r][3 * c + 2] = src[2];
        }
    }
    __syncthreads();
    for (int e = tid; e < PU_IH * PU_TX; e += 256) {
        const int r = e / PU_TX, ox = e % PU_TX;
        const int16_t* p = &tile[r][6 * ox];
#pragma unroll
        for (int c = 0; c < 3; ++c) hsum[r][ox][c] = (int)p[c] + 4 * (int)p[3 + c] + 6 * (int)p[6 + c] + 4 * (int)p[9 + c] + (int)p[12 + c];
    }
    __syncthreads();
    const int ox = tid & 31;
    const int x = ox0 + ox;
    if (x >= T.dw) return;
    for (int oy = tid >> 5; oy < PU_TY; oy += 8) {
        const int y = oy0 + oy;
        if (y >= T.dh) break;
        int16_t* o = T.g_out + ((size_t)y * T.dw + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int acc = hsum[2 * oy][ox][c] + 4 * hsum[2 * oy + 1][ox][c] + 6 * hsum[2 * oy + 2][ox][c] + 4 * hsum[2 * oy + 3][ox][c] + hsum[2 * oy + 4][ox][c];
            o[c] = (int16_t)sat16((acc + 128) >> 8);
        }
    }
}

// ---- pyrUp evaluated at one destination pixel ------------------------------------------------------------------
// per axis: even 2i: s[i-1] + 6 s[i] + s[i+1], odd 2i+1: 4 (s[i] + s[i+1]); s[-1] -> s[1], s[n] -> s[n-1]; (t + 32) >> 6
struct UpTaps { int i[3]; int w[3]; };

__device__ __forceinline__ UpTaps up_taps(int d, int n) {
    UpTaps t;
    const int i = d >> 1;
    if ((d & 1) == 0) {
        t.i[0] = i > 0 ? i - 1 : (n > 1 ? 1 : 0); t.w[0] = 1;
        t.i[1] = i; t.w[1] = 6;
        t.i[2] = i + 1 < n ? i + 1 : n - 1; t.w[2] = 1;
    } else {
        t.i[0] = i; t.w[0] = 4;
        t.i[1] = i + 1 < n ? i + 1 : n - 1; t.w[1] = 4;
        t.i[2] = i; t.w[2] = 0;
    }
    return t;
}

__device__ __forceinline__ void pyrup_at(const int16_t* __restrict__ s, int sh, int sw, int y, int x, int* out) {
    const UpTaps ty = up_taps(y, sh), tx = up_taps(x, sw);
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (ty.w[a] == 0) continue;
        const int16_t* row = s + (size_t)ty.i[a] * sw * 3;
        int r[3] = {0, 0, 0};
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (tx.w[b] == 0) continue;
            const int16_t* p = row + 3 * tx.i[b];
            r[0] += tx.w[b] * p[0]; r[1] += tx.w[b] * p[1]; r[2] += tx.w[b] * p[2];
        }
        acc[0] += ty.w[a] * r[0]; acc[1] += ty.w[a] * r[1]; acc[2] += ty.w[a] * r[2];
    }
    out[0] = sat16((acc[0] + 32) >> 6);
    out[1] = sat16((acc[1] + 32) >> 6);
    out[2] = sat16((acc[2] + 32) >> 6);
}

// ---- one pyramid level of blend(): gather + normalise + collapse -------------------------------------------------
struct ImgLevel {
    Level0 l0;                 // used when k == 0
    const int16_t* g;          // Gaussian level k   (k >= 1), dims h x w
    const int16_t* g_up;       // Gaussian level k+1 (k < nb), dims (h/2) x (w/2)
    const void* wgt;           // weight level k     (k >= 1)
    int x_tl, y_tl;            // position of the image's level-k array inside the panorama's level-k array
    int h, w;                  // level-k dims
};

struct LevelArgs {
    const ImgLevel* imgs; int n;
    int k, nb;
    int H, W;                          // panorama level-k dims (padded)
    int16_t* out;                      // collapsed level k (H x W x 3), k >= 1
    const int16_t* up; int uh, uw;     // collapsed level k+1
    // k == 0 only: final outputs
    int16_t* dst; size_t dstep; uint8_t* dmask; size_t mstep; int fw, fh;
    int xb, xe;                        // columns of this level computed by the launch (xb even)
    int sx0, sx1;                      // level 0: columns of the final ROI stored (dst column = x - sx0)
};

template <bool WF>
__global__ void k_blend_level(LevelArgs A) {
    const int x = A.xb + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int W = A.k == 0 ? min(A.fw, A.sx1) : min(A.W, A.xe), H = A.k == 0 ? A.fh : A.H;   // level 0 is cropped to the final ROI
    if (x >= W || y >= H || (A.k == 0 && x < A.sx0)) return;
    int acc[3] = {0, 0, 0};
    float wsum_f = 0.f;
    int wsum_s = 0;
    for (int i = 0; i < A.n; ++i) {
        const ImgLevel& I = A.imgs[i];
        const int lx = x - I.x_tl, ly = y - I.y_tl;
        if ((unsigned)lx >= (unsigned)I.w || (unsigned)ly >= (unsigned)I.h) continue;
        int g[3];
        if (A.k == 0) l0_pixel(I.l0, ly, lx, g);
        else { const int16_t* p = I.g + ((size_t)ly * I.w + lx) * 3; g[0] = p[0]; g[1] = p[1]; g[2] = p[2]; }
        if (A.k < A.nb) {   // Laplacian = g_k - pyrUp(g_{k+1}), saturating (cv::subtract)
            int u[3];
            pyrup_at(I.g_up, I.h >> 1, I.w >> 1, ly, lx, u);
            g[0] = sat16(g[0] - u[0]); g[1] = sat16(g[1] - u[1]); g[2] = sat16(g[2] - u[2]);
        }
        if (WF) {
            const float w = A.k == 0 ? l0_weight_f(I.l0, ly, lx) : reinterpret_cast<const float*>(I.wgt)[(size_t)ly * I.w + lx];
            // dst += static_cast<short>(src * w): truncation toward zero, int16 wrap-around add
            acc[0] += (int)(int16_t)__float2int_rz(__fmul_rn((float)g[0], w));
            acc[1] += (int)(int16_t)__float2int_rz(__fmul_rn((float)g[1], w));
            acc[2] += (int)(int16_t)__float2int_rz(__fmul_rn((float)g[2], w));
            wsum_f = __fadd_rn(wsum_f, w);
        } else {
            const int w = A.k == 0 ? l0_weight_s(I.l0, ly, lx) : (int)reinterpret_cast<const int16_t*>(I.wgt)[(size_t)ly * I.w + lx];
            acc[0] += (int)(int16_t)((g[0] * w) >> 8);
            acc[1] += (int)(int16_t)((g[1] * w) >> 8);
            acc[2] += (int)(int16_t)((g[2] * w) >> 8);
            wsum_s = (int)(int16_t)(wsum_s + w);
        }
    }
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d = (int)(int16_t)acc[c];   // the int16 accumulator of OpenCV
        if (WF) v[c] = (int)(int16_t)__float2int_rz(__fdiv_rn((float)d, __fadd_rn(wsum_f, IS_WEIGHT_EPS)));
        else v[c] = (int)(int16_t)((d * 256) / (wsum_s + 1));
    }
    if (A.k < A.nb) {   // restoreImageFromLaplacePyr: pyr[k] = pyrUp(pyr[k+1]) + pyr[k], saturating
        int u[3];
        pyrup_at(A.up, A.uh, A.uw, y, x, u);
        v[0] = sat16(v[0] + u[0]); v[1] = sat16(v[1] + u[1]); v[2] = sat16(v[2] + u[2]);
    }
    if (A.k > 0) {
        int16_t* o = A.out + ((size_t)y * A.W + x) * 3;
        o[0] = (int16_t)v[0]; o[1] = (int16_t)v[1]; o[2] = (int16_t)v[2];
    } else {
        const bool on = WF ? (wsum_f > IS_WEIGHT_EPS) : (wsum_s >= 1);
        int16_t* o = reinterpret_cast<int16_t*>(reinterpret_cast<char*>(A.dst) + (size_t)y * A.dstep) + 3 * (x - A.sx0);
        o[0] = on ? (int16_t)v[0] : 0; o[1] = on ? (int16_t)v[1] : 0; o[2] = on ? (int16_t)v[2] : 0;
        A.dmask[(size_t)y * A.mstep + (x - A.sx0)] = on ? 255 : 0;
    }
}


// ---- quad variant: one thread per 2x2 block of level-k pixels (k < nb) --------------------------------------------
// The four pixels of an even-aligned 2x2 block share one 3x3 neighbourhood of the next coarser level, so pyrUp
// costs 9 coarse loads per block instead of 9 per pixel (per axis: even = s[i-1] + 6 s[i] + s[i+1], odd =
// 4 (s[i] + s[i+1]); s[-1] -> s[1], s[n] -> s[n-1]).  Image rectangles are even-aligned on every level below the
// top one (offsets and sizes are multiples of 2^(nb-k)), so a block is entirely inside or outside an image.
__device__ __forceinline__ void pyrup_quad(const int16_t* __restrict__ s, int sh, int sw, int cy, int cx, int out[4][3]) {
    const int r0 = cy > 0 ? cy - 1 : (sh > 1 ? 1 : 0), r2 = cy + 1 < sh ? cy + 1 : sh - 1;
    const int c0 = cx > 0 ? cx - 1 : (sw > 1 ? 1 : 0), c2 = cx + 1 < sw ? cx + 1 : sw - 1;
    const int rows[3] = {r0, cy, r2};
    int e[3][3], o[3][3];   // per row: even-column and odd-column horizontal sums
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int16_t* row = s + (size_t)rows[a] * sw * 3;
        const int16_t* p0 = row + 3 * c0;
        const int16_t* p1 = row + 3 * cx;
        const int16_t* p2 = row + 3 * c2;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int v0 = p0[c], v1 = p1[c], v2 = p2[c];
            e[a][c] = v0 + 6 * v1 + v2;
            o[a][c] = 4 * (v1 + v2);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // every output is a weighted mean of int16 values with weights summing to 64, rounded: it cannot leave the int16 range,
        // so the saturate_cast of pyrUp is the identity here (a fifth of this kernel's instructions were these clamps)
        out[0][c] = (e[0][c] + 6 * e[1][c] + e[2][c] + 32) >> 6;   // (even y, even x)
        out[1][c] = (o[0][c] + 6 * o[1][c] + o[2][c] + 32) >> 6;   // (even y, odd x)
        out[2][c] = (4 * (e[1][c] + e[2][c]) + 32) >> 6;           // (odd y, even x)
        out[3][c] = (4 * (o[1][c] + o[2][c]) + 32) >> 6;           // (odd y, odd x)
    }
}

template <bool WF>
__device__ __forceinline__ void blend_quad_at(const LevelArgs& A, int x, int y) {   // the 2 x 2 block at even (x, y)
    const int qy = y >> 1;
    const int W = A.k == 0 ? min(A.fw, A.sx1) : min(A.W, A.xe), H = A.k == 0 ? A.fh : A.H;   // level 0 is cropped to the final ROI
    if (x >= W || y >= H) return;
    int acc[4][3];
    float wsum_f[4] = {0.f, 0.f, 0.f, 0.f};
    int wsum_s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[q][0] = 0; acc[q][1] = 0; acc[q][2] = 0; }
    for (int i = 0; i < A.n; ++i) {
        const ImgLevel& I = A.imgs[i];
        const int lx = x - I.x_tl, ly = y - I.y_tl;
        if ((unsigned)lx >= (unsigned)I.w || (unsigned)ly >= (unsigned)I.h) continue;
        // An image whose four weights are zero here adds short(g * 0) = 0 and w = 0 exactly: nothing of it needs to be read.  Away
        // from the seams that is every covering image but one.
        float wq_f[4] = {0.f, 0.f, 0.f, 0.f};
        int wq_s[4] = {0, 0, 0, 0};
        if (A.k > 0) {
            if (WF) {
                const float2 w0 = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(I.wgt) + (size_t)ly * I.w + lx);
                const float2 w1 = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(I.wgt) + (size_t)(ly + 1) * I.w + lx);
                wq_f[0] = w0.x; wq_f[1] = w0.y; wq_f[2] = w1.x; wq_f[3] = w1.y;
                if (w0.x == 0.f && w0.y == 0.f && w1.x == 0.f && w1.y == 0.f) continue;
            } else {
                const short2 w0 = *reinterpret_cast<const short2*>(reinterpret_cast<const int16_t*>(I.wgt) + (size_t)ly * I.w + lx);
                const short2 w1 = *reinterpret_cast<const short2*>(reinterpret_cast<const int16_t*>(I.wgt) + (size_t)(ly + 1) * I.w + lx);
                wq_s[0] = w0.x; wq_s[1] = w0.y; wq_s[2] = w1.x; wq_s[3] = w1.y;
                if ((w0.x | w0.y | w1.x | w1.y) == 0) continue;
            }
        }
        int u[4][3];
        pyrup_quad(I.g_up, I.h >> 1, I.w >> 1, ly >> 1, lx >> 1, u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int px = lx + (q & 1), py = ly + (q >> 1);
            int g[3];
            if (A.k == 0) l0_pixel(I.l0, py, px, g);
            else { const int16_t* p = I.g + ((size_t)py * I.w + px) * 3; g[0] = p[0]; g[1] = p[1]; g[2] = p[2]; }
            g[0] = sat16(g[0] - u[q][0]); g[1] = sat16(g[1] - u[q][1]); g[2] = sat16(g[2] - u[q][2]);
            if (WF) {
                const float w = A.k == 0 ? l0_weight_f(I.l0, py, px) : wq_f[q];
                acc[q][0] += (int)(int16_t)__float2int_rz(__fmul_rn((float)g[0], w));
                acc[q][1] += (int)(int16_t)__float2int_rz(__fmul_rn((float)g[1], w));
                acc[q][2] += (int)(int16_t)__float2int_rz(__fmul_rn((float)g[2], w));
                wsum_f[q] = __fadd_rn(wsum_f[q], w);
            } else {
                const int w = A.k == 0 ? l0_weight_s(I.l0, py, px) : wq_s[q];
                acc[q][0] += (int)(int16_t)((g[0] * w) >> 8);
                acc[q][1] += (int)(int16_t)((g[1] * w) >> 8);
                acc[q][2] += (int)(int16_t)((g[2] * w) >> 8);
                wsum_s[q] = (int)(int16_t)(wsum_s[q] + w);
            }
        }
    }
    int up[4][3];
    pyrup_quad(A.up, A.uh, A.uw, qy, x >> 1, up);
    int v[4][3];
    bool on[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int d = (int)(int16_t)acc[q][c];
            int n;
            // away from the seams the weight sum is exactly 1.0f on every level (pyrDown of a constant 1 is exact), and
            // short(d / (1 + 1e-5f)) == d - sign(d) for every int16 d (tests/test_blend_math.py): no IEEE division there
            if (WF) n = wsum_f[q] == 1.0f ? d - (d > 0) + (d < 0) : (int)(int16_t)__float2int_rz(__fdiv_rn((float)d, __fadd_rn(wsum_f[q], IS_WEIGHT_EPS)));
            else n = (int)(int16_t)((d * 256) / (wsum_s[q] + 1));
            v[q][c] = sat16(n + up[q][c]);
        }
        on[q] = WF ? (wsum_f[q] > IS_WEIGHT_EPS) : (wsum_s[q] >= 1);
    }
    if (A.k > 0) {   // W and H are even on these levels: the block is complete
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint32_t* o = reinterpret_cast<uint32_t*>(A.out + ((size_t)(y + r) * A.W + x) * 3);
            const int a = 2 * r, b = 2 * r + 1;
            o[0] = (uint32_t)(uint16_t)v[a][0] | ((uint32_t)(uint16_t)v[a][1] << 16);
            o[1] = (uint32_t)(uint16_t)v[a][2] | ((uint32_t)(uint16_t)v[b][0] << 16);
            o[2] = (uint32_t)(uint16_t)v[b][1] | ((uint32_t)(uint16_t)v[b][2] << 16);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (y + r >= H) break;
            int16_t* o = reinterpret_cast<int16_t*>(reinterpret_cast<char*>(A.dst) + (size_t)(y + r) * A.dstep) + 3 * (x - A.sx0);
            uint8_t* m = A.dmask + (size_t)(y + r) * A.mstep + (x - A.sx0);
            const int a = 2 * r, b = 2 * r + 1;
            int va[3], vb[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { va[c] = on[a] ? v[a][c] : 0; vb[c] = on[b] ? v[b][c] : 0; }
            if (x < A.sx0) {          // strip starting at an odd column: only the second pixel of the block belongs to it
                if (x + 1 < W) { o[3] = (int16_t)vb[0]; o[4] = (int16_t)vb[1]; o[5] = (int16_t)vb[2]; m[1] = on[b] ? 255 : 0; }
                continue;
            }
            if (x + 1 < W && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
                uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
                o32[0] = (uint32_t)(uint16_t)va[0] | ((uint32_t)(uint16_t)va[1] << 16);
                o32[1] = (uint32_t)(uint16_t)va[2] | ((uint32_t)(uint16_t)vb[0] << 16);
                o32[2] = (uint32_t)(uint16_t)vb[1] | ((uint32_t)(uint16_t)vb[2] << 16);
            } else {
                o[0] = (int16_t)va[0]; o[1] = (int16_t)va[1]; o[2] = (int16_t)va[2];
                if (x + 1 < W) { o[3] = (int16_t)vb[0]; o[4] = (int16_t)vb[1]; o[5] = (int16_t)vb[2]; }
            }
            m[0] = on[a] ? 255 : 0;
            if (x + 1 < W) m[1] = on[b] ? 255 : 0;
        }
    }
}

template <bool WF>
__global__ void __launch_bounds__(256) k_blend_level_quad(LevelArgs A) {
    const int qx = blockIdx.x * blockDim.x + threadIdx.x;
    const int qy = blockIdx.y * blockDim.y + threadIdx.y;
    blend_quad_at<WF>(A, A.xb + 2 * qx, 2 * qy);
}

// Level-0 bands (64 x 8 pixels at (x, y)) the tiled kernel below handed back: several contributing images or masks other
// than 0 / 255.  Same arithmetic as k_blend_level_quad; the list is produced on the device, so the grid is fixed and
// every CTA walks the list.
template <bool WF>
__global__ void __launch_bounds__(128) k_blend_l0_bands(LevelArgs A, const int2* __restrict__ bands, const int* __restrict__ count) {
    const int n = *count;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int2 o = bands[b];
        blend_quad_at<WF>(A, o.x + 2 * (threadIdx.x & 31), o.y + 2 * (threadIdx.x >> 5));
    }
}

// ---- level 0 of blend(), tiled: TMA-staged 2-D tiles, persistent CTAs walking 64 x 64 panorama tiles ------------------
// Level 0 carries most of the blend's bytes (u8 image + mask in, int16 panorama + mask out).  Facts used here:
//   * the level-0 weight of an image is its mask (x 1/255, or +1 for CV_16S) inside the image and 0 in the
//     copyMakeBorder frame; a pixel whose weight is 0 adds  short(laplacian * 0) = 0  and  w = 0  to the sums, so only
//     pixels with a non-zero mask matter and the BORDER_REFLECT view of the image is never needed;
//   * a coarse occupancy map of every mask (one byte per 64 x 32 cell, k_mask_summary) tells which images can
//     contribute to a tile at all, so the overlap zones do not fetch the image that lost the seam;
//   * mask 255 gives w = 255 * (1/255.f) = 1.0f exactly, and with a weight sum of exactly 1.0f the normalisation
//     short(d / (1.0f + 1e-5f)) equals d - sign(d) for every int16 d (checked exhaustively in tests/test_blend_math.py):
//     the common pixel needs no floating point at all.
// The tiles (collapsed level 1, and per contributing image: mask, u8 image, Gaussian level 1) arrive by
// cp.async.bulk.tensor.2d into shared memory, out-of-range parts zero-filled by the TMA unit.  Every thread owns a
// column of four vertically adjacent 2 x 2 quads and walks the six level-1 rows under them (pyrUp is separable:
// horizontal sums once per level-1 row, vertical combination per output row) in a rolled loop small enough for the
// instruction cache.
constexpr int L0_TW = 64, L0_TH = 64;                        // level-0 pixels per tile
constexpr int L0_QN = L0_TH / 16;                            // quads per thread: 8 consumer warps x 4 quads x 2 rows
constexpr int L0_UW = L0_TW / 2 + 2, L0_UH = L0_TH / 2 + 2;  // level-1 pixels per tile incl. the pyrUp halo: 34 x 34
// TMA wants the innermost box start 16-byte aligned (probed on B200: an unaligned start raises "illegal instruction"), so every
// box starts at the aligned address below the tile and is up to 15 bytes wider; the kernel indexes with the shift.
constexpr int L0_UROW = 112;                                 // int16 per shared-memory row of a level-1 tile: 3 * 34 = 102, + 7 shift, padded to 16 B
constexpr int L0_UBYTES = L0_UH * L0_UROW * 2;               // 7616
constexpr int L0_UALLOC = 7680;                              // ... rounded up to 128 B
constexpr int L0_IROW = 208;                                 // bytes per row of a u8 x 3 image tile: 192 + 15 shift, padded to 16 B
constexpr int L0_IBYTES = L0_TH * L0_IROW;                   // 13312
constexpr int L0_MROW = 80;                                  // bytes per row of a mask tile: 64 + 15 shift, padded to 16 B
constexpr int L0_MBYTES = L0_TH * L0_MROW;                   // 5120
constexpr int SUM_CW = 64, SUM_CH = 32;                      // cell of the mask occupancy map

struct L0Img {
    int X0, Y0;            // image origin (the real image, not the padded frame) in panorama level-0 coordinates
    int rows, cols;        // real image size
    int fx, fy;            // padded frame origin in panorama level-0 coordinates (multiples of 2^nb)
    int h1, w1;            // Gaussian level-1 dims of the frame
    const uint8_t* summary; int sum_w, sum_h;
};

struct L0Args {
    const CUtensorMap* maps;   // [0] collapsed level 1; [1 + 3 i ...]: mask, image, Gaussian level 1 of image i
    const L0Img* imgs; int n;
    int uh, uw;                // collapsed level-1 dims
    int16_t* dst; size_t dstep; uint8_t* dmask; size_t mstep;
    int W, H;                  // panorama columns [sx0, W) and rows [0, H) are stored
    int xb, sx0;               // first computed column (even), first stored column
    int ntx, ntiles;           // tiles per row, tiles in total
    int2* bands; int* nbands;  // out: 64 x 8 bands left to k_blend_l0_bands (several contributing images, masks other than 0 / 255)
};

// One warp per 64 x 32 cell (lane = row of the cell, 64 bytes as four 16-byte loads when the mask allows it).
// out[cell]: bit 0 = some pixel of the cell is non-zero, bit 1 = every pixel of the cell (inside the image) is 255.
__global__ void __launch_bounds__(256) k_mask_summary(const uint8_t* __restrict__ mask, size_t step, int rows, int cols, uint8_t* __restrict__ out,
                                                      int sw, int sh) {
    const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= sw * sh) return;
    const int lane = threadIdx.x & 31;
    const int cx = cell % sw, cy = cell / sw;
    const int y = cy * SUM_CH + lane, x0 = cx * SUM_CW;
    unsigned any = 0, all = 0xffffffffu;
    if (y < rows) {
        const uint8_t* p = mask + (size_t)y * step + x0;
        if (x0 + SUM_CW <= cols && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
            for (int k = 0; k < SUM_CW / 16; ++k) {
                const uint4 v = __ldg(q + k);
                any |= v.x | v.y | v.z | v.w;
                all &= v.x & v.y & v.z & v.w;
            }
        } else {
            for (int x = x0; x < min(x0 + SUM_CW, cols); ++x) {
                const unsigned m = p[x - x0];
                any |= m;
                if (m != 255u) all = 0u;
            }
        }
    }
    const bool w_any = __any_sync(0xffffffffu, any != 0);
    const bool w_all = __all_sync(0xffffffffu, all == 0xffffffffu);
    if (lane == 0) out[cell] = (uint8_t)((w_any ? 1 : 0) | (w_all ? 2 : 0));
}

// pyrUp is separable.  Horizontal sums of a level-1 row at three (border-mapped) columns: E = s[i-1] + 6 s[i] + s[i+1],
// O = s[i] + s[i+1] (the odd output is 4 * O; the factor is folded into the final shifts).
// pyrUp of a 2 x 2 quad from the horizontal sums of its three level-1 rows; u[q]: q = 0 (even y, even x), 1 (even y, odd x),
// 2 (odd y, even x), 3 (odd y, odd x).  Same integers as (t + 32) >> 6 of the separable [1 6 1] / [4 4] taps.
__device__ __forceinline__ void l0_quad_up(const int Ea[3], const int Oa[3], const int Eb[3], const int Ob[3], const int Ec[3], const int Oc[3],
                                           int u[4][3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        u[0][c] = (Ea[c] + 6 * Eb[c] + Ec[c] + 32) >> 6;
        u[1][c] = (Oa[c] + 6 * Ob[c] + Oc[c] + 8) >> 4;
        u[2][c] = (Eb[c] + Ec[c] + 8) >> 4;
        u[3][c] = (Ob[c] + Oc[c] + 2) >> 2;
    }
}
__device__ __forceinline__ void l0_hrow(const int16_t* a, const int16_t* b, const int16_t* c, int E[3], int O[3]) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int v0 = a[ch], v1 = b[ch], v2 = c[ch];
        E[ch] = v0 + 6 * v1 + v2;
        O[ch] = v1 + v2;
    }
}

// Persistent, warp-specialised pipeline: 8 consumer warps (one 64 x 8 pixel band of the tile each) + 1 producer warp.
// The producer finds the images that can contribute to the next tile (occupancy maps), then issues the tile's TMA loads
// into the free stage while the consumers still compute the previous one; full[] / empty[] mbarriers hand the two
// stages back and forth, no __syncthreads in steady state.  A tile with several contributing images (along a seam) takes one
// round per image: the first round stores every pixel (zeros where its mask is clear), the later rounds store the pixels
// of their own mask.  That is exact as long as the masks are 0 / 255 and no pixel is claimed twice; a band (64 x 8) that
// sees a fractional mask, or a pixel claimed by two images, goes on a device-side work list for k_blend_l0_bands, which
// runs afterwards and overwrites it with the general arithmetic.
constexpr int L0_STAGE_BYTES = L0_UALLOC + L0_MBYTES + L0_IBYTES + L0_UALLOC;   // collapsed level 1 | mask | image | Gaussian level 1 = 33792
constexpr int L0_OFF_MASK = L0_UALLOC, L0_OFF_IMG = L0_UALLOC + L0_MBYTES, L0_OFF_G1 = L0_UALLOC + L0_MBYTES + L0_IBYTES;
constexpr int L0_CONSUMERS = 256, L0_THREADS = L0_CONSUMERS + 32;
constexpr int L0_MAXCAND = 32;                                        // capacity of the per-tile candidate list (only its head is used)

struct L0Round {         // producer -> consumers, one per stage
    int x0t, y0t;        // tile origin (level-0 panorama coordinates)
    int ncand;           // image in this round (0 or 1); -1: no more work
    int flags;           // 1: first round of the tile; 8 / 16: the collapsed / the Gaussian level-1 neighbourhood lies inside its array
    int lx0;             // tile origin in image coordinates
    int ox, oy;          // level-1 origin of the Gaussian tile inside the image's frame
    int w1, h1;          // level-1 frame dims
};
constexpr int L0_SMEM_BYTES = 2 * L0_STAGE_BYTES + 64 + 2 * 64 + L0_MAXCAND * 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// consumer-only barrier (the producer warp never joins)
__device__ __forceinline__ void l0_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(L0_CONSUMERS) : "memory"); }

// pyrUp border rule (s[-1] -> s[1], s[n] -> s[n-1]) applied to a level-1 tile in shared memory, so that every thread can use
// the fixed-offset neighbourhood: the halo row / column that falls outside the array is overwritten with its mirror.
// (ox, oy): level-1 origin of the tile inside an array of w1 x h1.  All consumer threads call this.
__device__ __forceinline__ void l0_patch_border(unsigned char* tile, int ox, int oy, int w1, int h1, int ctid) {
    int16_t* base = reinterpret_cast<int16_t*>(tile) + ((6 * ox - ((6 * ox) & ~15)) >> 1);
    // rows
    if (ctid < 3 * L0_UW) {
        if (oy < 0 && 1 - oy < L0_UH) base[(-1 - oy) * L0_UROW + ctid] = base[(1 - oy) * L0_UROW + ctid];
        const int rb = h1 - oy;                        // first row below the array
        if (rb >= 1 && rb < L0_UH) base[rb * L0_UROW + ctid] = base[(rb - 1) * L0_UROW + ctid];
    }
    l0_consumer_sync();
    // columns
    if (ctid < 3 * L0_UH) {
        const int r = ctid / 3, c = ctid % 3;
        if (ox < 0 && 1 - ox < L0_UW) base[r * L0_UROW + 3 * (-1 - ox) + c] = base[r * L0_UROW + 3 * (1 - ox) + c];
        const int cb = w1 - ox;                        // first column right of the array
        if (cb >= 1 && cb < L0_UW) base[r * L0_UROW + 3 * cb + c] = base[r * L0_UROW + 3 * (cb - 1) + c];
    }
    l0_consumer_sync();
}

template <bool WF>
__global__ void __launch_bounds__(L0_THREADS, 3) k_blend_l0_tiled(L0Args A) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + 2 * L0_STAGE_BYTES);   // [2]
    uint64_t* empty = full + 2;                                                // [2]
    L0Round* info = reinterpret_cast<L0Round*>(sm + 2 * L0_STAGE_BYTES + 64);  // [2], 64 bytes apart
    int* s_list = reinterpret_cast<int*>(sm + 2 * L0_STAGE_BYTES + 64 + 2 * 64);
    static_assert(sizeof(L0Round) <= 64, "L0Round must fit its slot");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(&empty[0], L0_CONSUMERS / 32); mbar_init(&empty[1], L0_CONSUMERS / 32);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == L0_CONSUMERS / 32) {
        // ================================ producer warp ================================
        int uses[2] = {0, 0};
        int stage = 0;
        for (int t = blockIdx.x; t < A.ntiles; t += gridDim.x) {
            const int x0t = A.xb + (t % A.ntx) * L0_TW, y0t = (t / A.ntx) * L0_TH;
            // images whose mask has something inside this tile, in feed order
            int count = 0;
            for (int base = 0; base < A.n; base += 32) {
                bool cand = false;
                const int i = base + lane;
                if (i < A.n) {
                    const L0Img I = A.imgs[i];
                    const int lx_lo = max(x0t - I.X0, 0), lx_hi = min(x0t + L0_TW, I.X0 + I.cols) - I.X0 - 1;
                    const int ly_lo = max(y0t - I.Y0, 0), ly_hi = min(y0t + L0_TH, I.Y0 + I.rows) - I.Y0 - 1;
                    if (lx_lo <= lx_hi && ly_lo <= ly_hi) {
                        for (int cy = ly_lo / SUM_CH; cy <= ly_hi / SUM_CH; ++cy)
                            for (int cx = lx_lo / SUM_CW; cx <= lx_hi / SUM_CW; ++cx) cand = cand || (I.summary[(size_t)cy * I.sum_w + cx] & 1) != 0;
                    }
                }
                const unsigned b = __ballot_sync(0xffffffffu, cand);
                if (cand) {
                    const int pos = count + __popc(b & ((1u << lane) - 1u));
                    if (pos < L0_MAXCAND) s_list[pos] = i;
                }
                count += __popc(b);
            }
            __syncwarp();
            if (count > L0_MAXCAND) {   // more contributing images than the list holds: the whole tile goes to the general kernel
                if (lane < L0_TH / 8) {
                    const int pos = atomicAdd(A.nbands, 1);
                    A.bands[pos] = make_int2(x0t, y0t + 8 * lane);
                }
                continue;
            }
            // one round through a stage per contributing image (a tile nothing contributes to takes one round without loads)
            const int rounds = max(1, count);
            for (int r = 0; r < rounds; ++r) {
                if (uses[stage] > 0) mbar_wait(&empty[stage], (uint32_t)((uses[stage] - 1) & 1));   // the consumers are done with the stage
                if (lane == 0) {
                    unsigned char* st = sm + stage * L0_STAGE_BYTES;
                    L0Round& R = info[stage];
                    R.x0t = x0t; R.y0t = y0t; R.ncand = count > 0 ? 1 : 0;
                    int flags = r == 0 ? 1 : 0;
                    {
                        const int ox = (x0t >> 1) - 1, oy = (y0t >> 1) - 1;
                        if (ox >= 0 && oy >= 0 && ox + L0_UW <= A.uw && oy + L0_UH <= A.uh) flags |= 8;
                    }
                    if (count == 0) {
                        R.flags = flags;
                        mbar_arrive(&full[stage]);
                    } else {
                        const int idx = s_list[r];
                        const L0Img I = A.imgs[idx];
                        R.lx0 = x0t - I.X0;
                        R.ox = ((x0t - I.fx) >> 1) - 1;
                        R.oy = ((y0t - I.fy) >> 1) - 1;
                        R.w1 = I.w1; R.h1 = I.h1;
                        if (R.ox >= 0 && R.oy >= 0 && R.ox + L0_UW <= I.w1 && R.oy + L0_UH <= I.h1) flags |= 16;
                        R.flags = flags;
                        mbar_expect_tx(&full[stage], L0_UBYTES + L0_MBYTES + L0_IBYTES + L0_UBYTES);   // also publishes R (release)
                        tensormap_acquire(&A.maps[0]);
                        tma_load_2d(st, &A.maps[0], ((6 * ((x0t >> 1) - 1)) & ~15) >> 1, (y0t >> 1) - 1, &full[stage]);
                        const CUtensorMap* m = A.maps + 1 + 3 * idx;
                        tensormap_acquire(m); tensormap_acquire(m + 1); tensormap_acquire(m + 2);
                        tma_load_2d(st + L0_OFF_MASK, m, (x0t - I.X0) & ~15, y0t - I.Y0, &full[stage]);
                        tma_load_2d(st + L0_OFF_IMG, m + 1, (3 * (x0t - I.X0)) & ~15, y0t - I.Y0, &full[stage]);
                        tma_load_2d(st + L0_OFF_G1, m + 2, ((6 * R.ox) & ~15) >> 1, R.oy, &full[stage]);
                    }
                }
                uses[stage]++;
                stage ^= 1;
            }
            __syncwarp();   // lane 0 is done with s_list before the next tile's list is written
        }
        if (uses[stage] > 0) mbar_wait(&empty[stage], (uint32_t)((uses[stage] - 1) & 1));
        if (lane == 0) { info[stage].ncand = -1; mbar_arrive(&full[stage]); }
        return;
    }

    // ================================ consumer warps ================================
    // this thread's pixels: columns 2 tx, 2 tx + 1; rows 8 ty .. 8 ty + 7 as four quads; p = 4 * quad + 2 * (row in quad) + (column in quad)
    const int tx = lane, ty = warp;
    const int nbr_off = 3 * tx + L0_QN * ty * L0_UROW;        // level-1 neighbourhood of a quad column: fixed offsets from here
    int fuse[2] = {0, 0};
    int stage = 0;
    uint32_t claimed[L0_QN] = {};   // per quad: the mask bytes of the pixels an earlier round of this tile has stored
    bool dead = false;              // this warp's band is on the work list: the remaining rounds of the tile skip it
    for (;;) {
        mbar_wait(&full[stage], (uint32_t)(fuse[stage] & 1));
        fuse[stage]++;
        const L0Round& R = info[stage];
        const int ncand = R.ncand;
        if (ncand < 0) break;
        const int flags = R.flags;
        const bool first = (flags & 1) != 0;
        if (first) {
            dead = false;
#pragma unroll
            for (int q = 0; q < L0_QN; ++q) claimed[q] = 0;
        }
        const int x0t = R.x0t, y0t = R.y0t;
        unsigned char* st = sm + stage * L0_STAGE_BYTES;
        const int x = x0t + 2 * tx;
        const bool va = x >= A.sx0 && x < A.W, vb = x + 1 >= A.sx0 && x + 1 < A.W;
        int y = y0t + 2 * L0_QN * ty;
        char* orow = reinterpret_cast<char*>(A.dst) + (size_t)y * A.dstep + 6 * (ptrdiff_t)(x - A.sx0);
        uint8_t* mrow = A.dmask + (size_t)y * A.mstep + (x - A.sx0);
        if (ncand == 0) {
            // ---- nothing contributes: zeros
#pragma unroll 1
            for (int r = 0; r < 2 * L0_QN; ++r, orow += A.dstep, mrow += A.mstep) {
                if (y + r >= A.H) break;
                int16_t* o = reinterpret_cast<int16_t*>(orow);
                if (va) { o[0] = 0; o[1] = 0; o[2] = 0; mrow[0] = 0; }
                if (vb) { o[3] = 0; o[4] = 0; o[5] = 0; mrow[1] = 0; }
            }
        } else {
            // ---- one contributing image
            if ((flags & (8 | 16)) != (8 | 16)) {   // tile on the border of a level-1 array: mirror the halo in place (uniform per tile)
                if (!(flags & 8)) l0_patch_border(st, (x0t >> 1) - 1, (y0t >> 1) - 1, A.uw, A.uh, tid);
                if (!(flags & 16)) l0_patch_border(st + L0_OFF_G1, R.ox, R.oy, R.w1, R.h1, tid);
            }
            const int lx0 = R.lx0;
            const int msh = lx0 - (lx0 & ~15);               // byte shifts of the tiles inside their 16-byte aligned boxes
            const int ish = 3 * lx0 - ((3 * lx0) & ~15);
            const unsigned char* mp = st + L0_OFF_MASK + (2 * L0_QN * ty) * L0_MROW + msh + 2 * tx;
            uint32_t mw[L0_QN];                              // four mask bytes per quad
            bool binary = true;
#pragma unroll
            for (int q = 0; q < L0_QN; ++q) {
                mw[q] = (uint32_t)mp[(2 * q) * L0_MROW] | ((uint32_t)mp[(2 * q) * L0_MROW + 1] << 8) | ((uint32_t)mp[(2 * q + 1) * L0_MROW] << 16) |
                        ((uint32_t)mp[(2 * q + 1) * L0_MROW + 1] << 24);
                // every byte 0 or 255?  (b & 0x7f) == 0x7f * (b >> 7) per byte
                binary = binary && ((mw[q] & 0x7f7f7f7fu) == ((mw[q] >> 7) & 0x01010101u) * 0x7fu);
            }
            bool clash = false;                              // a pixel of this round's mask already stored by an earlier round
            uint32_t anyset = 0;
#pragma unroll
            for (int q = 0; q < L0_QN; ++q) { clash = clash || (mw[q] & claimed[q]) != 0; anyset |= mw[q]; }
            const bool ok = __all_sync(0xffffffffu, binary && !clash);
            if (dead) {
                // on the work list already
            } else if (!ok) {
                // fractional weights, or two images claim a pixel: this warp's band goes to the general kernel
                dead = true;
                if (lane == 0) {
                    const int pos = atomicAdd(A.nbands, 1);
                    A.bands[pos] = make_int2(x0t, y);
                }
            } else if (!first && __all_sync(0xffffffffu, anyset == 0)) {
                // a later round with nothing for this band
            } else {
#pragma unroll
                for (int q = 0; q < L0_QN; ++q) claimed[q] |= mw[q];
                // weight 1.0f (= 255 * (1/255.f)) where the mask is set: the accumulator is the Laplacian itself and
                // short(d / (1 + 1e-5f)) = d - sign(d).  CV_16S: weight 256, accumulator (d * 256) >> 8 = d, normalised (d << 8) / 257.
                const int oxg = R.ox, oxc = (x0t >> 1) - 1;
                const int16_t* g = reinterpret_cast<const int16_t*>(st + L0_OFF_G1) + ((6 * oxg - ((6 * oxg) & ~15)) >> 1) + nbr_off;
                const int16_t* k = reinterpret_cast<const int16_t*>(st) + ((6 * oxc - ((6 * oxc) & ~15)) >> 1) + nbr_off;
                const unsigned char* ip = st + L0_OFF_IMG + (2 * L0_QN * ty) * L0_IROW + ish + 6 * tx;
                // rolling horizontal sums of the Gaussian (G) and the collapsed (K) level-1 rows
                int GE0[3], GO0[3], GE1[3], GO1[3], KE0[3], KO0[3], KE1[3], KO1[3];
                l0_hrow(g, g + 3, g + 6, GE0, GO0);
                l0_hrow(g + L0_UROW, g + L0_UROW + 3, g + L0_UROW + 6, GE1, GO1);
                l0_hrow(k, k + 3, k + 6, KE0, KO0);
                l0_hrow(k + L0_UROW, k + L0_UROW + 3, k + L0_UROW + 6, KE1, KO1);
#pragma unroll 1
                for (int q = 0; q < L0_QN; ++q) {
                    g += L0_UROW; k += L0_UROW;
                    int GE2[3], GO2[3], KE2[3], KO2[3];
                    l0_hrow(g + L0_UROW, g + L0_UROW + 3, g + L0_UROW + 6, GE2, GO2);
                    l0_hrow(k + L0_UROW, k + L0_UROW + 3, k + L0_UROW + 6, KE2, KO2);
                    int ug[4][3], uk[4][3];
                    l0_quad_up(GE0, GO0, GE1, GO1, GE2, GO2, ug);
                    l0_quad_up(KE0, KO0, KE1, KO1, KE2, KO2, uk);
                    const uint32_t mq = mw[0];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        int v[2][3];
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const bool on = ((mq >> (16 * r + 8 * c2)) & 255u) != 0;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const int d = min((int)ip[r * L0_IROW + 3 * c2 + c] - ug[2 * r + c2][c], 32767);   // cv::subtract saturates
                                const int nrm = WF ? max(d - 1, min(d + 1, 0)) : (d * 256) / 257;                 // d - sign(d)
                                const int t = sat16(nrm + uk[2 * r + c2][c]);
                                v[c2][c] = on ? t : 0;
                            }
                        }
                        if (y + r < A.H) {
                            const uint32_t mm = (mq >> (16 * r)) & 0xffffu;   // the two mask bytes of this row (0 or 255 each)
                            int16_t* o = reinterpret_cast<int16_t*>(orow + (size_t)r * A.dstep);
                            uint8_t* mo = mrow + (size_t)r * A.mstep;
                            if (!first) {   // a later round of the tile: only the pixels of this image's mask
                                if (va && (mm & 0xffu)) { o[0] = (int16_t)v[0][0]; o[1] = (int16_t)v[0][1]; o[2] = (int16_t)v[0][2]; mo[0] = 255; }
                                if (vb && (mm >> 8)) { o[3] = (int16_t)v[1][0]; o[4] = (int16_t)v[1][1]; o[5] = (int16_t)v[1][2]; mo[1] = 255; }
                            } else {
                                if (va && vb && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
                                    uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
                                    o32[0] = (uint32_t)(uint16_t)v[0][0] | ((uint32_t)(uint16_t)v[0][1] << 16);
                                    o32[1] = (uint32_t)(uint16_t)v[0][2] | ((uint32_t)(uint16_t)v[1][0] << 16);
                                    o32[2] = (uint32_t)(uint16_t)v[1][1] | ((uint32_t)(uint16_t)v[1][2] << 16);
                                } else {
                                    if (va) { o[0] = (int16_t)v[0][0]; o[1] = (int16_t)v[0][1]; o[2] = (int16_t)v[0][2]; }
                                    if (vb) { o[3] = (int16_t)v[1][0]; o[4] = (int16_t)v[1][1]; o[5] = (int16_t)v[1][2]; }
                                }
                                if (va && vb && ((reinterpret_cast<uintptr_t>(mo) & 1) == 0)) {
                                    *reinterpret_cast<uint16_t*>(mo) = (uint16_t)mm;
                                } else {
                                    if (va) mo[0] = (uint8_t)(mm & 255u);
                                    if (vb) mo[1] = (uint8_t)(mm >> 8);
                                }
                            }
                        }
                    }
                    // roll on: next quad, next level-1 row
                    ip += 2 * L0_IROW; y += 2;
                    orow += 2 * A.dstep; mrow += 2 * A.mstep;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        GE0[c] = GE1[c]; GO0[c] = GO1[c]; GE1[c] = GE2[c]; GO1[c] = GO2[c];
                        KE0[c] = KE1[c]; KO0[c] = KO1[c]; KE1[c] = KE2[c]; KO1[c] = KO2[c];
                    }
#pragma unroll
                    for (int j = 0; j + 1 < L0_QN; ++j) mw[j] = mw[j + 1];
                }
            }
            if ((flags & (8 | 16)) != (8 | 16)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // patched bytes vs the next TMA write
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        stage ^= 1;
    }
}

// ---- pyrDown of level 0 through shared memory -------------------------------------------------------------------------
// One block = PD_TX x PD_TY outputs.  The (2 PD_TX + 3) x (2 PD_TY + 3) input tile is fetched once through the
// reflect-border view (REFLECT_101 of the padded frame, then BORDER_REFLECT into the fed image), packed as
// (b, g, r, mask) in one 32-bit word; the horizontal 5-tap sums go to shared memory, the vertical pass finishes.
constexpr int PD_TX = 32, PD_TY = 8, PD_IW = 2 * PD_TX + 3, PD_IH = 2 * PD_TY + 3;

template <bool WF, int PART>
__global__ void __launch_bounds__(PD_TX* PD_TY) k_pyrdown_l0_tiled(Level0 L, int16_t* __restrict__ g1, void* __restrict__ w1, int dh, int dw) {
    __shared__ uint32_t tile[PD_IH][PD_IW + 1];
    __shared__ int hsum[PART != PD_WEIGHT ? PD_IH : 1][PD_TX][3];
    __shared__ float hw_f[PART != PD_IMAGE ? PD_IH : 1][PD_TX];
    const int tid = threadIdx.y * PD_TX + threadIdx.x;
    const int ox0 = blockIdx.x * PD_TX, oy0 = blockIdx.y * PD_TY;
    if (PART == PD_WEIGHT && L.summary) {
        // Weights only: a tile whose whole 5 x 5 support is uniform needs no arithmetic.  All 255 -> every weight is exactly
        // 1.0f (255 * (1/255.f) = 1, row sums 16, (16 * 16) * (1/256) = 1; CV_16S: 256), all zero -> 0.  Uniformity comes from
        // the occupancy map; tiles that touch the frame border (reflection) take the general path.
        __shared__ int s_uni;
        if (tid == 0) {
            int uni = -1;                                           // -1: general path, 0 / 1: constant weight
            const int fx0 = 2 * ox0 - 2, fx1 = 2 * ox0 + 2 * PD_TX, fy0 = 2 * oy0 - 2, fy1 = 2 * oy0 + 2 * PD_TY;   // support in frame coordinates
            if (fx0 >= 0 && fy0 >= 0 && fx1 < L.width && fy1 < L.height) {
                const int mx0 = fx0 - L.left, mx1 = fx1 - L.left, my0 = fy0 - L.top, my1 = fy1 - L.top;            // ... in mask coordinates
                const int cx0 = max(mx0, 0), cx1 = min(mx1, L.cols - 1), cy0 = max(my0, 0), cy1 = min(my1, L.rows - 1);
                if (cx0 > cx1 || cy0 > cy1) {
                    uni = 0;                                        // entirely in the zero border
                } else {
                    unsigned orv = 0, andv = 3;
                    for (int cy = cy0 / SUM_CH; cy <= cy1 / SUM_CH; ++cy)
                        for (int cx = cx0 / SUM_CW; cx <= cx1 / SUM_CW; ++cx) { const unsigned v = L.summary[(size_t)cy * L.sum_w + cx]; orv |= v; andv &= v; }
                    const bool inside = mx0 >= 0 && my0 >= 0 && mx1 < L.cols && my1 < L.rows;
                    if (!(orv & 1)) uni = 0;
                    else if (inside && (andv & 2)) uni = 1;
                }
            }
            s_uni = uni;
        }
        __syncthreads();
        const int uni = s_uni;
        if (uni >= 0) {
            const int x = ox0 + threadIdx.x, y = oy0 + threadIdx.y;
            if (x < dw && y < dh) {
                const size_t o = (size_t)y * dw + x;
                if (WF) reinterpret_cast<float*>(w1)[o] = uni ? 1.0f : 0.f;
                else reinterpret_cast<int16_t*>(w1)[o] = uni ? (int16_t)256 : (int16_t)0;
            }
            return;
        }
    }
    for (int e = tid; e < PD_IH * PD_IW; e += PD_TX * PD_TY) {
        const int r = e / PD_IW, c = e % PD_IW;
        const int y = reflect101(2 * oy0 - 2 + r, L.height), x = reflect101(2 * ox0 - 2 + c, L.width);
        uint32_t word = 0;
        if (PART != PD_WEIGHT) {
            int v[3];
            l0_pixel(L, y, x, v);
            word = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16);
        }
        if (PART != PD_IMAGE) word |= (uint32_t)l0_mask(L, y, x) << 24;
        tile[r][c] = word;
    }
    __syncthreads();
    for (int e = tid; e < PD_IH * PD_TX; e += PD_TX * PD_TY) {
        const int r = e / PD_TX, ox = e % PD_TX;
        uint32_t p[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) p[b] = tile[r][2 * ox + b];
        const int kk[5] = {1, 4, 6, 4, 1};
        if (PART != PD_WEIGHT) {
            int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
            for (int b = 0; b < 5; ++b) { s0 += kk[b] * (int)(p[b] & 255); s1 += kk[b] * (int)((p[b] >> 8) & 255); s2 += kk[b] * (int)((p[b] >> 16) & 255); }
            hsum[r][ox][0] = s0; hsum[r][ox][1] = s1; hsum[r][ox][2] = s2;
        }
        if (PART != PD_IMAGE) {
            if (WF) {
                float wv[5];
#pragma unroll
                for (int b = 0; b < 5; ++b) wv[b] = __fmul_rn((float)(p[b] >> 24), (float)(1. / 255.));
                hw_f[r][ox] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wv[2], 6.f), __fmul_rn(__fadd_rn(wv[1], wv[3]), 4.f)), wv[0]), wv[4]);
            } else {
                int ws = 0;
#pragma unroll
                for (int b = 0; b < 5; ++b) { const int m = (int)(p[b] >> 24); ws += kk[b] * (m ? m + 1 : 0); }
                hw_f[r][ox] = __int_as_float(ws);
            }
        }
    }
    __syncthreads();
    const int ox = threadIdx.x, oy = threadIdx.y;
    const int x = ox0 + ox, y = oy0 + oy;
    if (x >= dw || y >= dh) return;
    const int kk[5] = {1, 4, 6, 4, 1};
    const size_t o = (size_t)y * dw + x;
    if (PART != PD_WEIGHT) {
        int acc[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            acc[0] += kk[a] * hsum[2 * oy + a][ox][0];
            acc[1] += kk[a] * hsum[2 * oy + a][ox][1];
            acc[2] += kk[a] * hsum[2 * oy + a][ox][2];
        }
        g1[3 * o] = (int16_t)sat16((acc[0] + 128) >> 8);
        g1[3 * o + 1] = (int16_t)sat16((acc[1] + 128) >> 8);
        g1[3 * o + 2] = (int16_t)sat16((acc[2] + 128) >> 8);
    }
    if (PART != PD_IMAGE) {
        if (WF) {
            const float r0 = hw_f[2 * oy][ox], r1 = hw_f[2 * oy + 1][ox], r2 = hw_f[2 * oy + 2][ox], r3 = hw_f[2 * oy + 3][ox], r4 = hw_f[2 * oy + 4][ox];
            const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2, 6.f), __fmul_rn(__fadd_rn(r1, r3), 4.f)), r0), r4);
            reinterpret_cast<float*>(w1)[o] = __fmul_rn(v, 1.f / 256.f);
        } else {
            int wacc = 0;
#pragma unroll
            for (int a = 0; a < 5; ++a) wacc += kk[a] * __float_as_int(hw_f[2 * oy + a][ox]);
            reinterpret_cast<int16_t*>(w1)[o] = (int16_t)sat16((wacc + 128) >> 8);
        }
    }
}

}  // namespace is

using namespace is;

struct FedImage {
    DevMat img, mask;              // device views (owned copies or borrowed)
    int tl_x = 0, tl_y = 0;
    int x_tl = 0, y_tl = 0;        // padded rect inside dst_roi_ (level 0)
    int top = 0, left = 0, height = 0, width = 0;
    std::vector<DevBuf> g, w;      // levels 1..nb at index k (index 0 unused)
    DevBuf summary;                // occupancy of the mask per 64 x 32 cell (k_mask_summary)
    int sum_w = 0, sum_h = 0;
    bool weights_pending = false;  // fed with the image pyramid only: weights + occupancy map are built when blend() starts
    bool upper_pending = false;    // image levels >= 2 not built yet (blender_build_upper_levels does all images of a blender at once)
    long long key = 0;             // feed order of the reference (ascending key, ties in call order)
};

struct is_blender {
    is_ctx* ctx = nullptr;
    int actual_num_bands = 5, num_bands = 5, weight_type = IS_WEIGHT_32F;
    bool prepared = false;
    is_rect roi_final{}, roi{};
    std::vector<FedImage> fed;
    is_ctx* side = nullptr;        // side stream with image pyramids in flight (deferred feeds), joined when blend() starts
    long long next_key = 0;
};

namespace is {

static Level0 level0_of(const FedImage& f) {
    Level0 L;
    L.img = f.img.data; L.istep = f.img.step; L.is_u8 = f.img.depth == IS_8U;
    L.mask = f.mask.ptr<uint8_t>(); L.mstep = f.mask.step;
    L.rows = f.img.rows; L.cols = f.img.cols;
    L.top = f.top; L.left = f.left; L.height = f.height; L.width = f.width;
    L.summary = f.summary.as<uint8_t>(); L.sum_w = f.sum_w;
    return L;
}

// The pyramid buffers of deferred feeds are written on the side stream but released into the block cache of the blender's
// context, which hands them out again on ctx->stream: that stream must first wait for the side stream (common.cuh: "only
// handed out again on the stream it was released on").
static void join_side(is_blender* b) {
    if (!b->side) return;
    if (stream_after(b->ctx, b->ctx->stream, b->side->stream) != IS_OK) cudaStreamSynchronize(b->side->stream);
    merge_child(b->ctx, b->side);
    b->side = nullptr;
}

int blender_prepare_roi(is_blender* b, is_rect dst_roi) {
    is_ctx* ctx = b->ctx;
    IS_REQUIRE(ctx, dst_roi.width > 0 && dst_roi.height > 0, IS_ERR_BAD_ARG, "empty destination ROI");
    join_side(b);          // pyramids of deferred feeds may still be in flight on the side stream: order their release after it
    b->fed.clear();
    b->roi_final = dst_roi;
    const double max_len = (double)std::max(dst_roi.width, dst_roi.height);
    b->num_bands = std::min(b->actual_num_bands, (int)std::ceil(std::log(max_len) / std::log(2.0)));
    IS_REQUIRE(ctx, b->num_bands >= 0 && b->num_bands < MAX_LEVELS, IS_ERR_BAD_ARG, "unsupported number of bands");
    const int m = 1 << b->num_bands;
    b->roi = dst_roi;
    b->roi.width += (m - dst_roi.width % m) % m;
    b->roi.height += (m - dst_roi.height % m) % m;
    b->prepared = true;
    return IS_OK;
}

// copy of a (host or device) mat into a fresh pitched device buffer
static int private_copy(is_ctx* ctx, const is_mat* m, DevMat* out) {
    IS_TRY(alloc_mat(ctx, m->rows, m->cols, m->channels, m->depth, out));
    IS_CUDA(ctx, cudaMemcpy2DAsync(out->data, out->step, m->data, m->step, out->row_bytes(), m->rows,
                                   m->device >= 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return IS_OK;
}

// per-level column ranges [xb, xe) a blend of the final-ROI columns [x0, x1) has to compute (pyrUp halo, even starts)
static void strip_ranges(const is_blender* b, int x0, int x1, std::vector<int>* xb, std::vector<int>* xe) {
    const int nb = b->num_bands;
    std::vector<int> W(nb + 1);
    W[0] = b->roi.width;
    for (int k = 1; k <= nb; ++k) W[k] = (W[k - 1] + 1) / 2;
    xb->assign(nb + 1, 0);
    xe->assign(nb + 1, 0);
    (*xb)[0] = x0 & ~1;
    (*xe)[0] = std::min(W[0], (x1 + 1) & ~1);
    for (int k = 1; k <= nb; ++k) {
        (*xb)[k] = std::max(0, ((*xb)[k - 1] >> 1) - 1) & ~1;
        (*xe)[k] = std::min(W[k], ((((*xe)[k - 1] - 1) >> 1) + 3) & ~1);
        if (W[k] & 1) (*xe)[k] = std::min(W[k], (((*xe)[k - 1] - 1) >> 1) + 2);   // only the top level can be odd
    }
}

// padded rectangle of a fed image inside dst_roi_ (MultiBandBlender::feed geometry): x_tl, y_tl, width, height, top, left
static void feed_geometry(const is_blender* b, int rows, int cols, int tl_x, int tl_y, int g[6]) {
    const int nb = b->num_bands;
    const is_rect R = b->roi;
    const int gap = 3 * (1 << nb);
    int tlx = std::max(R.x, tl_x - gap), tly = std::max(R.y, tl_y - gap);
    int brx = std::min(R.x + R.width, tl_x + cols + gap), bry = std::min(R.y + R.height, tl_y + rows + gap);
    tlx = R.x + (((tlx - R.x) >> nb) << nb);
    tly = R.y + (((tly - R.y) >> nb) << nb);
    int width = brx - tlx, height = bry - tly;
    const int m = 1 << nb;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    brx = tlx + width;
    bry = tly + height;
    const int dy = std::max(bry - (R.y + R.height), 0), dx = std::max(brx - (R.x + R.width), 0);
    tlx -= dx; tly -= dy;
    g[0] = tlx - R.x; g[1] = tly - R.y; g[2] = width; g[3] = height; g[4] = tl_y - tly; g[5] = tl_x - tlx;
}

// geometry of MultiBandBlender::feed + the (stream-ordered) allocation of the image's pyramid levels on the blender's stream
static int feed_prepare(is_blender* b, FedImage& f) {
    is_ctx* ctx = b->ctx;
    const int nb = b->num_bands;
    const is_rect R = b->roi;
    const int rows = f.img.rows, cols = f.img.cols;
    const int gap = 3 * (1 << nb);
    int tlx = std::max(R.x, f.tl_x - gap), tly = std::max(R.y, f.tl_y - gap);
    int brx = std::min(R.x + R.width, f.tl_x + cols + gap), bry = std::min(R.y + R.height, f.tl_y + rows + gap);
    tlx = R.x + (((tlx - R.x) >> nb) << nb);
    tly = R.y + (((tly - R.y) >> nb) << nb);
    int width = brx - tlx, height = bry - tly;
    const int m = 1 << nb;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    brx = tlx + width;
    bry = tly + height;
    const int dy = std::max(bry - (R.y + R.height), 0), dx = std::max(brx - (R.x + R.width), 0);
    tlx -= dx; brx -= dx;
    tly -= dy; bry -= dy;
    f.top = f.tl_y - tly;
    f.left = f.tl_x - tlx;
    f.width = width;
    f.height = height;
    f.x_tl = tlx - R.x;
    f.y_tl = tly - R.y;
    IS_REQUIRE(ctx, f.top >= 0 && f.left >= 0 && f.x_tl >= 0 && f.y_tl >= 0, IS_ERR_BAD_ARG, "image lies outside the prepared ROI");
    const size_t wsz = b->weight_type == IS_WEIGHT_32F ? sizeof(float) : sizeof(int16_t);
    f.g.resize(nb + 1);
    f.w.resize(nb + 1);
    int sh = height, sw = width;
    for (int k = 1; k <= nb; ++k) {
        const int dh = (sh + 1) / 2, dw = (sw + 1) / 2;
        IS_TRY(f.g[k].alloc(ctx, sizeof(int16_t) * 3 * (size_t)dh * dw));
        IS_TRY(f.w[k].alloc(ctx, wsz * (size_t)dh * dw));
        sh = dh; sw = dw;
    }
    if (nb >= 1) {
        f.sum_w = div_up(cols, SUM_CW); f.sum_h = div_up(rows, SUM_CH);
        IS_TRY(f.summary.alloc(ctx, (size_t)f.sum_w * f.sum_h));
    }
    return IS_OK;
}

template <bool WF, int PART>
static int feed_pyramid_t(is_ctx* sctx, const FedImage& f, int nb_all, bool level1_only) {
    const int nb = level1_only ? std::min(nb_all, 1) : nb_all;
    const size_t wsz = WF ? sizeof(float) : sizeof(int16_t);
    const double ib = PART != PD_WEIGHT ? 1. : 0., wb = PART != PD_IMAGE ? 1. : 0.;   // which bytes this launch accounts for
    int sh = f.height, sw = f.width;
    for (int k = 1; k <= nb; ++k) {
        const int dh = (sh + 1) / 2, dw = (sw + 1) / 2;
        // algorithmic bytes: level k-1 read once, level k written once
        const double in_px = k == 1 ? (double)f.img.rows * f.img.cols : (double)sh * sw;
        sctx->next_bytes = in_px * (k == 1 ? ib * (f.img.depth == IS_8U ? 3 : 6) + wb : ib * 6 + wb * (double)wsz) + (double)dh * dw * (ib * 6 + wb * (double)wsz);
        if (k == 1) {
            Level0 L = level0_of(f);
            dim3 tb(PD_TX, PD_TY), tg(div_up(dw, PD_TX), div_up(dh, PD_TY));
            IS_LAUNCH(sctx, (k_pyrdown_l0_tiled<WF, PART>), tg, tb, 0, L, f.g[1].as<int16_t>(), f.w[1].p, dh, dw);
        } else {
            dim3 block(32, 8), grid(div_up(dw, 32), div_up(dh, 8));
            IS_LAUNCH(sctx, (k_pyrdown<WF, PART>), grid, block, 0, f.g[k - 1].as<int16_t>(), f.w[k - 1].p, sh, sw, f.g[k].as<int16_t>(), f.w[k].p, dh, dw);
        }
        sh = dh; sw = dw;
    }
    return IS_OK;
}

// Gaussian pyramids (levels 1..nb) of the image and / or of its weight map, launched on sctx's stream
static int feed_pyramid(is_ctx* sctx, const is_blender* b, const FedImage& f, int part, bool level1_only = false) {
    const bool wf = b->weight_type == IS_WEIGHT_32F;
    const int nb = b->num_bands;
    switch (part) {
        case PD_BOTH: return wf ? feed_pyramid_t<true, PD_BOTH>(sctx, f, nb, level1_only) : feed_pyramid_t<false, PD_BOTH>(sctx, f, nb, level1_only);
        case PD_IMAGE: return wf ? feed_pyramid_t<true, PD_IMAGE>(sctx, f, nb, level1_only) : feed_pyramid_t<false, PD_IMAGE>(sctx, f, nb, level1_only);
        default: return wf ? feed_pyramid_t<true, PD_WEIGHT>(sctx, f, nb, level1_only) : feed_pyramid_t<false, PD_WEIGHT>(sctx, f, nb, level1_only);
    }
}

// image levels 2..nb of every image fed with the deferred path, one launch per level over all of them, on sctx's stream (which must
// be ordered after the level-1 launches: the side stream they ran on, or the blender's stream after the join)
int blender_build_upper_levels(is_blender* b, is_ctx* sctx) {
    const int nb = b->num_bands;
    std::vector<int> pend;
    for (size_t i = 0; i < b->fed.size(); ++i) if (b->fed[i].upper_pending) pend.push_back((int)i);
    if (pend.empty()) return IS_OK;
    for (int pi : pend) b->fed[(size_t)pi].upper_pending = false;
    if (nb < 2) return IS_OK;
    const int np = (int)pend.size();
    std::vector<ImageLevel> host((size_t)np * (nb - 1));
    std::vector<int> gx(nb + 1, 0), gy(nb + 1, 0);
    std::vector<double> bytes(nb + 1, 0.);
    for (int i = 0; i < np; ++i) {
        const FedImage& f = b->fed[(size_t)pend[i]];
        int sh = (f.height + 1) / 2, sw = (f.width + 1) / 2;
        for (int k = 2; k <= nb; ++k) {
            const int dh = (sh + 1) / 2, dw = (sw + 1) / 2;
            host[(size_t)(k - 2) * np + i] = ImageLevel{f.g[k - 1].as<int16_t>(), f.g[k].as<int16_t>(), sh, sw, dh, dw};
            gx[k] = std::max(gx[k], div_up(dw, PU_TX)); gy[k] = std::max(gy[k], div_up(dh, PU_TY));
            bytes[k] += ((double)sh * sw + (double)dh * dw) * 6.;
            sh = dh; sw = dw;
        }
    }
    DevBuf table;
    IS_TRY(table.alloc(sctx, sizeof(ImageLevel) * host.size()));
    IS_TRY(upload(sctx, table.p, host.data(), sizeof(ImageLevel) * host.size()));
    for (int k = 2; k <= nb; ++k) {
        dim3 grid(gx[k], gy[k], np);
        sctx->next_bytes = bytes[k];
        IS_LAUNCH(sctx, k_pyrdown_images_batch, grid, 256, 0, table.as<ImageLevel>() + (size_t)(k - 2) * np);
    }
    return IS_OK;
}

// which 64 x 32 cells of the mask hold anything: lets the level-0 blend skip images per tile
static int feed_summary(is_ctx* ctx, const FedImage& f) {
    if (!f.summary.p) return IS_OK;
    ctx->next_bytes = (double)f.img.rows * f.img.cols;
    IS_LAUNCH(ctx, k_mask_summary, div_up(f.sum_w * f.sum_h, 8), 256, 0, f.mask.ptr<uint8_t>(), f.mask.step, f.img.rows, f.img.cols, f.summary.as<uint8_t>(),
              f.sum_w, f.sum_h);
    return IS_OK;
}

int blender_feed_dev(is_blender* b, FedImage&& f) {
    if (f.key == 0) f.key = ++b->next_key; else b->next_key = std::max(b->next_key, f.key);
    IS_TRY(feed_prepare(b, f));
    IS_TRY(feed_summary(b->ctx, f));
    IS_TRY(feed_pyramid(b->ctx, b, f, PD_BOTH));
    b->fed.push_back(std::move(f));
    return IS_OK;
}

// Pipeline variant of feed(): the image pyramid is built on `side` (its stream must already be ordered after the image and
// after the blender's stream at this point); the mask is only read by blender_feed_weights().
int blender_feed_image(is_blender* b, is_ctx* side, const DevMat& img, const DevMat& mask, is_point tl) {
    FedImage f;
    f.tl_x = tl.x; f.tl_y = tl.y;
    f.img.data = img.data; f.img.rows = img.rows; f.img.cols = img.cols; f.img.channels = img.channels; f.img.depth = img.depth; f.img.step = img.step;
    f.mask.data = mask.data; f.mask.rows = mask.rows; f.mask.cols = mask.cols; f.mask.channels = 1; f.mask.depth = IS_8U; f.mask.step = mask.step;
    f.key = ++b->next_key;
    f.weights_pending = true;
    IS_TRY(feed_prepare(b, f));
    if (side != b->ctx) IS_TRY(stream_after(b->ctx, side->stream, b->ctx->stream));   // the allocations above are ordered on the blender's stream
    const int rc = feed_pyramid(side, b, f, PD_IMAGE, /*level1_only=*/true);      // levels >= 2: blender_build_upper_levels, all images at once
    if (rc != IS_OK) { if (b->ctx->last_error.empty()) b->ctx->last_error = side->last_error; return rc; }
    f.upper_pending = true;
    if (side != b->ctx) b->side = side;
    b->fed.push_back(std::move(f));
    return IS_OK;
}

// Pipeline variant of feed() fused with the warp: one kernel (launch_warp_g1, warp.cu) writes the warped image, its all-255 mask
// and level 1 of the image's Gaussian pyramid; the blender only learns the geometry and allocates the levels.  On the blender's
// stream; levels >= 2 follow with blender_build_upper_levels, the weights with blender_feed_weights.
int blender_feed_image_fused(is_blender* b, int proj, const WarpPlan& plan, const float* tables, const DevMat& src, const DevMat& img, const DevMat& mask, is_point tl) {
    is_ctx* ctx = b->ctx;
    IS_REQUIRE(ctx, b->num_bands >= 1, IS_ERR_INTERNAL, "fused feed needs at least one band");
    FedImage f;
    f.tl_x = tl.x; f.tl_y = tl.y;
    f.img.data = img.data; f.img.rows = img.rows; f.img.cols = img.cols; f.img.channels = img.channels; f.img.depth = img.depth; f.img.step = img.step;
    f.mask.data = mask.data; f.mask.rows = mask.rows; f.mask.cols = mask.cols; f.mask.channels = 1; f.mask.depth = IS_8U; f.mask.step = mask.step;
    f.key = ++b->next_key;
    f.weights_pending = true;
    f.upper_pending = true;
    IS_TRY(feed_prepare(b, f));
    IS_TRY(launch_warp_g1(ctx, proj, plan, tables, src, img, mask, f.top, f.left, f.height, f.width, f.g[1].as<int16_t>()));
    b->fed.push_back(std::move(f));
    return IS_OK;
}

// weight pyramids + occupancy maps of everything fed with blender_feed_image(), on the blender's stream
int blender_feed_weights(is_blender* b) {
    is_ctx* ctx = b->ctx;
    const int nb = b->num_bands, n = (int)b->fed.size();
    const bool wf = b->weight_type == IS_WEIGHT_32F;
    const size_t wsz = wf ? sizeof(float) : sizeof(int16_t);
    std::vector<int> pend;
    for (int i = 0; i < n; ++i) if (b->fed[i].weights_pending) pend.push_back(i);
    if (pend.empty()) return IS_OK;
    for (int pi : pend) {
        const FedImage& f = b->fed[pi];
        IS_TRY(feed_summary(ctx, f));
        if (nb >= 1) {   // level 1 from the mask (tiled kernel, constants for uniform tiles)
            const int dh = (f.height + 1) / 2, dw = (f.width + 1) / 2;
            ctx->next_bytes = (double)f.img.rows * f.img.cols + (double)dh * dw * (double)wsz;
            Level0 L = level0_of(f);
            dim3 tb(PD_TX, PD_TY), tg(div_up(dw, PD_TX), div_up(dh, PD_TY));
            if (wf) IS_LAUNCH(ctx, (k_pyrdown_l0_tiled<true, PD_WEIGHT>), tg, tb, 0, L, f.g[1].as<int16_t>(), f.w[1].p, dh, dw);
            else IS_LAUNCH(ctx, (k_pyrdown_l0_tiled<false, PD_WEIGHT>), tg, tb, 0, L, f.g[1].as<int16_t>(), f.w[1].p, dh, dw);
        }
    }
    for (int pi : pend) b->fed[pi].weights_pending = false;
    if (nb < 2) return IS_OK;
    // levels 2..nb: one launch per level over all pending images
    const int np = (int)pend.size();
    std::vector<WeightLevel> host((size_t)np * (nb - 1));
    std::vector<int> gx(nb + 1, 0), gy(nb + 1, 0);
    std::vector<double> bytes(nb + 1, 0.);
    for (int i = 0; i < np; ++i) {
        const FedImage& f = b->fed[pend[i]];
        int sh = (f.height + 1) / 2, sw = (f.width + 1) / 2;
        for (int k = 2; k <= nb; ++k) {
            const int dh = (sh + 1) / 2, dw = (sw + 1) / 2;
            host[(size_t)(k - 2) * np + i] = WeightLevel{f.w[k - 1].p, f.w[k].p, sh, sw, dh, dw};
            gx[k] = std::max(gx[k], div_up(dw, 32)); gy[k] = std::max(gy[k], div_up(dh, 8));
            bytes[k] += ((double)sh * sw + (double)dh * dw) * (double)wsz;
            sh = dh; sw = dw;
        }
    }
    DevBuf table;
    IS_TRY(table.alloc(ctx, sizeof(WeightLevel) * host.size()));
    IS_TRY(upload(ctx, table.p, host.data(), sizeof(WeightLevel) * host.size()));
    for (int k = 2; k <= nb; ++k) {
        dim3 block(32, 8), grid(gx[k], gy[k], np);
        ctx->next_bytes = bytes[k];
        const WeightLevel* t = table.as<WeightLevel>() + (size_t)(k - 2) * np;
        if (wf) IS_LAUNCH(ctx, k_pyrdown_weights_batch<true>, grid, block, 0, t);
        else IS_LAUNCH(ctx, k_pyrdown_weights_batch<false>, grid, block, 0, t);
    }
    return IS_OK;
}

static int make_map_2d(is_ctx* ctx, CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t rows, uint64_t stride_bytes,
                       uint32_t box_inner, uint32_t box_rows) {
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    // the driver entry point is looked up at run time: the library has no link-time dependency on libcuda.so.1, so it loads
    // (and exports its symbols) on a box without a GPU driver
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) return fail(ctx, IS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<encode_fn>(f);
    }
    const CUresult r = encode(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, IS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return IS_OK;
}

static bool tma_ok(const void* base, size_t stride) { return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride & 15) == 0; }

// level 0 through the tiled TMA kernel when every operand meets the TMA alignment rules (16-byte base and pitch);
// *done = false leaves the level to the generic quad kernel
static int blend_level0_tiled(is_blender* b, const LevelArgs& generic, const DevMat& dst, const DevMat& dmask, const int16_t* c1, int uh, int uw,
                              int xb, int xe, int sx0, int sx1, double bytes, bool* done) {
    is_ctx* ctx = b->ctx;
    *done = false;
    const int n = (int)b->fed.size();
    if (getenv("IS_BLEND_L0_GENERIC")) return IS_OK;
    if (b->num_bands < 1 || !tma_ok(c1, (size_t)uw * 6)) return IS_OK;
    for (const FedImage& f : b->fed) {
        if (f.img.depth != IS_8U) return IS_OK;
        if (!tma_ok(f.g[1].p, (size_t)(f.width >> 1) * 6) || !f.summary.p) return IS_OK;
    }
    // Borrowed caller buffers (e.g. dense torch tensors: pitch = 3 * cols) may miss TMA's 16-byte base / pitch rule: such an
    // image or mask is copied once into a pitched buffer (96 MB at HBM speed is ~30 us; the generic kernel would cost 1.4 ms)
    std::vector<DevMat> aligned_img((size_t)n), aligned_mask((size_t)n);
    std::vector<const void*> img_ptr((size_t)n), mask_ptr((size_t)n);
    std::vector<size_t> img_step((size_t)n), mask_step((size_t)n);
    for (int i = 0; i < n; ++i) {
        const FedImage& f = b->fed[i];
        img_ptr[i] = f.img.data; img_step[i] = f.img.step; mask_ptr[i] = f.mask.data; mask_step[i] = f.mask.step;
        if (!tma_ok(f.img.data, f.img.step)) {
            IS_TRY(alloc_mat(ctx, f.img.rows, f.img.cols, 3, IS_8U, &aligned_img[i]));
            IS_CUDA(ctx, cudaMemcpy2DAsync(aligned_img[i].data, aligned_img[i].step, f.img.data, f.img.step, (size_t)f.img.cols * 3, f.img.rows,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
            img_ptr[i] = aligned_img[i].data; img_step[i] = aligned_img[i].step;
        }
        if (!tma_ok(f.mask.data, f.mask.step)) {
            IS_TRY(alloc_mat(ctx, f.mask.rows, f.mask.cols, 1, IS_8U, &aligned_mask[i]));
            IS_CUDA(ctx, cudaMemcpy2DAsync(aligned_mask[i].data, aligned_mask[i].step, f.mask.data, f.mask.step, (size_t)f.mask.cols, f.mask.rows,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
            mask_ptr[i] = aligned_mask[i].data; mask_step[i] = aligned_mask[i].step;
        }
    }
    const size_t maps_bytes = sizeof(CUtensorMap) * (size_t)(1 + 3 * n);
    std::vector<unsigned char> host(maps_bytes + sizeof(L0Img) * (size_t)std::max(n, 1));
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(host.data());
    L0Img* imgs = reinterpret_cast<L0Img*>(host.data() + maps_bytes);
    IS_TRY(make_map_2d(ctx, &maps[0], CU_TENSOR_MAP_DATA_TYPE_UINT16, c1, (uint64_t)uw * 3, (uint64_t)uh, (uint64_t)uw * 6, L0_UROW, L0_UH));
    for (int i = 0; i < n; ++i) {
        const FedImage& f = b->fed[i];
        const int h1 = f.height >> 1, w1 = f.width >> 1;
        IS_TRY(make_map_2d(ctx, &maps[1 + 3 * i], CU_TENSOR_MAP_DATA_TYPE_UINT8, mask_ptr[i], (uint64_t)f.mask.cols, (uint64_t)f.mask.rows, mask_step[i],
                           L0_MROW, L0_TH));
        IS_TRY(make_map_2d(ctx, &maps[2 + 3 * i], CU_TENSOR_MAP_DATA_TYPE_UINT8, img_ptr[i], (uint64_t)f.img.cols * 3, (uint64_t)f.img.rows, img_step[i],
                           L0_IROW, L0_TH));
        IS_TRY(make_map_2d(ctx, &maps[3 + 3 * i], CU_TENSOR_MAP_DATA_TYPE_UINT16, f.g[1].p, (uint64_t)w1 * 3, (uint64_t)h1, (uint64_t)w1 * 6, L0_UROW, L0_UH));
        L0Img& I = imgs[i];
        I.fx = f.x_tl; I.fy = f.y_tl;
        I.X0 = f.x_tl + f.left; I.Y0 = f.y_tl + f.top;
        I.rows = f.img.rows; I.cols = f.img.cols;
        I.h1 = h1; I.w1 = w1;
        I.summary = f.summary.as<uint8_t>(); I.sum_w = f.sum_w; I.sum_h = f.sum_h;
    }
    DevBuf table;
    IS_TRY(table.alloc(ctx, host.size()));
    IS_TRY(upload(ctx, table.p, host.data(), host.size()));
    L0Args A;
    A.maps = table.as<CUtensorMap>();
    A.imgs = reinterpret_cast<const L0Img*>(table.as<unsigned char>() + maps_bytes);
    A.n = n;
    A.uh = uh; A.uw = uw;
    A.dst = dst.ptr<int16_t>(); A.dstep = dst.step; A.dmask = dmask.ptr<uint8_t>(); A.mstep = dmask.step;
    A.W = std::min(b->roi_final.width, sx1); A.H = b->roi_final.height;
    A.xb = xb; A.sx0 = sx0;
    const int gw = std::min(xe, A.W) - xb;
    if (gw <= 0) { *done = true; return IS_OK; }
    A.ntx = div_up(gw, L0_TW);
    A.ntiles = A.ntx * div_up(A.H, L0_TH);
    static int bps_table[64][2], sms_table[64];   // per device: resident CTAs per SM for the two instantiations
    const int wi = b->weight_type == IS_WEIGHT_32F ? 0 : 1;
    int* blocks_per_sm = bps_table[ctx->device & 63];
    int& sms = sms_table[ctx->device & 63];
    if (!blocks_per_sm[wi]) {
        const int dev = ctx->device;
        IS_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (wi == 0) {
            IS_CUDA(ctx, cudaFuncSetAttribute(k_blend_l0_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0_SMEM_BYTES));
            IS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[wi], k_blend_l0_tiled<true>, L0_THREADS, L0_SMEM_BYTES));
        } else {
            IS_CUDA(ctx, cudaFuncSetAttribute(k_blend_l0_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0_SMEM_BYTES));
            IS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[wi], k_blend_l0_tiled<false>, L0_THREADS, L0_SMEM_BYTES));
        }
        if (blocks_per_sm[wi] < 1) blocks_per_sm[wi] = 1;
    }
    // bands the tiled kernel leaves to the general one (device-side work list)
    DevBuf bands;
    IS_TRY(bands.alloc(ctx, sizeof(int2) * (size_t)A.ntiles * (L0_TH / 8) + 16));
    A.nbands = bands.as<int>();                                           // counter in the first 16 bytes
    A.bands = reinterpret_cast<int2*>(bands.as<unsigned char>() + 16);
    IS_CUDA(ctx, cudaMemsetAsync(bands.p, 0, 16, ctx->stream));
    const int grid = std::min(A.ntiles, sms * blocks_per_sm[wi]);   // persistent: every CTA walks tiles grid apart
    ctx->next_bytes = bytes;
    if (wi == 0) IS_LAUNCH(ctx, k_blend_l0_tiled<true>, grid, L0_THREADS, L0_SMEM_BYTES, A);
    else IS_LAUNCH(ctx, k_blend_l0_tiled<false>, grid, L0_THREADS, L0_SMEM_BYTES, A);
    const int ggrid = std::min(A.ntiles * (L0_TH / 8), sms * 8);
    if (wi == 0) IS_LAUNCH(ctx, k_blend_l0_bands<true>, ggrid, 128, 0, generic, A.bands, A.nbands);
    else IS_LAUNCH(ctx, k_blend_l0_bands<false>, ggrid, 128, 0, generic, A.bands, A.nbands);
    *done = true;
    return IS_OK;
}

int blender_blend_dev(is_blender* b, const DevMat& dst, const DevMat& dmask, int sx0, int sx1) {
    is_ctx* ctx = b->ctx;
    // deferred feeds: weights + occupancy maps from the masks as they are now, image pyramids of the side stream joined,
    // images in the reference's feed order
    IS_TRY(blender_feed_weights(b));
    {
        is_ctx* up = b->side ? b->side : ctx;
        const int rc = blender_build_upper_levels(b, up);               // no-op when the caller has done it already
        if (rc != IS_OK) { if (ctx->last_error.empty()) ctx->last_error = up->last_error; return rc; }
    }
    join_side(b);
    std::stable_sort(b->fed.begin(), b->fed.end(), [](const FedImage& a, const FedImage& c) { return a.key < c.key; });
    std::vector<int> xb, xe;
    strip_ranges(b, sx0, sx1, &xb, &xe);
    const int nb = b->num_bands, n = (int)b->fed.size();
    const bool wf = b->weight_type == IS_WEIGHT_32F;
    // panorama level dims
    std::vector<int> H(nb + 2), W(nb + 2);
    H[0] = b->roi.height; W[0] = b->roi.width;
    for (int k = 1; k <= nb; ++k) { H[k] = (H[k - 1] + 1) / 2; W[k] = (W[k - 1] + 1) / 2; }
    std::vector<DevBuf> out(nb + 1);
    for (int k = 1; k <= nb; ++k) IS_TRY(out[k].alloc(ctx, sizeof(int16_t) * 3 * (size_t)H[k] * W[k]));
    DevBuf table;
    IS_TRY(table.alloc(ctx, sizeof(ImgLevel) * (size_t)std::max(n, 1) * (nb + 1)));
    std::vector<ImgLevel> host((size_t)n * (nb + 1));
    for (int k = 0; k <= nb; ++k)
        for (int i = 0; i < n; ++i) {
            const FedImage& f = b->fed[i];
            ImgLevel& I = host[(size_t)k * n + i];
            I.l0 = level0_of(f);
            I.g = k >= 1 ? f.g[k].as<int16_t>() : nullptr;
            I.g_up = k < nb ? f.g[k + 1].as<int16_t>() : nullptr;
            I.wgt = k >= 1 ? f.w[k].p : nullptr;
            I.x_tl = f.x_tl >> k;        // x_tl /= 2 per level; exact: multiples of 2^nb
            I.y_tl = f.y_tl >> k;
            I.h = f.height >> k;
            I.w = f.width >> k;
        }
    if (n) IS_TRY(upload(ctx, table.p, host.data(), sizeof(ImgLevel) * host.size()));
    for (int k = nb; k >= 0; --k) {
        LevelArgs A;
        A.imgs = table.as<ImgLevel>() + (size_t)k * n; A.n = n;
        A.k = k; A.nb = nb;
        A.H = H[k]; A.W = W[k];
        A.out = k >= 1 ? out[k].as<int16_t>() : nullptr;
        A.up = k < nb ? out[k + 1].as<int16_t>() : nullptr;
        A.uh = k < nb ? H[k + 1] : 0; A.uw = k < nb ? W[k + 1] : 0;
        A.dst = dst.ptr<int16_t>(); A.dstep = dst.step; A.dmask = dmask.ptr<uint8_t>(); A.mstep = dmask.step;
        A.fw = b->roi_final.width; A.fh = b->roi_final.height;
        A.xb = xb[k]; A.xe = xe[k]; A.sx0 = sx0; A.sx1 = sx1;
        const int gw = std::max(0, xe[k] - xb[k]), gh = k == 0 ? A.fh : A.H;
        if (gw == 0) continue;
        dim3 block(32, 8), grid(div_up(gw, 32), div_up(gh, 8));
        {   // algorithmic bytes: every input of the level read once, the collapsed level written once
            const double wsz = wf ? 4 : 2;
            double bytes = 0;
            for (int i = 0; i < n; ++i) {
                const FedImage& f = b->fed[i];
                if (k == 0) bytes += (double)f.img.rows * f.img.cols * ((f.img.depth == IS_8U ? 3 : 6) + 1);
                else bytes += (double)(f.height >> k) * (f.width >> k) * (6 + wsz);
                if (k < nb) bytes += (double)(f.height >> (k + 1)) * (f.width >> (k + 1)) * 6;
            }
            if (k < nb) bytes += (double)H[k + 1] * W[k + 1] * 6;
            bytes += k == 0 ? (double)A.fw * A.fh * 7 : (double)H[k] * W[k] * 6;
            ctx->next_bytes = bytes * ((double)gw / (double)W[k]);   // strip: the share of the level this launch covers
        }
        if (k == 0 && nb >= 1) {
            bool done = false;
            const double bytes = ctx->next_bytes;
            IS_TRY(blend_level0_tiled(b, A, dst, dmask, out[1].as<int16_t>(), H[1], W[1], xb[0], xe[0], sx0, sx1, bytes, &done));
            if (done) continue;
            ctx->next_bytes = bytes;
        }
        if (k < nb) {   // 2x2 blocks share their pyrUp neighbourhood
            dim3 qgrid(div_up(div_up(gw, 2), 32), div_up(div_up(gh, 2), 8));
            if (wf) IS_LAUNCH(ctx, k_blend_level_quad<true>, qgrid, block, 0, A);
            else IS_LAUNCH(ctx, k_blend_level_quad<false>, qgrid, block, 0, A);
        } else {
            if (wf) IS_LAUNCH(ctx, k_blend_level<true>, grid, block, 0, A);
            else IS_LAUNCH(ctx, k_blend_level<false>, grid, block, 0, A);
        }
    }
    b->fed.clear();        // OpenCV releases the pyramids in blend()
    b->next_key = 0;
    b->prepared = false;
    return IS_OK;
}

}  // namespace is

extern "C" {

int is_blender_create(is_ctx* ctx, int num_bands, int weight_type, is_blender** out) {
    if (!ctx || !out) return IS_ERR_BAD_ARG;
    IS_REQUIRE(ctx, weight_type == IS_WEIGHT_32F || weight_type == IS_WEIGHT_16S, IS_ERR_BAD_ARG, "weight_type must be CV_32F or CV_16S");
    IS_REQUIRE(ctx, num_bands >= 0 && num_bands < MAX_LEVELS, IS_ERR_BAD_ARG, "num_bands out of range");
    is_blender* b = new is_blender();
    b->ctx = ctx;
    b->actual_num_bands = num_bands;
    b->num_bands = num_bands;
    b->weight_type = weight_type;
    *out = b;
    return IS_OK;
}

int is_blender_destroy(is_blender* b) {
    if (b) {
        cudaSetDevice(b->ctx->device);
        is::join_side(b);
        delete b;
    }
    return IS_OK;
}

int is_blender_prepare_roi(is_blender* b, is_rect dst_roi) {
    if (!b) return IS_ERR_BAD_ARG;
    return blender_prepare_roi(b, dst_roi);
}

int is_blender_prepare(is_blender* b, int n, const is_point* corners, const is_size* sizes) {   // Blender::prepare(corners, sizes) -> resultRoi
    if (!b) return IS_ERR_BAD_ARG;
    IS_REQUIRE(b->ctx, n > 0 && corners && sizes, IS_ERR_BAD_ARG, "prepare needs at least one image");
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        tlx = std::min(tlx, corners[i].x); tly = std::min(tly, corners[i].y);
        brx = std::max(brx, corners[i].x + sizes[i].width); bry = std::max(bry, corners[i].y + sizes[i].height);
    }
    return blender_prepare_roi(b, is_rect{tlx, tly, brx - tlx, bry - tly});
}

int is_blender_num_bands(const is_blender* b) { return b ? b->num_bands : IS_ERR_BAD_ARG; }

int is_blender_dst_size(const is_blender* b, is_size* size) {
    if (!b || !size || !b->prepared) return IS_ERR_BAD_ARG;
    size->width = b->roi_final.width;
    size->height = b->roi_final.height;
    return IS_OK;
}

int is_blender_feed(is_blender* b, const is_mat* img, const is_mat* mask, is_point tl, int flags) {
    if (!b) return IS_ERR_BAD_ARG;
    is_ctx* ctx = b->ctx;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, b->prepared, IS_ERR_ASSERT, "feed() before prepare()");
    IS_TRY(check_mat(ctx, img, "img"));
    IS_TRY(check_mat(ctx, mask, "mask"));
    IS_REQUIRE(ctx, img->channels == 3 && (img->depth == IS_16S || img->depth == IS_8U), IS_ERR_ASSERT, "img.type() == CV_16SC3 || img.type() == CV_8UC3");
    IS_REQUIRE(ctx, mask->depth == IS_8U && mask->channels == 1, IS_ERR_ASSERT, "mask.type() == CV_8U");
    IS_REQUIRE(ctx, mask->rows == img->rows && mask->cols == img->cols, IS_ERR_ASSERT, "mask.size() == img.size()");
    FedImage f;
    f.tl_x = tl.x; f.tl_y = tl.y;
    const bool borrow = (flags & IS_FEED_BORROW) != 0;
    if (borrow && img->device >= 0) IS_TRY(stage_in(ctx, img, &f.img)); else IS_TRY(private_copy(ctx, img, &f.img));
    if (borrow && mask->device >= 0) IS_TRY(stage_in(ctx, mask, &f.mask)); else IS_TRY(private_copy(ctx, mask, &f.mask));
    return blender_feed_dev(b, std::move(f));
}

int is_blender_feed_ex(is_blender* b, const is_mat* img, const is_mat* mask, is_point tl, int flags, long long key) {
    if (!b) return IS_ERR_BAD_ARG;
    is_ctx* ctx = b->ctx;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, b->prepared, IS_ERR_ASSERT, "feed() before prepare()");
    IS_REQUIRE(ctx, key > 0, IS_ERR_BAD_ARG, "key must be positive");
    if (!(flags & IS_FEED_DEFER_WEIGHTS)) {
        const size_t before = b->fed.size();
        IS_TRY(is_blender_feed(b, img, mask, tl, flags & IS_FEED_BORROW));
        if (b->fed.size() > before) b->fed.back().key = key;
        b->next_key = std::max(b->next_key, key);
        return IS_OK;
    }
    IS_TRY(check_mat(ctx, img, "img"));
    IS_TRY(check_mat(ctx, mask, "mask"));
    IS_REQUIRE(ctx, img->device >= 0 && mask->device >= 0, IS_ERR_BAD_ARG, "deferred feeds borrow device buffers");
    IS_REQUIRE(ctx, img->channels == 3 && (img->depth == IS_16S || img->depth == IS_8U), IS_ERR_ASSERT, "img.type() == CV_16SC3 || img.type() == CV_8UC3");
    IS_REQUIRE(ctx, mask->depth == IS_8U && mask->channels == 1 && mask->rows == img->rows && mask->cols == img->cols, IS_ERR_ASSERT,
               "mask.type() == CV_8U && mask.size() == img.size()");
    is_ctx* side = nullptr;
    IS_TRY(child_ctx(ctx, SIDE_PYRAMID, &side));
    side->ktiming = ctx->ktiming;
    DevMat im, mk;
    IS_TRY(stage_in(ctx, img, &im));
    IS_TRY(stage_in(ctx, mask, &mk));
    IS_TRY(blender_feed_image(b, side, im, mk, tl));
    b->fed.back().key = key;
    b->next_key = std::max(b->next_key, key);
    return IS_OK;
}

int is_blender_strip_needs(const is_blender* b, is_size img_size, is_point tl, int x0, int x1, int* needed) {
    if (!b || !needed || !b->prepared) return IS_ERR_BAD_ARG;
    std::vector<int> xb, xe;
    strip_ranges(b, x0, x1, &xb, &xe);
    int g[6];
    feed_geometry(b, img_size.height, img_size.width, tl.x, tl.y, g);
    *needed = 0;
    for (int k = 0; k <= b->num_bands; ++k) {
        const int lo = g[0] >> k, hi = (g[0] + g[2]) >> k;
        if (lo < xe[k] && hi > xb[k]) { *needed = 1; break; }
    }
    return IS_OK;
}

int is_blender_blend_strip(is_blender* b, int x0, int x1, is_mat* dst, is_mat* dst_mask) {
    if (!b) return IS_ERR_BAD_ARG;
    is_ctx* ctx = b->ctx;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, b->prepared, IS_ERR_ASSERT, "blend() before prepare()");
    IS_REQUIRE(ctx, 0 <= x0 && x0 < x1 && x1 <= b->roi_final.width, IS_ERR_BAD_ARG, "strip must lie inside the destination ROI");
    IS_TRY(check_mat(ctx, dst, "dst"));
    IS_TRY(check_mat(ctx, dst_mask, "dst_mask"));
    IS_REQUIRE(ctx, dst->depth == IS_16S && dst->channels == 3 && dst->rows == b->roi_final.height && dst->cols == x1 - x0, IS_ERR_BAD_ARG,
               "dst must be CV_16SC3, ROI height x strip width");
    IS_REQUIRE(ctx, dst_mask->depth == IS_8U && dst_mask->channels == 1 && dst_mask->rows == dst->rows && dst_mask->cols == dst->cols,
               IS_ERR_BAD_ARG, "dst_mask must be CV_8U of the strip size");
    DevMat d, m;
    IS_TRY(stage_out(ctx, dst, &d, false));
    IS_TRY(stage_out(ctx, dst_mask, &m, false));
    IS_TRY(blender_blend_dev(b, d, m, x0, x1));
    IS_TRY(commit(ctx, &d));
    IS_TRY(commit(ctx, &m));
    return IS_OK;
}

int is_blender_blend(is_blender* b, is_mat* dst, is_mat* dst_mask) {
    if (!b) return IS_ERR_BAD_ARG;
    is_ctx* ctx = b->ctx;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, b->prepared, IS_ERR_ASSERT, "blend() before prepare()");
    IS_TRY(check_mat(ctx, dst, "dst"));
    IS_TRY(check_mat(ctx, dst_mask, "dst_mask"));
    IS_REQUIRE(ctx, dst->depth == IS_16S && dst->channels == 3 && dst->rows == b->roi_final.height && dst->cols == b->roi_final.width,
               IS_ERR_BAD_ARG, "dst must be CV_16SC3 of is_blender_dst_size");
    IS_REQUIRE(ctx, dst_mask->depth == IS_8U && dst_mask->channels == 1 && dst_mask->rows == dst->rows && dst_mask->cols == dst->cols,
               IS_ERR_BAD_ARG, "dst_mask must be CV_8U of is_blender_dst_size");
    DevMat d, m;
    IS_TRY(stage_out(ctx, dst, &d, false));
    IS_TRY(stage_out(ctx, dst_mask, &m, false));
    IS_TRY(blender_blend_dev(b, d, m, 0, b->roi_final.width));
    IS_TRY(commit(ctx, &d));
    IS_TRY(commit(ctx, &m));
    return IS_OK;
}

}  // extern "C"
