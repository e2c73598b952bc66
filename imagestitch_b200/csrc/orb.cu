// orb.cu -- the ORB features finder of the reference on the device: gray conversion, the scale pyramid, FAST-9/16 with non-maximum
// suppression, Harris responses, intensity-centroid orientation, the pre-descriptor blur and the rBRIEF descriptors.
//
// Replaces find(InputArray image, ImageFeatures&) [FEAT]:948-1021 with detectAndCompute [FEAT]:727-946, computeKeyPoints :56-191,
// HarrisResponses :205-248, ICAngles :250-283, computeOrbDescriptors :288-418 (wta_k = 2) and the OpenCV calls they delegate to:
// cvtColor(BGR2GRAY), resize(INTER_LINEAR_EXACT), cv::FAST, KeyPointsFilter::runByImageBorder / retainBest, fastAtan2 and
// GaussianBlur on a sub-matrix (= sepFilter2D with the float kernel).  Key points (all fields, in the reference's order) and
// descriptors are bit-identical to the oracle, which is bit-identical to cv2.ORB (tests/test_oracle_orb.py).
//
// Shape of the computation: every per-pixel stage is ONE launch over all cells of the grid and all pyramid levels (a level table
// in the kernel parameters, blockIdx.z = entry); what the reference decides per key-point list -- the border filter, the two
// retainBest selections (std::nth_element + std::partition on the raster-ordered list, so ties and order come out as OpenCV's), the
// scaling of the points, cosf / sinf of the orientation (host libm, as the warp tables) -- stays on the host between four short
// round trips.  Frames around the levels ([FEAT]:776-840) are not built: no retained key point reads within 6 pixels of a level's
// edge (edgeThreshold 31 against a reach of 25).
#include "internal.cuh"
#include "hostpool.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <new>
#include <stdexcept>

namespace is {

// what the driver below needs from its surroundings (the host emulation defines the same names over plain host memory)
using OrbBuf = DevBuf;
static int orb_alloc(is_ctx* ctx, OrbBuf* b, size_t bytes) { return b->alloc(ctx, bytes); }
static int orb_h2d(is_ctx* ctx, void* dst, const void* src, size_t bytes) { return upload(ctx, dst, src, bytes); }
static int orb_d2h(is_ctx* ctx, void* dst, const void* src, size_t bytes) { return download(ctx, dst, src, bytes); }
static int orb_d2h_view(is_ctx* ctx, const void* src, size_t bytes, const void** view) { return download_view(ctx, src, bytes, view); }
static int orb_zero(is_ctx* ctx, void* dst, size_t bytes) {
    IS_CUDA(ctx, cudaMemsetAsync(dst, 0, bytes, ctx->stream));
    return IS_OK;
}
#define ORB_LAUNCH(ctx, kernel, grid, block, ...) IS_LAUNCH(ctx, kernel, grid, block, 0, __VA_ARGS__)
static void orb_parallel_for(is_ctx* ctx, size_t n, const std::function<void(size_t)>& fn) { host_pool(ctx)->run(n, fn); }
// IS_ORB_DUMP=<dir>: the intermediate buffers of a call as raw files (the host emulation writes the same files: diff them)
static int orb_dump(is_ctx* ctx, const char* name, const void* dev, size_t bytes) {
    const char* dir = getenv("IS_ORB_DUMP");
    if (!dir) return IS_OK;
    std::vector<uint8_t> h(bytes);
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    IS_CUDA(ctx, cudaMemcpy(h.data(), dev, bytes, cudaMemcpyDeviceToHost));
    const std::string path = std::string(dir) + "/" + name + ".bin";
    if (FILE* f = std::fopen(path.c_str(), "wb")) { std::fwrite(h.data(), 1, bytes, f); std::fclose(f); }
    return IS_OK;
}

// @emu-begin (tests/test_kernel_host_emulation.py compiles the marked region for the host; ORB_LAUNCH / orb_* memory helpers are
//             defined by whoever includes it)
static const int8_t kOrbPattern[1024] = {
#include "orb_pattern.inc"
};

constexpr int ORB_MAX_ENTRIES = 60;          // cells x levels
struct OrbEntry {
    int w, h;                 // level size
    int off;                  // first pixel in the pyramid-shaped buffers
    int coef;                 // resize tables of this level: [xofs(w) | xc1(w) | yofs(h) | yc1(h)]
    int x0, y0;               // level 0: the cell's origin in the gray image
};
struct OrbTable {
    int n, nlevels;
    OrbEntry e[ORB_MAX_ENTRIES];       // entry = cell * nlevels + level
};
struct OrbKp { int entry, x, y; };
struct OrbDescIn { int entry, cx, cy; float a, b; };
struct OrbConsts {
    float harris_k, scale_sq_sq;       // HarrisResponses [FEAT]:218-219
    float p1, p3, p5, p7, eps;         // cv::fastAtan2's polynomial (degrees), (float)DBL_EPSILON
    float gk[7];                       // getGaussianKernel(7, 2, CV_32F)
    int umax[18];                      // [FEAT]:85-100
    int half_patch;
};

struct KeyPt { float x, y, size, angle, response; int octave; };
struct OrbCand { float response; uint32_t yx; };      // a candidate during the selections: y << 16 | x

static inline int orb_cvround(float v) { return (int)lrintf(v); }

// resize(INTER_LINEAR_EXACT) coefficients: the source coordinate in double (OpenCV evaluates it in softdouble, i.e. IEEE), the
// fraction rounded to 8 bits; clamped offsets carry the weight 256 / 0
static void orb_linear_coeffs(int src, int dst, int* ofs, int* c1) {
    const double inv_scale = (double)dst / src;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        const double f = scale * ((double)d + 0.5) - 0.5;
        const int i = (int)std::floor(f);
        if (i >= 0 && src > 1) {
            if (i < src - 1) { ofs[d] = i; c1[d] = (int)lrint((f - (double)i) * 256.0); }
            else { ofs[d] = src - 1; c1[d] = 0; }
        } else { ofs[d] = 0; c1[d] = 0; }
    }
}

static void orb_consts(int patch, OrbConsts* c) {
    c->harris_k = 0.04f;
    float scale = 1.f / ((1 << 2) * 7 * 255.f);
    volatile float s2 = scale * scale, s3 = s2 * scale;
    c->scale_sq_sq = s3 * scale;
    const float deg = (float)(180 / 3.141592653589793238462643383279502884197169399375);
    c->p1 = 0.9997878412794807f * deg; c->p3 = -0.3258083974640975f * deg; c->p5 = 0.1555786518463281f * deg; c->p7 = -0.04432655554792128f * deg;
    c->eps = (float)DBL_EPSILON;
    {
        const double sigma = 2.0, scale2X = -0.5 / (sigma * sigma);
        double v[7], sum = 0;
        for (int i = 0; i < 7; ++i) { const double x = i - 3; v[i] = std::exp(scale2X * x * x); sum += v[i]; }
        sum = 1. / sum;
        for (int i = 0; i < 7; ++i) c->gk[i] = (float)(v[i] * sum);
    }
    const int half = patch / 2;
    c->half_patch = half;
    for (int i = 0; i < 18; ++i) c->umax[i] = 0;
    int v, v0, vmax = (int)std::floor(half * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(half * std::sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) c->umax[v] = (int)lrint(std::sqrt((double)half * half - v * v));
    for (v = half, v0 = 0; v >= vmin; --v) {
        while (c->umax[v0] == c->umax[v0 + 1]) ++v0;
        c->umax[v] = v0;
        ++v0;
    }
}

// KeyPointsFilter::retainBest: the n strongest and everything that ties with the n-th; the same std:: calls as OpenCV, on the
// same (raster) input order, so the surviving order is OpenCV's too
static void orb_retain_best(std::vector<OrbCand>& k, int n) {
    if (n < 0 || k.size() <= (size_t)n) return;
    if (n == 0) { k.clear(); return; }
    std::nth_element(k.begin(), k.begin() + n - 1, k.end(), [](const OrbCand& a, const OrbCand& b) { return a.response > b.response; });
    const float amb = k[(size_t)n - 1].response;
    auto e = std::partition(k.begin() + n, k.end(), [amb](const OrbCand& a) { return a.response >= amb; });
    k.resize((size_t)(e - k.begin()));
}

// ---- device ---------------------------------------------------------------------------------------------------------------

// cvtColor(BGR2GRAY / BGRA2GRAY), 8 bit (15-bit coefficients), written straight into level 0 of the pixel's cell
template <int CH>
__global__ void __launch_bounds__(256) k_orb_gray(const uint8_t* __restrict__ src, size_t sstep, OrbTable T, uint8_t* __restrict__ pyr) {
    const OrbEntry E = T.e[blockIdx.z * T.nlevels];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= E.w || y >= E.h) return;
    const uint8_t* s = src + (size_t)(E.y0 + y) * sstep + (size_t)(E.x0 + x) * CH;
    int g;
    if (CH == 1) g = s[0];
    else g = (s[0] * 3735 + s[1] * 19235 + s[2] * 9798 + (1 << 14)) >> 15;
    pyr[(size_t)E.off + (size_t)y * E.w + x] = (uint8_t)g;
}

// resize(prev, cur, INTER_LINEAR_EXACT): 8.8 fixed-point rows, 16.16 after the column step, one rounding
__global__ void __launch_bounds__(256) k_orb_resize(OrbTable T, int level, const int* __restrict__ coef, uint8_t* __restrict__ pyr) {
    const OrbEntry E = T.e[blockIdx.z * T.nlevels + level], S = T.e[blockIdx.z * T.nlevels + level - 1];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= E.w || y >= E.h) return;
    const int* c = coef + E.coef;
    const int xo = c[x], xc = c[E.w + x], yo = c[2 * E.w + y], yc = c[2 * E.w + E.h + y];
    const int xb = min(xo + 1, S.w - 1), yb = min(yo + 1, S.h - 1);
    const uint8_t* r0 = pyr + (size_t)S.off + (size_t)yo * S.w;
    const uint8_t* r1 = pyr + (size_t)S.off + (size_t)yb * S.w;
    const uint32_t h0 = (uint32_t)(256 - xc) * r0[xo] + (uint32_t)xc * r0[xb];
    const uint32_t h1 = (uint32_t)(256 - xc) * r1[xo] + (uint32_t)xc * r1[xb];
    const uint32_t v = (uint32_t)(256 - yc) * h0 + (uint32_t)yc * h1;
    pyr[(size_t)E.off + (size_t)y * E.w + x] = (uint8_t)min(255u, (v + (1u << 15)) >> 16);
}

// cv::FAST's corner test and cornerScore<16>: 0 for "no corner", else the largest threshold that keeps the pixel a corner
__global__ void __launch_bounds__(256) k_orb_fast(OrbTable T, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score, int threshold) {
    const OrbEntry E = T.e[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= E.w || y >= E.h) return;
    uint8_t out = 0;
    if (x >= 3 && y >= 3 && x < E.w - 3 && y < E.h - 3) {
        const uint8_t* p = pyr + (size_t)E.off + (size_t)y * E.w + x;
        const int w = E.w, v = p[0];
        int d[16];
        d[0] = v - p[3 * w]; d[1] = v - p[3 * w + 1]; d[2] = v - p[2 * w + 2]; d[3] = v - p[w + 3];
        d[4] = v - p[3]; d[5] = v - p[-w + 3]; d[6] = v - p[-2 * w + 2]; d[7] = v - p[-3 * w + 1];
        d[8] = v - p[-3 * w]; d[9] = v - p[-3 * w - 1]; d[10] = v - p[-2 * w - 2]; d[11] = v - p[-w - 3];
        d[12] = v - p[-3]; d[13] = v - p[w - 3]; d[14] = v - p[2 * w - 2]; d[15] = v - p[3 * w - 1];
        int nd[16];                                          // p - v, subtracted on its own: see the note at `best` below
        nd[0] = p[3 * w] - v; nd[1] = p[3 * w + 1] - v; nd[2] = p[2 * w + 2] - v; nd[3] = p[w + 3] - v;
        nd[4] = p[3] - v; nd[5] = p[-w + 3] - v; nd[6] = p[-2 * w + 2] - v; nd[7] = p[-3 * w + 1] - v;
        nd[8] = p[-3 * w] - v; nd[9] = p[-3 * w - 1] - v; nd[10] = p[-2 * w - 2] - v; nd[11] = p[-w - 3] - v;
        nd[12] = p[-3] - v; nd[13] = p[w - 3] - v; nd[14] = p[2 * w - 2] - v; nd[15] = p[3 * w - 1] - v;
        uint32_t dark = 0, bright = 0;                       // circle pixels darker than v - t / brighter than v + t
#pragma unroll
        for (int k = 0; k < 16; ++k) { dark |= (uint32_t)(d[k] > threshold) << k; bright |= (uint32_t)(nd[k] > threshold) << k; }
        dark |= dark << 16; bright |= bright << 16;          // nine contiguous set bits somewhere on the ring
        uint32_t a = dark & (dark >> 1); a &= a >> 2; a &= a >> 4; a &= dark >> 8;
        uint32_t b = bright & (bright >> 1); b &= b >> 2; b &= b >> 4; b &= bright >> 8;
        if ((a | b) & 0xffffu) {
            // score = max over the 16 arcs of nine of max(min d, min -d).  Written as minima over d and over nd = -d: the form
            // max(min d, -(max d)) is miscompiled by ptxas 12.9 for sm_100a -- it folds the maxima into VIMNMX3 and loses the
            // negation from the second arc on (seen on a B200: scores came out as max d; scripts/orb_debug.py found it).
            int best = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                int mn = d[k], mn2 = nd[k];
#pragma unroll
                for (int j = 1; j < 9; ++j) { mn = min(mn, d[(k + j) & 15]); mn2 = min(mn2, nd[(k + j) & 15]); }
                best = max(best, max(mn, mn2));
            }
            out = (uint8_t)(best - 1);
        }
    }
    score[(size_t)E.off + (size_t)y * E.w + x] = out;
}

// non-maximum suppression (strictly greater than the eight neighbours) + KeyPointsFilter::runByImageBorder; survivors are
// appended in any order as (y << 16 | x, entry << 8 | score): the host sorts them back into FAST's raster order
__global__ void __launch_bounds__(256) k_orb_nms(OrbTable T, const uint8_t* __restrict__ score, int border, uint32_t* __restrict__ list, unsigned* __restrict__ count, unsigned cap) {
    const OrbEntry E = T.e[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < border || y < border || x >= E.w - border || y >= E.h - border || x < 1 || y < 1 || x >= E.w - 1 || y >= E.h - 1) return;
    const uint8_t* s = score + (size_t)E.off + (size_t)y * E.w + x;
    const int c = s[0], w = E.w;
    if (c == 0) return;
    if (s[-1] >= c || s[1] >= c || s[-w - 1] >= c || s[-w] >= c || s[-w + 1] >= c || s[w - 1] >= c || s[w] >= c || s[w + 1] >= c) return;
    const unsigned at = atomicAdd(count, 1u);
    if (at < cap) { list[2 * at] = ((uint32_t)y << 16) | (uint32_t)x; list[2 * at + 1] = ((uint32_t)blockIdx.z << 8) | (uint32_t)c; }
}

__device__ __forceinline__ int orb_px(const uint8_t* __restrict__ lvl, int w, int h, int y, int x) {
    return lvl[(size_t)min(max(y, 0), h - 1) * w + min(max(x, 0), w - 1)];
}

// HarrisResponses [FEAT]:205-248: 7 x 7 block of Sobel products, integer sums, the response in the reference's float order
__global__ void __launch_bounds__(128) k_orb_harris(OrbTable T, const uint8_t* __restrict__ pyr, const OrbKp* __restrict__ kp, int n, OrbConsts C, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const OrbKp K = kp[i];
    const OrbEntry E = T.e[K.entry];
    const uint8_t* g = pyr + (size_t)E.off;
    int a = 0, b = 0, c = 0;
    for (int dy = -3; dy <= 3; ++dy)
        for (int dx = -3; dx <= 3; ++dx) {
            const int x = K.x + dx, y = K.y + dy;
            const int p00 = orb_px(g, E.w, E.h, y - 1, x - 1), p01 = orb_px(g, E.w, E.h, y - 1, x), p02 = orb_px(g, E.w, E.h, y - 1, x + 1);
            const int p10 = orb_px(g, E.w, E.h, y, x - 1), p12 = orb_px(g, E.w, E.h, y, x + 1);
            const int p20 = orb_px(g, E.w, E.h, y + 1, x - 1), p21 = orb_px(g, E.w, E.h, y + 1, x), p22 = orb_px(g, E.w, E.h, y + 1, x + 1);
            const int Ix = (p12 - p10) * 2 + (p02 - p00) + (p22 - p20);
            const int Iy = (p21 - p01) * 2 + (p20 - p00) + (p22 - p02);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    const float fa = __int2float_rn(a), fb = __int2float_rn(b), fc = __int2float_rn(c);
    const float s = __fadd_rn(fa, fb);
    const float r = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(C.harris_k, s), s));
    out[i] = __fmul_rn(r, C.scale_sq_sq);
}

// ICAngles [FEAT]:250-283: integer moments over the circular patch, cv::fastAtan2 (degrees)
__global__ void __launch_bounds__(128) k_orb_angle(OrbTable T, const uint8_t* __restrict__ pyr, const OrbKp* __restrict__ kp, int n, OrbConsts C, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const OrbKp K = kp[i];
    const OrbEntry E = T.e[K.entry];
    const uint8_t* g = pyr + (size_t)E.off;
    int m_01 = 0, m_10 = 0;
    for (int u = -C.half_patch; u <= C.half_patch; ++u) m_10 += u * orb_px(g, E.w, E.h, K.y, K.x + u);
    for (int v = 1; v <= C.half_patch; ++v) {
        int v_sum = 0;
        const int d = C.umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = orb_px(g, E.w, E.h, K.y + v, K.x + u), val_minus = orb_px(g, E.w, E.h, K.y - v, K.x + u);
            v_sum += val_plus - val_minus;
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    const float y = __int2float_rn(m_01), x = __int2float_rn(m_10);
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, C.eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(C.p7, c2), C.p5), c2), C.p3), c2), C.p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, C.eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(C.p7, c2), C.p5), c2), C.p3), c2), C.p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    out[i] = a;
}

// the pre-descriptor blur [FEAT]:921-926 = sepFilter2D with the float Gaussian: row sums left to right, every step a fused
// multiply-add as in OpenCV's AVX2 / FMA3 dispatch (what cv2 runs on current hosts; oracle/orb.cpp has the measurement) ...
__global__ void __launch_bounds__(256) k_orb_blur_rows(OrbTable T, const uint8_t* __restrict__ pyr, OrbConsts C, float* __restrict__ hbuf) {
    const OrbEntry E = T.e[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= E.w || y >= E.h) return;
    const uint8_t* row = pyr + (size_t)E.off + (size_t)y * E.w;
    float acc = __fmul_rn(C.gk[0], (float)row[max(x - 3, 0)]);
#pragma unroll
    for (int t = 1; t < 7; ++t) acc = __fmaf_rn(C.gk[t], (float)row[min(max(x - 3 + t, 0), E.w - 1)], acc);
    hbuf[(size_t)E.off + (size_t)y * E.w + x] = acc;
}

// ... columns symmetric from the centre outwards, saturate_cast<uchar> (round half to even)
__global__ void __launch_bounds__(256) k_orb_blur_cols(OrbTable T, const float* __restrict__ hbuf, OrbConsts C, uint8_t* __restrict__ blurred) {
    const OrbEntry E = T.e[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= E.w || y >= E.h) return;
    const float* col = hbuf + (size_t)E.off + x;
    float acc = __fmaf_rn(C.gk[3], col[(size_t)y * E.w], 0.f);
#pragma unroll
    for (int t = 1; t <= 3; ++t)
        acc = __fmaf_rn(C.gk[3 + t], __fadd_rn(col[(size_t)min(y + t, E.h - 1) * E.w], col[(size_t)max(y - t, 0) * E.w]), acc);
    blurred[(size_t)E.off + (size_t)y * E.w + x] = (uint8_t)min(255, max(0, __float2int_rn(acc)));
}

// computeOrbDescriptors [FEAT]:288-418, wta_k = 2: one thread per descriptor byte; a = cosf, b = sinf of the orientation come
// from the host's libm
__global__ void __launch_bounds__(256) k_orb_desc(OrbTable T, const uint8_t* __restrict__ blurred, const int8_t* __restrict__ pattern, const OrbDescIn* __restrict__ kp, int n,
                                                  uint8_t* __restrict__ desc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 5, byte = t & 31;
    if (i >= n) return;
    const OrbDescIn K = kp[i];
    const OrbEntry E = T.e[K.entry];
    const uint8_t* g = blurred + (size_t)E.off;
    const int8_t* p = pattern + 32 * byte;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int tv[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float px = (float)p[4 * k + 2 * q], py = (float)p[4 * k + 2 * q + 1];
            const float x = __fsub_rn(__fmul_rn(px, K.a), __fmul_rn(py, K.b));
            const float y = __fadd_rn(__fmul_rn(px, K.b), __fmul_rn(py, K.a));
            tv[q] = orb_px(g, E.w, E.h, K.cy + __float2int_rn(y), K.cx + __float2int_rn(x));
        }
        val |= (tv[0] < tv[1]) << k;
    }
    desc[(size_t)i * 32 + byte] = (uint8_t)val;
}

// ---- the driver: detectAndCompute over all cells at once ------------------------------------------------------------------------

struct OrbParams { int nfeatures; float scale_factor; int nlevels, grid_w, grid_h, edge_threshold, patch_size, fast_threshold; };

// src: device image (8UC1 / 8UC3 / 8UC4).  kps / desc: the reference's order (cells row-major, levels ascending).
static int orb_find_core(is_ctx* ctx, const uint8_t* src, size_t sstep, int rows, int cols, int ch, const OrbParams& P, std::vector<KeyPt>* kps, std::vector<uint8_t>* desc) {
    kps->clear(); desc->clear();
    const bool laps = getenv("IS_ORB_LAPS") != nullptr;          // phase times on stderr
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!laps) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[orb] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    const int ncells = P.grid_w * P.grid_h, nl = P.nlevels;
    if (ncells * nl > ORB_MAX_ENTRIES || nl < 1 || ncells < 1) return IS_ERR_UNSUPPORTED;
    OrbTable T;
    std::memset(&T, 0, sizeof(T));
    T.n = ncells * nl; T.nlevels = nl;
    std::vector<float> layer_scale((size_t)nl);
    for (int l = 0; l < nl; ++l) layer_scale[(size_t)l] = (float)std::pow((double)P.scale_factor, (double)l);   // getScale [FEAT]:721
    size_t total = 0, coef_total = 0;
    int maxw = 0, maxh = 0;
    for (int r = 0; r < P.grid_h; ++r)
        for (int c = 0; c < P.grid_w; ++c) {
            const int xl = c * cols / P.grid_w, yl = r * rows / P.grid_h, xr = (c + 1) * cols / P.grid_w, yr = (r + 1) * rows / P.grid_h;   // [FEAT]:991-994
            for (int l = 0; l < nl; ++l) {
                OrbEntry& E = T.e[(r * P.grid_w + c) * nl + l];
                E.w = orb_cvround((float)(xr - xl) / layer_scale[(size_t)l]);              // Size sz(cvRound(image.cols / scale), ...) [FEAT]:789
                E.h = orb_cvround((float)(yr - yl) / layer_scale[(size_t)l]);
                if (E.w < 1 || E.h < 1) return IS_ERR_BAD_ARG;
                E.x0 = xl; E.y0 = yl;
                E.off = (int)total; E.coef = (int)coef_total;
                total += (size_t)E.w * E.h;
                coef_total += 2 * (size_t)E.w + 2 * (size_t)E.h;
                if (total >= ((size_t)1 << 31)) return IS_ERR_UNSUPPORTED;
                maxw = std::max(maxw, E.w); maxh = std::max(maxh, E.h);
            }
        }
    if (maxh > 65535) return IS_ERR_UNSUPPORTED;
    OrbConsts C;
    orb_consts(P.patch_size, &C);
    std::vector<int> coef(coef_total);
    for (int cell = 0; cell < ncells; ++cell)
        for (int l = 1; l < nl; ++l) {
            const OrbEntry& E = T.e[cell * nl + l]; const OrbEntry& S = T.e[cell * nl + l - 1];
            orb_linear_coeffs(S.w, E.w, coef.data() + E.coef, coef.data() + E.coef + E.w);
            orb_linear_coeffs(S.h, E.h, coef.data() + E.coef + 2 * E.w, coef.data() + E.coef + 2 * E.w + E.h);
        }
    OrbBuf d_pyr, d_score, d_coef, d_list, d_count, d_h, d_blur, d_pat;
    const unsigned cap = (unsigned)(total / 4 + 1024);    // a corner beats its eight neighbours: at most one per 2 x 2 pixels
    IS_TRY(orb_alloc(ctx, &d_pyr, total + 64)); IS_TRY(orb_alloc(ctx, &d_score, total + 64)); IS_TRY(orb_alloc(ctx, &d_coef, coef_total * sizeof(int) + 16));
    IS_TRY(orb_alloc(ctx, &d_list, (size_t)cap * 8)); IS_TRY(orb_alloc(ctx, &d_count, 16)); IS_TRY(orb_alloc(ctx, &d_pat, 1024));
    IS_TRY(orb_h2d(ctx, d_coef.p, coef.data(), coef_total * sizeof(int)));
    IS_TRY(orb_h2d(ctx, d_pat.p, kOrbPattern, 1024));
    IS_TRY(orb_zero(ctx, d_count.p, 16));
    {   // gray level 0 of every cell, then the levels one after the other (each from the one below)
        int w0 = 0, h0 = 0;
        for (int cell = 0; cell < ncells; ++cell) { w0 = std::max(w0, T.e[cell * nl].w); h0 = std::max(h0, T.e[cell * nl].h); }
        dim3 grid((unsigned)div_up(w0, 256), (unsigned)h0, (unsigned)ncells);
        if (ch == 1) ORB_LAUNCH(ctx, k_orb_gray<1>, grid, 256, src, sstep, T, (uint8_t*)d_pyr.p);
        else if (ch == 3) ORB_LAUNCH(ctx, k_orb_gray<3>, grid, 256, src, sstep, T, (uint8_t*)d_pyr.p);
        else ORB_LAUNCH(ctx, k_orb_gray<4>, grid, 256, src, sstep, T, (uint8_t*)d_pyr.p);
        for (int l = 1; l < nl; ++l) {
            int wl = 0, hl = 0;
            for (int cell = 0; cell < ncells; ++cell) { wl = std::max(wl, T.e[cell * nl + l].w); hl = std::max(hl, T.e[cell * nl + l].h); }
            dim3 gl((unsigned)div_up(wl, 256), (unsigned)hl, (unsigned)ncells);
            ORB_LAUNCH(ctx, k_orb_resize, gl, 256, T, l, (const int*)d_coef.p, (uint8_t*)d_pyr.p);
        }
    }
    dim3 gall((unsigned)div_up(maxw, 256), (unsigned)maxh, (unsigned)T.n);
    ORB_LAUNCH(ctx, k_orb_fast, gall, 256, T, (const uint8_t*)d_pyr.p, (uint8_t*)d_score.p, P.fast_threshold);
    ORB_LAUNCH(ctx, k_orb_nms, gall, 256, T, (const uint8_t*)d_score.p, P.edge_threshold, (uint32_t*)d_list.p, (unsigned*)d_count.p, cap);
    IS_TRY(orb_dump(ctx, "pyr", d_pyr.p, total)); IS_TRY(orb_dump(ctx, "score", d_score.p, total));
    lap("tables, uploads, launches");
    unsigned found = 0;
    IS_TRY(orb_d2h(ctx, &found, d_count.p, sizeof(found)));
    lap("pyramid .. corners on the device");
    if (found > cap) return IS_ERR_NO_MEM;
    const uint32_t* list = nullptr;                          // a view of the staging buffer: valid until the next download (the responses)
    if (found) IS_TRY(orb_d2h_view(ctx, d_list.p, 2 * (size_t)found * sizeof(uint32_t), reinterpret_cast<const void**>(&list)));
    // the blur does not depend on the key points: queued behind the download, it runs while the host selects
    IS_TRY(orb_alloc(ctx, &d_h, total * sizeof(float) + 64)); IS_TRY(orb_alloc(ctx, &d_blur, total + 64));
    ORB_LAUNCH(ctx, k_orb_blur_rows, gall, 256, T, (const uint8_t*)d_pyr.p, C, (float*)d_h.p);
    ORB_LAUNCH(ctx, k_orb_blur_cols, gall, 256, T, (const float*)d_h.p, C, (uint8_t*)d_blur.p);
    IS_TRY(orb_dump(ctx, "blur", d_blur.p, total)); IS_TRY(orb_dump(ctx, "hbuf", d_h.p, total * sizeof(float)));
    lap("corner list download");
    // ---- computeKeyPoints [FEAT]:56-191 on the host: per entry the raster-ordered FAST points, retainBest(2 n)
    std::vector<int> per_level((size_t)nl);
    {
        float factor = (float)(1.0 / (double)P.scale_factor);
        float ndesired = P.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) { per_level[(size_t)l] = orb_cvround(ndesired); sum += per_level[(size_t)l]; ndesired *= factor; }
        per_level[(size_t)nl - 1] = std::max(P.nfeatures - sum, 0);
    }
    // FAST's raster order back from the unordered list: a counting sort over (entry, row) -- linear in the corners found, which a
    // textured 24 MP image has by the million --, then the few corners of a row by x.  The selection works on 8-byte records;
    // std::nth_element / std::partition permute by comparisons and positions only, so the order is the one they give cv::KeyPoints.
    std::vector<std::vector<OrbCand>> cand((size_t)T.n);
    {
        // split by entry as a parallel counting sort: every slice of the list counts, the counts are scanned slice-major per
        // entry, every slice scatters to its own places (order inside an entry does not matter yet: it is sorted below)
        constexpr size_t SLICES = 16;
        std::vector<size_t> cnt(SLICES * (size_t)T.n, 0);
        orb_parallel_for(ctx, SLICES, [&](size_t s) {
            const size_t i0 = (size_t)found * s / SLICES, i1 = (size_t)found * (s + 1) / SLICES;
            size_t* c = cnt.data() + s * (size_t)T.n;
            for (size_t i = i0; i < i1; ++i) ++c[list[2 * i + 1] >> 8];
        });
        std::vector<size_t> ebeg((size_t)T.n + 1, 0);
        for (int e = 0; e < T.n; ++e) {
            size_t at = ebeg[(size_t)e];
            for (size_t s = 0; s < SLICES; ++s) { const size_t c = cnt[s * (size_t)T.n + (size_t)e]; cnt[s * (size_t)T.n + (size_t)e] = at; at += c; }
            ebeg[(size_t)e + 1] = at;
        }
        std::vector<OrbCand> by_entry((size_t)found);
        orb_parallel_for(ctx, SLICES, [&](size_t s) {
            const size_t i0 = (size_t)found * s / SLICES, i1 = (size_t)found * (s + 1) / SLICES;
            size_t* at = cnt.data() + s * (size_t)T.n;
            for (size_t i = i0; i < i1; ++i) by_entry[at[list[2 * i + 1] >> 8]++] = OrbCand{(float)(list[2 * i + 1] & 255u), list[2 * i]};
        });
        lap("corners by entry");
        orb_parallel_for(ctx, (size_t)T.n, [&](size_t e) {       // the levels are independent of each other: one host thread each
            const OrbCand* in = by_entry.data() + ebeg[e];
            const size_t n = ebeg[e + 1] - ebeg[e];
            std::vector<uint32_t> start((size_t)T.e[e].h + 1, 0);
            for (size_t i = 0; i < n; ++i) ++start[(size_t)(in[i].yx >> 16) + 1];
            for (size_t r = 1; r < start.size(); ++r) start[r] += start[r - 1];
            std::vector<uint32_t> fill(start.begin(), start.end() - 1);
            std::vector<OrbCand>& out = cand[e];
            out.resize(n);
            for (size_t i = 0; i < n; ++i) out[fill[in[i].yx >> 16]++] = in[i];
            for (size_t r = 0; r + 1 < start.size(); ++r)
                if (start[r + 1] - start[r] > 1) std::sort(out.begin() + start[r], out.begin() + start[r + 1], [](const OrbCand& p, const OrbCand& q) { return p.yx < q.yx; });
            orb_retain_best(out, 2 * per_level[e % (size_t)nl]);
        });
    }
    std::vector<OrbKp> hk;
    for (int e = 0; e < T.n; ++e)
        for (const OrbCand& p : cand[(size_t)e]) hk.push_back(OrbKp{e, (int)(p.yx & 0xffffu), (int)(p.yx >> 16)});
    lap("raster order, retainBest(2 n)");
    if (hk.empty()) return IS_OK;
    OrbBuf d_kp, d_f;
    IS_TRY(orb_alloc(ctx, &d_kp, hk.size() * sizeof(OrbKp))); IS_TRY(orb_alloc(ctx, &d_f, hk.size() * sizeof(float)));
    IS_TRY(orb_h2d(ctx, d_kp.p, hk.data(), hk.size() * sizeof(OrbKp)));
    ORB_LAUNCH(ctx, k_orb_harris, dim3((unsigned)div_up((int)hk.size(), 128)), 128, T, (const uint8_t*)d_pyr.p, (const OrbKp*)d_kp.p, (int)hk.size(), C, (float*)d_f.p);
    std::vector<float> resp(hk.size());
    IS_TRY(orb_d2h(ctx, resp.data(), d_f.p, resp.size() * sizeof(float)));
    IS_TRY(orb_dump(ctx, "kp2n", d_kp.p, hk.size() * sizeof(OrbKp))); IS_TRY(orb_dump(ctx, "resp", d_f.p, hk.size() * sizeof(float)));
    {
        size_t at = 0;
        hk.clear();
        for (int e = 0; e < T.n; ++e) {
            for (OrbCand& p : cand[(size_t)e]) p.response = resp[at++];
            orb_retain_best(cand[(size_t)e], per_level[(size_t)(e % nl)]);
            for (const OrbCand& p : cand[(size_t)e]) hk.push_back(OrbKp{e, (int)(p.yx & 0xffffu), (int)(p.yx >> 16)});
        }
    }
    IS_TRY(orb_h2d(ctx, d_kp.p, hk.data(), hk.size() * sizeof(OrbKp)));
    ORB_LAUNCH(ctx, k_orb_angle, dim3((unsigned)div_up((int)hk.size(), 128)), 128, T, (const uint8_t*)d_pyr.p, (const OrbKp*)d_kp.p, (int)hk.size(), C, (float*)d_f.p);
    std::vector<float> ang(hk.size());
    IS_TRY(orb_d2h(ctx, ang.data(), d_f.p, ang.size() * sizeof(float)));
    IS_TRY(orb_dump(ctx, "kpn", d_kp.p, hk.size() * sizeof(OrbKp))); IS_TRY(orb_dump(ctx, "ang", d_f.p, hk.size() * sizeof(float)));
    // ---- the points in image coordinates, the descriptor inputs (host libm for cosf / sinf, [FEAT]:303-304)
    std::vector<OrbDescIn> din;
    {
        size_t at = 0;
        for (int e = 0; e < T.n; ++e) {
            const int l = e % nl;
            const float sf = layer_scale[(size_t)l];
            for (const OrbCand& c : cand[(size_t)e]) {
                KeyPt p{(float)(c.yx & 0xffffu), (float)(c.yx >> 16), 0.f, 0.f, c.response, l};
                p.angle = ang[at++];
                p.octave = l;
                p.size = P.patch_size * sf;
                p.x *= sf; p.y *= sf;                                            // allKeypoints[i].pt *= scale [FEAT]:186-189
                float scale = 1.f / sf;
                float angle = p.angle;
                angle *= (float)(3.1415926535897932384626433832795 / 180.f);
                OrbDescIn D;
                D.entry = e;
                D.cx = orb_cvround(p.x * scale); D.cy = orb_cvround(p.y * scale);
                D.a = (float)cosf(angle); D.b = (float)sinf(angle);
                din.push_back(D);
                p.x += (float)T.e[e].x0; p.y += (float)T.e[e].y0;               // kp->pt.x += xl [FEAT]:1005-1006
                kps->push_back(p);
            }
        }
    }
    OrbBuf d_din, d_desc;
    IS_TRY(orb_alloc(ctx, &d_din, din.size() * sizeof(OrbDescIn))); IS_TRY(orb_alloc(ctx, &d_desc, din.size() * 32));
    IS_TRY(orb_h2d(ctx, d_din.p, din.data(), din.size() * sizeof(OrbDescIn)));
    ORB_LAUNCH(ctx, k_orb_desc, dim3((unsigned)div_up((int)din.size() * 32, 256)), 256, T, (const uint8_t*)d_blur.p, (const int8_t*)d_pat.p, (const OrbDescIn*)d_din.p,
               (int)din.size(), (uint8_t*)d_desc.p);
    desc->resize(din.size() * 32);
    IS_TRY(orb_d2h(ctx, desc->data(), d_desc.p, desc->size()));
    lap("harris .. descriptors");
    if (laps) std::fprintf(stderr, "[orb] corners after NMS: %u, key points: %zu\n", found, kps->size());
    return IS_OK;
}
// @emu-end

}  // namespace is

using namespace is;

static int orb_find_entry(is_ctx* ctx, const is_mat* image, const is_orb_params* params, is_keypoint* keypoints, uint8_t* descriptors, int capacity, int* count) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, image, "image"));
    IS_REQUIRE(ctx, image->depth == IS_8U && (image->channels == 1 || image->channels == 3 || image->channels == 4), IS_ERR_ASSERT,
               "image must be CV_8UC1, CV_8UC3 or CV_8UC4");                       // CV_Assert [FEAT]:953
    IS_REQUIRE(ctx, count != nullptr && capacity >= 0 && (capacity == 0 || (keypoints && descriptors)), IS_ERR_BAD_ARG, "count / output buffers");
    OrbParams P{510, 1.3f, 5, 3, 1, 31, 31, 20};                                   // the reference's globals, [FEAT]:39-55
    if (params) {
        P.nfeatures = params->nfeatures; P.scale_factor = params->scale_factor; P.nlevels = params->nlevels;
        P.grid_w = params->grid_width; P.grid_h = params->grid_height;
    }
    IS_REQUIRE(ctx, P.nfeatures > 0 && P.scale_factor > 1.f && P.nlevels >= 1 && P.grid_w >= 1 && P.grid_h >= 1 && P.grid_w * P.grid_h * P.nlevels <= ORB_MAX_ENTRIES,
               IS_ERR_BAD_ARG, "ORB parameters (at most 60 cells x levels)");
    IS_REQUIRE(ctx, image->cols / P.grid_w >= 1 && image->rows / P.grid_h >= 1 && image->cols < 65536 && image->rows < 65536, IS_ERR_BAD_ARG, "image size");
    DevMat s;
    IS_TRY(stage_in(ctx, image, &s));
    std::vector<KeyPt> kps;
    std::vector<uint8_t> desc;
    ctx->last_error.clear();
    const int rc = orb_find_core(ctx, s.ptr<uint8_t>(), s.step, s.rows, s.cols, s.channels, P, &kps, &desc);
    if (rc != IS_OK) return ctx->last_error.empty() ? fail(ctx, rc, "ORB: a pyramid level of a grid cell is empty, or the pyramid exceeds 2^31 pixels") : rc;
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count = (int)kps.size();
    const int n = std::min(capacity, (int)kps.size());
    for (int i = 0; i < n; ++i) {
        keypoints[i].x = kps[(size_t)i].x; keypoints[i].y = kps[(size_t)i].y; keypoints[i].size = kps[(size_t)i].size; keypoints[i].angle = kps[(size_t)i].angle;
        keypoints[i].response = kps[(size_t)i].response; keypoints[i].octave = kps[(size_t)i].octave; keypoints[i].class_id = -1;
    }
    if (n) std::memcpy(descriptors, desc.data(), (size_t)n * 32);
    return IS_OK;
}

extern "C" {

int is_orb_find(is_ctx* ctx, const is_mat* image, const is_orb_params* params, is_keypoint* keypoints, uint8_t* descriptors, int capacity, int* count) {
    try {                                                     // no exception crosses the boundary (host vectors sized by the corner count)
        return orb_find_entry(ctx, image, params, keypoints, descriptors, capacity, count);
    } catch (const std::bad_alloc&) {
        return ctx ? fail(ctx, IS_ERR_NO_MEM, "ORB: host memory") : IS_ERR_NO_MEM;
    } catch (const std::exception& e) {
        return ctx ? fail(ctx, IS_ERR_INTERNAL, "ORB: %s", e.what()) : IS_ERR_INTERNAL;
    }
}

}  // extern "C"
