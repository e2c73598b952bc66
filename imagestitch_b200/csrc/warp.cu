// warp.cu -- cylindrical / spherical backward warp with OpenCV-exact fixed-point sampling.
//
// Replaces [WARP]:36-161 (setCameraParams, mapForward/mapBackward, detectResultRoi, buildMaps, warp) and
// the cv::remap call at [WARP]:157.  buildMaps and remap are fused: the kernel evaluates the backward map
// of a destination pixel and samples the source at once, so the two CV_32F maps (8 B per pixel, written
// and read back by the reference) never exist in HBM.
//
// Bit-exactness recipe ("hard parts"):
//   * sinf/cosf of the per-column angle u/scale (and, spherical, of the per-row angle pi - v/scale) are
//     evaluated by the host libm into O(W+H) tables -- the same libm the CPU path uses;
//   * everything per pixel is IEEE mul/add/div in the reference's association order ([WARP]:57-61) with
//     explicit round-to-nearest intrinsics, so nvcc cannot contract them into FMAs;
//   * sampling is cv::remap's 8-bit path: coordinates rounded half-even to 1/32 pixel, integer weights
//     (32-fy)(32-fx) ... (the 15-bit table of OpenCV divided by its common factor 32), BORDER_REFLECT /
//     BORDER_CONSTANT neighbour fetch, (sum + 512) >> 10.
#include "internal.cuh"
#include "hostpool.h"

#include <cmath>
#include <limits>
#include <new>

namespace is {

// @emu-begin (tests/test_kernel_host_emulation.py compiles the marked regions for the host)
struct Projector {
    float k[9], rinv[9], r_kinv[9], k_rinv[9];
};

// ---- host: camera parameters and result ROI -------------------------------------------------------

static bool invert3x3f(const float* s, float* d) {   // cv::invert on 3x3 CV_32F: double cofactors, float result
    double a = s[0], b = s[1], c = s[2], e = s[3], f = s[4], g = s[5], h = s[6], i = s[7], j = s[8];
    double det = a * (f * j - g * i) - b * (e * j - g * h) + c * (e * i - f * h);
    if (det == 0.) return false;
    double r = 1. / det;
    double t[9] = {(f * j - g * i) * r, (c * i - b * j) * r, (b * g - c * f) * r,
                   (g * h - e * j) * r, (a * j - c * h) * r, (c * e - a * g) * r,
                   (e * i - f * h) * r, (b * h - a * i) * r, (a * f - b * e) * r};
    for (int n = 0; n < 9; ++n) d[n] = (float)t[n];
    return true;
}

static void matmul3f(const float* A, const float* B, float* C) {   // float accumulation, left to right
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            volatile float acc = 0.f;   // volatile: one rounding per operation, no contraction by the host compiler
            for (int t = 0; t < 3; ++t) {
                volatile float prod = A[r * 3 + t] * B[t * 3 + c];
                acc = acc + prod;
            }
            C[r * 3 + c] = acc;
        }
}

static void set_camera(const float* K, const float* R, Projector* p) {   // [WARP]:90-120
    for (int n = 0; n < 9; ++n) p->k[n] = K[n];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) p->rinv[r * 3 + c] = R[c * 3 + r];
    float kinv[9];
    if (!invert3x3f(K, kinv)) std::memset(kinv, 0, sizeof(kinv));
    matmul3f(R, kinv, p->r_kinv);
    matmul3f(K, p->rinv, p->k_rinv);
}

static const float kPiF = static_cast<float>(3.14159265358979323846);

static inline void map_forward(int proj, const Projector& p, float scale, float x, float y, float* u, float* v) {
    volatile float x_ = p.r_kinv[0] * x + p.r_kinv[1] * y + p.r_kinv[2];
    volatile float y_ = p.r_kinv[3] * x + p.r_kinv[4] * y + p.r_kinv[5];
    volatile float z_ = p.r_kinv[6] * x + p.r_kinv[7] * y + p.r_kinv[8];
    if (proj == IS_PROJ_PLANE) {                    // cv::detail::PlaneProjector::mapForward with t = 0
        volatile float px = 0.f + x_ / z_ * (1 - 0.f), py = 0.f + y_ / z_ * (1 - 0.f);
        *u = scale * px;
        *v = scale * py;
        return;
    }
    if (proj == IS_PROJ_FISHEYE || proj == IS_PROJ_STEREOGRAPHIC) {   // FisheyeProjector / StereographicProjector::mapForward
        float u_ = atan2f(x_, z_);
        float v_ = kPiF - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        if (proj == IS_PROJ_FISHEYE) {
            volatile float r = scale * v_;
            *u = r * cosf(u_);
            *v = r * sinf(u_);
        } else {
            volatile float r = sinf(v_) / (1 - cosf(v_));
            volatile float sr = scale * r;
            *u = sr * cosf(u_);
            *v = sr * sinf(u_);
        }
        return;
    }
    *u = scale * atan2f(x_, z_);
    if (proj == IS_PROJ_CYLINDRICAL) {
        *v = scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    } else {
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = scale * (kPiF - acosf(w == w ? w : 0));
    }
}

// FisheyeProjector / StereographicProjector::mapBackward: atan2f / sinf / cosf of a per-PIXEL angle, so no O(W + H) table of
// host-libm values can carry them to the device (and the device's own libm rounds differently).  These two projectors --
// commented-out alternatives in the reference, [BLEND]:94-95 -- build their maps on the host pool with the host's libm,
// as buildMaps [WARP]:122-144 does, and the device samples through them (k_remap).
static inline void map_backward_host(int proj, const Projector& p, float scale, float u, float v, float* x, float* y) {
    u /= scale;
    v /= scale;
    float u_ = atan2f(v, u);
    volatile float rr = u * u + v * v;
    float r = sqrtf(rr);
    float v_;
    if (proj == IS_PROJ_FISHEYE) v_ = r;
    else { volatile float inv = 1.f / r; v_ = 2 * atanf(inv); }
    volatile float pv = kPiF - v_;
    float sinv = sinf(pv);
    volatile float x_ = sinv * sinf(u_);
    volatile float y_ = cosf(pv);
    volatile float z_ = sinv * cosf(u_);
    const float* m = p.k_rinv;
    volatile float a0 = m[0] * x_, a1 = m[1] * y_, a2 = m[2] * z_;
    volatile float b0 = m[3] * x_, b1 = m[4] * y_, b2 = m[5] * z_;
    volatile float c0 = m[6] * x_, c1 = m[7] * y_, c2 = m[8] * z_;
    volatile float X = a0 + a1; X = X + a2;
    volatile float Y = b0 + b1; Y = Y + b2;
    volatile float Z = c0 + c1; Z = Z + c2;
    if (Z > 0) { *x = X / Z; *y = Y / Z; }
    else { *x = -1.f; *y = -1.f; }
}

static inline bool proj_uses_maps(int proj) { return proj == IS_PROJ_FISHEYE || proj == IS_PROJ_STEREOGRAPHIC; }
static inline bool proj_known(int proj) { return proj >= IS_PROJ_CYLINDRICAL && proj <= IS_PROJ_STEREOGRAPHIC; }

// rows [y0, y1) of the two maps of a per-pixel projector (dense, w floats per row)
static void fill_map_rows(int proj, const Projector& p, float scale, int tl_x, int tl_y, int w, int y0, int y1, float* xmap, float* ymap) {
    for (int j = y0; j < y1; ++j)
        for (int i = 0; i < w; ++i)
            map_backward_host(proj, p, scale, (float)(tl_x + i), (float)(tl_y + j), xmap + (size_t)j * w + i, ymap + (size_t)j * w + i);
}

// detectResultRoi: the extrema of (u, v) over the source image lie on its border for every camera
// that does not look along the projection axis; that is the scan cv::detail::CylindricalWarper /
// SphericalWarper themselves use (detectResultRoiByBorder).  The reference's full scan
// ([WARP]:72-81) gives the same corners (tests/test_oracle_warp.py pins both).
struct RoiAcc { float tl_u, tl_v, br_u, br_v; };

// extrema over a part of the border: segment 0 / 1 = top + bottom rows [x0, x1), segment 2 / 3 = left + right columns [y0, y1)
static RoiAcc roi_segment(int proj, int w, int h, const Projector& p, float scale, bool rows, int a0, int a1) {
    RoiAcc r{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
    auto acc = [&](int x, int y) {
        float u, v;
        map_forward(proj, p, scale, (float)x, (float)y, &u, &v);
        r.tl_u = std::min(r.tl_u, u); r.tl_v = std::min(r.tl_v, v);
        r.br_u = std::max(r.br_u, u); r.br_v = std::max(r.br_v, v);
    };
    if (rows) for (int x = a0; x < a1; ++x) { acc(x, 0); acc(x, h - 1); }
    else for (int y = a0; y < a1; ++y) { acc(0, y); acc(w - 1, y); }
    return r;
}

// RotationWarperBase::detectResultRoi, the scan over ALL source pixels that FisheyeWarper / StereographicWarper inherit
// (and that [WARP]:72-81 spells out): rows [y0, y1)
static RoiAcc roi_rows(int proj, int w, const Projector& p, float scale, int y0, int y1) {
    RoiAcc r{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < w; ++x) {
            float u, v;
            map_forward(proj, p, scale, (float)x, (float)y, &u, &v);
            r.tl_u = std::min(r.tl_u, u); r.tl_v = std::min(r.tl_v, v);
            r.br_u = std::max(r.br_u, u); r.br_v = std::max(r.br_v, v);
        }
    return r;
}

// PlaneWarper::detectResultRoi: the four corners of the source
static RoiAcc roi_corners(int proj, int w, int h, const Projector& p, float scale) {
    RoiAcc r{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
    const int cx[4] = {0, 0, w - 1, w - 1}, cy[4] = {0, h - 1, 0, h - 1};
    for (int c = 0; c < 4; ++c) {
        float u, v;
        map_forward(proj, p, scale, (float)cx[c], (float)cy[c], &u, &v);
        r.tl_u = std::min(r.tl_u, u); r.tl_v = std::min(r.tl_v, v);
        r.br_u = std::max(r.br_u, u); r.br_v = std::max(r.br_v, v);
    }
    return r;
}

static void finish_roi(int proj, int w, int h, const Projector& p, float scale, float tl_u, float tl_v, float br_u, float br_v, int roi[4]) {
    if (proj == IS_PROJ_SPHERICAL) {   // SphericalWarper::detectResultRoi: poles inside the image widen the ROI
        tl_u = (float)(int)tl_u; tl_v = (float)(int)tl_v; br_u = (float)(int)br_u; br_v = (float)(int)br_v;
        for (int pole = 0; pole < 2; ++pole) {
            float s = pole ? -1.f : 1.f;
            float x = s * p.rinv[1], y = s * p.rinv[4], z = s * p.rinv[7];
            if (y > 0.f) {
                float x_ = (p.k[0] * x + p.k[1] * y) / z + p.k[2];
                float y_ = p.k[4] * y / z + p.k[5];
                if (x_ > 0.f && x_ < w && y_ > 0.f && y_ < h) {
                    float pv = pole ? 0.f : static_cast<float>(3.14159265358979323846 * scale);
                    tl_u = std::min(tl_u, 0.f); tl_v = std::min(tl_v, pv);
                    br_u = std::max(br_u, 0.f); br_v = std::max(br_v, pv);
                }
            }
        }
    }
    roi[0] = (int)tl_u; roi[1] = (int)tl_v; roi[2] = (int)br_u; roi[3] = (int)br_v;
}

// detectResultRoi: the extrema of (u, v) over the source image lie on its border for every camera
// that does not look along the projection axis; that is the scan cv::detail::CylindricalWarper /
// SphericalWarper themselves use (detectResultRoiByBorder).  The reference's full scan
// ([WARP]:72-81) gives the same corners (tests/test_oracle_warp.py pins both).  min / max are order independent, so the
// border may be scanned in pieces (detect_roi_parallel below).
static void detect_roi(int proj, int w, int h, const Projector& p, float scale, int roi[4]) {
    if (proj == IS_PROJ_PLANE || proj_uses_maps(proj)) {
        const RoiAcc a = proj == IS_PROJ_PLANE ? roi_corners(proj, w, h, p, scale) : roi_rows(proj, w, p, scale, 0, h);
        finish_roi(proj, w, h, p, scale, a.tl_u, a.tl_v, a.br_u, a.br_v, roi);
        return;
    }
    const RoiAcc a = roi_segment(proj, w, h, p, scale, true, 0, w), b = roi_segment(proj, w, h, p, scale, false, 0, h);
    finish_roi(proj, w, h, p, scale, std::min(a.tl_u, b.tl_u), std::min(a.tl_v, b.tl_v), std::max(a.br_u, b.br_u), std::max(a.br_v, b.br_v), roi);
}

// Per-column and per-row trigonometry of the backward map, host libm.  Layout: [sinu(w) | cosu(w) | rowA(h) | rowB(h)]
//   cylindrical: rowA = v / scale (y_), rowB unused
//   spherical:   rowA = sinf(pi - v/scale), rowB = cosf(pi - v/scale)
//   plane:       "sinu" = u / scale - t[0], "cosu" = 1 - t[2], rowA = v / scale - t[1] with t = 0 (PlaneProjector::mapBackward: the
//                point (u', v', 1 - t[2]) takes the place of the cylinder's (sin u, v', cos u))
static void fill_tables(int proj, float scale, int tl_x, int tl_y, int w, int h, float* t) {
    float* sinu = t; float* cosu = t + w; float* rowA = t + 2 * (size_t)w; float* rowB = rowA + h;
    for (int i = 0; i < w; ++i) {
        float u = (float)(tl_x + i) / scale;
        if (proj == IS_PROJ_PLANE) { sinu[i] = u - 0.f; cosu[i] = 1 - 0.f; continue; }
        sinu[i] = sinf(u);
        cosu[i] = cosf(u);
    }
    for (int j = 0; j < h; ++j) {
        float v = (float)(tl_y + j) / scale;
        if (proj == IS_PROJ_CYLINDRICAL) { rowA[j] = v; rowB[j] = 0.f; }
        else if (proj == IS_PROJ_PLANE) { rowA[j] = v - 0.f; rowB[j] = 0.f; }
        else { rowA[j] = sinf(kPiF - v); rowB[j] = cosf(kPiF - v); }
    }
}

// ---- device ------------------------------------------------------------------------------------------

template <int PROJ>
__device__ __forceinline__ void map_backward(const WarpParams& P, float su, float cu, float ra, float rb, float* x, float* y) {
    float x_, y_, z_;
    if (PROJ == IS_PROJ_CYLINDRICAL || PROJ == IS_PROJ_PLANE) { x_ = su; y_ = ra; z_ = cu; }
    else { x_ = __fmul_rn(ra, su); y_ = rb; z_ = __fmul_rn(ra, cu); }
    const float* m = P.k_rinv;
    float X = __fadd_rn(__fadd_rn(__fmul_rn(m[0], x_), __fmul_rn(m[1], y_)), __fmul_rn(m[2], z_));
    float Y = __fadd_rn(__fadd_rn(__fmul_rn(m[3], x_), __fmul_rn(m[4], y_)), __fmul_rn(m[5], z_));
    float Z = __fadd_rn(__fadd_rn(__fmul_rn(m[6], x_), __fmul_rn(m[7], y_)), __fmul_rn(m[8], z_));
    if (PROJ == IS_PROJ_PLANE || Z > 0.f) { *x = __fdiv_rn(X, Z); *y = __fdiv_rn(Y, Z); }   // the plane projector has no z > 0 guard
    else { *x = -1.f; *y = -1.f; }
}

__device__ __forceinline__ int reflect_idx(int p, int len) {   // cv::borderInterpolate, BORDER_REFLECT
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

__device__ __forceinline__ int clamp_short(int v) { return max(-32768, min(32767, v)); }

// cv::cvRound on the CPUs the reference runs on (cvtss2si / cvtps2dq): round half even and INT_MIN -- the "integer indefinite"
// value -- for NaN and for everything outside the int range, where cvt.rni.s32.f32 saturates and turns NaN into 0.  Only maps
// that are not bounded by construction need it (caller-supplied maps, the plane projector's unguarded division).
template <bool SAFE>
__device__ __forceinline__ int cv_round(float v) {
    if (SAFE && !(v >= -2147483648.f && v < 2147483648.f)) return -2147483647 - 1;
    return __float2int_rn(v);
}

// One destination pixel of cv::remap (8-bit).  out[c], c < CH.
template <int CH, int INTERP, int BORDER, bool SAFE = false>
__device__ __forceinline__ void sample(const uint8_t* __restrict__ src, size_t sstep, int sw, int sh, float x, float y, int* out) {
    if (INTERP == IS_INTER_NEAREST) {
        int sx = clamp_short(cv_round<SAFE>(x)), sy = clamp_short(cv_round<SAFE>(y));
        bool in = (unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh;
        if (!in) {
            if (BORDER == IS_BORDER_CONSTANT) {
#pragma unroll
                for (int c = 0; c < CH; ++c) out[c] = 0;
                return;
            }
            sx = reflect_idx(sx, sw);
            sy = reflect_idx(sy, sh);
        }
        const uint8_t* s = src + (size_t)sy * sstep + sx * CH;
#pragma unroll
        for (int c = 0; c < CH; ++c) out[c] = __ldg(s + c);
        return;
    }
    int ix = cv_round<SAFE>(__fmul_rn(x, 32.f)), iy = cv_round<SAFE>(__fmul_rn(y, 32.f));
    int sx = clamp_short(ix >> 5), sy = clamp_short(iy >> 5);
    int fx = ix & 31, fy = iy & 31;
    int w00 = (32 - fy) * (32 - fx), w01 = (32 - fy) * fx, w10 = fy * (32 - fx), w11 = fy * fx;
    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {   // all four neighbours inside
        const uint8_t* s0 = src + (size_t)sy * sstep + sx * CH;
        const uint8_t* s1 = s0 + sstep;
#pragma unroll
        for (int c = 0; c < CH; ++c)
            out[c] = (__ldg(s0 + c) * w00 + __ldg(s0 + CH + c) * w01 + __ldg(s1 + c) * w10 + __ldg(s1 + CH + c) * w11 + 512) >> 10;
        return;
    }
    if (BORDER == IS_BORDER_CONSTANT) {
        bool x0 = (unsigned)sx < (unsigned)sw, x1 = (unsigned)(sx + 1) < (unsigned)sw;
        bool y0 = (unsigned)sy < (unsigned)sh, y1 = (unsigned)(sy + 1) < (unsigned)sh;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            int v0 = (x0 && y0) ? __ldg(src + (size_t)sy * sstep + sx * CH + c) : 0;
            int v1 = (x1 && y0) ? __ldg(src + (size_t)sy * sstep + (sx + 1) * CH + c) : 0;
            int v2 = (x0 && y1) ? __ldg(src + (size_t)(sy + 1) * sstep + sx * CH + c) : 0;
            int v3 = (x1 && y1) ? __ldg(src + (size_t)(sy + 1) * sstep + (sx + 1) * CH + c) : 0;
            out[c] = (v0 * w00 + v1 * w01 + v2 * w10 + v3 * w11 + 512) >> 10;
        }
        return;
    }
    int sx0 = reflect_idx(sx, sw), sx1 = reflect_idx(sx + 1, sw);
    int sy0 = reflect_idx(sy, sh), sy1 = reflect_idx(sy + 1, sh);
    const uint8_t* r0 = src + (size_t)sy0 * sstep;
    const uint8_t* r1 = src + (size_t)sy1 * sstep;
#pragma unroll
    for (int c = 0; c < CH; ++c)
        out[c] = (__ldg(r0 + sx0 * CH + c) * w00 + __ldg(r0 + sx1 * CH + c) * w01 + __ldg(r1 + sx0 * CH + c) * w10 +
                  __ldg(r1 + sx1 * CH + c) * w11 + 512) >> 10;
}

constexpr int WARP_PX = 4;          // destination pixels per thread (consecutive in x)
constexpr int WARP_BX = 32, WARP_BY = 8;

// tables: [sinu(w) | cosu(w) | rowA(h) | rowB(h)]
template <int PROJ, int CH, int INTERP, int BORDER, bool WITH_MASK>
__global__ void __launch_bounds__(WARP_BX* WARP_BY)
k_warp(WarpParams P, const float* __restrict__ tables, const uint8_t* __restrict__ src, size_t sstep,
       uint8_t* __restrict__ dst, size_t dstep, uint8_t* __restrict__ mask, size_t mstep) {
    const int x0 = (blockIdx.x * WARP_BX + threadIdx.x) * WARP_PX;
    const int y = blockIdx.y * WARP_BY + threadIdx.y;
    if (y >= P.dst_h || x0 >= P.dst_w) return;
    const float* sinu = tables;
    const float* cosu = tables + P.dst_w;
    const float ra = __ldg(tables + 2 * (size_t)P.dst_w + y);
    const float rb = __ldg(tables + 2 * (size_t)P.dst_w + P.dst_h + y);
    int px[WARP_PX][CH];
    int mk[WARP_PX];
#pragma unroll
    for (int i = 0; i < WARP_PX; ++i) {
        int x = min(x0 + i, P.dst_w - 1);
        float sx, sy;
        map_backward<PROJ>(P, __ldg(sinu + x), __ldg(cosu + x), ra, rb, &sx, &sy);
        sample<CH, INTERP, BORDER, PROJ == IS_PROJ_PLANE>(src, sstep, P.src_w, P.src_h, sx, sy, px[i]);
        if (WITH_MASK) {   // INTER_NEAREST + BORDER_CONSTANT on an all-255 source mask
            int nx = clamp_short(cv_round<PROJ == IS_PROJ_PLANE>(sx)), ny = clamp_short(cv_round<PROJ == IS_PROJ_PLANE>(sy));
            mk[i] = ((unsigned)nx < (unsigned)P.src_w && (unsigned)ny < (unsigned)P.src_h) ? 255 : 0;
        }
    }
    uint8_t* d = dst + (size_t)y * dstep + (size_t)x0 * CH;
    const bool full = x0 + WARP_PX <= P.dst_w;
    if (full && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
        if (CH == 3) {
            uint32_t w0 = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
            uint32_t w1 = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
            uint32_t w2 = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
            uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
            d32[0] = w0; d32[1] = w1; d32[2] = w2;
        } else {
            *reinterpret_cast<uint32_t*>(d) = px[0][0] | (px[1][0] << 8) | (px[2][0] << 16) | (px[3][0] << 24);
        }
    } else {
#pragma unroll
        for (int i = 0; i < WARP_PX; ++i)
            if (x0 + i < P.dst_w)
#pragma unroll
                for (int c = 0; c < CH; ++c) d[i * CH + c] = (uint8_t)px[i][c];
    }
    if (WITH_MASK) {
        uint8_t* m = mask + (size_t)y * mstep + x0;
        if (full && ((reinterpret_cast<uintptr_t>(m) & 3) == 0)) {
            *reinterpret_cast<uint32_t*>(m) = mk[0] | (mk[1] << 8) | (mk[2] << 16) | (mk[3] << 24);
        } else {
#pragma unroll
            for (int i = 0; i < WARP_PX; ++i)
                if (x0 + i < P.dst_w) m[i] = (uint8_t)mk[i];
        }
    }
}

template <int PROJ>
__global__ void k_build_maps(WarpParams P, const float* __restrict__ tables, float* __restrict__ xmap, size_t xstep,
                             float* __restrict__ ymap, size_t ystep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.dst_w || y >= P.dst_h) return;
    float sx, sy;
    map_backward<PROJ>(P, __ldg(tables + x), __ldg(tables + P.dst_w + x), __ldg(tables + 2 * (size_t)P.dst_w + y),
                       __ldg(tables + 2 * (size_t)P.dst_w + P.dst_h + y), &sx, &sy);
    reinterpret_cast<float*>(reinterpret_cast<char*>(xmap) + (size_t)y * xstep)[x] = sx;
    reinterpret_cast<float*>(reinterpret_cast<char*>(ymap) + (size_t)y * ystep)[x] = sy;
}

// cv::remap [WARP]:157 through caller-supplied (or host-built: fisheye / stereographic) CV_32F maps: the sampler of k_warp
// behind two map loads; WITH_MASK adds the all-255 mask's INTER_NEAREST + BORDER_CONSTANT warp through the same maps.
template <int CH, int INTERP, int BORDER, bool WITH_MASK>
__global__ void __launch_bounds__(WARP_BX* WARP_BY)
k_remap(int dst_w, int dst_h, int src_w, int src_h, const float* __restrict__ xmap, size_t xstep, const float* __restrict__ ymap, size_t ystep,
        const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst, size_t dstep, uint8_t* __restrict__ mask, size_t mstep) {
    const int x0 = (blockIdx.x * WARP_BX + threadIdx.x) * WARP_PX;
    const int y = blockIdx.y * WARP_BY + threadIdx.y;
    if (y >= dst_h || x0 >= dst_w) return;
    const float* mx = reinterpret_cast<const float*>(reinterpret_cast<const char*>(xmap) + (size_t)y * xstep);
    const float* my = reinterpret_cast<const float*>(reinterpret_cast<const char*>(ymap) + (size_t)y * ystep);
    uint8_t* d = dst + (size_t)y * dstep + (size_t)x0 * CH;
#pragma unroll
    for (int i = 0; i < WARP_PX; ++i) {
        if (x0 + i >= dst_w) break;
        const float sx = __ldg(mx + x0 + i), sy = __ldg(my + x0 + i);
        int px[CH];
        sample<CH, INTERP, BORDER, true>(src, sstep, src_w, src_h, sx, sy, px);
#pragma unroll
        for (int c = 0; c < CH; ++c) d[i * CH + c] = (uint8_t)px[c];
        if (WITH_MASK) {
            const int nx = clamp_short(cv_round<true>(sx)), ny = clamp_short(cv_round<true>(sy));
            mask[(size_t)y * mstep + x0 + i] = ((unsigned)nx < (unsigned)src_w && (unsigned)ny < (unsigned)src_h) ? 255 : 0;
        }
    }
}

// @emu-end
// ---- fused: warp + all-255 mask + level 1 of the image's Gaussian pyramid ------------------------------------------------
// What MultiBandBlender::feed does first with a warped image is copyMakeBorder(BORDER_REFLECT) into a frame padded to a
// multiple of 2^bands, then pyrDown.  A block of this kernel owns a 64 x 32 rectangle of that frame: it evaluates the warp for
// the 67 x 35 frame pixels under its 32 x 16 level-1 outputs (the 5 x 5 support reaches two pixels beyond; frame pixels in the
// padding are reflections of image pixels, i.e. the warp evaluated at the reflected image coordinates), keeps them in shared
// memory, writes the image pixels it owns (u8 x 3 + mask, four pixels = three 32-bit words per thread) and finishes the
// separable [1 4 6 4 1] reduction from shared memory.  The warped image is therefore never read back for its first pyramid level
// (489 MB per C2 step in round 1), and the 8 % of a frame that is padding costs warp arithmetic instead of a second kernel.
// Source pixels: the 2 x 2 neighbourhood is six consecutive bytes per row, fetched as aligned 32-bit words (three per row at
// most) instead of twelve byte loads.
// @emu-g1-begin (tests/test_kernel_host_emulation.py runs the fused kernel block by block on the multi-threaded host emulator)
struct WarpG1Args {
    WarpParams P;
    const float* tables;
    const uint8_t* src; size_t sstep;
    uint8_t* dst; size_t dstep; uint8_t* mask; size_t mstep;
    int top, left, height, width;          // padded frame; the image sits at (left, top) inside it
    int16_t* g1; int dh, dw;
    int wide_ok;                           // source base and pitch are 4-byte aligned
};

__device__ __forceinline__ int reflect101_i(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// cv::remap INTER_LINEAR + BORDER_REFLECT on 8UC3: same integers as sample<3, LINEAR, REFLECT>; returns b | g << 8 | r << 16
template <bool SAFE = false>
__device__ __forceinline__ uint32_t sample3_packed(const uint8_t* __restrict__ src, size_t sstep, int sw, int sh, float x, float y, bool wide_ok) {
    const int ix = cv_round<SAFE>(__fmul_rn(x, 32.f)), iy = cv_round<SAFE>(__fmul_rn(y, 32.f));
    const int sx = clamp_short(ix >> 5), sy = clamp_short(iy >> 5);
    const int fx = ix & 31, fy = iy & 31;
    const int w00 = (32 - fy) * (32 - fx), w01 = (32 - fy) * fx, w10 = fy * (32 - fx), w11 = fy * fx;
    int out[3];
    if (wide_ok && sx >= 0 && sx < sw - 3 && (unsigned)sy < (unsigned)(sh - 1)) {
        const int o = 3 * sx, a = o & ~3, sh8 = 8 * (o & 3);
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(src + (size_t)sy * sstep + a);
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(src + (size_t)(sy + 1) * sstep + a);
        const uint32_t a0 = __ldg(r0), a1 = __ldg(r0 + 1), a2 = __ldg(r0 + 2);
        const uint32_t b0 = __ldg(r1), b1 = __ldg(r1 + 1), b2 = __ldg(r1 + 2);
        const uint32_t t0 = __funnelshift_r(a0, a1, sh8), t1 = __funnelshift_r(a1, a2, sh8);   // bytes 0..3 and 4..7 of the top row's window
        const uint32_t u0 = __funnelshift_r(b0, b1, sh8), u1 = __funnelshift_r(b1, b2, sh8);
        // window bytes: [0..2] left pixel, [3..5] right pixel
        const int tl[3] = {(int)(t0 & 255u), (int)((t0 >> 8) & 255u), (int)((t0 >> 16) & 255u)};
        const int tr[3] = {(int)(t0 >> 24), (int)(t1 & 255u), (int)((t1 >> 8) & 255u)};
        const int bl[3] = {(int)(u0 & 255u), (int)((u0 >> 8) & 255u), (int)((u0 >> 16) & 255u)};
        const int br[3] = {(int)(u0 >> 24), (int)(u1 & 255u), (int)((u1 >> 8) & 255u)};
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = (tl[c] * w00 + tr[c] * w01 + bl[c] * w10 + br[c] * w11 + 512) >> 10;
    } else if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        const uint8_t* s0 = src + (size_t)sy * sstep + sx * 3;
        const uint8_t* s1 = s0 + sstep;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = (__ldg(s0 + c) * w00 + __ldg(s0 + 3 + c) * w01 + __ldg(s1 + c) * w10 + __ldg(s1 + 3 + c) * w11 + 512) >> 10;
    } else {
        const int sx0 = reflect_idx(sx, sw), sx1 = reflect_idx(sx + 1, sw);
        const int sy0 = reflect_idx(sy, sh), sy1 = reflect_idx(sy + 1, sh);
        const uint8_t* r0 = src + (size_t)sy0 * sstep;
        const uint8_t* r1 = src + (size_t)sy1 * sstep;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[c] = (__ldg(r0 + sx0 * 3 + c) * w00 + __ldg(r0 + sx1 * 3 + c) * w01 + __ldg(r1 + sx0 * 3 + c) * w10 + __ldg(r1 + sx1 * 3 + c) * w11 + 512) >> 10;
    }
    return (uint32_t)out[0] | ((uint32_t)out[1] << 8) | ((uint32_t)out[2] << 16);
}

// cv::remap INTER_LINEAR on 8UC3 for a sample whose 2 x 2 neighbourhood lies inside the source, with the interpolation split into
// its horizontal and vertical step: top = tl (32 - fx) + tr fx, bottom likewise, out = (top (32 - fy) + bottom fy + 512) >> 10 --
// the same integer as sum(w p) with w = (32 - fy)(32 - fx) ...  A channel's two taps of a row are bytes c and c + 3 of the
// six-byte window: one funnel shift + one four-way byte dot product (weights 32 - fx, 0, 0, fx) per channel and row.
__device__ __forceinline__ uint32_t sample3_dp_taps(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1, uint32_t b2, int sh8, int fx, int fy) {
    const uint32_t t0 = __funnelshift_r(a0, a1, sh8), t1 = __funnelshift_r(a1, a2, sh8);   // window bytes 0..3, 4..7 of the upper row
    const uint32_t u0 = __funnelshift_r(b0, b1, sh8), u1 = __funnelshift_r(b1, b2, sh8);
    const uint32_t wx = (uint32_t)(32 - fx) | ((uint32_t)fx << 24);
    const uint32_t wy = (uint32_t)(32 - fy) | ((uint32_t)fy << 8);                          // dp2a: low halves of a x bytes 0, 1 of b
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t top = __dp4a(__funnelshift_r(t0, t1, 8 * c), wx, 0u);                 // bytes c .. c + 3 of the window: taps at c and c + 3
        const uint32_t bot = __dp4a(__funnelshift_r(u0, u1, 8 * c), wx, 0u);
        const uint32_t v = __dp2a_lo(top | (bot << 16), wy, 512u) >> 10;
        out |= v << (8 * c);
    }
    return out;
}
__device__ __forceinline__ uint32_t sample3_dp(const uint8_t* __restrict__ src, unsigned sstep, int sx, int sy, int fx, int fy) {
    const int o = 3 * sx, a = o & ~3, sh8 = 8 * (o & 3);
    const uint32_t* r0 = reinterpret_cast<const uint32_t*>(src + ((size_t)((unsigned)sy * sstep) + (unsigned)a));   // sources are below 4 GB
    const uint32_t* r1 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(r0) + sstep);
    const uint32_t a0 = __ldg(r0), a1 = __ldg(r0 + 1), a2 = __ldg(r0 + 2);
    const uint32_t b0 = __ldg(r1), b1 = __ldg(r1 + 1), b2 = __ldg(r1 + 2);
    return sample3_dp_taps(a0, a1, a2, b0, b1, b2, sh8, fx, fy);
}

// One thread per frame column of the block's tile, walking down its rows: everything that depends on the column only (table
// entries, the reflected image column, for the cylinder also the products m * sin u and m * cos u of the backward map) is
// computed once per thread, everything that depends on the row only is uniform over the block.
constexpr int WG_TX = 62, WG_TY = 16, WG_IW = 2 * WG_TX + 3, WG_IH = 2 * WG_TY + 3;       // 127 x 35 frame pixels under 62 x 16 outputs
constexpr int WG_THREADS = 128;

template <int PROJ>
__global__ void __launch_bounds__(WG_THREADS) k_warp_g1(WarpG1Args A) {
    __shared__ uint32_t tile[WG_IH][WG_IW + 1];                  // b | g << 8 | r << 16 | mask << 24
    __shared__ int16_t hsum[WG_IH][WG_TX][3];                     // at most 16 * 255
    __shared__ __align__(16) float row_t[WG_IH][4];                             // per tile row: the row's terms of the backward map
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * WG_TX, oy0 = blockIdx.y * WG_TY;
    const int fx_lo = 2 * ox0 - 2, fy_lo = 2 * oy0 - 2;
    const int cols = A.P.dst_w, rows = A.P.dst_h;
    const float* m = A.P.k_rinv;
    if (tid < WG_IH) {                                            // what depends on the row only, once per block
        const int fy = reflect101_i(fy_lo + tid, A.height);      // pyrDown's BORDER_REFLECT_101 on the frame
        const int iy = reflect_idx(fy - A.top, rows);            // copyMakeBorder's BORDER_REFLECT into the image
        const float ra = __ldg(A.tables + 2 * (size_t)cols + iy);
        if (PROJ == IS_PROJ_CYLINDRICAL || PROJ == IS_PROJ_PLANE) {   // (x_, y_, z_) = (sin u, v / scale, cos u) -- plane: (u', v', 1) --: the middle product of every row of k_rinv
            row_t[tid][0] = __fmul_rn(m[1], ra); row_t[tid][1] = __fmul_rn(m[4], ra); row_t[tid][2] = __fmul_rn(m[7], ra); row_t[tid][3] = 0.f;
        } else {
            row_t[tid][0] = ra; row_t[tid][1] = __ldg(A.tables + 2 * (size_t)cols + rows + iy); row_t[tid][2] = 0.f; row_t[tid][3] = 0.f;
        }
    }
    __syncthreads();
    if (tid < WG_IW) {
        const int fx = reflect101_i(fx_lo + tid, A.width);
        const int ix = reflect_idx(fx - A.left, cols);
        const float su = __ldg(A.tables + ix), cu = __ldg(A.tables + cols + ix);
        // cylinder: the first and third product of every row of k_rinv depend on the column only
        const float ax = __fmul_rn(m[0], su), cx = __fmul_rn(m[2], cu), ay = __fmul_rn(m[3], su), cy = __fmul_rn(m[5], cu), az = __fmul_rn(m[6], su), cz = __fmul_rn(m[8], cu);
        const bool wide = A.wide_ok != 0;
        const uint8_t* src = A.src;
        const unsigned sstep = (unsigned)A.sstep;
        const int sw = A.P.src_w, sh = A.P.src_h;
        // Software pipeline over the rows: the six loads of row r + 1 are in flight while row r is interpolated (the first use of a
        // loaded word was where this kernel waited: 29 % of its stall samples).
        struct Taps { uint32_t a0, a1, a2, b0, b1, b2; float sx, sy; int fx, fy, sh8; bool fast, inside; };
        auto issue = [&](int r, Taps& T) {
            const float4 rt = *reinterpret_cast<const float4*>(row_t[r]);
            float sx, sy;
            constexpr bool SAFE = PROJ == IS_PROJ_PLANE;         // no z > 0 guard there: the quotients may be anything
            if (PROJ == IS_PROJ_CYLINDRICAL || PROJ == IS_PROJ_PLANE) {
                const float X = __fadd_rn(__fadd_rn(ax, rt.x), cx);
                const float Y = __fadd_rn(__fadd_rn(ay, rt.y), cy);
                const float Z = __fadd_rn(__fadd_rn(az, rt.z), cz);
                if (PROJ == IS_PROJ_PLANE || Z > 0.f) { sx = __fdiv_rn(X, Z); sy = __fdiv_rn(Y, Z); }
                else { sx = -1.f; sy = -1.f; }
            } else {
                map_backward<PROJ>(A.P, su, cu, rt.x, rt.y, &sx, &sy);
            }
            const int qx = cv_round<SAFE>(__fmul_rn(sx, 32.f)), qy = cv_round<SAFE>(__fmul_rn(sy, 32.f));
            const int px = qx >> 5, py = qy >> 5;                 // inside the source here, so the remap's saturation to short is the identity
            T.sx = sx; T.sy = sy; T.fx = qx & 31; T.fy = qy & 31;
            T.fast = wide && px >= 0 && px < sw - 3 && (unsigned)py < (unsigned)(sh - 1);
            // the all-255 mask: INTER_NEAREST + BORDER_CONSTANT (source sizes are below 32768: saturating the rounded coordinate to short changes nothing)
            const int nx = cv_round<SAFE>(sx), ny = cv_round<SAFE>(sy);
            T.inside = (unsigned)nx < (unsigned)sw && (unsigned)ny < (unsigned)sh;
            if (T.fast) {
                const int o = 3 * px, a = o & ~3;
                T.sh8 = 8 * (o & 3);
                const uint32_t* r0 = reinterpret_cast<const uint32_t*>(src + ((size_t)((unsigned)py * sstep) + (unsigned)a));   // sources are below 4 GB
                const uint32_t* r1 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(r0) + sstep);
                T.a0 = __ldg(r0); T.a1 = __ldg(r0 + 1); T.a2 = __ldg(r0 + 2);
                T.b0 = __ldg(r1); T.b1 = __ldg(r1 + 1); T.b2 = __ldg(r1 + 2);
            }
        };
        auto finish = [&](int r, const Taps& T) {
            uint32_t w;
            if (T.fast) w = sample3_dp_taps(T.a0, T.a1, T.a2, T.b0, T.b1, T.b2, T.sh8, T.fx, T.fy);
            else w = sample3_packed<PROJ == IS_PROJ_PLANE>(src, A.sstep, sw, sh, T.sx, T.sy, false);
            if (T.inside) w |= 0xff000000u;
            tile[r][tid] = w;
        };
        Taps t0, t1, t2;                                         // the loads of two rows ahead are in flight
        issue(0, t0);
        issue(1, t1);
#pragma unroll 1
        for (int r = 0; r < WG_IH; r += 3) {
            if (r + 2 < WG_IH) issue(r + 2, t2);
            finish(r, t0);
            if (r + 1 < WG_IH) {
                if (r + 3 < WG_IH) issue(r + 3, t0);
                finish(r + 1, t1);
            }
            if (r + 2 < WG_IH) {
                if (r + 4 < WG_IH) issue(r + 4, t1);
                finish(r + 2, t2);
            }
        }
    }
    __syncthreads();
    // the image pixels of this block's frame rectangle: four pixels (12 bytes) per thread where the address allows it
    {
        const int x_img0 = max(2 * ox0 - A.left, 0), x_img1 = min(2 * ox0 + 2 * WG_TX - A.left, cols);   // image columns owned
        const int y_img0 = max(2 * oy0 - A.top, 0), y_img1 = min(2 * oy0 + 2 * WG_TY - A.top, rows);
        if (x_img0 < x_img1 && y_img0 < y_img1) {
            const int q0 = x_img0 >> 2, q1 = (x_img1 + 3) >> 2;                                         // quads of four image columns
            const int nq = q1 - q0, nrow = y_img1 - y_img0;
            const bool aligned = ((reinterpret_cast<uintptr_t>(A.dst) | A.dstep) & 3) == 0 && ((reinterpret_cast<uintptr_t>(A.mask) | A.mstep) & 3) == 0;
            for (int e = tid >> 5; e < nrow * ((nq + 31) >> 5); e += WG_THREADS >> 5) {        // one warp per row (and per 32 quads of it: a tile row has 31 or 32)
                const int chunks = (nq + 31) >> 5;
                const int ry = chunks == 1 ? e : e / chunks, qq = (chunks == 1 ? 0 : (e % chunks) << 5) + (tid & 31);
                if (qq >= nq) continue;
                const int yy = y_img0 + ry, x4 = 4 * (q0 + qq);
                const uint32_t* t = &tile[yy + A.top - fy_lo][0] + (A.left - fx_lo);                    // t[image x] = frame pixel
                uint8_t* d = A.dst + (size_t)yy * A.dstep + 3 * (size_t)x4;
                uint8_t* m = A.mask + (size_t)yy * A.mstep + x4;
                if (aligned && x4 >= x_img0 && x4 + 4 <= x_img1) {
                    const uint32_t p0 = t[x4], p1 = t[x4 + 1], p2 = t[x4 + 2], p3 = t[x4 + 3];
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
                    d32[0] = (p0 & 0xffffffu) | (p1 << 24);
                    d32[1] = ((p1 >> 8) & 0xffffu) | (p2 << 16);
                    d32[2] = ((p2 >> 16) & 0xffu) | (p3 << 8);
                    *reinterpret_cast<uint32_t*>(m) = (p0 >> 24) | ((p1 >> 24) << 8) | ((p2 >> 24) << 16) | ((p3 >> 24) << 24);
                } else {
                    for (int k = 0; k < 4; ++k) {
                        const int x = x4 + k;
                        if (x < x_img0 || x >= x_img1) continue;
                        const uint32_t p = t[x];
                        d[3 * k] = (uint8_t)p; d[3 * k + 1] = (uint8_t)(p >> 8); d[3 * k + 2] = (uint8_t)(p >> 16);
                        m[k] = (uint8_t)(p >> 24);
                    }
                }
            }
        }
    }
    // pyrDown: horizontal 5-tap sums, then the vertical pass ((s + 128) >> 8, exact integers)
    for (int e = tid; e < WG_IH * WG_TX; e += WG_THREADS) {
        const int r = e / WG_TX, ox = e % WG_TX;
        const uint32_t* p = &tile[r][2 * ox];
        const uint32_t p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3], p4 = p[4];
        // channel c of four neighbouring pixels gathered into one word (three byte permutes), the taps 1 4 6 4 as one byte dot product
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t w01 = __byte_perm(p0, p1, (uint32_t)c | ((uint32_t)(4 + c) << 4));
            const uint32_t w23 = __byte_perm(p2, p3, (uint32_t)c | ((uint32_t)(4 + c) << 4));
            const uint32_t w = __byte_perm(w01, w23, 0x5410u);
            hsum[r][ox][c] = (int16_t)(__dp4a(w, 0x04060401u, (p4 >> (8 * c)) & 255u));
        }
    }
    __syncthreads();
    for (int e = tid; e < WG_TY * WG_TX; e += WG_THREADS) {
        const int oy = e / WG_TX, ox = e % WG_TX;
        const int x = ox0 + ox, y = oy0 + oy;
        if (x >= A.dw || y >= A.dh) continue;
        int16_t* o = A.g1 + ((size_t)y * A.dw + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int acc = hsum[2 * oy][ox][c] + 4 * hsum[2 * oy + 1][ox][c] + 6 * hsum[2 * oy + 2][ox][c] + 4 * hsum[2 * oy + 3][ox][c] + hsum[2 * oy + 4][ox][c];
            o[c] = (int16_t)((acc + 128) >> 8);                                                         // at most 255: no saturation needed
        }
    }
}
// @emu-g1-end

// ---- host drivers ----------------------------------------------------------------------------------

struct PlanKey { int proj, w, h; float K[9], R[9], scale; };
struct PlanEntry { PlanKey key; WarpPlan plan; };

static PlanKey plan_key(int proj, int src_w, int src_h, const float* K, const float* R, float scale) {
    PlanKey key;
    std::memset(&key, 0, sizeof(key));
    key.proj = proj; key.w = src_w; key.h = src_h; key.scale = scale;
    std::memcpy(key.K, K, sizeof(key.K));
    std::memcpy(key.R, R, sizeof(key.R));
    return key;
}

static bool plan_lookup(is_ctx* ctx, const PlanKey& key, WarpPlan* plan) {
    const size_t nent = ctx->plan_cache.size() / sizeof(PlanEntry);
    for (size_t e = 0; e < nent; ++e) {
        const PlanEntry* ent = reinterpret_cast<const PlanEntry*>(ctx->plan_cache.data()) + e;
        if (std::memcmp(&ent->key, &key, sizeof(PlanKey)) == 0) { *plan = ent->plan; return true; }
    }
    return false;
}

static void plan_store(is_ctx* ctx, const PlanKey& key, const WarpPlan& plan) {
    if (ctx->plan_cache.size() / sizeof(PlanEntry) >= 256) ctx->plan_cache.clear();
    PlanEntry ent;
    std::memset(&ent, 0, sizeof(ent));
    ent.key = key;
    ent.plan = plan;
    const unsigned char* raw = reinterpret_cast<const unsigned char*>(&ent);
    ctx->plan_cache.insert(ctx->plan_cache.end(), raw, raw + sizeof(PlanEntry));
}

static int plan_from_roi(is_ctx* ctx, const Projector& p, int src_w, int src_h, float scale, WarpPlan* plan) {
    std::memcpy(plan->P.k_rinv, p.k_rinv, sizeof(p.k_rinv));
    plan->P.scale = scale;
    plan->P.tl_x = plan->roi[0];
    plan->P.tl_y = plan->roi[1];
    plan->P.dst_w = plan->roi[2] - plan->roi[0] + 1;
    plan->P.dst_h = plan->roi[3] - plan->roi[1] + 1;
    plan->P.src_w = src_w;
    plan->P.src_h = src_h;
    IS_REQUIRE(ctx, plan->P.dst_w > 0 && plan->P.dst_h > 0 && plan->P.dst_w < (1 << 24) && plan->P.dst_h < (1 << 24),
               IS_ERR_BAD_ARG, "degenerate warp ROI");
    return IS_OK;
}

// the full forward scan of the per-pixel projectors (fisheye, stereographic), rows spread over the host pool
static void detect_roi_full_parallel(is_ctx* ctx, int proj, int w, int h, const Projector& p, float scale, int roi[4]) {
    const size_t pieces = (size_t)std::min(h, 64);
    std::vector<RoiAcc> acc(pieces);
    host_pool(ctx)->run(pieces, [&](size_t j) {
        acc[j] = roi_rows(proj, w, p, scale, (int)((long long)h * (long long)j / (long long)pieces), (int)((long long)h * (long long)(j + 1) / (long long)pieces));
    });
    RoiAcc a = acc[0];
    for (size_t q = 1; q < pieces; ++q) {
        const RoiAcc& b = acc[q];
        a.tl_u = std::min(a.tl_u, b.tl_u); a.tl_v = std::min(a.tl_v, b.tl_v); a.br_u = std::max(a.br_u, b.br_u); a.br_v = std::max(a.br_v, b.br_v);
    }
    finish_roi(proj, w, h, p, scale, a.tl_u, a.tl_v, a.br_u, a.br_v, roi);
}

int warp_plan(is_ctx* ctx, int proj, int src_w, int src_h, const float* K, const float* R, float scale, WarpPlan* plan) {
    IS_REQUIRE(ctx, proj_known(proj), IS_ERR_BAD_ARG, "unknown projection");
    IS_REQUIRE(ctx, K && R, IS_ERR_ASSERT, "K and R must be 3x3 CV_32F");
    IS_REQUIRE(ctx, src_w > 0 && src_h > 0 && scale > 0.f, IS_ERR_BAD_ARG, "empty source or non-positive scale");
    const PlanKey key = plan_key(proj, src_w, src_h, K, R, scale);
    if (plan_lookup(ctx, key, plan)) return IS_OK;
    Projector p;
    set_camera(K, R, &p);
    if (proj_uses_maps(proj)) detect_roi_full_parallel(ctx, proj, src_w, src_h, p, scale, plan->roi);
    else detect_roi(proj, src_w, src_h, p, scale, plan->roi);
    IS_TRY(plan_from_roi(ctx, p, src_w, src_h, scale, plan));
    plan_store(ctx, key, *plan);
    return IS_OK;
}

// The plans of all images of a panorama at once: the border scans of the images not yet in the memo are cut into pieces and
// spread over the context's host threads (libm atan2f / sqrtf per border pixel: ~0.25 ms per 24 MP image on one core).
int warp_plan_many(is_ctx* ctx, int proj, int n, const int* src_w, const int* src_h, const float* const* K, const float* const* R, float scale, WarpPlan* plans) {
    IS_REQUIRE(ctx, proj_known(proj), IS_ERR_BAD_ARG, "unknown projection");
    IS_REQUIRE(ctx, scale > 0.f, IS_ERR_BAD_ARG, "non-positive scale");
    if (proj != IS_PROJ_CYLINDRICAL && proj != IS_PROJ_SPHERICAL) {      // four corners (plane) or a scan that is spread over the pool per image
        for (int i = 0; i < n; ++i) IS_TRY(warp_plan(ctx, proj, src_w[i], src_h[i], K[i], R[i], scale, &plans[i]));
        return IS_OK;
    }
    std::vector<int> todo;
    std::vector<PlanKey> keys((size_t)n);
    for (int i = 0; i < n; ++i) {
        IS_REQUIRE(ctx, K[i] && R[i], IS_ERR_ASSERT, "K and R must be 3x3 CV_32F");
        IS_REQUIRE(ctx, src_w[i] > 0 && src_h[i] > 0, IS_ERR_BAD_ARG, "empty source");
        keys[(size_t)i] = plan_key(proj, src_w[i], src_h[i], K[i], R[i], scale);
        if (!plan_lookup(ctx, keys[(size_t)i], &plans[i])) todo.push_back(i);
    }
    if (todo.empty()) return IS_OK;
    constexpr int PIECES = 8;                             // per image: 4 pieces of the rows, 4 of the columns
    std::vector<Projector> proj_of(todo.size());
    std::vector<RoiAcc> acc(todo.size() * PIECES);
    for (size_t t = 0; t < todo.size(); ++t) set_camera(K[todo[t]], R[todo[t]], &proj_of[t]);
    host_pool(ctx)->run(todo.size() * PIECES, [&](size_t job) {
        const size_t t = job / PIECES;
        const int piece = (int)(job % PIECES), i = todo[t];
        const bool rows = piece < PIECES / 2;
        const int len = rows ? src_w[i] : src_h[i], q = piece % (PIECES / 2);
        acc[job] = roi_segment(proj, src_w[i], src_h[i], proj_of[t], scale, rows, (int)((long long)len * q / (PIECES / 2)), (int)((long long)len * (q + 1) / (PIECES / 2)));
    });
    for (size_t t = 0; t < todo.size(); ++t) {
        const int i = todo[t];
        RoiAcc a = acc[t * PIECES];
        for (int q = 1; q < PIECES; ++q) {
            const RoiAcc& b = acc[t * PIECES + q];
            a.tl_u = std::min(a.tl_u, b.tl_u); a.tl_v = std::min(a.tl_v, b.tl_v); a.br_u = std::max(a.br_u, b.br_u); a.br_v = std::max(a.br_v, b.br_v);
        }
        finish_roi(proj, src_w[i], src_h[i], proj_of[t], scale, a.tl_u, a.tl_v, a.br_u, a.br_v, plans[i].roi);
        IS_TRY(plan_from_roi(ctx, proj_of[t], src_w[i], src_h[i], scale, &plans[i]));
        plan_store(ctx, keys[(size_t)i], plans[i]);
    }
    return IS_OK;
}

// xmap | ymap (dense, dst_h x dst_w floats each) of a per-pixel projector, built by the host pool
static void build_maps_host(is_ctx* ctx, int proj, const WarpPlan& plan, float* xmap, float* ymap) {
    Projector p;
    std::memset(&p, 0, sizeof(p));
    std::memcpy(p.k_rinv, plan.P.k_rinv, sizeof(p.k_rinv));
    const int w = plan.P.dst_w, h = plan.P.dst_h;
    const size_t pieces = (size_t)std::min(h, 256);
    host_pool(ctx)->run(pieces, [&](size_t j) {
        fill_map_rows(proj, p, plan.P.scale, plan.P.tl_x, plan.P.tl_y, w, (int)((long long)h * (long long)j / (long long)pieces),
                      (int)((long long)h * (long long)(j + 1) / (long long)pieces), xmap, ymap);
    });
}

int upload_tables(is_ctx* ctx, int proj, const WarpPlan& plan, DevBuf* buf) {
    const int w = plan.P.dst_w, h = plan.P.dst_h;
    if (proj_uses_maps(proj)) {
        const size_t px = (size_t)w * (size_t)h;
        IS_REQUIRE(ctx, px < ((size_t)1 << 31), IS_ERR_NO_MEM, "maps of a per-pixel projector: destination of 2^31 pixels or more (a pole inside the image?)");
        IS_TRY(buf->alloc(ctx, 2 * px * sizeof(float)));
        std::vector<float> maps;
        try { maps.resize(2 * px); } catch (const std::bad_alloc&) { return fail(ctx, IS_ERR_NO_MEM, "host memory for the maps of a per-pixel projector (%zu pixels)", px); }
        build_maps_host(ctx, proj, plan, maps.data(), maps.data() + px);
        IS_CUDA(ctx, cudaMemcpyAsync(buf->p, maps.data(), 2 * px * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // pageable source: gone when this returns
        return IS_OK;
    }
    const size_t n = 2 * (size_t)w + 2 * (size_t)h;
    IS_TRY(buf->alloc(ctx, n * sizeof(float)));
    void* stage = nullptr;
    IS_TRY(pinned_alloc(ctx, n * sizeof(float), &stage));
    fill_tables(proj, plan.P.scale, plan.P.tl_x, plan.P.tl_y, w, h, (float*)stage);
    IS_TRY(copy_small(ctx, buf->p, stage, n * sizeof(float), cudaMemcpyHostToDevice));
    return IS_OK;
}

template <int PROJ, int CH, int INTERP, int BORDER, bool WITH_MASK>
static int launch_warp_t(is_ctx* ctx, const WarpPlan& plan, const float* tables, const DevMat& src, const DevMat& dst,
                         const DevMat* mask) {
    dim3 block(WARP_BX, WARP_BY);
    dim3 grid(div_up(plan.P.dst_w, WARP_BX * WARP_PX), div_up(plan.P.dst_h, WARP_BY));
    // algorithmic bytes (B_warp_only): source read once, warped image (+ mask) written once
    ctx->next_bytes = (double)CH * plan.P.src_w * plan.P.src_h + (double)(CH + (WITH_MASK ? 1 : 0)) * plan.P.dst_w * plan.P.dst_h;
    IS_LAUNCH(ctx, (k_warp<PROJ, CH, INTERP, BORDER, WITH_MASK>), grid, block, 0, plan.P, tables, src.ptr<uint8_t>(), src.step,
              dst.ptr<uint8_t>(), dst.step, mask ? mask->ptr<uint8_t>() : nullptr, mask ? mask->step : 0);
    return IS_OK;
}

template <int CH, int INTERP, int BORDER, bool WITH_MASK>
static int launch_remap_t(is_ctx* ctx, int src_w, int src_h, const float* xmap, size_t xstep, const float* ymap, size_t ystep, const DevMat& src, const DevMat& dst,
                          const DevMat* mask) {
    dim3 block(WARP_BX, WARP_BY);
    dim3 grid(div_up(dst.cols, WARP_BX * WARP_PX), div_up(dst.rows, WARP_BY));
    // algorithmic bytes: the two maps and the source read once, the destination (+ mask) written once
    ctx->next_bytes = (double)CH * src_w * src_h + (double)(8 + CH + (WITH_MASK ? 1 : 0)) * dst.cols * dst.rows;
    IS_LAUNCH(ctx, (k_remap<CH, INTERP, BORDER, WITH_MASK>), grid, block, 0, dst.cols, dst.rows, src_w, src_h, xmap, xstep, ymap, ystep, src.ptr<uint8_t>(),
              src.step, dst.ptr<uint8_t>(), dst.step, mask ? mask->ptr<uint8_t>() : nullptr, mask ? mask->step : 0);
    return IS_OK;
}

bool warp_fusable(int proj) { return proj == IS_PROJ_CYLINDRICAL || proj == IS_PROJ_SPHERICAL || proj == IS_PROJ_PLANE; }

// cv::remap of a device-resident 8-bit image through device-resident maps
int launch_remap(is_ctx* ctx, const float* xmap, size_t xstep, const float* ymap, size_t ystep, const DevMat& src, int interp, int border, const DevMat& dst,
                 const DevMat* mask) {
    const int ch = src.channels;
#define IS_REMAP_CASE(C, I, B, M) \
    if (ch == C && interp == I && border == B && (mask != nullptr) == M) return launch_remap_t<C, I, B, M>(ctx, src.cols, src.rows, xmap, xstep, ymap, ystep, src, dst, mask);
    IS_REMAP_CASE(3, IS_INTER_LINEAR, IS_BORDER_REFLECT, true)
    IS_REMAP_CASE(3, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    IS_REMAP_CASE(3, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)
    IS_REMAP_CASE(3, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)
    IS_REMAP_CASE(3, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
    IS_REMAP_CASE(1, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    IS_REMAP_CASE(1, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)
    IS_REMAP_CASE(1, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)
    IS_REMAP_CASE(1, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
#undef IS_REMAP_CASE
    return fail(ctx, IS_ERR_UNSUPPORTED, "remap: unsupported combination (channels=%d interp=%d border=%d)", ch, interp, border);
}

// warp + mask + Gaussian level 1 of the padded frame (top, left, height, width) in one pass; g1: (height / 2) x (width / 2) x 3 int16
int launch_warp_g1(is_ctx* ctx, int proj, const WarpPlan& plan, const float* tables, const DevMat& src, const DevMat& dst, const DevMat& mask,
                   int top, int left, int height, int width, int16_t* g1) {
    WarpG1Args A;
    A.P = plan.P; A.tables = tables;
    A.src = src.ptr<uint8_t>(); A.sstep = src.step;
    A.dst = dst.ptr<uint8_t>(); A.dstep = dst.step; A.mask = mask.ptr<uint8_t>(); A.mstep = mask.step;
    A.top = top; A.left = left; A.height = height; A.width = width;
    A.g1 = g1; A.dh = (height + 1) / 2; A.dw = (width + 1) / 2;
    IS_REQUIRE(ctx, plan.P.src_w < 32768 && plan.P.src_h < 32768 && (size_t)src.step * (size_t)plan.P.src_h < ((size_t)1 << 32), IS_ERR_UNSUPPORTED,
               "fused warp: source larger than 32767 pixels a side or 4 GB");
    A.wide_ok = ((reinterpret_cast<uintptr_t>(src.data) | src.step) & 3) == 0 ? 1 : 0;
    dim3 grid(div_up(A.dw, WG_TX), div_up(A.dh, WG_TY));
    // algorithmic bytes: source read once, warped image + mask and level 1 written once
    ctx->next_bytes = 3. * plan.P.src_w * plan.P.src_h + 4. * plan.P.dst_w * plan.P.dst_h + 6. * A.dh * A.dw;
    IS_REQUIRE(ctx, proj == IS_PROJ_CYLINDRICAL || proj == IS_PROJ_SPHERICAL || proj == IS_PROJ_PLANE, IS_ERR_UNSUPPORTED, "fused warp: table projectors only");
    if (proj == IS_PROJ_CYLINDRICAL) IS_LAUNCH(ctx, k_warp_g1<IS_PROJ_CYLINDRICAL>, grid, WG_THREADS, 0, A);
    else if (proj == IS_PROJ_PLANE) IS_LAUNCH(ctx, k_warp_g1<IS_PROJ_PLANE>, grid, WG_THREADS, 0, A);
    else IS_LAUNCH(ctx, k_warp_g1<IS_PROJ_SPHERICAL>, grid, WG_THREADS, 0, A);
    return IS_OK;
}

// device-resident src/dst; used by is_warp* and by the pipeline
int launch_warp(is_ctx* ctx, int proj, const WarpPlan& plan, const float* tables, const DevMat& src, int interp, int border,
                const DevMat& dst, const DevMat* mask) {
    const int ch = src.channels;
    if (proj_uses_maps(proj)) {                              // `tables` = xmap | ymap of upload_tables
        const size_t step = (size_t)plan.P.dst_w * sizeof(float);
        return launch_remap(ctx, tables, step, tables + (size_t)plan.P.dst_w * (size_t)plan.P.dst_h, step, src, interp, border, dst, mask);
    }
#define IS_WARP_CASE(PJ, C, I, B, M) \
    if (proj == PJ && ch == C && interp == I && border == B && (mask != nullptr) == M) \
        return launch_warp_t<PJ, C, I, B, M>(ctx, plan, tables, src, dst, mask);
#define IS_WARP_CASES(PJ)                                                   \
    IS_WARP_CASE(PJ, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, true)           \
    IS_WARP_CASE(PJ, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)          \
    IS_WARP_CASE(PJ, 3, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)         \
    IS_WARP_CASE(PJ, 3, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)         \
    IS_WARP_CASE(PJ, 3, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)        \
    IS_WARP_CASE(PJ, 1, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)          \
    IS_WARP_CASE(PJ, 1, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)         \
    IS_WARP_CASE(PJ, 1, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)         \
    IS_WARP_CASE(PJ, 1, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
    IS_WARP_CASES(IS_PROJ_CYLINDRICAL)
    IS_WARP_CASES(IS_PROJ_SPHERICAL)
    IS_WARP_CASES(IS_PROJ_PLANE)
#undef IS_WARP_CASES
#undef IS_WARP_CASE
    return fail(ctx, IS_ERR_UNSUPPORTED, "warp: unsupported combination (channels=%d interp=%d border=%d)", ch, interp, border);
}

}  // namespace is

using namespace is;

extern "C" {

int is_warp_roi(is_ctx* ctx, int projection, is_size src_size, const float K[9], const float R[9], float scale,
                is_point* dst_tl, is_size* dst_size) {
    if (!ctx) return IS_ERR_BAD_ARG;
    WarpPlan plan;
    IS_TRY(warp_plan(ctx, projection, src_size.width, src_size.height, K, R, scale, &plan));
    if (dst_tl) { dst_tl->x = plan.roi[0]; dst_tl->y = plan.roi[1]; }
    if (dst_size) { dst_size->width = plan.P.dst_w; dst_size->height = plan.P.dst_h; }
    return IS_OK;
}

int is_build_maps(is_ctx* ctx, int projection, is_size src_size, const float K[9], const float R[9], float scale,
                  is_mat* xmap, is_mat* ymap, is_rect* dst_roi) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    WarpPlan plan;
    IS_TRY(warp_plan(ctx, projection, src_size.width, src_size.height, K, R, scale, &plan));
    IS_TRY(check_mat(ctx, xmap, "xmap"));
    IS_TRY(check_mat(ctx, ymap, "ymap"));
    IS_REQUIRE(ctx, xmap->depth == IS_32F && xmap->channels == 1 && ymap->depth == IS_32F && ymap->channels == 1, IS_ERR_BAD_ARG,
               "maps must be 1-channel IS_32F");
    IS_REQUIRE(ctx, xmap->rows == plan.P.dst_h && xmap->cols == plan.P.dst_w && ymap->rows == plan.P.dst_h && ymap->cols == plan.P.dst_w,
               IS_ERR_BAD_ARG, "maps must have the size reported by is_warp_roi");
    if (dst_roi) { dst_roi->x = plan.roi[0]; dst_roi->y = plan.roi[1]; dst_roi->width = plan.roi[2] - plan.roi[0]; dst_roi->height = plan.roi[3] - plan.roi[1]; }
    if (proj_uses_maps(projection)) {                       // host-built maps: straight into host mats, one copy for device mats
        const size_t px = (size_t)plan.P.dst_w * (size_t)plan.P.dst_h, row = (size_t)plan.P.dst_w * sizeof(float);
        std::vector<float> maps;
        try { maps.resize(2 * px); } catch (const std::bad_alloc&) { return fail(ctx, IS_ERR_NO_MEM, "host memory for the maps of a per-pixel projector (%zu pixels)", px); }
        build_maps_host(ctx, projection, plan, maps.data(), maps.data() + px);
        is_mat* out[2] = {xmap, ymap};
        for (int k = 0; k < 2; ++k) {
            const float* m = maps.data() + (size_t)k * px;
            if (out[k]->device < 0) {
                for (int y = 0; y < plan.P.dst_h; ++y) std::memcpy((char*)out[k]->data + (size_t)y * out[k]->step, m + (size_t)y * plan.P.dst_w, row);
            } else {
                IS_CUDA(ctx, cudaMemcpy2DAsync(out[k]->data, out[k]->step, m, row, row, (size_t)plan.P.dst_h, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return IS_OK;
    }
    DevBuf tables;
    IS_TRY(upload_tables(ctx, projection, plan, &tables));
    DevMat dx, dy;
    IS_TRY(stage_out(ctx, xmap, &dx, false));
    IS_TRY(stage_out(ctx, ymap, &dy, false));
    dim3 block(32, 8), grid(div_up(plan.P.dst_w, 32), div_up(plan.P.dst_h, 8));
    if (projection == IS_PROJ_PLANE)
        IS_LAUNCH(ctx, k_build_maps<IS_PROJ_PLANE>, grid, block, 0, plan.P, tables.as<float>(), dx.ptr<float>(), dx.step,
                  dy.ptr<float>(), dy.step);
    else if (projection == IS_PROJ_CYLINDRICAL)
        IS_LAUNCH(ctx, k_build_maps<IS_PROJ_CYLINDRICAL>, grid, block, 0, plan.P, tables.as<float>(), dx.ptr<float>(), dx.step,
                  dy.ptr<float>(), dy.step);
    else
        IS_LAUNCH(ctx, k_build_maps<IS_PROJ_SPHERICAL>, grid, block, 0, plan.P, tables.as<float>(), dx.ptr<float>(), dx.step,
                  dy.ptr<float>(), dy.step);
    IS_TRY(commit(ctx, &dx));
    IS_TRY(commit(ctx, &dy));
    return IS_OK;
}

int is_remap(is_ctx* ctx, const is_mat* src, const is_mat* xmap, const is_mat* ymap, int interp, int border, is_mat* dst) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, src, "src"));
    IS_TRY(check_mat(ctx, xmap, "xmap"));
    IS_TRY(check_mat(ctx, ymap, "ymap"));
    IS_TRY(check_mat(ctx, dst, "dst"));
    IS_REQUIRE(ctx, src->depth == IS_8U && (src->channels == 1 || src->channels == 3), IS_ERR_UNSUPPORTED, "src must be 8UC1 or 8UC3");
    IS_REQUIRE(ctx, src->cols < 32768 && src->rows < 32768, IS_ERR_UNSUPPORTED, "cv::remap addresses sources of less than 32768 pixels a side");
    IS_REQUIRE(ctx, xmap->depth == IS_32F && xmap->channels == 1 && ymap->depth == IS_32F && ymap->channels == 1, IS_ERR_BAD_ARG, "maps must be 1-channel IS_32F");
    IS_REQUIRE(ctx, xmap->rows == ymap->rows && xmap->cols == ymap->cols, IS_ERR_BAD_ARG, "xmap and ymap must have the same size");
    IS_REQUIRE(ctx, dst->depth == IS_8U && dst->channels == src->channels && dst->rows == xmap->rows && dst->cols == xmap->cols, IS_ERR_BAD_ARG,
               "dst must have the type of src and the size of the maps");
    IS_REQUIRE(ctx, (interp == IS_INTER_NEAREST || interp == IS_INTER_LINEAR) && (border == IS_BORDER_CONSTANT || border == IS_BORDER_REFLECT),
               IS_ERR_UNSUPPORTED, "interp must be NEAREST/LINEAR and border CONSTANT/REFLECT");
    DevMat s, mx, my, d;
    IS_TRY(stage_in(ctx, src, &s));
    IS_TRY(stage_in(ctx, xmap, &mx));
    IS_TRY(stage_in(ctx, ymap, &my));
    IS_TRY(stage_out(ctx, dst, &d, false));
    IS_TRY(launch_remap(ctx, mx.ptr<float>(), mx.step, my.ptr<float>(), my.step, s, interp, border, d, nullptr));
    IS_TRY(commit(ctx, &d));
    return IS_OK;
}

static int warp_common(is_ctx* ctx, int projection, const is_mat* src, const float* K, const float* R, float scale,
                       int interp, int border, is_mat* dst, is_mat* dst_mask, is_point* dst_tl) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, src, "src"));
    IS_TRY(check_mat(ctx, dst, "dst"));
    IS_REQUIRE(ctx, src->depth == IS_8U && (src->channels == 1 || src->channels == 3), IS_ERR_UNSUPPORTED, "src must be 8UC1 or 8UC3");
    IS_REQUIRE(ctx, dst->depth == IS_8U && dst->channels == src->channels, IS_ERR_BAD_ARG, "dst must have the type of src");
    IS_REQUIRE(ctx, (interp == IS_INTER_NEAREST || interp == IS_INTER_LINEAR) && (border == IS_BORDER_CONSTANT || border == IS_BORDER_REFLECT),
               IS_ERR_UNSUPPORTED, "interp must be NEAREST/LINEAR and border CONSTANT/REFLECT");
    WarpPlan plan;
    IS_TRY(warp_plan(ctx, projection, src->cols, src->rows, K, R, scale, &plan));
    IS_REQUIRE(ctx, dst->rows == plan.P.dst_h && dst->cols == plan.P.dst_w, IS_ERR_BAD_ARG, "dst must have the size reported by is_warp_roi");
    if (dst_mask) {
        IS_TRY(check_mat(ctx, dst_mask, "dst_mask"));
        IS_REQUIRE(ctx, src->channels == 3, IS_ERR_BAD_ARG, "is_warp_with_mask needs a 3-channel source");
        IS_REQUIRE(ctx, dst_mask->depth == IS_8U && dst_mask->channels == 1 && dst_mask->rows == dst->rows && dst_mask->cols == dst->cols,
                   IS_ERR_BAD_ARG, "dst_mask must be 8UC1 of the dst size");
    }
    DevBuf tables;
    IS_TRY(upload_tables(ctx, projection, plan, &tables));
    DevMat s, d, m;
    IS_TRY(stage_in(ctx, src, &s));
    IS_TRY(stage_out(ctx, dst, &d, false));
    if (dst_mask) IS_TRY(stage_out(ctx, dst_mask, &m, false));
    IS_TRY(launch_warp(ctx, projection, plan, tables.as<float>(), s, interp, border, d, dst_mask ? &m : nullptr));
    IS_TRY(commit(ctx, &d));
    if (dst_mask) IS_TRY(commit(ctx, &m));
    if (dst_tl) { dst_tl->x = plan.roi[0]; dst_tl->y = plan.roi[1]; }
    return IS_OK;
}

int is_warp(is_ctx* ctx, int projection, const is_mat* src, const float K[9], const float R[9], float scale, int interp,
            int border, is_mat* dst, is_point* dst_tl) {
    return warp_common(ctx, projection, src, K, R, scale, interp, border, dst, nullptr, dst_tl);
}

int is_warp_with_mask(is_ctx* ctx, int projection, const is_mat* src, const float K[9], const float R[9], float scale,
                      is_mat* dst, is_mat* dst_mask, is_point* dst_tl) {
    if (!dst_mask) return ctx ? fail(ctx, IS_ERR_BAD_ARG, "dst_mask is null") : IS_ERR_BAD_ARG;
    return warp_common(ctx, projection, src, K, R, scale, IS_INTER_LINEAR, IS_BORDER_REFLECT, dst, dst_mask, dst_tl);
}

}  // extern "C"
