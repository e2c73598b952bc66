// oracle/warp.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the reference's rotation warper:
//   setCameraParams   [WARP]:90-120
//   mapForward        [WARP]:37-45      (cylindrical)   / SURVEY.md a25 (spherical, OpenCV warpers_inl.hpp)
//   plane / fisheye / stereographic: the projectors behind the warper creators the reference keeps commented out at
//                     [BLEND]:91-95 (cv::PlaneWarper, FisheyeWarper, StereographicWarper; OpenCV stitching, un-vendored:
//                     PlaneProjector, FisheyeProjector, StereographicProjector of warpers_inl.hpp, T = 0), pinned to cv2
//   mapBackward       [WARP]:47-63
//   detectResultRoi   [WARP]:64-88
//   buildMaps         [WARP]:122-144
//   warp -> cv::remap [WARP]:145-161    (OpenCV imgproc remap, 8-bit, fixed-point)
//
// Must be compiled with -ffp-contract=off: every float expression below is evaluated in the
// reference's association order with one rounding per operation.
#include "oracle.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Projector {
    float k[9], rinv[9], r_kinv[9], k_rinv[9];
    float scale;
    int proj;
};

// cv::invert (DECOMP_LU) closed form for a 3x3 CV_32F matrix: determinant and cofactors in double,
// result rounded to float.
bool invert3x3(const float* s, float* d) {
    double det = (double)s[0] * ((double)s[4] * s[8] - (double)s[5] * s[7]) -
                 (double)s[1] * ((double)s[3] * s[8] - (double)s[5] * s[6]) +
                 (double)s[2] * ((double)s[3] * s[7] - (double)s[4] * s[6]);
    if (det == 0.) return false;
    double id = 1. / det;
    double t[9];
    t[0] = ((double)s[4] * s[8] - (double)s[5] * s[7]) * id;
    t[1] = ((double)s[2] * s[7] - (double)s[1] * s[8]) * id;
    t[2] = ((double)s[1] * s[5] - (double)s[2] * s[4]) * id;
    t[3] = ((double)s[5] * s[6] - (double)s[3] * s[8]) * id;
    t[4] = ((double)s[0] * s[8] - (double)s[2] * s[6]) * id;
    t[5] = ((double)s[2] * s[3] - (double)s[0] * s[5]) * id;
    t[6] = ((double)s[3] * s[7] - (double)s[4] * s[6]) * id;
    t[7] = ((double)s[1] * s[6] - (double)s[0] * s[7]) * id;
    t[8] = ((double)s[0] * s[4] - (double)s[1] * s[3]) * id;
    for (int i = 0; i < 9; ++i) d[i] = (float)t[i];
    return true;
}

// 3x3 float product, float accumulation left to right (cv::gemm small-matrix path).
void mul3x3(const float* a, const float* b, float* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int t = 0; t < 3; ++t) s += a[i * 3 + t] * b[t * 3 + j];
            c[i * 3 + j] = s;
        }
}

void setCameraParams(Projector& p, const float* K, const float* R) {   // [WARP]:90-120
    std::memcpy(p.k, K, sizeof(p.k));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) p.rinv[i * 3 + j] = R[j * 3 + i];   // Rinv = R.t()   :103
    float kinv[9];
    if (!invert3x3(K, kinv)) std::memset(kinv, 0, sizeof(kinv));
    mul3x3(R, kinv, p.r_kinv);                                          // :108
    mul3x3(K, p.rinv, p.k_rinv);                                        // :113
}

const float PI_F = static_cast<float>(3.1415926535897932384626433832795);

inline void mapForward(const Projector& p, float x, float y, float& u, float& v) {
    const float* r = p.r_kinv;
    float x_ = r[0] * x + r[1] * y + r[2];
    float y_ = r[3] * x + r[4] * y + r[5];
    float z_ = r[6] * x + r[7] * y + r[8];
    if (p.proj == ORC_PROJ_CYLINDRICAL) {          // [WARP]:43-44
        u = p.scale * atan2f(x_, z_);
        v = p.scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    } else if (p.proj == ORC_PROJ_PLANE) {          // PlaneProjector::mapForward with t = (0, 0, 0)
        x_ = 0.f + x_ / z_ * (1 - 0.f);
        y_ = 0.f + y_ / z_ * (1 - 0.f);
        u = p.scale * x_;
        v = p.scale * y_;
    } else if (p.proj == ORC_PROJ_FISHEYE) {        // FisheyeProjector::mapForward
        float u_ = atan2f(x_, z_);
        float v_ = PI_F - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        float r = p.scale * v_;
        u = r * cosf(u_);
        v = r * sinf(u_);
    } else if (p.proj == ORC_PROJ_STEREOGRAPHIC) {  // StereographicProjector::mapForward
        float u_ = atan2f(x_, z_);
        float v_ = PI_F - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        float r = sinf(v_) / (1 - cosf(v_));
        u = p.scale * r * cosf(u_);
        v = p.scale * r * sinf(u_);
    } else {                                        // SphericalProjector::mapForward
        u = p.scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        v = p.scale * (PI_F - acosf(w == w ? w : 0));
    }
}

inline void mapBackward(const Projector& p, float u, float v, float& x, float& y) {
    u /= p.scale;                                   // [WARP]:49-50
    v /= p.scale;
    float x_, y_, z_;
    if (p.proj == ORC_PROJ_CYLINDRICAL) {           // [WARP]:52-54
        x_ = sinf(u);
        y_ = v;
        z_ = cosf(u);
    } else if (p.proj == ORC_PROJ_PLANE) {          // PlaneProjector::mapBackward, t = 0: no z > 0 guard
        const float* m = p.k_rinv;
        u = u - 0.f;
        v = v - 0.f;
        float z;
        x = m[0] * u + m[1] * v + m[2] * (1 - 0.f);
        y = m[3] * u + m[4] * v + m[5] * (1 - 0.f);
        z = m[6] * u + m[7] * v + m[8] * (1 - 0.f);
        x /= z;
        y /= z;
        return;
    } else if (p.proj == ORC_PROJ_FISHEYE) {        // FisheyeProjector::mapBackward
        float u_ = atan2f(v, u);
        float v_ = sqrtf(u * u + v * v);
        float sinv = sinf(PI_F - v_);
        x_ = sinv * sinf(u_);
        y_ = cosf(PI_F - v_);
        z_ = sinv * cosf(u_);
    } else if (p.proj == ORC_PROJ_STEREOGRAPHIC) {  // StereographicProjector::mapBackward
        float u_ = atan2f(v, u);
        float r = sqrtf(u * u + v * v);
        float v_ = 2 * atanf(1.f / r);
        float sinv = sinf(PI_F - v_);
        x_ = sinv * sinf(u_);
        y_ = cosf(PI_F - v_);
        z_ = sinv * cosf(u_);
    } else {                                        // SphericalProjector::mapBackward
        float sinv = sinf(PI_F - v);
        x_ = sinv * sinf(u);
        y_ = cosf(PI_F - v);
        z_ = sinv * cosf(u);
    }
    const float* m = p.k_rinv;
    float z;
    x = m[0] * x_ + m[1] * y_ + m[2] * z_;          // [WARP]:57-59
    y = m[3] * x_ + m[4] * y_ + m[5] * z_;
    z = m[6] * x_ + m[7] * y_ + m[8] * z_;
    if (z > 0) { x /= z; y /= z; }                  // [WARP]:61-62
    else x = y = -1;
}

// cv::cvRound on x86 is _mm_cvtss_si32 / cvtps2dq: round-half-even, and the "integer indefinite" value INT_MIN for NaN and for
// everything outside the int range (lrintf alone would give the 64-bit result's low half there)
inline int cvRound(float v) {
    if (!(v >= -2147483648.f && v < 2147483648.f)) return -2147483647 - 1;
    return (int)lrintf(v);
}

inline short saturate_short(int v) {
    return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
}

// cv::borderInterpolate for BORDER_REFLECT
inline int reflect(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// cv::initInterTab2D(INTER_LINEAR, fixpt=true): 32x32 entries of 4 short weights that sum to 32768.
const short* bilinearTab() {
    static short tab[32 * 32 * 4];
    static bool inited = false;
    if (inited) return tab;
    float t1[32][2];
    for (int i = 0; i < 32; ++i) {
        float x = i * (1.f / 32);
        t1[i][0] = 1.f - x;
        t1[i][1] = x;
    }
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            short* it = tab + (i * 32 + j) * 4;
            int isum = 0;
            for (int k1 = 0; k1 < 2; ++k1)
                for (int k2 = 0; k2 < 2; ++k2) {
                    float v = t1[i][k1] * t1[j][k2];
                    it[k1 * 2 + k2] = saturate_short(cvRound(v * 32768));
                    isum += it[k1 * 2 + k2];
                }
            if (isum != 32768) {
                // OpenCV's fix-up (initInterTab2D) scans k1,k2 in [ksize/2, ksize/2+2); for the 2x2
                // bilinear table that is entry (1,1) only.  It triggers for (fy,fx) = (0,0) alone, where
                // saturate_cast<short>(32768) = 32767: the table becomes {32767,0,0,1}, which still
                // reproduces src(sy,sx) exactly for 8-bit data.
                int diff = isum - 32768;
                it[3] = (short)(it[3] - diff);
            }
        }
    inited = true;
    return tab;
}

}  // namespace

extern "C" {

void orc_camera_params(const float K[9], const float R[9], float k_rinv[9], float r_kinv[9]) {
    Projector p;
    setCameraParams(p, K, R);
    std::memcpy(k_rinv, p.k_rinv, sizeof(p.k_rinv));
    std::memcpy(r_kinv, p.r_kinv, sizeof(p.r_kinv));
}

void orc_detect_roi(int proj, int src_w, int src_h, const float K[9], const float R[9],
                    float scale, int full_scan, int roi[4]) {
    Projector p;
    p.scale = scale;
    p.proj = proj;
    setCameraParams(p, K, R);
    float tl_uf = (std::numeric_limits<float>::max)();      // [WARP]:66-69
    float tl_vf = (std::numeric_limits<float>::max)();
    float br_uf = -(std::numeric_limits<float>::max)();
    float br_vf = -(std::numeric_limits<float>::max)();
    float u, v;
    auto acc = [&](int x, int y) {
        mapForward(p, static_cast<float>(x), static_cast<float>(y), u, v);
        tl_uf = (std::min)(tl_uf, u); tl_vf = (std::min)(tl_vf, v);
        br_uf = (std::max)(br_uf, u); br_vf = (std::max)(br_vf, v);
    };
    if (proj == ORC_PROJ_PLANE) {                           // PlaneWarper::detectResultRoi: the four corners
        acc(0, 0); acc(0, src_h - 1); acc(src_w - 1, 0); acc(src_w - 1, src_h - 1);
    } else if (full_scan || proj == ORC_PROJ_FISHEYE || proj == ORC_PROJ_STEREOGRAPHIC) {   // [WARP]:72-81 (min/max are order-independent); RotationWarperBase::detectResultRoi
#pragma omp parallel for schedule(static) reduction(min : tl_uf, tl_vf) reduction(max : br_uf, br_vf)
        for (int y = 0; y < src_h; ++y)
            for (int x = 0; x < src_w; ++x) {
                float uu, vv;
                mapForward(p, static_cast<float>(x), static_cast<float>(y), uu, vv);
                tl_uf = (std::min)(tl_uf, uu); tl_vf = (std::min)(tl_vf, vv);
                br_uf = (std::max)(br_uf, uu); br_vf = (std::max)(br_vf, vv);
            }
    } else {                                                // RotationWarperBase::detectResultRoiByBorder
        for (int x = 0; x < src_w; ++x) { acc(x, 0); acc(x, src_h - 1); }
        for (int y = 0; y < src_h; ++y) { acc(0, y); acc(src_w - 1, y); }
    }
    if (proj == ORC_PROJ_SPHERICAL) {
        // cv::detail::SphericalWarper::detectResultRoi: widen when a pole projects inside the image.
        // OpenCV works on the int-truncated border ROI there.
        tl_uf = static_cast<float>(static_cast<int>(tl_uf));
        tl_vf = static_cast<float>(static_cast<int>(tl_vf));
        br_uf = static_cast<float>(static_cast<int>(br_uf));
        br_vf = static_cast<float>(static_cast<int>(br_vf));
        for (int pole = 0; pole < 2; ++pole) {
            float sgn = pole ? -1.f : 1.f;
            float x = sgn * p.rinv[1], y = sgn * p.rinv[4], z = sgn * p.rinv[7];
            if (y > 0.f) {
                float x_ = (p.k[0] * x + p.k[1] * y) / z + p.k[2];
                float y_ = p.k[4] * y / z + p.k[5];
                if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) {
                    float pv = pole ? 0.f : static_cast<float>(3.1415926535897932384626433832795 * p.scale);
                    tl_uf = (std::min)(tl_uf, 0.f); tl_vf = (std::min)(tl_vf, pv);
                    br_uf = (std::max)(br_uf, 0.f); br_vf = (std::max)(br_vf, pv);
                }
            }
        }
    }
    roi[0] = static_cast<int>(tl_uf);                       // [WARP]:83-86 (truncate toward zero)
    roi[1] = static_cast<int>(tl_vf);
    roi[2] = static_cast<int>(br_uf);
    roi[3] = static_cast<int>(br_vf);
}

void orc_build_maps(int proj, const float K[9], const float R[9], float scale,
                    const int roi[4], float* xmap, float* ymap) {
    Projector p;
    p.scale = scale;
    p.proj = proj;
    setCameraParams(p, K, R);
    const int tlx = roi[0], tly = roi[1], brx = roi[2], bry = roi[3];
    const size_t w = (size_t)(brx - tlx + 1);
    float x, y;
#pragma omp parallel for schedule(static) private(x, y)
    for (int v = tly; v <= bry; ++v)                        // [WARP]:133-141
        for (int u = tlx; u <= brx; ++u) {
            mapBackward(p, static_cast<float>(u), static_cast<float>(v), x, y);
            xmap[(size_t)(v - tly) * w + (u - tlx)] = x;
            ymap[(size_t)(v - tly) * w + (u - tlx)] = y;
        }
}

void orc_remap_u8(const uint8_t* src, int src_h, int src_w, int ch, size_t src_step,
                  const float* xmap, const float* ymap, int h, int w,
                  int interp, int border, uint8_t* dst) {
    const short* wtab = bilinearTab();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const float* mx = xmap + (size_t)y * w;
        const float* my = ymap + (size_t)y * w;
        uint8_t* d = dst + (size_t)y * w * ch;
        for (int x = 0; x < w; ++x, d += ch) {
            if (interp == ORC_INTER_NEAREST) {
                // remap(): XY = saturate_cast<short>(map) ; remapNearest()
                int sx = saturate_short(cvRound(mx[x]));
                int sy = saturate_short(cvRound(my[x]));
                if ((unsigned)sx < (unsigned)src_w && (unsigned)sy < (unsigned)src_h) {
                    for (int c = 0; c < ch; ++c) d[c] = src[(size_t)sy * src_step + sx * ch + c];
                } else if (border == ORC_BORDER_CONSTANT) {
                    for (int c = 0; c < ch; ++c) d[c] = 0;
                } else {
                    sx = reflect(sx, src_w);
                    sy = reflect(sy, src_h);
                    for (int c = 0; c < ch; ++c) d[c] = src[(size_t)sy * src_step + sx * ch + c];
                }
                continue;
            }
            // INTER_LINEAR: 1/32-pixel fixed point coordinates, 15-bit weights (remapBilinear)
            int ix = cvRound(mx[x] * 32);
            int iy = cvRound(my[x] * 32);
            int sx = saturate_short(ix >> 5);
            int sy = saturate_short(iy >> 5);
            const short* wt = wtab + ((iy & 31) * 32 + (ix & 31)) * 4;
            if (border == ORC_BORDER_CONSTANT) {
                if (sx >= src_w || sx + 1 < 0 || sy >= src_h || sy + 1 < 0) {
                    for (int c = 0; c < ch; ++c) d[c] = 0;
                    continue;
                }
                bool x0 = (unsigned)sx < (unsigned)src_w, x1 = (unsigned)(sx + 1) < (unsigned)src_w;
                bool y0 = (unsigned)sy < (unsigned)src_h, y1 = (unsigned)(sy + 1) < (unsigned)src_h;
                for (int c = 0; c < ch; ++c) {
                    int v0 = (x0 && y0) ? src[(size_t)sy * src_step + sx * ch + c] : 0;
                    int v1 = (x1 && y0) ? src[(size_t)sy * src_step + (sx + 1) * ch + c] : 0;
                    int v2 = (x0 && y1) ? src[(size_t)(sy + 1) * src_step + sx * ch + c] : 0;
                    int v3 = (x1 && y1) ? src[(size_t)(sy + 1) * src_step + (sx + 1) * ch + c] : 0;
                    int s = v0 * wt[0] + v1 * wt[1] + v2 * wt[2] + v3 * wt[3];
                    int r = (s + (1 << 14)) >> 15;
                    d[c] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
                }
                continue;
            }
            int sx0 = reflect(sx, src_w), sx1 = reflect(sx + 1, src_w);
            int sy0 = reflect(sy, src_h), sy1 = reflect(sy + 1, src_h);
            const uint8_t* r0 = src + (size_t)sy0 * src_step;
            const uint8_t* r1 = src + (size_t)sy1 * src_step;
            for (int c = 0; c < ch; ++c) {
                int s = r0[sx0 * ch + c] * wt[0] + r0[sx1 * ch + c] * wt[1] +
                        r1[sx0 * ch + c] * wt[2] + r1[sx1 * ch + c] * wt[3];
                int r = (s + (1 << 14)) >> 15;
                d[c] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
        }
    }
}

}  // extern "C"
