// oracle/blend.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the OpenCV arithmetic the reference's blend stage delegates to (the multi-band
// blender calls are present but commented out in every main: [SEAM]:1244-1246, [WARP]:271-273):
//   cv::pyrDown / cv::pyrUp            (imgproc/pyramids.cpp)
//   cv::detail::MultiBandBlender       (stitching/blenders.cpp)
//   createLaplacePyr / restoreImageFromLaplacePyr / normalizeUsingWeightMap
// Call shape = the reference's blender calls: prepare / feed / blend ([SEAM]:1252,1271,1280).
//
// Float pyrDown (CV_32F weight maps) follows the scalar association order of pyramids.cpp; OpenCV's
// SIMD build can differ from it by 1-2 ulp which is why the CV_32F-weight blend is
// compared against cv2 with a tolerance while the CV_16S-weight blend is bit-exact.
//
// Compile with -ffp-contract=off.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

inline int reflect(int p, int len) {   // BORDER_REFLECT
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

inline int16_t sat16(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

// pyrDown, 16S: 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (s + 128) >> 8
void pyrDownS16(const int16_t* src, int h, int w, int ch, int16_t* dst) {
    const int dh = (h + 1) / 2, dw = (w + 1) / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y) {
        std::vector<int> row((size_t)5 * dw * ch);
        for (int a = 0; a < 5; ++a) {
            int sy = reflect101(2 * y + a - 2, h);
            const int16_t* s = src + (size_t)sy * w * ch;
            int* r = row.data() + (size_t)a * dw * ch;
            for (int x = 0; x < dw; ++x)
                for (int c = 0; c < ch; ++c) {
                    int acc = 0;
                    for (int b = 0; b < 5; ++b) acc += k[b] * s[reflect101(2 * x + b - 2, w) * ch + c];
                    r[x * ch + c] = acc;
                }
        }
        int16_t* d = dst + (size_t)y * dw * ch;
        for (int i = 0; i < dw * ch; ++i) {
            int acc = 0;
            for (int a = 0; a < 5; ++a) acc += k[a] * row[(size_t)a * dw * ch + i];
            d[i] = sat16((acc + 128) >> 8);
        }
    }
}

// pyrUp, 16S, to (dh, dw) with dh in {2h-1, 2h, 2h+1}... (OpenCV requires |dw - 2w| == dw % 2)
// per axis: even = s[i-1] + 6 s[i] + s[i+1], odd = 4 (s[i] + s[i+1]); s[-1] -> s[1], s[n] -> s[n-1]; (t + 32) >> 6
void pyrUpS16(const int16_t* src, int h, int w, int ch, int dh, int dw, int16_t* dst) {
    auto sidx = [](int i, int n) { return i < 0 ? (n > 1 ? 1 : 0) : (i >= n ? n - 1 : i); };
    std::vector<int> tmp((size_t)h * dw * ch);   // horizontal pass
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const int16_t* s = src + (size_t)y * w * ch;
        int* t = tmp.data() + (size_t)y * dw * ch;
        for (int x = 0; x < dw; ++x) {
            int i = x >> 1;
            for (int c = 0; c < ch; ++c) {
                if ((x & 1) == 0)
                    t[x * ch + c] = s[sidx(i - 1, w) * ch + c] + 6 * s[i * ch + c] + s[sidx(i + 1, w) * ch + c];
                else
                    t[x * ch + c] = 4 * (s[i * ch + c] + s[sidx(i + 1, w) * ch + c]);
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y) {
        int i = y >> 1;
        const int* t0 = tmp.data() + (size_t)sidx(i - 1, h) * dw * ch;
        const int* t1 = tmp.data() + (size_t)i * dw * ch;
        const int* t2 = tmp.data() + (size_t)sidx(i + 1, h) * dw * ch;
        int16_t* d = dst + (size_t)y * dw * ch;
        for (int k = 0; k < dw * ch; ++k) {
            int v = (y & 1) == 0 ? t0[k] + 6 * t1[k] + t2[k] : 4 * (t1[k] + t2[k]);
            d[k] = sat16((v + 32) >> 6);
        }
    }
}

// pyrDown, 32F single channel; scalar order of pyramids.cpp:
//   row[x] = src[2x]*6 + (src[2x-1] + src[2x+1])*4 + src[2x-2] + src[2x+2]
//   dst[x] = (row2*6 + (row1 + row3)*4 + row0 + row4) * (1/256)
void pyrDownF32(const float* src, int h, int w, float* dst) {
    const int dh = (h + 1) / 2, dw = (w + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y) {
        std::vector<float> row((size_t)5 * dw);
        for (int a = 0; a < 5; ++a) {
            int sy = reflect101(2 * y + a - 2, h);
            const float* s = src + (size_t)sy * w;
            float* r = row.data() + (size_t)a * dw;
            for (int x = 0; x < dw; ++x) {
                float s0 = s[reflect101(2 * x - 2, w)], s1 = s[reflect101(2 * x - 1, w)], s2 = s[2 * x < w ? 2 * x : reflect101(2 * x, w)];
                float s3 = s[reflect101(2 * x + 1, w)], s4 = s[reflect101(2 * x + 2, w)];
                r[x] = s2 * 6 + (s1 + s3) * 4 + s0 + s4;
            }
        }
        float* d = dst + (size_t)y * dw;
        const float* r0 = row.data(), *r1 = r0 + dw, *r2 = r1 + dw, *r3 = r2 + dw, *r4 = r3 + dw;
        for (int x = 0; x < dw; ++x) d[x] = (r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x]) * (1.f / 256.f);
    }
}

struct Mat16 { int rows = 0, cols = 0; std::vector<int16_t> v; void create(int r, int c) { rows = r; cols = c; v.assign((size_t)r * c * 3, 0); } };
struct MatW {   // weight plane: float or int16 stored as float / int
    int rows = 0, cols = 0; std::vector<float> f; std::vector<int16_t> s;
    void create(int r, int c, bool isf) { rows = r; cols = c; if (isf) f.assign((size_t)r * c, 0.f); else s.assign((size_t)r * c, 0); }
};

const float WEIGHT_EPS = 1e-5f;

}  // namespace

struct orc_mb {
    int actual_num_bands, num_bands, weight_type;
    int roi_final[4];     // x,y,w,h as given
    int roi[4];           // padded
    std::vector<Mat16> dst_pyr_laplace;
    std::vector<MatW> dst_band_weights;
};

extern "C" {

void orc_pyr_down_s16(const int16_t* src, int h, int w, int ch, int16_t* dst) { pyrDownS16(src, h, w, ch, dst); }
void orc_pyr_up_s16(const int16_t* src, int h, int w, int ch, int dh, int dw, int16_t* dst) { pyrUpS16(src, h, w, ch, dh, dw, dst); }
void orc_pyr_down_f32(const float* src, int h, int w, float* dst) { pyrDownF32(src, h, w, dst); }

orc_mb* orc_mb_create(int num_bands, int weight_type) {
    orc_mb* b = new orc_mb();
    b->actual_num_bands = num_bands;
    b->num_bands = num_bands;
    b->weight_type = weight_type;
    return b;
}

void orc_mb_destroy(orc_mb* b) { delete b; }

int orc_mb_num_bands(const orc_mb* b) { return b->num_bands; }

// MultiBandBlender::prepare(Rect dst_roi)
void orc_mb_prepare(orc_mb* b, const int dst_roi[4]) {
    std::memcpy(b->roi_final, dst_roi, sizeof(int) * 4);
    std::memcpy(b->roi, dst_roi, sizeof(int) * 4);
    double max_len = static_cast<double>(std::max(dst_roi[2], dst_roi[3]));
    b->num_bands = std::min(b->actual_num_bands, static_cast<int>(std::ceil(std::log(max_len) / std::log(2.0))));
    const int nb = b->num_bands;
    b->roi[2] += ((1 << nb) - b->roi[2] % (1 << nb)) % (1 << nb);
    b->roi[3] += ((1 << nb) - b->roi[3] % (1 << nb)) % (1 << nb);
    b->dst_pyr_laplace.assign(nb + 1, Mat16());
    b->dst_band_weights.assign(nb + 1, MatW());
    const bool isf = b->weight_type == ORC_WEIGHT_32F;
    b->dst_pyr_laplace[0].create(b->roi[3], b->roi[2]);
    b->dst_band_weights[0].create(b->roi[3], b->roi[2], isf);
    for (int i = 1; i <= nb; ++i) {
        b->dst_pyr_laplace[i].create((b->dst_pyr_laplace[i - 1].rows + 1) / 2, (b->dst_pyr_laplace[i - 1].cols + 1) / 2);
        b->dst_band_weights[i].create((b->dst_band_weights[i - 1].rows + 1) / 2, (b->dst_band_weights[i - 1].cols + 1) / 2, isf);
    }
}

// MultiBandBlender::feed(img CV_16SC3, mask CV_8U, tl)
void orc_mb_feed(orc_mb* b, const int16_t* img, const uint8_t* mask, int rows, int cols, int tlx, int tly) {
    const int nb = b->num_bands;
    const int rx = b->roi[0], ry = b->roi[1], rbx = b->roi[0] + b->roi[2], rby = b->roi[1] + b->roi[3];
    int gap = 3 * (1 << nb);
    int tlnx = std::max(rx, tlx - gap), tlny = std::max(ry, tly - gap);
    int brnx = std::min(rbx, tlx + cols + gap), brny = std::min(rby, tly + rows + gap);
    tlnx = rx + (((tlnx - rx) >> nb) << nb);
    tlny = ry + (((tlny - ry) >> nb) << nb);
    int width = brnx - tlnx, height = brny - tlny;
    width += ((1 << nb) - width % (1 << nb)) % (1 << nb);
    height += ((1 << nb) - height % (1 << nb)) % (1 << nb);
    brnx = tlnx + width;
    brny = tlny + height;
    int dy = std::max(brny - rby, 0), dx = std::max(brnx - rbx, 0);
    tlnx -= dx; brnx -= dx;
    tlny -= dy; brny -= dy;
    int top = tly - tlny, left = tlx - tlnx;
    // (bottom/right follow from the padded size)

    // copyMakeBorder(img, BORDER_REFLECT)
    std::vector<std::vector<int16_t>> pyr(nb + 1);
    std::vector<int> ph(nb + 1), pw(nb + 1);
    ph[0] = height; pw[0] = width;
    pyr[0].resize((size_t)height * width * 3);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; ++y) {
        int sy = reflect(y - top, rows);
        for (int x = 0; x < width; ++x) {
            int sx = reflect(x - left, cols);
            const int16_t* s = img + ((size_t)sy * cols + sx) * 3;
            int16_t* d = pyr[0].data() + ((size_t)y * width + x) * 3;
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        }
    }
    // createLaplacePyr (16S branch)
    for (int i = 0; i < nb; ++i) {
        ph[i + 1] = (ph[i] + 1) / 2; pw[i + 1] = (pw[i] + 1) / 2;
        pyr[i + 1].resize((size_t)ph[i + 1] * pw[i + 1] * 3);
        pyrDownS16(pyr[i].data(), ph[i], pw[i], 3, pyr[i + 1].data());
    }
    {
        std::vector<int16_t> tmp;
        for (int i = 0; i < nb; ++i) {
            tmp.resize(pyr[i].size());
            pyrUpS16(pyr[i + 1].data(), ph[i + 1], pw[i + 1], 3, ph[i], pw[i], tmp.data());
            for (size_t k = 0; k < tmp.size(); ++k) pyr[i][k] = sat16((int)pyr[i][k] - (int)tmp[k]);   // cv::subtract saturates
        }
    }
    // weight pyramid
    const bool isf = b->weight_type == ORC_WEIGHT_32F;
    std::vector<std::vector<float>> wf(nb + 1);
    std::vector<std::vector<int16_t>> ws(nb + 1);
    if (isf) {
        wf[0].assign((size_t)height * width, 0.f);            // copyMakeBorder CONSTANT 0
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                // convertTo(CV_32F, 1./255.): cvtScale 8u->32f works in float: src * (float)alpha
                wf[0][(size_t)(y + top) * width + x + left] = (float)mask[(size_t)y * cols + x] * (float)(1. / 255.);
        for (int i = 0; i < nb; ++i) {
            wf[i + 1].resize((size_t)ph[i + 1] * pw[i + 1]);
            pyrDownF32(wf[i].data(), ph[i], pw[i], wf[i + 1].data());
        }
    } else {
        ws[0].assign((size_t)height * width, 0);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                int m = mask[(size_t)y * cols + x];
                ws[0][(size_t)(y + top) * width + x + left] = (int16_t)(m ? m + 1 : 0);
            }
        for (int i = 0; i < nb; ++i) {
            ws[i + 1].resize((size_t)ph[i + 1] * pw[i + 1]);
            pyrDownS16(ws[i].data(), ph[i], pw[i], 1, ws[i + 1].data());
        }
    }
    int y_tl = tlny - ry, y_br = brny - ry, x_tl = tlnx - rx, x_br = brnx - rx;
    for (int i = 0; i <= nb; ++i) {
        int rcw = x_br - x_tl, rch = y_br - y_tl;
        Mat16& D = b->dst_pyr_laplace[i];
        MatW& W = b->dst_band_weights[i];
#pragma omp parallel for schedule(static)
        for (int y = 0; y < rch; ++y)
            for (int x = 0; x < rcw; ++x) {
                const int16_t* s = pyr[i].data() + ((size_t)y * pw[i] + x) * 3;
                int16_t* d = D.v.data() + ((size_t)(y + y_tl) * D.cols + (x + x_tl)) * 3;
                if (isf) {
                    float wv = wf[i][(size_t)y * pw[i] + x];
                    d[0] = (int16_t)(d[0] + static_cast<short>(s[0] * wv));
                    d[1] = (int16_t)(d[1] + static_cast<short>(s[1] * wv));
                    d[2] = (int16_t)(d[2] + static_cast<short>(s[2] * wv));
                    W.f[(size_t)(y + y_tl) * W.cols + (x + x_tl)] += wv;
                } else {
                    int wv = ws[i][(size_t)y * pw[i] + x];
                    d[0] = (int16_t)(d[0] + short((s[0] * wv) >> 8));
                    d[1] = (int16_t)(d[1] + short((s[1] * wv) >> 8));
                    d[2] = (int16_t)(d[2] + short((s[2] * wv) >> 8));
                    int16_t& dw = W.s[(size_t)(y + y_tl) * W.cols + (x + x_tl)];
                    dw = (int16_t)(dw + wv);
                }
            }
        x_tl /= 2; y_tl /= 2;
        x_br /= 2; y_br /= 2;
    }
}

// MultiBandBlender::blend
void orc_mb_blend(orc_mb* b, int16_t* dst, uint8_t* dst_mask) {
    const int nb = b->num_bands;
    const bool isf = b->weight_type == ORC_WEIGHT_32F;
    for (int i = 0; i <= nb; ++i) {   // normalizeUsingWeightMap
        Mat16& D = b->dst_pyr_laplace[i];
        MatW& W = b->dst_band_weights[i];
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < (size_t)D.rows * D.cols; ++p) {
            int16_t* d = D.v.data() + p * 3;
            if (isf) {
                float w = W.f[p] + WEIGHT_EPS;
                d[0] = static_cast<short>(d[0] / w);
                d[1] = static_cast<short>(d[1] / w);
                d[2] = static_cast<short>(d[2] / w);
            } else {
                int w = W.s[p] + 1;
                d[0] = static_cast<short>((d[0] << 8) / w);
                d[1] = static_cast<short>((d[1] << 8) / w);
                d[2] = static_cast<short>((d[2] << 8) / w);
            }
        }
    }
    std::vector<int16_t> tmp;   // restoreImageFromLaplacePyr
    for (int i = nb; i > 0; --i) {
        Mat16& S = b->dst_pyr_laplace[i];
        Mat16& D = b->dst_pyr_laplace[i - 1];
        tmp.resize(D.v.size());
        pyrUpS16(S.v.data(), S.rows, S.cols, 3, D.rows, D.cols, tmp.data());
        for (size_t k = 0; k < tmp.size(); ++k) D.v[k] = sat16((int)tmp[k] + (int)D.v[k]);   // cv::add saturates
    }
    const int fw = b->roi_final[2], fh = b->roi_final[3];
    Mat16& D0 = b->dst_pyr_laplace[0];
    MatW& W0 = b->dst_band_weights[0];
    for (int y = 0; y < fh; ++y)
        for (int x = 0; x < fw; ++x) {
            size_t p = (size_t)y * D0.cols + x;
            bool on = isf ? (W0.f[p] > WEIGHT_EPS) : ((double)W0.s[p] > (double)WEIGHT_EPS);
            dst_mask[(size_t)y * fw + x] = on ? 255 : 0;
            for (int c = 0; c < 3; ++c) dst[((size_t)y * fw + x) * 3 + c] = on ? D0.v[p * 3 + c] : 0;
        }
}

}  // extern "C"
