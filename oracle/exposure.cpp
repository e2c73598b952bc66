// oracle/exposure.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the gain exposure compensator the reference's mains run between the warp loop and the seam finder:
//   Ptr<ExposureCompensator> compensator = ExposureCompensator::createDefault(ExposureCompensator::GAIN);
//   compensator->feed(corners, images_warped, masks_warped);                    [BLEND]:117-123, [SEAM]:1165-1171
//   compensator->apply(img_idx, corners[img_idx], img_warped, mask_warped);     (compositing loop)
// The arithmetic lives in un-vendored OpenCV 3.4.2 (modules/stitching/src/exposure_compensate.cpp,
// cv::detail::GainCompensator); restated here from its published algorithm and pinned against OpenCV 4.13
// (tests/test_oracle_cv2.py): for every pair i <= j with overlapping rectangles, N = max(1, #pixels set in both masks),
// I(i,j) = mean over those pixels of sqrt(b^2 + g^2 + r^2) of image i (double, summed in raster order); then the
// normal equations with alpha = 0.01, beta = 100, solved by LU with partial pivoting; apply() = multiply(image, gain):
// saturate_cast<uchar>(double(pixel) * gain) with round-half-even.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <vector>

extern "C" {

// images: n pointers to u8 BGR (rows[i] x cols[i], dense), masks: n pointers to u8; corners_xy[2n]; gains[n] out
int orc_gain_feed(int n, const uint8_t* const* images, const uint8_t* const* masks, const int* rows, const int* cols, const int* corners_xy,
                  double* gains) {
    std::vector<double> N((size_t)n * n, 0.), I((size_t)n * n, 0.);
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            const int x0 = std::max(corners_xy[2 * i], corners_xy[2 * j]), y0 = std::max(corners_xy[2 * i + 1], corners_xy[2 * j + 1]);
            const int x1 = std::min(corners_xy[2 * i] + cols[i], corners_xy[2 * j] + cols[j]);
            const int y1 = std::min(corners_xy[2 * i + 1] + rows[i], corners_xy[2 * j + 1] + rows[j]);
            if (!(x0 < x1 && y0 < y1)) continue;                         // overlapRoi
            long long cnt = 0;
            double s1 = 0., s2 = 0.;
            for (int y = y0; y < y1; ++y) {
                const size_t o1 = (size_t)(y - corners_xy[2 * i + 1]) * cols[i] + (x0 - corners_xy[2 * i]);
                const size_t o2 = (size_t)(y - corners_xy[2 * j + 1]) * cols[j] + (x0 - corners_xy[2 * j]);
                const uint8_t *m1 = masks[i] + o1, *m2 = masks[j] + o2, *p1 = images[i] + 3 * o1, *p2 = images[j] + 3 * o2;
                for (int x = 0; x < x1 - x0; ++x) {
                    if (m1[x] != 255 || m2[x] != 255) continue;          // (submask1 == 255) & (submask2 == 255)
                    ++cnt;
                    s1 += std::sqrt((double)(p1[3 * x] * p1[3 * x] + p1[3 * x + 1] * p1[3 * x + 1] + p1[3 * x + 2] * p1[3 * x + 2]));
                    s2 += std::sqrt((double)(p2[3 * x] * p2[3 * x] + p2[3 * x + 1] * p2[3 * x + 1] + p2[3 * x + 2] * p2[3 * x + 2]));
                }
            }
            const double nn = (double)std::max<long long>(1, cnt);
            N[(size_t)i * n + j] = N[(size_t)j * n + i] = nn;
            I[(size_t)i * n + j] = s1 / nn;
            I[(size_t)j * n + i] = s2 / nn;
        }
    const double alpha = 0.01, beta = 100.;
    std::vector<double> A((size_t)n * n, 0.), b(n, 0.);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            b[i] += beta * N[(size_t)i * n + j];
            A[(size_t)i * n + i] += beta * N[(size_t)i * n + j];
            if (j == i) continue;
            A[(size_t)i * n + i] += 2 * alpha * I[(size_t)i * n + j] * I[(size_t)i * n + j] * N[(size_t)i * n + j];
            A[(size_t)i * n + j] -= 2 * alpha * I[(size_t)i * n + j] * I[(size_t)j * n + i] * N[(size_t)i * n + j];
        }
    // cv::solve(A, b, gains, DECOMP_LU): Gaussian elimination with partial pivoting
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r) if (std::fabs(A[(size_t)r * n + k]) > std::fabs(A[(size_t)piv * n + k])) piv = r;
        if (std::fabs(A[(size_t)piv * n + k]) < 1e-300) return -1;
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
    return 0;
}

// cv::multiply(image, gain, image) on CV_8UC3 with a double scalar: double product, cvRound (half to even), saturate
// (pinned against OpenCV 4.13 on all 256 values x 20000 gains: a float product differs in about 1 gain out of 2000)
void orc_gain_apply(uint8_t* image, size_t count, double gain) {
    for (size_t k = 0; k < count; ++k) {
        const double v = (double)image[k] * gain;
        const long r = lrint(v);
        image[k] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

}  // extern "C"
