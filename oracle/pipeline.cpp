// oracle/pipeline.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// The reference's composite call sequence on the CPU, stage by stage, as every main() runs it
// ([BLEND]:99-110 warp loop, [SEAM]:1188-1192 convertTo(CV_32F) + find, [SEAM]:1263 convertTo(CV_16S),
// blender prepare/feed/blend [SEAM]:1252,1271,1280 with the multi-band blender of [SEAM]:1244-1246).
// Used by tests as the end-to-end checker and by bench.py as the timed CPU baseline ("port").
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// corners_xy[2n], sizes_wh[2n], pano_roi[4] (x,y,w,h) = cv::detail::resultRoi(corners, sizes)
int orc_pipeline_plan(int n, int proj, const int* src_rows, const int* src_cols, const float* K, const float* R,
                      float scale, int* corners_xy, int* sizes_wh, int* pano_roi) {
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        int roi[4];
        orc_detect_roi(proj, src_cols[i], src_rows[i], K + 9 * i, R + 9 * i, scale, 0, roi);
        corners_xy[2 * i] = roi[0];
        corners_xy[2 * i + 1] = roi[1];
        sizes_wh[2 * i] = roi[2] - roi[0] + 1;        // dst.create(roi.height + 1, roi.width + 1)  [WARP]:150
        sizes_wh[2 * i + 1] = roi[3] - roi[1] + 1;
        tlx = std::min(tlx, roi[0]); tly = std::min(tly, roi[1]);
        brx = std::max(brx, roi[0] + sizes_wh[2 * i]); bry = std::max(bry, roi[1] + sizes_wh[2 * i + 1]);
    }
    pano_roi[0] = tlx; pano_roi[1] = tly; pano_roi[2] = brx - tlx; pano_roi[3] = bry - tly;
    return 0;
}

// seam: 0 = none (warped masks fed as they are), 1 = DP seam finder (COLOR).
// warped_out / masks_out: optional arrays of n caller buffers (sizes from orc_pipeline_plan) receiving
// the warped 8-bit images and the final (post-seam) masks.  stage_seconds[4] = warp, seam, blend, total.
int orc_pipeline_run(int n, int proj, const uint8_t* const* srcs, const int* src_rows, const int* src_cols,
                     const float* K, const float* R, float scale, int seam, int num_bands, int weight_type,
                     const int* corners_xy, const int* sizes_wh, const int* pano_roi,
                     uint8_t* const* warped_out, uint8_t* const* masks_out,
                     int16_t* pano, uint8_t* pano_mask, double* stage_seconds) {
    return orc_pipeline_run_ex(n, proj, srcs, src_rows, src_cols, K, R, scale, seam, num_bands, weight_type, 0, corners_xy, sizes_wh, pano_roi,
                               warped_out, masks_out, pano, pano_mask, stage_seconds, nullptr);
}

int orc_pipeline_run_ex(int n, int proj, const uint8_t* const* srcs, const int* src_rows, const int* src_cols,
                        const float* K, const float* R, float scale, int seam, int num_bands, int weight_type, int exposure_gain,
                        const int* corners_xy, const int* sizes_wh, const int* pano_roi,
                        uint8_t* const* warped_out, uint8_t* const* masks_out,
                        int16_t* pano, uint8_t* pano_mask, double* stage_seconds, double* gains_out) {
    return orc_pipeline_run_ex2(n, proj, srcs, src_rows, src_cols, K, R, scale, seam, num_bands, weight_type, exposure_gain, 0, 0.02f, 0, corners_xy,
                                sizes_wh, pano_roi, warped_out, masks_out, pano, pano_mask, stage_seconds, gains_out);
}

// blender: 0 = multi-band (num_bands, weight_type), 1 = feather (sharpness).  seam_dilate > 0: masks = dilate(masks, d x d) & warped
// masks before the blender's feed ([SEAM]:1257-1270).  With blender = 1, seam_dilate = 20, sharpness = 0.1 this is the sequence the
// reference's mains execute.
int orc_pipeline_run_ex2(int n, int proj, const uint8_t* const* srcs, const int* src_rows, const int* src_cols,
                         const float* K, const float* R, float scale, int seam, int num_bands, int weight_type, int exposure_gain,
                         int blender, float sharpness, int seam_dilate,
                         const int* corners_xy, const int* sizes_wh, const int* pano_roi,
                         uint8_t* const* warped_out, uint8_t* const* masks_out,
                         int16_t* pano, uint8_t* pano_mask, double* stage_seconds, double* gains_out) {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    std::vector<std::vector<uint8_t>> warped(n), masks(n);
    for (int i = 0; i < n; ++i) {                                // [BLEND]:100-110
        const int w = sizes_wh[2 * i], h = sizes_wh[2 * i + 1];
        int roi[4] = {corners_xy[2 * i], corners_xy[2 * i + 1], corners_xy[2 * i] + w - 1, corners_xy[2 * i + 1] + h - 1};
        std::vector<float> xmap((size_t)h * w), ymap((size_t)h * w);
        orc_build_maps(proj, K + 9 * i, R + 9 * i, scale, roi, xmap.data(), ymap.data());
        warped[i].resize((size_t)h * w * 3);
        orc_remap_u8(srcs[i], src_rows[i], src_cols[i], 3, (size_t)src_cols[i] * 3, xmap.data(), ymap.data(), h, w,
                     ORC_INTER_LINEAR, ORC_BORDER_REFLECT, warped[i].data());
        // second warp call: the all-255 source mask, INTER_NEAREST + BORDER_CONSTANT (maps are rebuilt
        // by the reference; they are identical so they are reused here)
        std::vector<uint8_t> ones((size_t)src_rows[i] * src_cols[i], 255);
        masks[i].resize((size_t)h * w);
        orc_remap_u8(ones.data(), src_rows[i], src_cols[i], 1, (size_t)src_cols[i], xmap.data(), ymap.data(), h, w,
                     ORC_INTER_NEAREST, ORC_BORDER_CONSTANT, masks[i].data());
    }
    std::vector<int> rows(n), cols(n);
    for (int i = 0; i < n; ++i) { cols[i] = sizes_wh[2 * i]; rows[i] = sizes_wh[2 * i + 1]; }
    std::vector<double> gains(n, 1.0);
    if (exposure_gain) {                                         // compensator->feed(corners, images_warped, masks_warped)  [BLEND]:117-123
        std::vector<const uint8_t*> ip(n), mp(n);
        for (int i = 0; i < n; ++i) { ip[i] = warped[i].data(); mp[i] = masks[i].data(); }
        if (orc_gain_feed(n, ip.data(), mp.data(), rows.data(), cols.data(), corners_xy, gains.data())) return -1;
    }
    if (gains_out) for (int i = 0; i < n; ++i) gains_out[i] = gains[i];
    // compensator->apply(i, corners[i], images_warped[i], masks_warped[i]) in place, BEFORE convertTo(CV_32F) and find():
    // the seam finder and the blender both see the compensated images ([SEAM]:1165-1171, 1188-1192; [BLEND]:117-123, 138-140)
    if (exposure_gain)
        for (int i = 0; i < n; ++i) orc_gain_apply(warped[i].data(), warped[i].size(), gains[i]);
    auto t1 = clk::now();
    std::vector<std::vector<uint8_t>> wmask0;                    // masks_warped: the seam finder changes `masks` in place
    if (seam_dilate > 0) wmask0 = masks;
    if (seam) {                                                  // [SEAM]:1188-1192
        std::vector<std::vector<float>> imgf(n);
        std::vector<const void*> ip(n);
        std::vector<uint8_t*> mp(n);
        for (int i = 0; i < n; ++i) {
            imgf[i].resize(warped[i].size());
            const uint8_t* s = warped[i].data();
            float* d = imgf[i].data();
            const size_t cnt = warped[i].size();
#pragma omp parallel for schedule(static)
            for (size_t k = 0; k < cnt; ++k) d[k] = (float)s[k];
            ip[i] = imgf[i].data();
            mp[i] = masks[i].data();
        }
        int rc = orc_dp_seam_find(n, ip.data(), 0, rows.data(), cols.data(), corners_xy, mp.data(), seam == 2 ? ORC_COST_COLOR_GRAD : ORC_COST_COLOR,
                                  nullptr, 0, nullptr);        // seam: 1 = DP with COLOR, 2 = DP with COLOR_GRAD
        if (rc) return rc;
    }
    auto t2 = clk::now();
    orc_mb* mb = blender == 0 ? orc_mb_create(num_bands, weight_type) : nullptr;
    orc_fb* fb = blender == 1 ? orc_fb_create(sharpness) : nullptr;
    if (mb) orc_mb_prepare(mb, pano_roi); else orc_fb_prepare(fb, pano_roi);
    for (int i = 0; i < n; ++i) {                                // [SEAM]:1263,1271
        if (seam_dilate > 0) {                                   // dilate(masks_seam) & masks_warped  [SEAM]:1264-1269
            std::vector<uint8_t> dil(masks[i].size());
            orc_dilate_rect(masks[i].data(), rows[i], cols[i], seam_dilate, seam_dilate, dil.data());
            for (size_t k = 0; k < dil.size(); ++k) masks[i][k] = dil[k] & wmask0[i][k];
        }
        std::vector<int16_t> s16(warped[i].size());
        const uint8_t* s = warped[i].data();
        int16_t* d = s16.data();
        const size_t cnt = warped[i].size();
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < cnt; ++k) d[k] = (int16_t)s[k];
        if (mb) orc_mb_feed(mb, s16.data(), masks[i].data(), rows[i], cols[i], corners_xy[2 * i], corners_xy[2 * i + 1]);
        else orc_fb_feed(fb, s16.data(), masks[i].data(), rows[i], cols[i], corners_xy[2 * i], corners_xy[2 * i + 1]);
    }
    if (mb) { orc_mb_blend(mb, pano, pano_mask); orc_mb_destroy(mb); }   // [SEAM]:1280
    else { orc_fb_blend(fb, pano, pano_mask); orc_fb_destroy(fb); }
    auto t3 = clk::now();
    for (int i = 0; i < n; ++i) {
        if (warped_out && warped_out[i]) std::memcpy(warped_out[i], warped[i].data(), warped[i].size());
        if (masks_out && masks_out[i]) std::memcpy(masks_out[i], masks[i].data(), masks[i].size());
    }
    if (stage_seconds) {
        stage_seconds[0] = std::chrono::duration<double>(t1 - t0).count();
        stage_seconds[1] = std::chrono::duration<double>(t2 - t1).count();
        stage_seconds[2] = std::chrono::duration<double>(t3 - t2).count();
        stage_seconds[3] = std::chrono::duration<double>(t3 - t0).count();
    }
    return 0;
}

}  // extern "C"
