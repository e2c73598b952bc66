// imagestitch.hpp -- header-only C++ shim over the C ABI (imagestitch.h), shaped like the cv::detail
// interfaces the reference calls and re-implements, so a main() of the reference switches stage by stage:
//
//   Ptr<RotationWarper> warper = warper_creator->create(scale);                 [BLEND]:99
//   corners[i] = warper->warp(img, K, R, INTER_LINEAR, BORDER_REFLECT, dst);    [BLEND]:105   -> is::RotationWarper::warp
//   seam_finder->find(images_warped_f, corners, masks_warped);                  [SEAM]:1192   -> is::DpSeamFinder::find
//   blender->prepare(corners, sizes); feed(img_s, mask, corner); blend(r, m);   [SEAM]:1252,1271,1280 -> is::MultiBandBlender
//
// No OpenCV dependency: images are is_mat descriptors (a cv::Mat converts with is::from_cv).
// Errors become is::Error exceptions carrying the status code (the reference's CV_Assert / CV_Error throw too).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "imagestitch.h"

namespace is {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        int s = is_ctx_create(device, &h_);
        if (s != IS_OK) throw Error(s, "is_ctx_create failed (no CUDA device? there is no CPU fallback)");
    }
    ~Context() { is_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    is_ctx* get() const { return h_; }
    void check(int status) const {
        if (status < 0) throw Error(status, is_ctx_last_error(h_));
    }
    void synchronize() const { check(is_ctx_synchronize(h_)); }

private:
    is_ctx* h_ = nullptr;
};

// cv::detail::RotationWarper as restated by the reference ([WARP]:122 buildMaps, :145 warp)
class RotationWarper {
public:
    RotationWarper(Context& ctx, is_projection projection, float scale) : ctx_(ctx), proj_(projection), scale_(scale) {}
    // size the destination the reference allocates inside warp() (dst.create(roi.height + 1, roi.width + 1))
    is_point warpRoi(is_size src_size, const float K[9], const float R[9], is_size* dst_size) const {
        is_point tl{};
        ctx_.check(is_warp_roi(ctx_.get(), proj_, src_size, K, R, scale_, &tl, dst_size));
        return tl;
    }
    is_rect buildMaps(is_size src_size, const float K[9], const float R[9], is_mat& xmap, is_mat& ymap) const {
        is_rect roi{};
        ctx_.check(is_build_maps(ctx_.get(), proj_, src_size, K, R, scale_, &xmap, &ymap, &roi));
        return roi;
    }
    is_point warp(const is_mat& src, const float K[9], const float R[9], int interp_mode, int border_mode, is_mat& dst) const {
        is_point tl{};
        ctx_.check(is_warp(ctx_.get(), proj_, &src, K, R, scale_, interp_mode, border_mode, &dst, &tl));
        return tl;
    }
    // cv::remap of [WARP]:157 on its own (any maps of the destination's size)
    static void remap(Context& ctx, const is_mat& src, const is_mat& xmap, const is_mat& ymap, int interp_mode, int border_mode, is_mat& dst) {
        ctx.check(is_remap(ctx.get(), &src, &xmap, &ymap, interp_mode, border_mode, &dst));
    }
    // image (INTER_LINEAR, BORDER_REFLECT) and all-255 mask (INTER_NEAREST, BORDER_CONSTANT) of [BLEND]:105,109 in one pass
    is_point warpWithMask(const is_mat& src, const float K[9], const float R[9], is_mat& dst, is_mat& dst_mask) const {
        is_point tl{};
        ctx_.check(is_warp_with_mask(ctx_.get(), proj_, &src, K, R, scale_, &dst, &dst_mask, &tl));
        return tl;
    }

private:
    Context& ctx_;
    is_projection proj_;
    float scale_;
};

// cv::detail::DpSeamFinder == find() of [SEAM]:87; masks are modified in place
class DpSeamFinder {
public:
    explicit DpSeamFinder(Context& ctx, is_seam_cost cost = IS_COST_COLOR) : ctx_(ctx), cost_(cost) {}
    void find(const std::vector<is_mat>& src, const std::vector<is_point>& corners, std::vector<is_mat>& masks) const {
        if (src.empty()) return;   // [SEAM]:94-95
        if (src.size() != corners.size() || src.size() != masks.size()) throw Error(IS_ERR_BAD_ARG, "find: size mismatch");
        ctx_.check(is_seam_dp_find(ctx_.get(), (int)src.size(), src.data(), corners.data(), masks.data(), cost_));
    }

private:
    Context& ctx_;
    is_seam_cost cost_;
};

// cv::detail::MultiBandBlender (Blender::createDefault(MULTI_BAND) + setNumBands), [SEAM]:1244-1252,1271,1280
class MultiBandBlender {
public:
    MultiBandBlender(Context& ctx, int try_gpu = 0, int num_bands = 5, is_weight_type weight_type = IS_WEIGHT_32F) : ctx_(ctx) {
        (void)try_gpu;
        ctx_.check(is_blender_create(ctx_.get(), num_bands, weight_type, &h_));
    }
    ~MultiBandBlender() { is_blender_destroy(h_); }
    MultiBandBlender(const MultiBandBlender&) = delete;
    MultiBandBlender& operator=(const MultiBandBlender&) = delete;
    void prepare(const std::vector<is_point>& corners, const std::vector<is_size>& sizes) {
        ctx_.check(is_blender_prepare(h_, (int)corners.size(), corners.data(), sizes.data()));
    }
    void prepare(is_rect dst_roi) { ctx_.check(is_blender_prepare_roi(h_, dst_roi)); }
    int numBands() const { return is_blender_num_bands(h_); }
    is_size dstSize() const {
        is_size s{};
        ctx_.check(is_blender_dst_size(h_, &s));
        return s;
    }
    void feed(const is_mat& img, const is_mat& mask, is_point tl, int flags = IS_FEED_COPY) { ctx_.check(is_blender_feed(h_, &img, &mask, tl, flags)); }
    void blend(is_mat& dst, is_mat& dst_mask) { ctx_.check(is_blender_blend(h_, &dst, &dst_mask)); }

private:
    Context& ctx_;
    is_blender* h_ = nullptr;
};

// cv::detail::GainCompensator (ExposureCompensator::createDefault(GAIN)): feed [BLEND]:117-123, apply [SEAM]:1165-1171
class GainCompensator {
public:
    explicit GainCompensator(Context& ctx) : ctx_(ctx) {}
    void feed(const std::vector<is_point>& corners, const std::vector<is_mat>& images, const std::vector<is_mat>& masks) {
        if (images.size() != corners.size() || images.size() != masks.size()) throw Error(IS_ERR_BAD_ARG, "feed: size mismatch");
        gains_.assign(images.size(), 1.0);
        ctx_.check(is_gain_feed(ctx_.get(), (int)images.size(), corners.data(), images.data(), masks.data(), gains_.data()));
    }
    void apply(int index, is_point /*corner*/, is_mat& image, const is_mat& /*mask*/) const { ctx_.check(is_gain_apply(ctx_.get(), &image, gains_.at(index))); }
    const std::vector<double>& gains() const { return gains_; }

private:
    Context& ctx_;
    std::vector<double> gains_;
};

// dilate(mask, mask, getStructuringElement(MORPH_RECT, Size(kw, kh))); mask &= and_mask   ([SEAM]:1258-1269)
inline void dilateAnd(Context& ctx, is_mat& mask, int kw, int kh, const is_mat* and_mask = nullptr) {
    ctx.check(is_mask_dilate_and(ctx.get(), &mask, kw, kh, and_mask));
}

// cv::detail::FeatherBlender (Blender::createDefault(FEATHER) + setSharpness), the blender the mains run: [SEAM]:1249-1252,1271,1280
class FeatherBlender {
public:
    explicit FeatherBlender(Context& ctx, float sharpness = 0.02f) : ctx_(ctx) { ctx_.check(is_feather_create(ctx_.get(), sharpness, &h_)); }
    ~FeatherBlender() { is_feather_destroy(h_); }
    FeatherBlender(const FeatherBlender&) = delete;
    FeatherBlender& operator=(const FeatherBlender&) = delete;
    void prepare(const std::vector<is_point>& corners, const std::vector<is_size>& sizes) {
        ctx_.check(is_feather_prepare(h_, (int)corners.size(), corners.data(), sizes.data()));
    }
    void prepare(is_rect dst_roi) { ctx_.check(is_feather_prepare_roi(h_, dst_roi)); }
    is_size dstSize() const {
        is_size s{};
        ctx_.check(is_feather_dst_size(h_, &s));
        return s;
    }
    void feed(const is_mat& img, const is_mat& mask, is_point tl) { ctx_.check(is_feather_feed(h_, &img, &mask, tl)); }
    void blend(is_mat& dst, is_mat& dst_mask) { ctx_.check(is_feather_blend(h_, &dst, &dst_mask)); }

private:
    Context& ctx_;
    is_feather_blender* h_ = nullptr;
};

// cv::detail::ImageFeatures / OrbFeaturesFinder as the mains use them ([BLEND]:36-41; the reference's own find() is [FEAT]:948):
// OrbFeaturesFinder(grid_size = Size(3, 1), nfeatures = 1500, scaleFactor = 1.3f, nlevels = 5) hands
// nfeatures * (99 + grid.area()) / 100 / grid.area() = 510 features per grid cell to ORB ([FEAT]:39-44)
struct ImageFeatures {
    int img_idx = -1;
    is_size img_size{};
    std::vector<is_keypoint> keypoints;
    std::vector<uint8_t> descriptors;          // 32 bytes per key point (cv::Mat keypoints.size() x 32, CV_8U)
};

class OrbFeaturesFinder {
public:
    explicit OrbFeaturesFinder(Context& ctx, is_size grid_size = is_size{3, 1}, int nfeatures = 1500, float scaleFactor = 1.3f, int nlevels = 5)
        : ctx_(ctx), prm_{nfeatures * (99 + grid_size.width * grid_size.height) / 100 / (grid_size.width * grid_size.height), scaleFactor, nlevels,
                          grid_size.width, grid_size.height} {}
    void operator()(const is_mat& image, ImageFeatures& features) {
        const int cap = (2 * prm_.nfeatures + 64) * prm_.grid_width * prm_.grid_height;
        features.keypoints.resize((size_t)cap);
        features.descriptors.resize((size_t)cap * 32);
        int n = 0;
        ctx_.check(is_orb_find(ctx_.get(), &image, &prm_, features.keypoints.data(), features.descriptors.data(), cap, &n));
        if (n > cap) n = cap;
        features.keypoints.resize((size_t)n);
        features.descriptors.resize((size_t)n * 32);
        features.img_size = is_size{image.cols, image.rows};
    }

private:
    Context& ctx_;
    is_orb_params prm_;
};

// cv::imread / cv::imwrite for bitmaps ([BLEND]:31-34, 717): the caller allocates, as everywhere at this boundary
inline is_size bmpSize(Context& ctx, const std::string& path) {
    is_size s{};
    ctx.check(is_bmp_info(ctx.get(), path.c_str(), &s, nullptr));
    return s;
}
inline void imread(Context& ctx, const std::string& path, is_mat& dst) { ctx.check(is_imread_bmp(ctx.get(), path.c_str(), &dst)); }
inline void imwrite(Context& ctx, const std::string& path, const is_mat& src) { ctx.check(is_imwrite_bmp(ctx.get(), path.c_str(), &src)); }

// the whole composite sequence (every main() of the reference)
inline void stitch(Context& ctx, const std::vector<is_mat>& images, const std::vector<is_camera>& cameras, const is_pipeline_config& cfg,
                   is_mat& pano, is_mat& pano_mask, const is_registration_hooks* hooks = nullptr) {
    ctx.check(is_pipeline_run(ctx.get(), (int)images.size(), images.data(), cameras.empty() ? nullptr : cameras.data(), hooks, &cfg, &pano,
                              &pano_mask, nullptr));
}

#ifdef OPENCV_CORE_MAT_HPP
// With OpenCV headers available (the reference's build): describe a host cv::Mat without copying.
inline is_mat from_cv(const cv::Mat& m) {
    is_mat r;
    r.data = m.data; r.rows = m.rows; r.cols = m.cols; r.channels = m.channels(); r.depth = m.depth();
    r.step = m.step; r.device = -1;
    return r;
}
#endif

}  // namespace is
