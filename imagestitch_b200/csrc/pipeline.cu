// pipeline.cu -- the composite call sequence of every main() of the reference:
//   detect -> match -> homography/estimate   host control flow ([BLEND]:36-64), supplied by the caller as
//                                            cameras or through is_registration_hooks
//   warp loop                                [BLEND]:99-110   -> one fused image+mask kernel per image
//   seam                                     [SEAM]:1188-1192 -> is::seam_find_device (8-bit warped images:
//                                            identical costs to the reference's CV_32F copies)
//   blend                                    [SEAM]:1244-1280 -> multi-band blender, images borrowed in place
// The device boundary is exactly the warp loop: sources + K, R, scale go in, the panorama comes out.
#include "internal.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <mutex>

using namespace is;

// Result downloads of panoramas stitched concurrently on one device (one context per host thread) take turns on the host link:
// two 760 MB downloads at once each run at half speed and finish together, which also drags the callers into lock step
// (upload | compute | download, all at the same time).  One at a time, a download runs under the other caller's upload +
// compute and the link stays busy in both directions.
static std::mutex g_d2h_turn[64];

// IS_PIPELINE_DEBUG=1: host time stamps (ms since the first call in the process) of the phases of is_pipeline_run, per context
struct PipeStamp {
    bool on; const void* who;
    PipeStamp(const void* w) : on(getenv("IS_PIPELINE_DEBUG") != nullptr), who(w) {}
    static double now() {
        static const auto t0 = std::chrono::steady_clock::now();
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void mark(const char* what) const { if (on) fprintf(stderr, "[pipeline %p] %9.3f ms  %s\n", who, now(), what); }
};

extern "C" {

int is_pipeline_plan(is_ctx* ctx, int n, const is_size* src_sizes, const is_camera* cameras, const is_pipeline_config* cfg,
                     is_point* corners, is_size* sizes, is_rect* pano_roi) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_REQUIRE(ctx, n > 0 && src_sizes && cameras && cfg, IS_ERR_BAD_ARG, "null argument");
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    std::vector<WarpPlan> plans((size_t)n);
    {
        std::vector<int> sw((size_t)n), sh((size_t)n);
        std::vector<const float*> Kp((size_t)n), Rp((size_t)n);
        for (int i = 0; i < n; ++i) { sw[(size_t)i] = src_sizes[i].width; sh[(size_t)i] = src_sizes[i].height; Kp[(size_t)i] = cameras[i].K; Rp[(size_t)i] = cameras[i].R; }
        IS_TRY(warp_plan_many(ctx, cfg->projection, n, sw.data(), sh.data(), Kp.data(), Rp.data(), cfg->scale, plans.data()));
    }
    for (int i = 0; i < n; ++i) {
        const WarpPlan& plan = plans[(size_t)i];
        if (corners) corners[i] = is_point{plan.roi[0], plan.roi[1]};
        if (sizes) sizes[i] = is_size{plan.P.dst_w, plan.P.dst_h};
        tlx = std::min(tlx, plan.roi[0]); tly = std::min(tly, plan.roi[1]);
        brx = std::max(brx, plan.roi[0] + plan.P.dst_w); bry = std::max(bry, plan.roi[1] + plan.P.dst_h);
    }
    if (pano_roi) *pano_roi = is_rect{tlx, tly, brx - tlx, bry - tly};   // cv::detail::resultRoi(corners, sizes)
    return IS_OK;
}

// the images are validated BEFORE any hook sees them; the hooks run in the order of the mains (detect per image, match, estimate)
static int run_registration(is_ctx* ctx, int n, const is_mat* images, const is_registration_hooks* hooks, is_camera* cams, float* scale, bool* estimated) {
    for (int i = 0; i < n; ++i) {
        IS_TRY(check_mat(ctx, &images[i], "image"));
        IS_REQUIRE(ctx, images[i].depth == IS_8U && images[i].channels == 3, IS_ERR_BAD_ARG, "source images must be CV_8UC3");
    }
    *estimated = false;
    if (hooks && hooks->detect)
        for (int i = 0; i < n; ++i) {
            int rc = hooks->detect(hooks->user, i, &images[i]);
            if (rc) return fail(ctx, rc, "detect hook failed for image %d", i);
        }
    if (hooks && hooks->match) {
        int rc = hooks->match(hooks->user, n);
        if (rc) return fail(ctx, rc, "match hook failed");
    }
    if (hooks && hooks->estimate) {
        int rc = hooks->estimate(hooks->user, n, cams, scale);
        if (rc) return fail(ctx, rc, "estimate hook failed");
        *estimated = true;
    }
    return IS_OK;
}

int is_pipeline_estimate(is_ctx* ctx, int n, const is_mat* images, const is_registration_hooks* hooks, is_camera* cameras, float* scale) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_REQUIRE(ctx, n > 0 && images && hooks && hooks->estimate && cameras && scale, IS_ERR_BAD_ARG, "null argument (an estimate hook is required)");
    bool estimated = false;
    return run_registration(ctx, n, images, hooks, cameras, scale, &estimated);
}

int is_pipeline_run(is_ctx* ctx, int n, const is_mat* images, const is_camera* cameras_in, const is_registration_hooks* hooks,
                    const is_pipeline_config* cfg_in, is_mat* pano, is_mat* pano_mask, is_mat* seam_masks) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_REQUIRE(ctx, n > 0 && images && cfg_in && pano && pano_mask, IS_ERR_BAD_ARG, "null argument");
    is_pipeline_config cfg = *cfg_in;
    const PipeStamp stamp(ctx);
    stamp.mark("enter");
    std::vector<is_camera> cams(n);
    // ---- host registration stages (control flow around the GPU path)
    bool estimated = false;
    IS_TRY(run_registration(ctx, n, images, hooks, cams.data(), &cfg.scale, &estimated));
    if (!estimated) {
        IS_REQUIRE(ctx, cameras_in, IS_ERR_BAD_ARG, "cameras are required when no estimate hook is given");
        for (int i = 0; i < n; ++i) cams[i] = cameras_in[i];
    }
    // ---- geometry
    std::vector<WarpPlan> plans(n);
    std::vector<is_point> corners(n);
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    {
        std::vector<int> sw((size_t)n), sh((size_t)n);
        std::vector<const float*> Kp((size_t)n), Rp((size_t)n);
        for (int i = 0; i < n; ++i) { sw[(size_t)i] = images[i].cols; sh[(size_t)i] = images[i].rows; Kp[(size_t)i] = cams[i].K; Rp[(size_t)i] = cams[i].R; }
        IS_TRY(warp_plan_many(ctx, cfg.projection, n, sw.data(), sh.data(), Kp.data(), Rp.data(), cfg.scale, plans.data()));
    }
    for (int i = 0; i < n; ++i) {
        corners[i] = is_point{plans[i].roi[0], plans[i].roi[1]};
        tlx = std::min(tlx, plans[i].roi[0]); tly = std::min(tly, plans[i].roi[1]);
        brx = std::max(brx, plans[i].roi[0] + plans[i].P.dst_w); bry = std::max(bry, plans[i].roi[1] + plans[i].P.dst_h);
    }
    const is_rect roi{tlx, tly, brx - tlx, bry - tly};
    IS_TRY(check_mat(ctx, pano, "pano"));
    IS_TRY(check_mat(ctx, pano_mask, "pano_mask"));
    IS_REQUIRE(ctx, pano->depth == IS_16S && pano->channels == 3 && pano->rows == roi.height && pano->cols == roi.width, IS_ERR_BAD_ARG,
               "pano must be CV_16SC3 of the planned panorama size");
    IS_REQUIRE(ctx, pano_mask->depth == IS_8U && pano_mask->channels == 1 && pano_mask->rows == roi.height && pano_mask->cols == roi.width,
               IS_ERR_BAD_ARG, "pano_mask must be CV_8U of the planned panorama size");
    if (seam_masks)
        for (int i = 0; i < n; ++i) {
            IS_TRY(check_mat(ctx, &seam_masks[i], "seam_mask"));
            IS_REQUIRE(ctx, seam_masks[i].depth == IS_8U && seam_masks[i].channels == 1 && seam_masks[i].rows == plans[i].P.dst_h &&
                                seam_masks[i].cols == plans[i].P.dst_w, IS_ERR_BAD_ARG, "seam_masks[i] must be CV_8U of the warped size");
        }
    // ---- device.  Three streams: the caller's (warp, seam, weights, blend), a copy stream that uploads host sources one
    //      image ahead of the warp, and a side stream that builds the Gaussian image pyramids (they do not depend on the
    //      seam masks) while the latency-bound seam stage leaves the SMs mostly idle.
    ctx->last_gains.clear();
    is_ctx *side = nullptr, *copy = nullptr;
    IS_TRY(child_ctx(ctx, SIDE_PYRAMID, &side));
    IS_TRY(child_ctx(ctx, SIDE_COPY, &copy));
    if (getenv("IS_PIPELINE_SERIAL")) side = ctx;           // tuning knob: everything on the caller's stream
    side->ktiming = ctx->ktiming;
    std::vector<DevMat> src(n), warped(n), masks(n), wmask0(cfg.seam_dilate > 0 ? n : 0);
    is_blender* bl = nullptr;
    IS_TRY(is_blender_create(ctx, cfg.num_bands, cfg.weight_type, &bl));
    // destroyed before the buffers above: on an error path the side streams may still be using them
    struct Guard {
        is_blender* b; is_ctx* side; is_ctx* copy;
        ~Guard() { cudaStreamSynchronize(side->stream); cudaStreamSynchronize(copy->stream); is_blender_destroy(b); }
    } guard{bl, side, copy};
    IS_TRY(is_blender_prepare_roi(bl, roi));
    const bool multiband = cfg.blender == IS_BLEND_MULTI_BAND;
    IS_REQUIRE(ctx, multiband || cfg.blender == IS_BLEND_FEATHER, IS_ERR_BAD_ARG, "unknown blender type");
    IS_REQUIRE(ctx, cfg.seam_dilate >= 0 && cfg.seam_dilate <= 64, IS_ERR_BAD_ARG, "seam_dilate must be 0..64");
    IS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    bool any_host = false;
    for (int i = 0; i < n; ++i) {
        if (images[i].device >= 0) { IS_TRY(stage_in(ctx, &images[i], &src[i])); continue; }
        any_host = true;
        IS_TRY(alloc_mat(ctx, images[i].rows, images[i].cols, images[i].channels, images[i].depth, &src[i]));
    }
    if (any_host) IS_TRY(stream_after(ctx, copy->stream, ctx->stream));   // the allocations are ordered on the caller's stream
    for (int i = 0; i < n; ++i) {
        if (images[i].device < 0) {
            IS_CUDA(ctx, cudaMemcpy2DAsync(src[i].data, src[i].step, images[i].data, images[i].step, src[i].row_bytes(), images[i].rows,
                                           cudaMemcpyHostToDevice, copy->stream));
            IS_TRY(stream_after(ctx, ctx->stream, copy->stream));
        }
        IS_TRY(alloc_mat(ctx, plans[i].P.dst_h, plans[i].P.dst_w, 3, IS_8U, &warped[i]));
        IS_TRY(alloc_mat(ctx, plans[i].P.dst_h, plans[i].P.dst_w, 1, IS_8U, &masks[i]));
        DevBuf tables;
        IS_TRY(upload_tables(ctx, cfg.projection, plans[i], &tables));
        // feed(): geometry + level 1 of the image pyramid now, the other levels on the side stream, weights after the seam stage
        const bool fused = multiband && cfg.exposure == IS_EXPOSURE_NONE && is_blender_num_bands(bl) >= 1 && warp_fusable(cfg.projection) && !getenv("IS_WARP_UNFUSED");
        if (fused) {
            IS_TRY(blender_feed_image_fused(bl, cfg.projection, plans[i], tables.as<float>(), src[i], warped[i], masks[i], corners[i]));
            continue;
        }
        IS_TRY(launch_warp(ctx, cfg.projection, plans[i], tables.as<float>(), src[i], IS_INTER_LINEAR, IS_BORDER_REFLECT, warped[i], &masks[i]));
        if (multiband && cfg.exposure == IS_EXPOSURE_NONE) IS_TRY(blender_feed_image(bl, side, warped[i], masks[i], corners[i]));
    }
    // ---- exposure: compensator->feed(corners, images_warped, masks_warped), then compensator->apply(i, ...) IN PLACE on
    //      images_warped[i] ([BLEND]:117-123, [SEAM]:1165-1171) -- before convertTo(CV_32F) and find(): the seam finder and the
    //      blender both see the compensated images.  The image pyramids follow on the side stream.
    if (cfg.exposure == IS_EXPOSURE_GAIN) {
        std::vector<double> gains(n, 1.0);
        IS_TRY(gain_feed_device(ctx, n, warped.data(), masks.data(), corners.data(), gains.data()));
        for (int i = 0; i < n; ++i) {
            IS_TRY(gain_apply_device(ctx, warped[i], warped[i], gains[i]));
            if (multiband) IS_TRY(blender_feed_image(bl, side, warped[i], masks[i], corners[i]));
        }
        ctx->last_gains = gains;
    } else {
        IS_REQUIRE(ctx, cfg.exposure == IS_EXPOSURE_NONE, IS_ERR_BAD_ARG, "unknown exposure mode");
    }
    // image levels >= 2 of all images, one launch per level on the side stream.  With a seam stage to come they are queued from
    // inside it, behind its first two kernels (is_ctx::deferred_side_work); otherwise now.
    auto upper_levels = [ctx, side, bl]() -> int {
        if (side != ctx) IS_TRY(stream_after(ctx, side->stream, ctx->stream));   // level 1 may have come from the fused warp kernel (caller's stream)
        const int rc = blender_build_upper_levels(bl, side);
        if (rc != IS_OK && ctx->last_error.empty()) ctx->last_error = side->last_error;
        return rc;
    };
    struct ClearDeferred { is_ctx* c; ~ClearDeferred() { c->deferred_side_work = nullptr; } } clear_deferred{ctx};
    if (multiband) {
        if (cfg.seam == IS_SEAM_DP && side != ctx) ctx->deferred_side_work = upper_levels;
        else IS_TRY(upper_levels());
    }
    if (cfg.seam_dilate > 0)   // masks_warped of [SEAM]:1267: the seam finder changes `masks` in place
        for (int i = 0; i < n; ++i) {
            IS_TRY(alloc_mat(ctx, masks[i].rows, masks[i].cols, 1, IS_8U, &wmask0[i]));
            IS_CUDA(ctx, cudaMemcpy2DAsync(wmask0[i].data, wmask0[i].step, masks[i].data, masks[i].step, (size_t)masks[i].cols, masks[i].rows,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
        }
    IS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    stamp.mark("warps queued");
    // ---- seam
    if (cfg.seam == IS_SEAM_DP) {
        IS_REQUIRE(ctx, cfg.seam_cost == IS_COST_COLOR || cfg.seam_cost == IS_COST_COLOR_GRAD, IS_ERR_BAD_ARG, "unknown seam cost function");
        IS_TRY(seam_find_device(ctx, n, warped.data(), corners.data(), masks.data(), cfg.seam_cost));
        if (ctx->deferred_side_work) {                       // the seam stage had nothing to query (no overlapping pair) or took another path
            std::function<int()> f = std::move(ctx->deferred_side_work);
            ctx->deferred_side_work = nullptr;
            IS_TRY(f());
        }
    } else {
        IS_REQUIRE(ctx, cfg.seam == IS_SEAM_NONE, IS_ERR_BAD_ARG, "unknown seam mode");
    }
    IS_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    stamp.mark("seam done");
    // ---- blend.  masks_seam[k] = dilate(masks_seam[k]) & masks_warped[k] ([SEAM]:1264-1269) opens a band on either side of
    //      the seam for the blender to work in; the seam masks handed back to the caller are the dilated ones, as in the mains.
    if (cfg.seam_dilate > 0)
        for (int i = 0; i < n; ++i) IS_TRY(mask_dilate_and_device(ctx, masks[i], cfg.seam_dilate, cfg.seam_dilate, &wmask0[i]));
    if (multiband) IS_TRY(blender_feed_weights(bl));
    if (side != ctx) {
        IS_TRY(stream_after(ctx, ctx->stream, side->stream));   // join: the image pyramids / compensated copies are complete
        merge_child(ctx, side);
    }
    DevMat dp, dm;
    IS_TRY(stage_out(ctx, pano, &dp, false));
    IS_TRY(stage_out(ctx, pano_mask, &dm, false));
    if (multiband) IS_TRY(blender_blend_dev(bl, dp, dm, 0, roi.width));
    else IS_TRY(feather_blend_device(ctx, cfg.sharpness, roi, n, warped.data(), masks.data(),
                                     corners.data(), dp, dm));
    IS_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    stamp.mark("blend queued");
    if (stamp.on) { cudaStreamSynchronize(ctx->stream); stamp.mark("blend done"); }
    // ---- results
    {
        const bool to_host = dp.host != nullptr || dm.host != nullptr;
        if (to_host) IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));            // wait for the blend outside the turn, hold it for the copy only
        std::unique_lock<std::mutex> turn(g_d2h_turn[ctx->device & 63], std::defer_lock);
        if (to_host) turn.lock();
        IS_TRY(commit(ctx, &dp));
        IS_TRY(commit(ctx, &dm));
    }
    stamp.mark("results copied");
    if (seam_masks)
        for (int i = 0; i < n; ++i)
            IS_CUDA(ctx, cudaMemcpy2DAsync(seam_masks[i].data, seam_masks[i].step, masks[i].data, masks[i].step, (size_t)masks[i].cols, masks[i].rows,
                                           seam_masks[i].device >= 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    for (int k = 0; k < 3; ++k) {
        if (cudaEventElapsedTime(&ms, ctx->ev[k], ctx->ev[k + 1]) == cudaSuccess) ctx->timings[k] = ms;
    }
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]) == cudaSuccess) ctx->timings[3] = ms;
    return IS_OK;
}

int is_pipeline_last_gains(is_ctx* ctx, int n, double* gains) {
    if (!ctx || !gains || n < 0) return IS_ERR_BAD_ARG;
    IS_REQUIRE(ctx, (size_t)n == ctx->last_gains.size(), IS_ERR_BAD_ARG, "the last is_pipeline_run did not compute gains for n images");
    for (int i = 0; i < n; ++i) gains[i] = ctx->last_gains[i];
    return IS_OK;
}

int is_pipeline_last_timings(is_ctx* ctx, float ms[4]) {
    if (!ctx || !ms) return IS_ERR_BAD_ARG;
    for (int k = 0; k < 4; ++k) ms[k] = ctx->timings[k];
    return IS_OK;
}

}  // extern "C"
