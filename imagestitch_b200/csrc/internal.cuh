// internal.cuh -- declarations shared between the translation units of the library (not part of the ABI).
#pragma once

#include "common.cuh"

namespace is {

// @emu-begin (tests/test_kernel_host_emulation.py compiles the marked regions for the host)
struct WarpParams {
    float k_rinv[9];
    float scale;
    int tl_x, tl_y;          // dst top-left in panorama coordinates
    int dst_w, dst_h;
    int src_w, src_h;
};

struct WarpPlan {
    WarpParams P;
    int roi[4];              // tl_x, tl_y, br_x, br_y as detectResultRoi returns them
};

// @emu-end
// warp.cu
int warp_plan(is_ctx* ctx, int proj, int src_w, int src_h, const float* K, const float* R, float scale, WarpPlan* plan);
int warp_plan_many(is_ctx* ctx, int proj, int n, const int* src_w, const int* src_h, const float* const* K, const float* const* R, float scale, WarpPlan* plans);
int upload_tables(is_ctx* ctx, int proj, const WarpPlan& plan, DevBuf* buf);
int launch_warp(is_ctx* ctx, int proj, const WarpPlan& plan, const float* tables, const DevMat& src, int interp, int border,
                const DevMat& dst, const DevMat* mask);

int launch_remap(is_ctx* ctx, const float* xmap, size_t xstep, const float* ymap, size_t ystep, const DevMat& src, int interp, int border, const DevMat& dst,
                 const DevMat* mask);
bool warp_fusable(int proj);     // the projector's backward map comes from O(W + H) tables: k_warp_g1 can evaluate it

int launch_warp_g1(is_ctx* ctx, int proj, const WarpPlan& plan, const float* tables, const DevMat& src, const DevMat& dst, const DevMat& mask,
                   int top, int left, int height, int width, int16_t* g1);

// blend.cu
int blender_feed_image_fused(is_blender* b, int proj, const WarpPlan& plan, const float* tables, const DevMat& src, const DevMat& img, const DevMat& mask, is_point tl);
int blender_feed_image(is_blender* b, is_ctx* side, const DevMat& img, const DevMat& mask, is_point tl);
int blender_feed_weights(is_blender* b);
int blender_build_upper_levels(is_blender* b, is_ctx* side);
int blender_blend_dev(is_blender* b, const DevMat& dst, const DevMat& dmask, int sx0, int sx1);

// feather.cu
int mask_dilate_and_device(is_ctx* ctx, const DevMat& mask, int kw, int kh, const DevMat* andm);
int feather_blend_device(is_ctx* ctx, float sharpness, is_rect roi, int n, const DevMat* imgs, const DevMat* masks, const is_point* corners,
                         const DevMat& dst, const DevMat& dmask);

// exposure.cu
int gain_feed_device(is_ctx* ctx, int n, const DevMat* images, const DevMat* masks, const is_point* corners, double* gains);
int gain_apply_device(is_ctx* ctx, const DevMat& src, const DevMat& dst, double gain);

// seam.cu
int seam_find_device(is_ctx* ctx, int n, const DevMat* images, const is_point* corners, const DevMat* masks, int cost_fn = IS_COST_COLOR);

}  // namespace is
