// tests/emu/warp_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked regions of imagestitch_b200/csrc/internal.cuh (WarpParams, WarpPlan) and warp.cu (camera
// products, ROI scan, trig tables, map_backward, the cv::remap sampler, k_warp, k_build_maps): the whole arithmetic of
// is_warp / is_warp_with_mask / is_build_maps, driven the way warp_plan / upload_tables / launch_warp drive it.
#include "cuda_host_emul.h"

#include <limits>
#include <vector>

#include "../../include/imagestitch.h"

namespace is {
#include "warp_regions.inc"
}
using namespace is;

static inline unsigned div_up(int a, int b) { return (unsigned)((a + b - 1) / b); }

static void plan_of(int proj, int src_w, int src_h, const float* K, const float* R, float scale, WarpPlan* plan) {   // warp_plan without the cache
    Projector p;
    set_camera(K, R, &p);
    detect_roi(proj, src_w, src_h, p, scale, plan->roi);
    std::memcpy(plan->P.k_rinv, p.k_rinv, sizeof(p.k_rinv));
    plan->P.scale = scale;
    plan->P.tl_x = plan->roi[0];
    plan->P.tl_y = plan->roi[1];
    plan->P.dst_w = plan->roi[2] - plan->roi[0] + 1;
    plan->P.dst_h = plan->roi[3] - plan->roi[1] + 1;
    plan->P.src_w = src_w;
    plan->P.src_h = src_h;
}

extern "C" void emu_warp_roi(int proj, int src_w, int src_h, const float* K, const float* R, float scale, int roi[4]) {
    WarpPlan plan;
    plan_of(proj, src_w, src_h, K, R, scale, &plan);
    std::memcpy(roi, plan.roi, sizeof(plan.roi));
}

template <int PROJ, int CH, int INTERP, int BORDER, bool WITH_MASK>
static void run(const WarpPlan& plan, const float* tables, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, uint8_t* mask, size_t mstep) {
    dim3 block(WARP_BX, WARP_BY), grid(div_up(plan.P.dst_w, WARP_BX * WARP_PX), div_up(plan.P.dst_h, WARP_BY));
    emu_launch(grid, block, [&] { k_warp<PROJ, CH, INTERP, BORDER, WITH_MASK>(plan.P, tables, src, sstep, dst, dstep, mask, mstep); });
}

template <int CH, int INTERP, int BORDER, bool WITH_MASK>
static void run_remap(int dw, int dh, int sw, int sh, const float* xm, const float* ym, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, uint8_t* mask, size_t mstep) {
    dim3 block(WARP_BX, WARP_BY), grid(div_up(dw, WARP_BX * WARP_PX), div_up(dh, WARP_BY));
    const size_t step = sizeof(float) * (size_t)dw;
    emu_launch(grid, block, [&] { k_remap<CH, INTERP, BORDER, WITH_MASK>(dw, dh, sw, sh, xm, step, ym, step, src, sstep, dst, dstep, mask, mstep); });
}

// cv::remap through dense maps (is_remap / launch_remap)
extern "C" int emu_remap(const uint8_t* src, int src_h, int src_w, int ch, size_t sstep, const float* xm, const float* ym, int dh, int dw, int interp, int border,
                         uint8_t* dst, size_t dstep, uint8_t* mask, size_t mstep) {
#define EMU_RCASE(C, I, B, M) if (ch == C && interp == I && border == B && (mask != nullptr) == M) { run_remap<C, I, B, M>(dw, dh, src_w, src_h, xm, ym, src, sstep, dst, dstep, mask, mstep); return 0; }
    EMU_RCASE(3, IS_INTER_LINEAR, IS_BORDER_REFLECT, true)
    EMU_RCASE(3, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    EMU_RCASE(3, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)
    EMU_RCASE(3, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)
    EMU_RCASE(3, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
    EMU_RCASE(1, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    EMU_RCASE(1, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)
    EMU_RCASE(1, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)
    EMU_RCASE(1, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
#undef EMU_RCASE
    return -1;
}

// dst: (roi.h + 1) x (roi.w + 1) x ch, dstep bytes per row; mask (optional, ch == 3 linear/reflect only): same size, 1 channel
extern "C" int emu_warp(int proj, const uint8_t* src, int src_h, int src_w, int ch, size_t sstep, const float* K, const float* R, float scale, int interp,
                        int border, uint8_t* dst, size_t dstep, uint8_t* mask, size_t mstep) {
    WarpPlan plan;
    plan_of(proj, src_w, src_h, K, R, scale, &plan);
    if (proj_uses_maps(proj)) {                              // upload_tables + launch_warp of the per-pixel projectors
        Projector p;
        std::memcpy(p.k_rinv, plan.P.k_rinv, sizeof(p.k_rinv));
        const size_t px = (size_t)plan.P.dst_w * (size_t)plan.P.dst_h;
        std::vector<float> maps(2 * px);
        fill_map_rows(proj, p, scale, plan.P.tl_x, plan.P.tl_y, plan.P.dst_w, 0, plan.P.dst_h, maps.data(), maps.data() + px);
        return emu_remap(src, src_h, src_w, ch, sstep, maps.data(), maps.data() + px, plan.P.dst_h, plan.P.dst_w, interp, border, dst, dstep, mask, mstep);
    }
    std::vector<float> tables(2 * (size_t)plan.P.dst_w + 2 * (size_t)plan.P.dst_h);
    fill_tables(proj, plan.P.scale, plan.P.tl_x, plan.P.tl_y, plan.P.dst_w, plan.P.dst_h, tables.data());
    const float* t = tables.data();
#define EMU_CASE(PR, C, I, B, M) if (proj == PR && ch == C && interp == I && border == B && (mask != nullptr) == M) { run<PR, C, I, B, M>(plan, t, src, sstep, dst, dstep, mask, mstep); return 0; }
    EMU_CASE(IS_PROJ_CYLINDRICAL, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, true)
    EMU_CASE(IS_PROJ_SPHERICAL, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, true)
    EMU_CASE(IS_PROJ_CYLINDRICAL, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    EMU_CASE(IS_PROJ_SPHERICAL, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    EMU_CASE(IS_PROJ_CYLINDRICAL, 1, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
    EMU_CASE(IS_PROJ_SPHERICAL, 1, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
    EMU_CASE(IS_PROJ_PLANE, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, true)
    EMU_CASE(IS_PROJ_PLANE, 3, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
    EMU_CASE(IS_PROJ_PLANE, 1, IS_INTER_NEAREST, IS_BORDER_CONSTANT, false)
    EMU_CASE(IS_PROJ_CYLINDRICAL, 3, IS_INTER_LINEAR, IS_BORDER_CONSTANT, false)
    EMU_CASE(IS_PROJ_CYLINDRICAL, 3, IS_INTER_NEAREST, IS_BORDER_REFLECT, false)
    EMU_CASE(IS_PROJ_CYLINDRICAL, 1, IS_INTER_LINEAR, IS_BORDER_REFLECT, false)
#undef EMU_CASE
    return -1;
}

extern "C" void emu_build_maps(int proj, int src_w, int src_h, const float* K, const float* R, float scale, float* xmap, float* ymap) {
    WarpPlan plan;
    plan_of(proj, src_w, src_h, K, R, scale, &plan);
    if (proj_uses_maps(proj)) {
        Projector p;
        std::memcpy(p.k_rinv, plan.P.k_rinv, sizeof(p.k_rinv));
        fill_map_rows(proj, p, scale, plan.P.tl_x, plan.P.tl_y, plan.P.dst_w, 0, plan.P.dst_h, xmap, ymap);
        return;
    }
    std::vector<float> tables(2 * (size_t)plan.P.dst_w + 2 * (size_t)plan.P.dst_h);
    fill_tables(proj, plan.P.scale, plan.P.tl_x, plan.P.tl_y, plan.P.dst_w, plan.P.dst_h, tables.data());
    dim3 block(32, 8), grid(div_up(plan.P.dst_w, 32), div_up(plan.P.dst_h, 8));
    const size_t step = sizeof(float) * (size_t)plan.P.dst_w;
    if (proj == IS_PROJ_PLANE) emu_launch(grid, block, [&] { k_build_maps<IS_PROJ_PLANE>(plan.P, tables.data(), xmap, step, ymap, step); });
    else if (proj == IS_PROJ_CYLINDRICAL) emu_launch(grid, block, [&] { k_build_maps<IS_PROJ_CYLINDRICAL>(plan.P, tables.data(), xmap, step, ymap, step); });
    else emu_launch(grid, block, [&] { k_build_maps<IS_PROJ_SPHERICAL>(plan.P, tables.data(), xmap, step, ymap, step); });
}
