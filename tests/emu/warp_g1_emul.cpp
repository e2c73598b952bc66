// tests/emu/warp_g1_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the fused warp + mask + Gaussian level 1 kernel (k_warp_g1 of imagestitch_b200/csrc/warp.cu) with everything it
// uses from the regions above it, run block by block on the multi-threaded emulator with the geometry launch_warp_g1 /
// blender_feed_image_fused give it: the padded frame (top, left, height, width) of MultiBandBlender::feed around the warped image.
#include "cuda_host_emul_mt.h"

#include <limits>
#include <vector>

#include "../../include/imagestitch.h"

namespace is {
#include "warp_g1_regions.inc"
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

// dst: dst_h x dst_w x 3 (dense), mask: dst_h x dst_w, g1: ((height + 1) / 2) x ((width + 1) / 2) x 3 int16; wide_ok as launch_warp_g1 sets it
extern "C" int emu_warp_g1(int proj, const uint8_t* src, int src_h, int src_w, size_t sstep, const float* K, const float* R, float scale, uint8_t* dst, uint8_t* mask,
                           int top, int left, int height, int width, int16_t* g1) {
    WarpPlan plan;
    plan_of(proj, src_w, src_h, K, R, scale, &plan);
    std::vector<float> tables(2 * (size_t)plan.P.dst_w + 2 * (size_t)plan.P.dst_h);
    fill_tables(proj, plan.P.scale, plan.P.tl_x, plan.P.tl_y, plan.P.dst_w, plan.P.dst_h, tables.data());
    WarpG1Args A;
    A.P = plan.P; A.tables = tables.data();
    A.src = src; A.sstep = sstep;
    A.dst = dst; A.dstep = 3 * (size_t)plan.P.dst_w; A.mask = mask; A.mstep = (size_t)plan.P.dst_w;
    A.top = top; A.left = left; A.height = height; A.width = width;
    A.g1 = g1; A.dh = (height + 1) / 2; A.dw = (width + 1) / 2;
    A.wide_ok = ((reinterpret_cast<uintptr_t>(src) | sstep) & 3) == 0 ? 1 : 0;
    const unsigned gx = div_up(A.dw, WG_TX), gy = div_up(A.dh, WG_TY);
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
            auto body = [&] {
                blockIdx = EmuDim3(bx, by);                          // emu_launch_mt numbers blocks along x only
                if (proj == IS_PROJ_CYLINDRICAL) k_warp_g1<IS_PROJ_CYLINDRICAL>(A);
                else if (proj == IS_PROJ_PLANE) k_warp_g1<IS_PROJ_PLANE>(A);
                else k_warp_g1<IS_PROJ_SPHERICAL>(A);
            };
            emu_launch_mt(1, WG_THREADS, body);
        }
    return 0;
}
