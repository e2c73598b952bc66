// linblend.cu -- the reference's hand-written pair blend ([BLEND]:141-717) on the GPU:
//   costV overlap cost map (:206-261) -> greedy per-row seam (:268-307) -> gray + 4-way overlap classification
//   (:311-470) -> per-row left/right scan (:483-526, warp-shuffle max reductions) -> seam-guided linear
//   weights + fix-up (:529-572) -> three-region composite (:579-711).
//
// Out-of-row reads of the reference (cv::Mat_ rows are contiguous, so they land in the neighbouring row)
// are reproduced with flat indexing; reads outside a whole buffer give 0 and image rows past the end of
// an image are skipped -- the same defined behaviour as oracle/linblend.cpp.
#include "common.cuh"

#include <algorithm>

namespace is {

// @emu-begin (tests/test_kernel_host_emulation.py compiles the marked region for the host)
struct LinGeo {
    int panoBr, panoHe, dx2, dy, dy1, dy2, height, width, IB;
    int rows1, cols1, rows2, cols2;
    int overlap;
};

static LinGeo lin_geometry(int rows1, int cols1, int rows2, int cols2, is_point tl1, is_point tl2) {
    LinGeo g;
    g.rows1 = rows1; g.cols1 = cols1; g.rows2 = rows2; g.cols2 = cols2;
    g.panoBr = tl2.x - tl1.x + cols2;                                                    // :152
    g.panoHe = std::max(tl1.y + rows1, tl2.y + rows2) - std::min(tl1.y, tl2.y);          // :153
    g.dx2 = tl2.x - tl1.x;
    g.dy = tl2.y - tl1.y;
    g.dy1 = g.dy < 0 ? -g.dy : 0;
    g.dy2 = g.dy > 0 ? g.dy : 0;
    const int itx = std::max(tl1.x, tl2.x), ity = std::max(tl1.y, tl2.y);
    const int ibx = std::min(tl1.x + cols1, tl2.x + cols2), iby = std::min(tl1.y + rows1, tl2.y + rows2);
    g.overlap = !(itx >= ibx || ity >= iby);
    g.height = iby - ity;
    g.width = ibx - itx;
    g.IB = cols1 - g.dx2;                                                                // :191
    return g;
}

struct FImg {
    const float* p; size_t step;
    __device__ __forceinline__ const float* row(int y) const { return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) + (size_t)y * step); }
};

__device__ __forceinline__ float sq(float v) { return __fmul_rn(v, v); }

// costV: panoHe x (IB + 2), zero outside the computed interior
__global__ void k_lin_cost(FImg a, FImg b, LinGeo g, float* __restrict__ costV) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int W = g.IB + 2;
    if (x >= W || y >= g.panoHe) return;
    int y0, y1, off2;
    if (g.dy > 0) { y0 = g.dy2; y1 = g.panoHe - g.dy2; off2 = g.dy2; }
    else if (g.dy < 0) { y0 = g.dy1; y1 = g.panoHe - g.dy1; off2 = g.dy1; }   // as written in the reference (:232-236)
    else { y0 = 0; y1 = min(g.rows1, g.rows2); off2 = 0; }
    float v = 0.f;
    if (y >= y0 && y < y1 && y < g.rows1 && y - off2 >= 0 && y - off2 < g.rows2 && x >= 1 && x < g.IB - 1 &&
        x + g.dx2 + 1 < g.cols1 && x < g.cols2) {
        const float* p1 = a.row(y);
        const float* p2 = b.row(y - off2);
        const float* q1 = p1 + 3 * (x + g.dx2);
        const float* q2 = p2 + 3 * x;
        float d0 = __fadd_rn(__fadd_rn(sq(__fsub_rn(q1[0], q2[0])), sq(__fsub_rn(q1[1], q2[1]))), sq(__fsub_rn(q1[2], q2[2])));
        float d1 = __fadd_rn(__fadd_rn(sq(__fsub_rn(q1[3], q2[-3])), sq(__fsub_rn(q1[4], q2[-2]))), sq(__fsub_rn(q1[5], q2[-1])));
        v = __fmul_rn(__fadd_rn(d0, d1), 0.5f);
    }
    costV[(size_t)y * W + x] = v;
}

// greedy seam: sequential over rows; one thread, window staged through shared memory by the block
__global__ void k_lin_seam(const float* __restrict__ costV, int He, int W, int start_x, int* __restrict__ seam) {
    constexpr int BT = 32, WW = 2 * BT + 3;
    __shared__ float win[BT * WW];
    __shared__ int cur;
    const long total = (long)He * W;
    if (threadIdx.x == 0) { cur = start_x; seam[0] = start_x; }
    __syncthreads();
    for (int ylo = 1; ylo < He; ylo += BT) {
        const int nrow = min(BT, He - ylo);
        const int cx = cur, wl = cx - BT - 1;
        for (int e = threadIdx.x; e < nrow * WW; e += blockDim.x) {
            int r = e / WW, c = e % WW;
            long i = (long)(ylo + r) * W + (wl + c);           // flat semantics of the continuous cv::Mat_
            win[e] = (i < 0 || i >= total) ? 0.f : costV[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int px = cx;
            for (int r = 0; r < nrow; ++r) {
                const float* p = win + r * WW + (px - wl);
                float a = p[-1], b = p[0], c = p[1];
                if (a == b && a == c) { }
                else if (a <= b && a <= c) px -= 1;
                else if (b <= a && b <= c) { }
                else if (c <= a && c <= b) px += 1;
                seam[ylo + r] = px;
            }
            cur = px;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float gray3(const float* p) {   // cvtColor RGB2GRAY on float: c0*0.299 + c1*0.587 + c2*0.114
    return __fadd_rn(__fadd_rn(__fmul_rn(p[0], 0.299f), __fmul_rn(p[1], 0.587f)), __fmul_rn(p[2], 0.114f));
}

// masks m1, m2: height x (width + 2)
__global__ void k_lin_classify(FImg a, FImg b, LinGeo g, float* __restrict__ m1, float* __restrict__ m2) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int W = g.width + 2;
    if (x >= W || y >= g.height) return;
    float v1, v2;
    if (x == 0 || x == g.width + 1) { v1 = 128.f; v2 = 128.f; }
    else {
        const float thr = g.dy == 0 ? 10.f : 20.f;
        const float g1 = gray3(a.row(g.dy > 0 ? y + g.dy2 : y) + 3 * (x + g.dx2 - 1));
        const float g2 = gray3(b.row(g.dy < 0 ? y + g.dy1 : y) + 3 * (x - 1));
        const bool on1 = g1 >= thr, on2 = g2 >= thr;
        if (on1 && on2) { v1 = 255.f; v2 = 255.f; }
        else if (on1) { v1 = 1.f; v2 = 0.f; }
        else if (on2) { v1 = 0.f; v2 = 1.f; }
        else { v1 = 1.f; v2 = 1.f; }
    }
    m1[(size_t)y * W + x] = v1;
    m2[(size_t)y * W + x] = v2;
}

// one warp per row: left = last x matching (:496 | :501), right = last x matching (:513 | :518)
__global__ void k_lin_rowscan(const float* __restrict__ m2, int height, int width, int* __restrict__ left, int* __restrict__ right) {
    const int W = width + 2;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= height) return;
    const int lane = threadIdx.x & 31;
    const long total = (long)height * W;
    auto get = [&](long x) -> float { long i = (long)y * W + x; return (i < 0 || i >= total) ? 0.f : m2[i]; };
    int l = 0, r = 0;
    for (int x = 1 + lane; x < width + 1; x += 32) {
        const float c0 = get(x - 1), c1 = get(x), c2 = get(x + 1);
        if (c1 == 255.f && c0 == 0.f && c2 == 1.f) l = x;
        if ((c1 == 255.f && c0 == 0.f && c2 == 255.f) ||
            (c0 == 128.f && c1 == 255.f && c2 == 255.f && get(x + 2) == 255.f && get(x + 3) == 255.f)) l = x;
        if (c0 == 0.f && c1 == 255.f && c2 == 1.f) r = x;
        if (c0 == 255.f && c1 == 255.f && (c2 == 1.f || c2 == 128.f)) r = x;
    }
    for (int o = 16; o; o >>= 1) {   // row reduction: "last matching x" = max over the lanes
        l = max(l, __shfl_xor_sync(0xffffffffu, l, o));
        r = max(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    if (lane == 0) { left[y] = l; right[y] = r; }
}

__global__ void k_lin_weights(float* __restrict__ m1, float* __restrict__ m2, LinGeo g, const int* __restrict__ left, const int* __restrict__ right,
                              const int* __restrict__ seam) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int W = g.width + 2;
    if (x >= g.width + 1 || y >= g.height) return;
    const size_t i = (size_t)y * W + x;
    float p1 = m1[i], p2 = m2[i];
    if (x >= 1 && p2 == 255.f) {                                               // :531-552
        const int l = left[y], r = right[y], sx = seam[y + g.dy2 + g.dy1];
        if (l && l == r) { p1 = 1.f; p2 = 0.f; }
        else if (x <= sx + 1) {
            p1 = (float)(1 - 0.5 * (x - l) / (sx + 1 - l));
            p2 = __fsub_rn(1.f, p1);
        } else if (x > sx + 1 && x <= r) {
            p1 = (float)(0.5 * (r - x) / (r - sx - 1));
            p2 = __fsub_rn(1.f, p1);
        }
    }
    if (p1 == 255.f) { p1 = 1.f; p2 = 0.f; }                                   // :560-572
    m1[i] = p1;
    m2[i] = p2;
}

__global__ void k_lin_composite(FImg a, FImg b, LinGeo g, const float* __restrict__ m1, const float* __restrict__ m2, float* pano, size_t pstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.panoBr || y >= g.panoHe) return;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (x < g.dx2) {                                       // image 1 only
        const int sy = g.dy < 0 ? y - g.dy1 : y;
        if (sy >= 0 && sy < g.rows1) { const float* p = a.row(sy) + 3 * x; o0 = p[0]; o1 = p[1]; o2 = p[2]; }
    }
    if (x >= g.cols1) {                                    // image 2 only (rows as the reference copies them, :598,:641,:684)
        const int sy = g.dy > 0 ? y - g.dy2 : y;
        if (y >= (g.dy > 0 ? g.dy2 : 0) && y < g.rows2 && sy >= 0) { const float* p = b.row(sy) + 3 * (x - g.dx2); o0 = p[0]; o1 = p[1]; o2 = p[2]; }
    }
    if (x >= g.dx2 && x < g.dx2 + g.width) {               // overlap
        const int oy = g.dy > 0 ? y - g.dy2 : (g.dy < 0 ? y - g.dy1 : y);
        if (oy >= 0 && oy < g.height) {
            const float* p1 = a.row(g.dy > 0 ? oy + g.dy2 : oy) + 3 * x;
            const float* p2 = b.row(g.dy < 0 ? oy + g.dy1 : oy) + 3 * (x - g.dx2);
            const size_t mi = (size_t)oy * (g.width + 2) + (x - g.dx2 + 1);
            const float w1 = m1[mi], w2 = m2[mi];
            o0 = __fadd_rn(__fmul_rn(p1[0], w1), __fmul_rn(p2[0], w2));
            o1 = __fadd_rn(__fmul_rn(p1[1], w1), __fmul_rn(p2[1], w2));
            o2 = __fadd_rn(__fmul_rn(p1[2], w1), __fmul_rn(p2[2], w2));
        }
    }
    float* o = reinterpret_cast<float*>(reinterpret_cast<char*>(pano) + (size_t)y * pstep) + 3 * x;
    o[0] = o0; o[1] = o1; o[2] = o2;
}

// @emu-end
}  // namespace is

using namespace is;

extern "C" {

int is_linear_blend_size(is_size size1, is_size size2, is_point tl1, is_point tl2, is_size* pano_size) {
    if (!pano_size) return IS_ERR_BAD_ARG;
    LinGeo g = lin_geometry(size1.height, size1.width, size2.height, size2.width, tl1, tl2);
    pano_size->width = g.panoBr;
    pano_size->height = g.panoHe;
    return IS_OK;
}

int is_linear_blend_pair(is_ctx* ctx, const is_mat* img1, const is_mat* img2, is_point tl1, is_point tl2, is_mat* pano, int* seam_x) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, img1, "img1"));
    IS_TRY(check_mat(ctx, img2, "img2"));
    IS_TRY(check_mat(ctx, pano, "pano"));
    IS_REQUIRE(ctx, img1->depth == IS_32F && img1->channels == 3 && img2->depth == IS_32F && img2->channels == 3, IS_ERR_BAD_ARG, "images must be CV_32FC3");
    LinGeo g = lin_geometry(img1->rows, img1->cols, img2->rows, img2->cols, tl1, tl2);
    if (!g.overlap) return 1;                                                            // [BLEND]:182-183
    IS_REQUIRE(ctx, g.dx2 >= 0, IS_ERR_BAD_ARG, "image 1 must be the left image (tl2.x >= tl1.x)");
    IS_REQUIRE(ctx, pano->depth == IS_32F && pano->channels == 3 && pano->rows == g.panoHe && pano->cols == g.panoBr, IS_ERR_BAD_ARG,
               "pano must be CV_32FC3 of is_linear_blend_size");
    DevMat a, b, p;
    IS_TRY(stage_in(ctx, img1, &a));
    IS_TRY(stage_in(ctx, img2, &b));
    IS_TRY(stage_out(ctx, pano, &p, false));
    FImg fa{a.ptr<float>(), a.step}, fb{b.ptr<float>(), b.step};
    const int CW = g.IB + 2, MW = g.width + 2;
    IS_REQUIRE(ctx, CW > 0, IS_ERR_BAD_ARG, "degenerate overlap");
    DevBuf costV, seam, m1, m2, left, right;
    IS_TRY(costV.alloc(ctx, sizeof(float) * (size_t)g.panoHe * CW));
    IS_TRY(seam.alloc(ctx, sizeof(int) * (size_t)g.panoHe));
    IS_TRY(m1.alloc(ctx, sizeof(float) * (size_t)g.height * MW));
    IS_TRY(m2.alloc(ctx, sizeof(float) * (size_t)g.height * MW));
    IS_TRY(left.alloc(ctx, sizeof(int) * (size_t)g.height));
    IS_TRY(right.alloc(ctx, sizeof(int) * (size_t)g.height));
    dim3 block(32, 8);
    IS_LAUNCH(ctx, k_lin_cost, dim3(div_up(CW, 32), div_up(g.panoHe, 8)), block, 0, fa, fb, g, costV.as<float>());
    IS_LAUNCH(ctx, k_lin_seam, 1, 256, 0, costV.as<float>(), g.panoHe, CW, g.IB / 2, seam.as<int>());
    IS_LAUNCH(ctx, k_lin_classify, dim3(div_up(MW, 32), div_up(g.height, 8)), block, 0, fa, fb, g, m1.as<float>(), m2.as<float>());
    IS_LAUNCH(ctx, k_lin_rowscan, div_up(g.height, 8), 256, 0, m2.as<float>(), g.height, g.width, left.as<int>(), right.as<int>());
    IS_LAUNCH(ctx, k_lin_weights, dim3(div_up(g.width + 1, 32), div_up(g.height, 8)), block, 0, m1.as<float>(), m2.as<float>(), g, left.as<int>(),
              right.as<int>(), seam.as<int>());
    IS_LAUNCH(ctx, k_lin_composite, dim3(div_up(g.panoBr, 32), div_up(g.panoHe, 8)), block, 0, fa, fb, g, m1.as<float>(), m2.as<float>(),
              p.ptr<float>(), p.step);
    if (seam_x) IS_TRY(download(ctx, seam_x, seam.p, sizeof(int) * (size_t)g.panoHe));
    IS_TRY(commit(ctx, &p));
    return IS_OK;
}

}  // extern "C"
