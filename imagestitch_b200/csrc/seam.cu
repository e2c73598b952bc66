// seam.cu -- dynamic-programming seam finder on the GPU.
//
// Replaces [SEAM]:87-1093 (find / process / findComponents / findEdges / resolveConflicts /
// getSeamTips / computeCosts / estimateSeam / updateLabelsUsingSeam), i.e. cv::detail::DpSeamFinder.
//
// Split of work:
//   device  every per-pixel loop of the reference: union-frame mask paste + contour masks ([SEAM]:153-186),
//           component labelling ([SEAM]:205-256: union-find CCL whose roots are the raster-first pixel
//           of each component, so ranking the roots reproduces floodFill's numbering), contour lists in
//           raster order (order-preserving compaction), cost maps ([SEAM]:733-803), the DP forward pass
//           and back-track ([SEAM]:846-947), the flood fill of updateLabelsUsingSeam ([SEAM]:978-981),
//           relabelling and the final mask update ([SEAM]:527-545);
//   host    the tiny irregular part: component graph / edge set / conflict loop ([SEAM]:311-546),
//           seam tips ([SEAM]:607-706) and the order-dependent walk over the contour of
//           updateLabelsUsingSeam ([SEAM]:983-1085), all O(perimeter).
//
// Exactness: labels, seams and masks are integers; DP costs are IEEE float adds in the reference's
// association order (cost + costV first, then + costH, [SEAM]:900-904) and candidates are compared
// lexicographically as (cost, step) like std::min_element over std::pair<float,int> ([SEAM]:909).
#include "internal.cuh"
#include "hostpool.h"
#include "tma.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <limits>
#include <cmath>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <thread>
#include <tuple>
#include <utility>

namespace is {

enum { ST_FIRST = 1, ST_SECOND = 2, ST_INTERS = 4 };

struct ContourRec {
    int x, y;        // union-frame coordinates
    int label;
    int nl[4];       // labels of the left, up, right, down neighbours; -1 outside the frame
    int flags;       // bit0: closeToContour(contour1mask_), bit1: closeToContour(contour2mask_)
};

// =====================================================================================================
// device kernels
// =====================================================================================================

// @emu-begin (tests/test_kernel_host_emulation.py compiles the marked regions for the host)
struct MaskView {
    const uint8_t* p; size_t step; int rows, cols; int ox, oy;   // (ox, oy): position inside the union frame
    __device__ __forceinline__ int at(int ux, int uy) const {
        int x = ux - ox, y = uy - oy;
        if ((unsigned)x >= (unsigned)cols || (unsigned)y >= (unsigned)rows) return 0;
        return p[(size_t)y * step + x];
    }
};
// @emu-end

// [SEAM]:153-186 + :205-218.  cls bits: 0 mask1, 1 mask2, 2 contour1mask_, 3 contour2mask_.
__global__ void k_classify(MaskView m1, MaskView m2, uint8_t* __restrict__ cls, int uw, int uh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= uw || y >= uh) return;
    int c = 0;
    if (m1.at(x, y)) {
        c |= 1;
        if (x == 0 || !m1.at(x - 1, y) || x == uw - 1 || !m1.at(x + 1, y) || y == 0 || !m1.at(x, y - 1) || y == uh - 1 || !m1.at(x, y + 1)) c |= 4;
    }
    if (m2.at(x, y)) {
        c |= 2;
        if (x == 0 || !m2.at(x - 1, y) || x == uw - 1 || !m2.at(x + 1, y) || y == 0 || !m2.at(x, y - 1) || y == uh - 1 || !m2.at(x, y + 1)) c |= 8;
    }
    cls[(size_t)y * uw + x] = (uint8_t)c;
}

// ---- connected components: union-find, root = smallest linear index of the component ------------------
// `klass` image: pixels are 4-connected when they carry the same non-zero (klass & kmask).

// one block per row: parent = index of the start of the horizontal run the pixel belongs to
__global__ void k_ccl_rows(const uint8_t* __restrict__ klass, int kmask, int* __restrict__ parent, int w) {
    const int y = blockIdx.x;
    const uint8_t* row = klass + (size_t)y * w;
    int* prow = parent + (size_t)y * w;
    __shared__ int warp_max[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < w; base += blockDim.x) {
        const int x = base + threadIdx.x;
        int k = 0, start = -1;
        if (x < w) {
            k = row[x] & kmask;
            int kprev = x > 0 ? (row[x - 1] & kmask) : -1;
            if (k != kprev) start = x;
        }
        // inclusive max-scan of `start` over the block
        int v = start;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = max(v, t);
        }
        if (lane == 31) warp_max[wid] = v;
        __syncthreads();
        if (wid == 0) {
            int t = lane < nw ? warp_max[lane] : -1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t = max(t, u);
            }
            warp_max[lane] = t;
        }
        __syncthreads();
        int prefix = wid > 0 ? warp_max[wid - 1] : -1;
        v = max(max(v, prefix), carry_s);
        if (x < w) prow[x] = k ? y * w + v : -1;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = v;   // run start carried into the next chunk
        __syncthreads();
    }
}

__device__ __forceinline__ int uf_find(volatile int* parent, int i) {
    int p = parent[i];
    while (p != i) { i = p; p = parent[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        int old = atomicMin(parent + b, a);
        if (old == b) return;
        b = old;
    }
}

// vertical merges, once per pair of vertically adjacent runs (at the larger of the two run starts)
__global__ void k_ccl_merge(const uint8_t* __restrict__ klass, int kmask, int* parent, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (x >= w || y >= h) return;
    const uint8_t* r1 = klass + (size_t)y * w;
    const uint8_t* r0 = r1 - w;
    const int k = r1[x] & kmask;
    if (!k || (r0[x] & kmask) != k) return;
    const bool start1 = x == 0 || (r1[x - 1] & kmask) != k;
    const bool start0 = x == 0 || (r0[x - 1] & kmask) != k;
    if (start1 || start0) uf_union(parent, y * w + x, (y - 1) * w + x);
}

__global__ void k_ccl_flatten(int* parent, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = parent[i];
    if (p < 0) return;
    // chase without writing intermediate nodes: every node ends at its root after this kernel because
    // roots are fixed points and are never modified here
    volatile int* vp = parent;
    while (true) { int q = vp[p]; if (q == p) break; p = q; }
    parent[i] = p;
}

// roots (unordered): out[2k] = index, out[2k+1] = klass at the root
__global__ void k_collect_roots(const int* __restrict__ parent, const uint8_t* __restrict__ klass, size_t n, int* out, int* count, int cap) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (parent[i] == (int)i) {
        int k = atomicAdd(count, 1);
        if (k < cap) { out[2 * k] = (int)i; out[2 * k + 1] = klass[i]; }
    }
}

__global__ void k_scatter_ids(const int* __restrict__ roots, int n, int* labels) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) labels[roots[k]] = k + 1;
}

// labels[i] = id stored at the root; roots keep their value; background 0
__global__ void k_labels_from_roots(const int* __restrict__ parent, int* labels, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = parent[i];
    if (p < 0) labels[i] = 0;
    else if (p != (int)i) labels[i] = labels[p];
}

// ---- contour lists (raster order) over a window of the union frame ---------------------------------------

// The union frame of a pair.  Component labels are stored only for a window of it (the intersection rectangle grown
// by one pixel when the run-based labelling is used, the whole frame in the dense fallback): every label the
// algorithm reads after labelling lies there (INTERS components and the 4-neighbours of their pixels).
// @emu-begin
struct Frame {
    int uw, uh;              // union frame size
    int wx, wy, ww, wh;      // label window: labels[(y - wy) * ww + (x - wx)]
    MaskView m1, m2;         // the two masks placed in the frame
};

__device__ __forceinline__ size_t lidx(const Frame& f, int x, int y) { return (size_t)(y - f.wy) * f.ww + (x - f.wx); }

__device__ __forceinline__ int lab(const int* __restrict__ labels, const Frame& f, int x, int y) {
    if ((unsigned)(x - f.wx) >= (unsigned)f.ww || (unsigned)(y - f.wy) >= (unsigned)f.wh) return -2;   // never a label
    return labels[lidx(f, x, y)];
}

// @emu-end
__device__ __forceinline__ int class_at(const Frame& f, int x, int y) { return (f.m1.at(x, y) ? 1 : 0) | (f.m2.at(x, y) ? 2 : 0); }

__device__ __forceinline__ bool is_contour(const int* __restrict__ labels, const Frame& f, int x, int y, int l) {
    return (x == 0 || lab(labels, f, x - 1, y) != l) || (x == f.uw - 1 || lab(labels, f, x + 1, y) != l) ||
           (y == 0 || lab(labels, f, x, y - 1) != l) || (y == f.uh - 1 || lab(labels, f, x, y + 1) != l);
}

constexpr int CT_THREADS = 256, CT_PER_THREAD = 8, CT_CHUNK = CT_THREADS * CT_PER_THREAD;
constexpr int CT_FILTER_INTERS = -3;   // select the pixels lying in both masks (the INTERS components at labelling time)

__device__ __forceinline__ bool ct_select(int l, int fa, int fb, const Frame& f, int x, int y) {
    if (l <= 0) return false;
    if (fa == CT_FILTER_INTERS) return class_at(f, x, y) == 3;
    return fa == 0 || l == fa || l == fb;
}

// pass 1: number of selected contour pixels per chunk of the window (window-raster order)
__global__ void k_contour_count(const int* __restrict__ labels, Frame f, int wx, int wy, int ww, int wh, int fa, int fb, int* counts) {
    const size_t total = (size_t)ww * wh;
    size_t e0 = (size_t)blockIdx.x * CT_CHUNK + (size_t)threadIdx.x * CT_PER_THREAD;
    int c = 0;
    for (int k = 0; k < CT_PER_THREAD; ++k) {
        size_t e = e0 + k;
        if (e >= total) break;
        int x = wx + (int)(e % ww), y = wy + (int)(e / ww);
        int l = lab(labels, f, x, y);
        if (ct_select(l, fa, fb, f, x, y) && is_contour(labels, f, x, y, l)) ++c;
    }
    __shared__ int red[CT_THREADS / 32];
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < CT_THREADS / 32; ++i) s += red[i];
        counts[blockIdx.x] = s;
    }
}

// exclusive scan of the chunk counts (single block); offsets[n] = total
__global__ void k_scan_counts(const int* __restrict__ counts, int n, int* offsets) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < n ? counts[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_sum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int t = lane < nw ? warp_sum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
            warp_sum[lane] = t;
        }
        __syncthreads();
        int incl = s + (wid > 0 ? warp_sum[wid - 1] : 0) + carry;
        if (i < n) offsets[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

// contour{1,2}mask_ of [SEAM]:165-186, evaluated on demand: a mask pixel with a 4-neighbour outside the mask or the frame
__device__ __forceinline__ bool mask_contour(const MaskView& m, const Frame& f, int x, int y) {
    return m.at(x, y) && (x == 0 || !m.at(x - 1, y) || x == f.uw - 1 || !m.at(x + 1, y) || y == 0 || !m.at(x, y - 1) || y == f.uh - 1 || !m.at(x, y + 1));
}

__device__ __forceinline__ bool close_to(const MaskView& m, const Frame& f, int x, int y) {   // closeToContour [SEAM]:584-604
    for (int dy = -2; dy <= 2; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= f.uh) continue;
        for (int dx = -2; dx <= 2; ++dx) {
            int xx = x + dx;
            if (xx >= 0 && xx < f.uw && mask_contour(m, f, xx, yy)) return true;
        }
    }
    return false;
}

// pass 2: write the records in window-raster order
__global__ void k_contour_write(const int* __restrict__ labels, Frame f, int wx, int wy, int ww, int wh, int fa, int fb,
                                const int* __restrict__ offsets, ContourRec* out, int cap) {
    const size_t total = (size_t)ww * wh;
    size_t e0 = (size_t)blockIdx.x * CT_CHUNK + (size_t)threadIdx.x * CT_PER_THREAD;
    unsigned sel = 0;
    int c = 0;
    for (int k = 0; k < CT_PER_THREAD; ++k) {
        size_t e = e0 + k;
        if (e >= total) break;
        int x = wx + (int)(e % ww), y = wy + (int)(e / ww);
        int l = lab(labels, f, x, y);
        if (ct_select(l, fa, fb, f, x, y) && is_contour(labels, f, x, y, l)) { sel |= 1u << k; ++c; }
    }
    // exclusive scan of c over the block
    __shared__ int warp_sum[CT_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int s = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_sum[wid] = s;
    __syncthreads();
    int prefix = 0;
    for (int i = 0; i < wid; ++i) prefix += warp_sum[i];
    int pos = offsets[blockIdx.x] + prefix + s - c;
    for (int k = 0; k < CT_PER_THREAD; ++k) {
        if (!(sel & (1u << k))) continue;
        size_t e = e0 + k;
        int x = wx + (int)(e % ww), y = wy + (int)(e / ww);
        if (pos < cap) {
            ContourRec r;
            r.x = x; r.y = y; r.label = lab(labels, f, x, y);
            r.nl[0] = x > 0 ? lab(labels, f, x - 1, y) : -1;
            r.nl[1] = y > 0 ? lab(labels, f, x, y - 1) : -1;
            r.nl[2] = x < f.uw - 1 ? lab(labels, f, x + 1, y) : -1;
            r.nl[3] = y < f.uh - 1 ? lab(labels, f, x, y + 1) : -1;
            r.flags = 0;                                  // filled by k_contour_flags
            out[pos] = r;
        }
        ++pos;
    }
}

// closeToContour flags of the records: one warp per record, lanes 0..24 test one position of the 5x5 window each
__global__ void k_contour_flags(ContourRec* recs, int n, Frame f) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    const int x = recs[r].x + (lane % 5) - 2, y = recs[r].y + (lane / 5) - 2;
    const bool in = lane < 25 && x >= 0 && x < f.uw && y >= 0 && y < f.uh;
    const unsigned b1 = __ballot_sync(0xffffffffu, in && mask_contour(f.m1, f, x, y));
    const unsigned b2 = __ballot_sync(0xffffffffu, in && mask_contour(f.m2, f, x, y));
    if (lane == 0) recs[r].flags = (b1 ? 1 : 0) | (b2 ? 2 : 0);
}

__global__ void k_relabel_rect(int* labels, Frame f, int x0, int y0, int w, int h, int from, int to) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int* p = labels + lidx(f, x0 + x, y0 + y);
    if (*p == from) *p = to;
}

// ---- cost maps ------------------------------------------------------------------------------------------------

// @emu-begin
template <typename T> struct ImgView {
    const T* p; size_t step; int rows, cols; int dx, dy;   // image coords = union coords + (dx, dy)   ([SEAM]:752-753)
    int cn = 3;                                            // elements per pixel: 3, or 4 with the fourth one skipped (diffL2Square4, [SEAM]:722-730)
    __device__ __forceinline__ const T* px(int ux, int uy) const {
        return reinterpret_cast<const T*>(reinterpret_cast<const char*>(p) + (size_t)(uy + dy) * step) + cn * (ux + dx);
    }
};

// diffL2Square3 [SEAM]:713-718: ((d0^2 + d1^2) + d2^2), float
template <typename T>
__device__ __forceinline__ float diff3(const T* a, const T* b) {
    float d0 = __fsub_rn((float)a[0], (float)b[0]), d1 = __fsub_rn((float)a[1], (float)b[1]), d2 = __fsub_rn((float)a[2], (float)b[2]);
    return __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
}

#define IS_BAD_REGION_COST 195075.f   // detail::normL2(Point3f(255,255,255)): squared norm ([SEAM]:754)

__device__ __forceinline__ int label_at(const int* __restrict__ labels, Frame f, int x, int y) {
    if ((unsigned)x >= (unsigned)f.uw || (unsigned)y >= (unsigned)f.uh) return -1;
    return lab(labels, f, x, y);
}

// COLOR_GRAD ([SEAM]:549-572): Sobel gradients of the gray images over the intersection rectangle of the pair (every cell
// with the component's label lies inside it).  Window coordinates = union coordinates - (ox, oy).
struct GradView {
    const float* gx1; const float* gy1; const float* gx2; const float* gy2;
    int pitch, ox, oy;
    __device__ __forceinline__ size_t at(int ux, int uy) const { return (size_t)(uy - oy) * pitch + (ux - ox); }
};

// cvtColor(BGR2GRAY) on CV_32F in the association of OpenCV's FMA vector body (the oracle's `sobelPair`)
template <typename T>
__device__ __forceinline__ float gray_at(const ImgView<T>& a, int ix, int iy) {   // image coordinates
    const T* p = reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.p) + (size_t)iy * a.step) + a.cn * ix;
    return __fmaf_rn((float)p[2], 0.299f, __fmaf_rn((float)p[0], 0.114f, __fmul_rn((float)p[1], 0.587f)));
}

__device__ __forceinline__ int reflect101(int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); }

// Sobel(gray, CV_32F, 1, 0) and (0, 1), ksize 3, BORDER_REFLECT_101 at the IMAGE border: row filter first,
// [1 2 1] as (a + c) + 2b, [-1 0 1] as c - a.  One thread per window pixel; the 3x3 gray values are recomputed.
template <typename T>
__global__ void k_sobel_window(ImgView<T> a, int ox, int oy, int ww, int wh, float* __restrict__ gx, float* __restrict__ gy, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ww || y >= wh) return;
    const int ix = ox + x + a.dx, iy = oy + y + a.dy;
    const int xs0 = reflect101(ix - 1, a.cols), xs2 = reflect101(ix + 1, a.cols);
    const int ys[3] = {reflect101(iy - 1, a.rows), iy, reflect101(iy + 1, a.rows)};
    float d[3], sm[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float l = gray_at(a, xs0, ys[k]), c = gray_at(a, ix, ys[k]), r = gray_at(a, xs2, ys[k]);
        d[k] = __fsub_rn(r, l);
        sm[k] = __fadd_rn(__fadd_rn(l, r), __fmul_rn(c, 2.f));
    }
    gx[(size_t)y * pitch + x] = __fadd_rn(__fadd_rn(d[0], d[2]), __fmul_rn(d[1], 2.f));
    gy[(size_t)y * pitch + x] = __fsub_rn(sm[2], sm[0]);
}

template <typename T, bool GRAD>
__device__ __forceinline__ float cost_v(const ImgView<T>& a, const ImgView<T>& b, const int* __restrict__ labels, Frame f, int l, int x, int y,
                                        const GradView& g) {
    if (label_at(labels, f, x, y) == l && x > 0 && label_at(labels, f, x - 1, y) == l) {
        float c = __fmul_rn(__fadd_rn(diff3(a.px(x - 1, y), b.px(x, y)), diff3(a.px(x, y), b.px(x - 1, y))), 0.5f);
        if (GRAD) {                                                                        // [SEAM]:767-772
            const size_t i = g.at(x, y);
            const float cg = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(fabsf(g.gx1[i]), fabsf(g.gx1[i - 1])), fabsf(g.gx2[i])), fabsf(g.gx2[i - 1])), 1.f);
            c = __fdiv_rn(c, cg);
        }
        return c;
    }
    return IS_BAD_REGION_COST;
}

template <typename T, bool GRAD>
__device__ __forceinline__ float cost_h(const ImgView<T>& a, const ImgView<T>& b, const int* __restrict__ labels, Frame f, int l, int x, int y,
                                        const GradView& g) {
    if (label_at(labels, f, x, y) == l && y > 0 && label_at(labels, f, x, y - 1) == l) {
        float c = __fmul_rn(__fadd_rn(diff3(a.px(x, y - 1), b.px(x, y)), diff3(a.px(x, y), b.px(x, y - 1))), 0.5f);
        if (GRAD) {                                                                        // [SEAM]:792-797
            const size_t i = g.at(x, y);
            const float cg = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(fabsf(g.gy1[i]), fabsf(g.gy1[i - g.pitch])), fabsf(g.gy2[i])), fabsf(g.gy2[i - g.pitch])), 1.f);
            c = __fdiv_rn(c, cg);
        }
        return c;
    }
    return IS_BAD_REGION_COST;
}

// reference layout ([SEAM]:756-802): costV h x (w+1), costH (h+1) x w  -- parity entry point
template <typename T>
__global__ void k_cost_maps(ImgView<T> a, ImgView<T> b, const int* __restrict__ labels, Frame f, int l, int rx, int ry, int rw, int rh,
                            float* costV, size_t vstep, float* costH, size_t hstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x > rw || y > rh) return;
    const GradView g{};
    if (y < rh) reinterpret_cast<float*>(reinterpret_cast<char*>(costV) + (size_t)y * vstep)[x] = cost_v<T, false>(a, b, labels, f, l, rx + x, ry + y, g);
    if (x < rw) reinterpret_cast<float*>(reinterpret_cast<char*>(costH) + (size_t)y * hstep)[x] = cost_h<T, false>(a, b, labels, f, l, rx + x, ry + y, g);
}

// DP layout: P[step][lane] = cost of advancing one step at `lane`, Q[step][lane] = cost of the crossing
// between lane and lane+1 at `step`.  Vertical seam: step = y, lane = x, P = costV, Q = costH; horizontal
// seam: step = x, lane = y, P = costH, Q = costV.  P = +inf marks a cell outside the component.
template <typename T, bool GRAD>
__global__ void k_cost_pq(ImgView<T> a, ImgView<T> b, const int* __restrict__ labels, Frame f, int l, int rx, int ry, int rw, int rh,
                          int horizontal, float* __restrict__ P, float* __restrict__ Q, int pitch, GradView g) {
    const int lanes = horizontal ? rh : rw, steps = horizontal ? rw : rh;
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    const int step = blockIdx.y * blockDim.y + threadIdx.y;
    if (lane >= pitch || step >= steps) return;
    if (lane >= lanes) {   // padding lanes: outside the component
        P[(size_t)step * pitch + lane] = __int_as_float(0x7f800000);
        Q[(size_t)step * pitch + lane] = 0.f;
        return;
    }
    const int x = rx + (horizontal ? step : lane), y = ry + (horizontal ? lane : step);
    float p, q;
    if (horizontal) { p = cost_h<T, GRAD>(a, b, labels, f, l, x, y, g); q = cost_v<T, GRAD>(a, b, labels, f, l, x, y, g); }
    else { p = cost_v<T, GRAD>(a, b, labels, f, l, x, y, g); q = cost_h<T, GRAD>(a, b, labels, f, l, x, y, g); }
    if (lab(labels, f, x, y) != l) p = __int_as_float(0x7f800000);   // +inf: the cell can never be on a path
    P[(size_t)step * pitch + lane] = p;
    Q[(size_t)step * pitch + lane] = q;
}

// @emu-end
// ---- DP forward pass + back-track: one CTA per seam --------------------------------------------------------
// Thread t owns LPT consecutive lanes.  t[] = cost of the previous step + P of the previous step (the first
// add of every candidate).  Candidates ([SEAM]:899-904 / :869-874):
//   1: t[lane]                2: t[lane-1] + Q[step][lane-1]          3: t[lane+1] + Q[step][lane]
// Unreachable cells carry +inf, which never wins against a finite candidate and never ties with one.
//
// The pass is latency bound (one dependent step per row of the overlap), so the cost rows are taken off the
// critical path: they stream from HBM/L2 into a shared-memory ring with TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx), G rows of P and Q per stage, D stages in flight, issued by one thread.  Inside a step
// neighbouring threads exchange the running cost of their edge lanes through shared memory; one
// __syncthreads per step is the only synchronisation on the chain.
// @emu-dp-begin (tests/test_kernel_host_emulation.py runs this region on the multi-threaded block emulator)
struct DpArgs {
    const float* P; const float* Q;   // [steps][pitch]
    uint8_t* control;                 // [steps][lanes]
    int lanes, pitch, steps;
    int s0, lane0;                    // source (step, lane)
    int s1, lane1;                    // destination
    int* seam_lane;                   // out: lane of the seam at step s0..s1 (index step - s0)
    int* reached;                     // out: 1 when the destination is reachable
    int G, D;                         // rows per ring stage, number of stages
};

// Rows of P, Q and control are padded to `pitch` = nt * LPT lanes.  Cells outside the component (and the padding
// lanes) carry P = +inf: whatever cost reaches them, the running value t = cost + P leaving them is +inf, so no path
// continues through them and the back-track never visits them -- the reference's `labels_ == l` test ([SEAM]:897) and
// its `x > 0` / `x < roi.width - 1` guards without a branch.  (Their own control byte is arbitrary and never read.)
template <int LPT>
__device__ __forceinline__ void seam_dp_body(const DpArgs& A) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const int nt = blockDim.x, tid = threadIdx.x;
    const int row_f = A.pitch;                             // floats per row
    const int stage_f = 2 * A.G * row_f;                   // floats per stage: G rows of P, then G rows of Q
    float* ring = reinterpret_cast<float*>(sm_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)A.D * stage_f);
    // running value of the first / last lane of every thread, double buffered (static: plain LDS/STS addressing)
    __shared__ float edgeL[2][1024];
    __shared__ float edgeR[2][1024];
    const int l0 = tid * LPT;
    const float INF = __int_as_float(0x7f800000);
    const int R = A.s1 - A.s0;                             // DP steps to run: rows s0+1 .. s1
    const int NG = (R + A.G - 1) / A.G;

    auto issue = [&](int g) {                              // thread 0: load group g into stage g % D
        const int stage = g % A.D;
        const int rows = min(A.G, R - g * A.G);
        const uint32_t bytes = (uint32_t)(rows * row_f) * (uint32_t)sizeof(float);
        const size_t goff = (size_t)(A.s0 + 1 + g * A.G) * row_f;
        mbar_expect_tx(&bars[stage], 2 * bytes);
        bulk_g2s(ring + (size_t)stage * stage_f, A.P + goff, bytes, &bars[stage]);
        bulk_g2s(ring + (size_t)stage * stage_f + A.G * row_f, A.Q + goff, bytes, &bars[stage]);
    };
    if (tid == 0) {
        for (int d = 0; d < A.D; ++d) mbar_init(&bars[d], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int g = 0; g < min(A.D, NG); ++g) issue(g);

    float t[LPT];
    // source step
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
        const int lane = l0 + j;
        const float c = lane == A.lane0 ? 0.f : INF;
        t[j] = __fadd_rn(c, A.P[(size_t)A.s0 * row_f + lane]);
    }
    // double-buffered edge exchange: this thread writes its own slots, reads its neighbours' slots
    const bool has_left = tid > 0, has_right = tid < nt - 1;
    const int tl_i = max(tid - 1, 0), tr_i = min(tid + 1, nt - 1);
    int cur = 0;
    edgeL[cur][tid] = t[0];
    edgeR[cur][tid] = t[LPT - 1];
    __syncthreads();
    uint8_t* ctl_row = A.control + (size_t)(A.s0 + 1) * row_f + l0;
    for (int g = 0; g < NG; ++g) {
        const int stage = g % A.D;
        mbar_wait(&bars[stage], (uint32_t)((g / A.D) & 1));
        const float* Pr = ring + (size_t)stage * stage_f + l0;
        const float* Qr = Pr + A.G * row_f;
        const int rows = min(A.G, R - g * A.G);
        for (int rr = 0; rr < rows; ++rr, Pr += row_f, Qr += row_f, ctl_row += row_f) {
            const float el = edgeR[cur][tl_i], er = edgeL[cur][tr_i];
            const float tl_edge = has_left ? el : INF;
            const float tr_edge = has_right ? er : INF;
            float p[LPT], q[LPT];
#pragma unroll
            for (int j = 0; j < LPT; j += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(Pr + j);
                const float4 qv = *reinterpret_cast<const float4*>(Qr + j);
                p[j] = pv.x; p[j + 1] = pv.y; p[j + 2] = pv.z; p[j + 3] = pv.w;
                q[j] = qv.x; q[j + 1] = qv.y; q[j + 2] = qv.z; q[j + 3] = qv.w;
            }
            const float qleft = has_left ? Qr[-1] : 0.f;
            float tn[LPT];
            uint32_t ctl_pack[LPT / 4];
#pragma unroll
            for (int j = 0; j < LPT / 4; ++j) ctl_pack[j] = 0;
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                const float tleft = j > 0 ? t[j - 1] : tl_edge;
                const float tright = j < LPT - 1 ? t[j + 1] : tr_edge;
                const float ql = j > 0 ? q[j - 1] : qleft;
                const float c2 = __fadd_rn(tleft, ql);
                const float c3 = __fadd_rn(tright, q[j]);
                float best = t[j]; uint32_t ctl = 1;
                if (c2 < best) { best = c2; ctl = 2; }
                if (c3 < best) { best = c3; ctl = 3; }
                tn[j] = __fadd_rn(best, p[j]);
                ctl_pack[j / 4] |= ctl << (8 * (j & 3));
            }
#pragma unroll
            for (int j = 0; j < LPT / 4; ++j) reinterpret_cast<uint32_t*>(ctl_row)[j] = ctl_pack[j];
#pragma unroll
            for (int j = 0; j < LPT; ++j) t[j] = tn[j];
            cur ^= 1;
            edgeL[cur][tid] = t[0];
            edgeR[cur][tid] = t[LPT - 1];
            __syncthreads();
        }
        // every thread is past its reads of this stage (barrier above): refill it with group g + D
        if (tid == 0 && g + A.D < NG) issue(g + A.D);
    }
    // destination reachable ([SEAM]:918) <=> its running value is finite (the destination is a cell of the component,
    // so its own P is finite)
    __shared__ int reached_s;
    if (tid == 0) reached_s = (A.s1 == A.s0) ? (A.lane1 == A.lane0) : 0;
    __syncthreads();
    if (A.s1 > A.s0 && A.lane1 >= l0 && A.lane1 < l0 + LPT) {
        float tv = INF;
#pragma unroll
        for (int j = 0; j < LPT; ++j) if (l0 + j == A.lane1) tv = t[j];
        if (tv < INF) reached_s = 1;
    }
    __syncthreads();
    if (tid == 0) *A.reached = reached_s;
    if (!reached_s) return;
    // back-track ([SEAM]:923-947), staged through shared memory in chunks of BT steps
    constexpr int BT = 32;
    uint8_t* win = reinterpret_cast<uint8_t*>(sm_raw);    // [BT][2*BT+1]; the ring is idle now
    __shared__ int cur_lane_s;
    if (tid == 0) { cur_lane_s = A.lane1; }
    __syncthreads();
    for (int shi = A.s1; shi > A.s0; shi -= BT) {
        const int slo = max(A.s0 + 1, shi - BT + 1);   // steps slo..shi
        const int cl = cur_lane_s;
        const int wl = cl - BT;                         // window lanes wl .. wl + 2*BT
        const int nrow = shi - slo + 1;
        for (int e = tid; e < nrow * (2 * BT + 1); e += nt) {
            int r = e / (2 * BT + 1), c = e % (2 * BT + 1);
            int lane = wl + c;
            win[e] = (lane >= 0 && lane < A.lanes) ? A.control[(size_t)(slo + r) * row_f + lane] : 0;
        }
        __syncthreads();
        if (tid == 0) {
            int lane = cl;
            for (int s = shi; s >= slo; --s) {
                A.seam_lane[s - A.s0] = lane;
                int c = win[(s - slo) * (2 * BT + 1) + (lane - wl)];
                if (c == 2) lane--;
                else if (c == 3) lane++;
            }
            cur_lane_s = lane;
        }
        __syncthreads();
    }
    if (tid == 0) A.seam_lane[0] = cur_lane_s;
}

template <int LPT>
__global__ void __launch_bounds__(1024) k_seam_dp(DpArgs A) { seam_dp_body<LPT>(A); }

// all seams of a call in one launch: one CTA per seam (the seams of a launch share LPT and the block size)
template <int LPT>
__global__ void __launch_bounds__(1024) k_seam_dp_batch(const DpArgs* __restrict__ table) { seam_dp_body<LPT>(table[blockIdx.x]); }

// @emu-dp-end

// ---- DP forward pass, second formulation: warp-private windows with redundant halos -------------------------------------
// The pass above pays one __syncthreads per step (~420 cycles per step measured on B200, whatever the number of warps).
// The arithmetic itself is ~11 instructions per lane and step, i.e. ~130 issue cycles per step for a 1500-lane overlap on one
// SM -- so the barrier, not the work, sets the pace.  Here every warp owns OWN = 32 LPT - 2 H lanes but computes a window of
// 32 LPT lanes: H lanes of halo on either side, recomputed redundantly from the same inputs (bit-identical by construction).
// A lane at the rim of a window has no valid neighbour, which invalidates one more halo lane per step, so after K = H steps
// the owned lanes are still exact and the halos are spent: only then do the warps meet (one __syncthreads per H steps) and
// refresh their windows from each other's owned lanes through shared memory.  Inside a window neighbours talk through warp
// shuffles; the cost rows come straight from L2 into a register ring R steps ahead (every thread reads only its own lanes'
// costs, so no staging through shared memory is needed); the control bytes of the owned lanes go out as one 32-bit store.
// The back-track is a separate, parallel pair of kernels (k_bt_compose / k_bt_walk).
constexpr int DP_L2_AHEAD = 24;                              // steps between a cost row's L2 prefetch and its use in k_seam_fwd
// The dynamic shared-memory limit of a kernel is an attribute of the FUNCTION, shared by every context and host thread of the
// process: it is always set to this one value (setting it to what a launch needs let a concurrent caller with a smaller seam
// lower it between another thread's cudaFuncSetAttribute and its launch -> "invalid argument").
constexpr int DP_SMEM_MAX = 200 * 1024;
constexpr int DP_ROW_PAD = 8;                                 // spare rows behind the cost tables (k_seam_fwd prefetches past the last step)
__device__ unsigned g_dp_inf_row[16] = {0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u,
                                        0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u};

// Thread-block clusters (CL > 1): the windows of one seam are spread over the CL CTAs of a cluster -- a CTA of four warps keeps one
// window per scheduler, so a step costs the issue slots of ONE window instead of sixteen, and a 1500-lane seam runs on four SMs.
// At an exchange the owners of a CTA's outermost H lanes also store them into the neighbouring CTA's halo slots through
// distributed shared memory (st.shared::cluster), and the block barrier becomes a cluster barrier (arrive.release /
// wait.acquire).  The arithmetic of a lane does not depend on where its window runs: the seams are the same bit for bit.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f4(const float* local, uint32_t rank, const float4& v) {   // the same offset in CTA `rank`'s shared memory
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The same store as an asynchronous one that reports its bytes to an mbarrier of the receiving CTA (st.async ...
// mbarrier::complete_tx): the receiver waits for the bytes it expects, nobody fences.  A cluster barrier's release has to drain
// the sender's outstanding global stores (the control bytes) first -- a quarter of the kernel's stall samples were that membar.
__device__ __forceinline__ void st_async_f4(const float* local, const uint64_t* local_bar, uint32_t rank, const float4& v) {
    uint32_t remote, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w), "r"(rbar)
                 : "memory");
}

template <int LPT, int H, int R, int CL, bool ASY = false>
__device__ __forceinline__ void seam_fwd_body(const DpArgs* __restrict__ table) {
    static_assert(H % LPT == 0 && H % R == 0 && LPT % 4 == 0 && LPT <= 16, "halo / ring shapes");
    extern __shared__ __align__(16) float T_sm[];             // [2][span + 2 H]: running costs of this CTA's lanes (+ halos) at the last exchange
    const DpArgs A = table[blockIdx.x / CL];
    constexpr int WIN = 32 * LPT, OWN = WIN - 2 * H, K = H;
    const int tid = threadIdx.x, lane_id = tid & 31, warp = tid >> 5;
    const int rank = CL > 1 ? (int)cluster_rank() : 0;
    const int span = CL > 1 ? (int)(blockDim.x >> 5) * OWN : A.pitch;        // lanes this CTA owns (CL == 1: the whole table)
    const int cta_base = rank * span;
    const int l0 = cta_base + warp * OWN - H + lane_id * LPT; // first lane of this thread (may lie outside the table)
    const bool live = l0 >= 0 && l0 < A.lanes;                // lanes [l0, l0 + LPT) exist (pitch is a multiple of LPT)
    const bool owned = live && lane_id * LPT >= H && lane_id * LPT < WIN - H;
    const float INF = __int_as_float(0x7f800000);
    const int tstride = span + 2 * H;
    __shared__ uint64_t halo_bar[2];                          // ASY: one per buffer, completed by the neighbours' halo bytes
    if (CL > 1) {
        if (ASY && tid == 0) { mbar_init(&halo_bar[0], 1); mbar_init(&halo_bar[1], 1); mbar_fence_init(); }
        cluster_barrier();                                    // every CTA of the cluster is running (and its barriers exist) before anyone stores into its shared memory
    }
    const size_t pitch = (size_t)A.pitch;
    const int lc = live ? l0 : 0;
    // Threads outside the table read a constant row of +inf with stride 0: their lanes stay unreachable without a branch.
    const float4* fp = live ? reinterpret_cast<const float4*>(A.P + (size_t)(A.s0 + 1) * pitch + lc) : reinterpret_cast<const float4*>(g_dp_inf_row);
    const float4* fq = live ? reinterpret_cast<const float4*>(A.Q + (size_t)(A.s0 + 1) * pitch + lc) : reinterpret_cast<const float4*>(g_dp_inf_row);
    const int row4 = live ? (A.pitch >> 2) : 0;               // float4 per row
    uint32_t* cp = reinterpret_cast<uint32_t*>(A.control + (size_t)(A.s0 + 1) * pitch + lc);
    const int crow = A.pitch >> 2;
    float t[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
        const float c = l0 + j == A.lane0 ? 0.f : INF;
        t[j] = live ? __fadd_rn(c, __ldg(A.P + (size_t)A.s0 * pitch + l0 + j)) : INF;
    }
    // one DP step from the cost rows (pv, qv) of this thread's lanes
    auto dp_step = [&](const float4* pv, const float4* qv) {
        float p[LPT], q[LPT];
#pragma unroll
        for (int v = 0; v < LPT / 4; ++v) {
            p[4 * v] = pv[v].x; p[4 * v + 1] = pv[v].y; p[4 * v + 2] = pv[v].z; p[4 * v + 3] = pv[v].w;
            q[4 * v] = qv[v].x; q[4 * v + 1] = qv[v].y; q[4 * v + 2] = qv[v].z; q[4 * v + 3] = qv[v].w;
        }
        float tl_edge = __shfl_up_sync(0xffffffffu, t[LPT - 1], 1);
        float tr_edge = __shfl_down_sync(0xffffffffu, t[0], 1);
        const float qleft = __shfl_up_sync(0xffffffffu, q[LPT - 1], 1);
        if (lane_id == 0) tl_edge = INF;
        if (lane_id == 31) tr_edge = INF;
        float tn[LPT];
        uint32_t ctl_pack[LPT / 4];
#pragma unroll
        for (int j = 0; j < LPT / 4; ++j) ctl_pack[j] = 0;
#pragma unroll
        for (int j = 0; j < LPT; ++j) {
            const float tleft = j > 0 ? t[j - 1] : tl_edge;
            const float tright = j < LPT - 1 ? t[j + 1] : tr_edge;
            const float ql = j > 0 ? q[j - 1] : qleft;
            const float c2 = __fadd_rn(tleft, ql);
            const float c3 = __fadd_rn(tright, q[j]);
            // Costs are non-negative floats or +inf (never NaN, never -0): their bit patterns order like signed integers, and
            // the DPX minimum hands back "first operand <= second" with the minimum -- (cost, step) lexicographic as in
            // std::min_element over pair<float, int> ([SEAM]:909): a later candidate wins only when strictly smaller.
            bool keep1, keep2;
            const int b1 = __vibmin_s32(__float_as_int(t[j]), __float_as_int(c2), &keep1);
            const int b2 = __vibmin_s32(b1, __float_as_int(c3), &keep2);
            uint32_t ctl = keep1 ? 1u << (8 * (j & 3)) : 2u << (8 * (j & 3));
            ctl = keep2 ? ctl : 3u << (8 * (j & 3));
            tn[j] = __fadd_rn(__int_as_float(b2), p[j]);
            ctl_pack[j / 4] |= ctl;
        }
        if (owned) {
#pragma unroll
            for (int j = 0; j < LPT / 4; ++j) cp[j] = ctl_pack[j];
        }
        cp += crow;
#pragma unroll
        for (int j = 0; j < LPT; ++j) t[j] = tn[j];
    };
    // refresh every window from the owners of its lanes (one __syncthreads)
    int buf = 0, nex = 0;                                     // nex: exchanges so far (the phase of halo_bar[buf] is nex / 2)
    auto exchange = [&]() {
        float* T = T_sm + buf * tstride + H - cta_base;       // T[lane], lane in [cta_base - H, cta_base + span + H)
        if (owned) {
#pragma unroll
            for (int v = 0; v < LPT / 4; ++v) {
                const float4 x = make_float4(t[4 * v], t[4 * v + 1], t[4 * v + 2], t[4 * v + 3]);
                reinterpret_cast<float4*>(T + l0)[v] = x;
                if (CL > 1 && !ASY) {                         // the CTA's outermost H lanes are the neighbours' halos
                    const int rel = l0 + 4 * v - cta_base;
                    if (rank > 0 && rel < H) st_cluster_f4(T + cta_base + span + rel, (uint32_t)(rank - 1), x);        // its lanes [base' + span', +H)
                    if (rank < CL - 1 && rel >= span - H) st_cluster_f4(T + cta_base - span + rel, (uint32_t)(rank + 1), x);   // its lanes [base' - H, base')
                }
            }
        }
        if (CL > 1 && !ASY) cluster_barrier(); else __syncthreads();
        if (CL > 1 && ASY) {
            // After the block barrier every thread of this CTA has read the halos of two exchanges ago: only now may a neighbour be
            // given cause to overwrite them (it sends exchange e + 1 after it has received this CTA's exchange e).  The senders
            // are fixed threads -- dead lanes travel as +inf -- so every CTA receives exactly the bytes it expects.
            uint64_t* bar = &halo_bar[buf];
            const int wl = lane_id * LPT, nw = (int)(blockDim.x >> 5);
#pragma unroll
            for (int v = 0; v < LPT / 4; ++v) {
                const float4 x = make_float4(t[4 * v], t[4 * v + 1], t[4 * v + 2], t[4 * v + 3]);
                const int rel = l0 + 4 * v - cta_base;
                if (rank > 0 && warp == 0 && wl >= H && wl < 2 * H) st_async_f4(T + cta_base + span + rel, bar, (uint32_t)(rank - 1), x);
                if (rank < CL - 1 && warp == nw - 1 && wl >= WIN - 2 * H && wl < WIN - H) st_async_f4(T + cta_base - span + rel, bar, (uint32_t)(rank + 1), x);
            }
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)(((rank > 0) + (rank < CL - 1)) * H * sizeof(float)));
            mbar_wait(bar, (uint32_t)((nex >> 1) & 1));
            ++nex;
        }
        if (live) {
#pragma unroll
            for (int v = 0; v < LPT / 4; ++v) {
                const float4 x = reinterpret_cast<const float4*>(T + l0)[v];
                t[4 * v] = x.x; t[4 * v + 1] = x.y; t[4 * v + 2] = x.z; t[4 * v + 3] = x.w;
            }
        }
        buf ^= 1;
    };
    const int nsteps = A.s1 - A.s0;
    const int rem = nsteps % K, ngroups = nsteps / K;
    // The odd steps come FIRST (plain loads, their latency exposed once per seam), so that every later group is complete and
    // runs fully unrolled over the register ring without a bounds test.
    if (rem) {
        for (int i = 0; i < rem; ++i) {
            float4 pv[LPT / 4], qv[LPT / 4];
#pragma unroll
            for (int v = 0; v < LPT / 4; ++v) { pv[v] = __ldg(fp + v); qv[v] = __ldg(fq + v); }
            fp += row4; fq += row4;
            dp_step(pv, qv);
        }
        if (ngroups) exchange();
    }
    // register ring of cost rows, R steps ahead; the fetch pointers never stop: the tables carry DP_ROW_PAD spare rows
    float4 rp[R][LPT / 4], rq[R][LPT / 4];
    // The tables come from DRAM (a call's cost maps exceed the L2): a register ring of R steps covers an L2 hit, not a DRAM access
    // (ncu: half of the stall samples were the first use of a loaded row).  So the rows are pulled into the L2 DP_L2_AHEAD steps
    // early with prefetch.global.L2 (one per 128-byte line: every 8 / LPT * 4-th lane), and the ring only has to hide the L2.
    constexpr int LINE_LANES = 128 / (4 * LPT) > 0 ? 128 / (4 * LPT) : 1;               // threads sharing a 128-byte line of a row
    const bool pf_lane = live && (lane_id % LINE_LANES) == 0;
    const size_t pf_off = (size_t)DP_L2_AHEAD * (size_t)(A.pitch >> 2);
    const float4* table_end_p = reinterpret_cast<const float4*>(A.P + (size_t)(A.s1 + 1) * pitch);
    auto fetch = [&](int slot) {
#pragma unroll
        for (int v = 0; v < LPT / 4; ++v) { rp[slot][v] = __ldg(fp + v); rq[slot][v] = __ldg(fq + v); }
        if (pf_lane && fp + pf_off < table_end_p) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fp + pf_off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fq + pf_off));
        }
        fp += row4; fq += row4;
    };
    if (ngroups) {
        if (pf_lane)                                                                      // the rows between the ring and the prefetch distance
            for (int d = R; d < DP_L2_AHEAD; ++d)
                if (fp + (size_t)d * (size_t)(A.pitch >> 2) < table_end_p) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(fp + (size_t)d * (size_t)(A.pitch >> 2)));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(fq + (size_t)d * (size_t)(A.pitch >> 2)));
                }
#pragma unroll
        for (int r = 0; r < R; ++r) fetch(r);
    }
    for (int g = 0; g < ngroups; ++g) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float4 pv[LPT / 4], qv[LPT / 4];
#pragma unroll
            for (int v = 0; v < LPT / 4; ++v) { pv[v] = rp[k % R][v]; qv[v] = rq[k % R][v]; }
            fetch(k % R);
            dp_step(pv, qv);
        }
        if (g + 1 < ngroups) exchange();
    }
    // destination reachable ([SEAM]:918) <=> its running value is finite
    if (nsteps == 0) { if (tid == 0) *A.reached = (A.lane1 == A.lane0) ? 1 : 0; }
    else if (owned && A.lane1 >= l0 && A.lane1 < l0 + LPT) {
        float tv = INF;
#pragma unroll
        for (int j = 0; j < LPT; ++j) if (l0 + j == A.lane1) tv = t[j];
        *A.reached = tv < INF ? 1 : 0;
    }
    if (CL > 1) cluster_barrier();                            // nobody leaves while a neighbour may still store into its shared memory
}

template <int LPT, int H, int R, int MAXT = 512>
__global__ void __launch_bounds__(MAXT) k_seam_fwd(const DpArgs* __restrict__ table) { seam_fwd_body<LPT, H, R, 1>(table); }

// MAXT = 128 / 256: at most four / eight windows per CTA, which leaves the registers for a ring of 16 / 8 steps.  With one or two
// windows per scheduler nothing else hides the latency of a cost row, and the time of a seam falls almost in proportion to the
// depth of the ring (measured: 1500 lanes x 4029 steps, 4 CTAs: 1.72 / 1.23 / 0.62 ms at 4 / 8 / 16 steps).
template <int LPT, int H, int R, int CL, int MAXT, bool ASY>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(MAXT) k_seam_fwd_cluster(const DpArgs* __restrict__ table) { seam_fwd_body<LPT, H, R, CL, ASY>(table); }

// Back-track ([SEAM]:923-947) in parallel.  Steps s0+1 .. s1 are cut into chunks of BT_CHUNK; k_bt_compose walks every lane
// of every chunk upwards through the control bytes (BT_CHUNK dependent byte loads per thread, all chunks of all seams at
// once) and records where it leaves the chunk; k_bt_walk chains the chunk maps from the destination (one short serial chain
// per seam) and then lets one thread per chunk write the seam lanes of its steps.
constexpr int BT_CHUNK = 64;
struct BtArgs { DpArgs A; short* map; int nchunks; };         // map[chunk][pitch]: lane at the step below the chunk, given the lane at its top step

__device__ __forceinline__ int bt_step(const uint8_t* __restrict__ control, size_t pitch, int lanes, int step, int lane) {
    const int c = control[(size_t)step * pitch + lane];
    lane += (c == 3) - (c == 2);
    return min(max(lane, 0), lanes - 1);                      // lanes off the optimal path may hold anything: stay inside the table
}

__global__ void __launch_bounds__(256) k_bt_compose(const BtArgs* __restrict__ table) {
    const BtArgs& B = table[blockIdx.z];
    const int c = blockIdx.y;
    if (c >= B.nchunks || !*B.A.reached) return;
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= B.A.lanes) return;
    const int hi = B.A.s1 - c * BT_CHUNK, lo = max(B.A.s0, hi - BT_CHUNK);   // steps hi, hi-1, ..., lo+1
    int cur = lane;
    for (int s = hi; s > lo; --s) cur = bt_step(B.A.control, (size_t)B.A.pitch, B.A.lanes, s, cur);
    B.map[(size_t)c * B.A.pitch + lane] = (short)cur;
}

__global__ void __launch_bounds__(256) k_bt_walk(const BtArgs* __restrict__ table) {
    const BtArgs& B = table[blockIdx.x];
    if (!*B.A.reached) return;
    extern __shared__ int entry[];                            // [nchunks]
    if (threadIdx.x == 0) {
        int cur = B.A.lane1;
        for (int c = 0; c < B.nchunks; ++c) { entry[c] = cur; cur = B.map[(size_t)c * B.A.pitch + cur]; }
        B.A.seam_lane[0] = cur;                               // lane at the source step
    }
    __syncthreads();
    for (int c = threadIdx.x; c < B.nchunks; c += blockDim.x) {
        const int hi = B.A.s1 - c * BT_CHUNK, lo = max(B.A.s0, hi - BT_CHUNK);
        int cur = entry[c];
        for (int s = hi; s > lo; --s) {
            B.A.seam_lane[s - B.A.s0] = cur;
            cur = bt_step(B.A.control, (size_t)B.A.pitch, B.A.lanes, s, cur);
        }
    }
}

// ---- updateLabelsUsingSeam: device part -------------------------------------------------------------------------
// sub-frame klass: 0 = not comp1, 1 = interior of comp1, 2 = painted (contour of comp1 or seam)  [SEAM]:963-970
__global__ void k_uls_class(const int* __restrict__ labels, Frame f, int l1, int bx, int by, int bw, int bh, uint8_t* klass) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= bw || y >= bh) return;
    const int ux = bx + x, uy = by + y;
    int k = 0;
    if (lab(labels, f, ux, uy) == l1) k = is_contour(labels, f, ux, uy, l1) ? 2 : 1;
    klass[(size_t)y * bw + x] = (uint8_t)k;
}

// seam pixels -> painted.  seam_lane[i] is the lane at step s0+i; (rx, ry) = bbox top-left in union coords
__global__ void k_uls_paint_seam(const int* __restrict__ seam_lane, int n, int s0, int horizontal, int bw, uint8_t* klass) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int step = s0 + i, lane = seam_lane[i];
    const int x = horizontal ? step : lane, y = horizontal ? lane : step;
    klass[(size_t)y * bw + x] = 2;
}

__device__ __forceinline__ int uls_value(const uint8_t* __restrict__ klass, const int* __restrict__ parent, int bw, int bh, int x, int y) {
    if ((unsigned)x >= (unsigned)bw || (unsigned)y >= (unsigned)bh) return -1;   // outside the mask
    const size_t i = (size_t)y * bw + x;
    const int k = klass[i];
    if (k == 0) return 0;
    if (k == 2) return -255;                                                      // still "255" in the reference's mask
    return parent[i] + 1;                                                         // root index + 1 of the interior component
}

// for every contour pixel of comp1 (bbox coordinates in pts): the 8 neighbours in the reference's order
// dx = {-1,+1,0,0,-1,+1,-1,+1}, dy = {0,0,-1,+1,-1,-1,+1,+1}   [SEAM]:989-990
__global__ void k_uls_gather8(const uint8_t* __restrict__ klass, const int* __restrict__ parent, int bw, int bh,
                              const int2* __restrict__ pts, int n, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dx[8] = {-1, +1, 0, 0, -1, +1, -1, +1};
    const int dy[8] = {0, 0, -1, +1, -1, -1, +1, +1};
    const int2 p = pts[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(size_t)i * 8 + j] = uls_value(klass, parent, bw, bh, p.x + dx[j], p.y + dy[j]);
}

// for every seam pixel: its position and the value of the neighbour the reference inspects ([SEAM]:1016,1029)
__global__ void k_uls_gather_seam(const uint8_t* __restrict__ klass, const int* __restrict__ parent, int bw, int bh,
                                  const int* __restrict__ seam_lane, int n, int s0, int horizontal, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int step = s0 + i, lane = seam_lane[i];
    const int x = horizontal ? step : lane, y = horizontal ? lane : step;
    out[3 * i] = x;
    out[3 * i + 1] = y;
    out[3 * i + 2] = horizontal ? uls_value(klass, parent, bw, bh, x, y + 1) : uls_value(klass, parent, bw, bh, x + 1, y);
}

// interior pixels whose component is adjacent to comp2 become l2 ([SEAM]:1089-1092); adj_roots sorted
__global__ void k_uls_apply(const uint8_t* __restrict__ klass, const int* __restrict__ parent, int bw, int bh,
                            const int* __restrict__ adj_roots, int nadj, int* labels, Frame f, int bx, int by, int l2) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= bw || y >= bh) return;
    const size_t i = (size_t)y * bw + x;
    if (klass[i] != 1) return;
    const int r = parent[i];
    int lo = 0, hi = nadj - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int v = adj_roots[mid];
        if (v == r) { labels[lidx(f, bx + x, by + y)] = l2; return; }
        if (v < r) lo = mid + 1; else hi = mid - 1;
    }
}

__global__ void k_scatter_label(const int2* __restrict__ pts, int n, int* labels, Frame f, int l) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) labels[lidx(f, pts[i].x, pts[i].y)] = l;
}

// ---- final mask update [SEAM]:527-545 --------------------------------------------------------------------------
// A pixel of `dst` is cleared when its label's state has `bit` and the other image's mask is set there; both masks
// being set means the pixel lies in the intersection rectangle (ix, iy, iw, ih in frame coordinates), so only that is
// visited.  dst: mask placed at (dox, doy) in the frame.
__global__ void k_mask_update(uint8_t* dst, size_t dstep, int dox, int doy, MaskView other, const int* __restrict__ labels, Frame f,
                              const int* __restrict__ states, int bit, int ix, int iy, int iw, int ih) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= iw || y >= ih) return;
    const int ux = ix + x, uy = iy + y;
    const int l = lab(labels, f, ux, uy);
    if (l > 0 && (states[l - 1] & bit) && other.at(ux, uy)) dst[(size_t)(uy - doy) * dstep + (ux - dox)] = 0;
}

// ---- run-based component labelling ---------------------------------------------------------------------------------
// Warped masks are made of a few long horizontal runs per row.  Instead of labelling every pixel of the union frame
// (3 x int32 passes over tens of megapixels per pair) the rows are reduced to their class change points on the device
// (class = mask1 | mask2 << 1), the host unites the runs of adjacent rows (a few thousand runs), and labels are
// materialised only in the window the algorithm reads.  Component numbering = rank of the component's first run in
// raster order = floodFill's numbering ([SEAM]:222-256).
struct ChangePt { int x, cls; };

// One warp per frame row.  The change points of row y go to out[y * cap ...] in increasing x; counts[y] is the true
// number (rows with more than `cap` set *overflow and the caller falls back to the dense labelling).
constexpr int ROW_CAP = 32;

// One warp per mask row: the x positions where (mask != 0) toggles, in increasing x; the state left of x = 0 is "outside".
// 16 pixels per lane and trip (one 16-byte load when the row allows it), so a 6000-pixel row is 12 dependent trips.
// counts[y] is the true number of toggles; rows with more than `cap` set *overflow (the caller falls back to the dense
// labelling).  The union-frame class change points of a pair are the merge of its two masks' toggles (host, label_runs).
__global__ void k_mask_row_toggles(const uint8_t* __restrict__ mask, size_t step, int rows, int cols, int cap, int* __restrict__ counts,
                                   int* __restrict__ xs, int* __restrict__ overflow) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* row = mask + (size_t)y * step;
    const bool aligned = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
    int* out = xs + (size_t)y * cap;
    int n = 0;
    unsigned carry = 0;                                  // state of the pixel left of the current 512-pixel chunk
    for (int base = 0; base < cols; base += 512) {
        const int x0 = base + 16 * lane;
        unsigned bits = 0;                               // bit i: pixel x0 + i is inside the mask
        if (x0 + 16 <= cols && aligned) {
            const uint4 v = *reinterpret_cast<const uint4*>(row + x0);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned t = ((((w[k] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w[k]) & 0x80808080u) >> 7;   // 1 in every non-zero byte
                bits |= (((t * 0x01020408u) >> 24) & 0xfu) << (4 * k);
            }
        } else {
            for (int i = 0; i < 16; ++i)
                if (x0 + i < cols && row[x0 + i]) bits |= 1u << i;
        }
        unsigned left = __shfl_up_sync(0xffffffffu, bits >> 15, 1);
        if (lane == 0) left = carry;
        unsigned tg = (bits ^ ((bits << 1) | (left & 1u))) & 0xffffu;
        if (x0 + 16 > cols) tg &= x0 < cols ? (1u << (cols - x0)) - 1u : 0u;   // nothing is reported at or beyond the row end
        const int c = __popc(tg);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int k = n + incl - c;
        while (tg) {
            const int i = __ffs(tg) - 1;
            tg &= tg - 1;
            if (k < cap) out[k] = x0 + i;
            ++k;
        }
        n += __shfl_sync(0xffffffffu, incl, 31);
        carry = __shfl_sync(0xffffffffu, bits >> 15, 31);
    }
    if (lane == 0) {
        counts[y] = n;
        if (n > cap) *overflow = 1;
    }
}

// labels of the window from the change points: label of the last change point at or before x in row y
__global__ void k_label_window(Frame f, int cap, const int* __restrict__ counts, const ChangePt* __restrict__ cps, const int* __restrict__ cp_label,
                               int* __restrict__ labels) {
    const int x = f.wx + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = f.wy + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= f.wx + f.ww || y >= f.wy + f.wh) return;
    int lo = y * cap, hi = y * cap + counts[y] - 1, best = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (cps[mid].x <= x) { best = mid; lo = mid + 1; } else hi = mid - 1;
    }
    labels[lidx(f, x, y)] = best >= 0 ? cp_label[best] : 0;
}

// =====================================================================================================
// host logic
// =====================================================================================================

struct Pt { int x, y; };

// open-addressing (key -> int) map for the few thousand contour / seam pixels of updateLabelsUsingSeam
struct FlatMap {
    std::vector<long long> keys;
    std::vector<int> vals;
    size_t mask = 0;
    explicit FlatMap(size_t expected) {
        size_t cap = 64;
        while (cap < expected * 4) cap <<= 1;
        keys.assign(cap, -1);
        vals.assign(cap, 0);
        mask = cap - 1;
    }
    inline int& operator[](long long k) {
        size_t h = ((unsigned long long)k * 0x9E3779B97F4A7C15ull) >> 20 & mask;
        while (keys[h] != k) {
            if (keys[h] < 0) { keys[h] = k; vals[h] = 0; break; }
            h = (h + 1) & mask;
        }
        return vals[h];
    }
};

#include "seam_dp_launch.inl"

struct DbgTimer {   // IS_SEAM_DEBUG=1: host wall-clock per section (includes the device work it waits for)
    bool on; std::chrono::steady_clock::time_point t;
    DbgTimer() : on(getenv("IS_SEAM_DEBUG") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[seam] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};


struct TraceSink {
    int32_t* buf = nullptr; size_t cap = 0; size_t len = 0;
};

class PairSeam {
public:
    PairSeam(is_ctx* c, bool u8, TraceSink* tr, int cost = IS_COST_COLOR) : ctx(c), is_u8(u8), trace(tr), cost_fn(cost) {}
    // in1/in2: masks as the pair sees them; out1/out2: where the masks with this pair's clears go (may alias in1/in2).
    // structure_only: stop after the component / contour analysis (used to validate a speculative run).
    int process(const DevMat& image1, const DevMat& image2, Pt tl1, Pt tl2, const DevMat& in1, const DevMat& in2, const DevMat& out1,
                const DevMat& out2, int pi, int pj, bool structure_only = false);
    // Everything the pair does after labelling is a function of these two (and of the images): the raster-ordered
    // contour records of the INTERS components (pixels, labels, neighbour labels, contour-proximity flags) and the
    // states of the components they mention.
    std::vector<ContourRec> fp_records;
    std::vector<int> fp_states;
    bool same_structure(const PairSeam& o) const;

private:
    is_ctx* ctx;
    bool is_u8;
    TraceSink* trace;
    int cost_fn;
    DevBuf grad;                          // COLOR_GRAD: gx1, gy1, gx2, gy2 over the intersection window
    GradView gview{};
    int compute_gradients(const Pt& iTl, const Pt& iBr);
    int pair_i = 0, pair_j = 0;
    Pt unionTl{}, tl1_{}, tl2_{};
    int uw = 0, uh = 0;
    const DevMat* img1 = nullptr; const DevMat* img2 = nullptr;
    DevBuf cls, parent, labels;
    // host mirrors
    int ncomps = 0;
    std::vector<int> states;
    std::vector<Pt> tls, brs;
    std::vector<std::vector<ContourRec>> contours;
    std::set<std::pair<int, int>> edges;
    // scratch
    DevBuf counts, offsets, recs;

    Frame fr{};
    Frame frame() const { return fr; }
    int label_runs(const Pt& iTl, const Pt& iBr, bool* done);
    std::vector<char> stale;   // components whose bbox / contour recomputation is pending (resolve_conflicts)
    int label_dense();
    void release_device() { cls.release(); parent.release(); labels.release(); counts.release(); offsets.release(); recs.release(); grad.release(); }
    int ccl(const uint8_t* klass, int kmask, int w, int h, int* parent_out);
    int collect_roots(const int* parent_d, const uint8_t* klass_d, size_t n, std::vector<std::pair<int, int>>* roots);
    int extract_contours(int wx, int wy, int ww, int wh, int fa, int fb, std::vector<ContourRec>* out);
    void find_edges();
    bool has_only_one_neighbor(int comp);
    bool get_seam_tips(int c1, int c2, Pt* p1, Pt* p2);
    int estimate_and_update(int c1, int c2, Pt p1, Pt p2);
    int refresh_component(int c);
    int resolve_conflicts(const DevMat& in1, const DevMat& in2, const DevMat& out1, const DevMat& out2);
};

bool PairSeam::same_structure(const PairSeam& o) const {
    if (fp_records.size() != o.fp_records.size()) return false;
    if (!fp_records.empty() && std::memcmp(fp_records.data(), o.fp_records.data(), sizeof(ContourRec) * fp_records.size()) != 0) return false;
    auto state_of = [](const std::vector<int>& st, int label) { return (label >= 1 && label <= (int)st.size()) ? st[label - 1] : -1; };
    for (const ContourRec& r : fp_records) {
        if (state_of(fp_states, r.label) != state_of(o.fp_states, r.label)) return false;
        for (int k = 0; k < 4; ++k)
            if (r.nl[k] > 0 && state_of(fp_states, r.nl[k]) != state_of(o.fp_states, r.nl[k])) return false;
    }
    return true;
}

int PairSeam::ccl(const uint8_t* klass, int kmask, int w, int h, int* parent_out) {
    IS_LAUNCH(ctx, k_ccl_rows, h, 256, 0, klass, kmask, parent_out, w);
    if (h > 1) {
        dim3 block(64, 4), grid(div_up(w, 64), div_up(h - 1, 4));
        IS_LAUNCH(ctx, k_ccl_merge, grid, block, 0, klass, kmask, parent_out, w, h);
    }
    const size_t n = (size_t)w * h;
    IS_LAUNCH(ctx, k_ccl_flatten, (unsigned)((n + 255) / 256), 256, 0, parent_out, n);
    return IS_OK;
}

// roots sorted by index (raster order of the first pixel) with the klass byte at the root
int PairSeam::collect_roots(const int* parent_d, const uint8_t* klass_d, size_t n, std::vector<std::pair<int, int>>* roots) {
    int cap = 4096;
    while (true) {
        DevBuf out;
        IS_TRY(out.alloc(ctx, sizeof(int) * (2 * (size_t)cap + 1)));   // [2 cap] = number of roots found: one download for both
        int* cnt = out.as<int>() + 2 * (size_t)cap;
        IS_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int), ctx->stream));
        IS_LAUNCH(ctx, k_collect_roots, (unsigned)((n + 255) / 256), 256, 0, parent_d, klass_d, n, out.as<int>(), cnt, cap);
        std::vector<int> h(2 * (size_t)cap + 1);
        IS_TRY(download(ctx, h.data(), out.p, sizeof(int) * h.size()));
        const int count = h[2 * (size_t)cap];
        if (count > cap) { cap = count; continue; }
        roots->resize(count);
        for (int k = 0; k < count; ++k) (*roots)[k] = {h[2 * k], h[2 * k + 1]};
        std::sort(roots->begin(), roots->end());
        return IS_OK;
    }
}

int PairSeam::extract_contours(int wx, int wy, int ww, int wh, int fa, int fb, std::vector<ContourRec>* out) {
    out->clear();
    if (ww <= 0 || wh <= 0) return IS_OK;
    const size_t total = (size_t)ww * wh;
    const int nblocks = (int)((total + CT_CHUNK - 1) / CT_CHUNK);
    IS_TRY(counts.alloc(ctx, sizeof(int) * (size_t)nblocks));
    IS_TRY(offsets.alloc(ctx, sizeof(int) * ((size_t)nblocks + 1)));
    IS_LAUNCH(ctx, k_contour_count, nblocks, CT_THREADS, 0, labels.as<int>(), frame(), wx, wy, ww, wh, fa, fb, counts.as<int>());
    IS_LAUNCH(ctx, k_scan_counts, 1, 1024, 0, counts.as<int>(), nblocks, offsets.as<int>());
    int n = 0;
    IS_TRY(download(ctx, &n, offsets.as<int>() + nblocks, sizeof(int)));
    if (n == 0) return IS_OK;
    IS_TRY(recs.alloc(ctx, sizeof(ContourRec) * (size_t)n));
    IS_LAUNCH(ctx, k_contour_write, nblocks, CT_THREADS, 0, labels.as<int>(), frame(), wx, wy, ww, wh, fa, fb, offsets.as<int>(),
              recs.as<ContourRec>(), n);
    IS_LAUNCH(ctx, k_contour_flags, div_up(n, 8), 256, 0, recs.as<ContourRec>(), n, frame());
    out->resize(n);
    IS_TRY(download(ctx, out->data(), recs.p, sizeof(ContourRec) * (size_t)n));
    return IS_OK;
}

// [SEAM]:311-392.  An edge exists when at least one contour pixel touches the other component; consecutive contour
// pixels mostly repeat the same neighbour, so repeated inserts are skipped.
void PairSeam::find_edges() {
    edges.clear();
    for (int ci = 0; ci < ncomps; ++ci) {
        int last = -1;
        for (const ContourRec& r : contours[ci]) {
            const int l = ci + 1;
            for (int k = 0; k < 4; ++k) {
                const int nl = r.nl[k];
                if (nl > 0 && nl != l && nl != last) {
                    edges.insert({ci, nl - 1});
                    edges.insert({nl - 1, ci});
                    last = nl;
                }
            }
        }
    }
}

// [SEAM]:575-581
bool PairSeam::has_only_one_neighbor(int comp) {
    auto begin = edges.lower_bound({comp, INT_MIN});
    auto end = edges.upper_bound({comp, INT_MAX});
    return ++begin == end;
}

static inline double round_half_even(double v) { return std::nearbyint(v); }   // cvRound

// [SEAM]:607-706
bool PairSeam::get_seam_tips(int c1, int c2, Pt* p1, Pt* p2) {
    std::vector<Pt> special;
    const int l2 = c2 + 1;
    for (const ContourRec& r : contours[c1])
        if ((r.flags & 3) == 3 && (r.nl[0] == l2 || r.nl[1] == l2 || r.nl[2] == l2 || r.nl[3] == l2)) special.push_back({r.x, r.y});
    if (special.size() < 2) return false;
    // cv::partition with ClosePoints(10): connected components of "dist^2 < 100", classes numbered by first member
    const int n = (int)special.size();
    std::vector<int> uf(n);
    for (int i = 0; i < n; ++i) uf[i] = i;
    auto find = [&](int i) { while (uf[i] != i) { uf[i] = uf[uf[i]]; i = uf[i]; } return i; };
    // points are in raster order: only rows within 10 of each other can be close
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n && special[j].y - special[i].y < 10; ++j) {
            int dx = special[i].x - special[j].x, dy = special[i].y - special[j].y;
            if (dx * dx + dy * dy < 100) { int a = find(i), b = find(j); if (a != b) uf[b] = a; }
        }
    std::vector<int> cls_of_root(n, -1), lab(n);
    int nlabels = 0;
    for (int i = 0; i < n; ++i) { int r = find(i); if (cls_of_root[r] < 0) cls_of_root[r] = nlabels++; lab[i] = cls_of_root[r]; }
    if (nlabels < 2) return false;
    std::vector<long long> sumx(nlabels, 0), sumy(nlabels, 0);
    std::vector<std::vector<Pt>> points(nlabels);
    for (int i = 0; i < n; ++i) { sumx[lab[i]] += special[i].x; sumy[lab[i]] += special[i].y; points[lab[i]].push_back(special[i]); }
    int idx[2] = {-1, -1};
    double maxDist = -std::numeric_limits<double>::max();
    for (int i = 0; i < nlabels - 1; ++i)
        for (int j = i + 1; j < nlabels; ++j) {
            double s1 = (double)points[i].size(), s2 = (double)points[j].size();
            double cx1 = round_half_even((int)sumx[i] / s1), cy1 = round_half_even((int)sumy[i] / s1);
            double cx2 = round_half_even((int)sumx[j] / s2), cy2 = round_half_even((int)sumy[j] / s2);
            double dist = (cx1 - cx2) * (cx1 - cx2) + (cy1 - cy2) * (cy1 - cy2);
            if (dist > maxDist) { maxDist = dist; idx[0] = i; idx[1] = j; }
        }
    Pt p[2];
    for (int i = 0; i < 2; ++i) {
        const std::vector<Pt>& pts = points[idx[i]];
        double size = (double)pts.size();
        double cx = round_half_even((int)sumx[idx[i]] / size), cy = round_half_even((int)sumy[idx[i]] / size);
        size_t closest = 0;
        double minDist = std::numeric_limits<double>::max();
        for (size_t j = 0; j < pts.size(); ++j) {
            double dist = (pts[j].x - cx) * (pts[j].x - cx) + (pts[j].y - cy) * (pts[j].y - cy);
            if (dist < minDist) { minDist = dist; closest = j; }
        }
        p[i] = pts[closest];
    }
    *p1 = p[0];
    *p2 = p[1];
    return true;
}

// bbox + raster-ordered contour of component c, rescanning its previous bbox ([SEAM]:483-513)
int PairSeam::refresh_component(int c) {
    const int x0 = tls[c].x, y0 = tls[c].y, x1 = brs[c].x, y1 = brs[c].y;
    tls[c] = {INT_MAX, INT_MAX};
    brs[c] = {INT_MIN, INT_MIN};
    contours[c].clear();
    if (x1 <= x0 || y1 <= y0) return IS_OK;
    IS_TRY(extract_contours(x0, y0, x1 - x0, y1 - y0, c + 1, c + 1, &contours[c]));
    for (const ContourRec& r : contours[c]) {   // the extreme pixels of a component are contour pixels
        tls[c].x = std::min(tls[c].x, r.x); tls[c].y = std::min(tls[c].y, r.y);
        brs[c].x = std::max(brs[c].x, r.x + 1); brs[c].y = std::max(brs[c].y, r.y + 1);
    }
    return IS_OK;
}

template <typename T>
static ImgView<T> make_view(const DevMat& m, int dx, int dy) {
    return ImgView<T>{m.ptr<T>(), m.step, m.rows, m.cols, dx, dy, m.channels};
}

// computeGradients [SEAM]:549-572, restricted to the intersection rectangle (the only place costs are evaluated)
int PairSeam::compute_gradients(const Pt& iTl, const Pt& iBr) {
    const int ww = iBr.x - iTl.x, wh = iBr.y - iTl.y;
    const int pitch = (ww + 31) & ~31;
    const size_t plane = (size_t)pitch * wh;
    IS_TRY(grad.alloc(ctx, sizeof(float) * plane * 4));
    float* g = grad.as<float>();
    gview = GradView{g, g + plane, g + 2 * plane, g + 3 * plane, pitch, iTl.x - unionTl.x, iTl.y - unionTl.y};
    const int dx1 = unionTl.x - tl1_.x, dy1 = unionTl.y - tl1_.y, dx2 = unionTl.x - tl2_.x, dy2 = unionTl.y - tl2_.y;
    dim3 block(64, 4), grid(div_up(ww, 64), div_up(wh, 4));
    ctx->next_bytes = (double)ww * wh * ((is_u8 ? 3 : 12) + 8);
    if (is_u8) {
        IS_LAUNCH(ctx, k_sobel_window<uint8_t>, grid, block, 0, make_view<uint8_t>(*img1, dx1, dy1), gview.ox, gview.oy, ww, wh, g, g + plane, pitch);
        ctx->next_bytes = (double)ww * wh * (3 + 8);
        IS_LAUNCH(ctx, k_sobel_window<uint8_t>, grid, block, 0, make_view<uint8_t>(*img2, dx2, dy2), gview.ox, gview.oy, ww, wh, g + 2 * plane, g + 3 * plane, pitch);
    } else {
        IS_LAUNCH(ctx, k_sobel_window<float>, grid, block, 0, make_view<float>(*img1, dx1, dy1), gview.ox, gview.oy, ww, wh, g, g + plane, pitch);
        ctx->next_bytes = (double)ww * wh * (12 + 8);
        IS_LAUNCH(ctx, k_sobel_window<float>, grid, block, 0, make_view<float>(*img2, dx2, dy2), gview.ox, gview.oy, ww, wh, g + 2 * plane, g + 3 * plane, pitch);
    }
    return IS_OK;
}

// estimateSeam [SEAM]:806-957 followed by updateLabelsUsingSeam [SEAM]:960-1093
int PairSeam::estimate_and_update(int c1, int c2, Pt p1, Pt p2) {
    const int l1 = c1 + 1, l2 = c2 + 1;
    const int rx = tls[c1].x, ry = tls[c1].y, rw = brs[c1].x - rx, rh = brs[c1].y - ry;
    Pt src{p1.x - rx, p1.y - ry}, dst{p2.x - rx, p2.y - ry};
    bool swapped = false;
    const bool horizontal = std::abs(dst.x - src.x) > std::abs(dst.y - src.y);
    if (horizontal) { if (src.x > dst.x) { std::swap(src, dst); swapped = true; } }
    else if (src.y > dst.y) { std::swap(src, dst); swapped = true; }
    const int lanes = horizontal ? rh : rw, steps = horizontal ? rw : rh;

    DevBuf P, Q, control, seam_lane;
    int lpt = lanes <= 4096 ? 4 : (lanes <= 8192 ? 8 : 16);
    if (const char* e = getenv("IS_DP_LPT")) {             // tuning knob (4, 8 or 16 lanes per thread)
        const int v = atoi(e);
        if ((v == 4 || v == 8 || v == 16) && lanes <= 1024 * v) lpt = v;
    }
    const int nt = std::min(1024, div_up(div_up(lanes, lpt), 32) * 32);
    const int pitch = nt * lpt;                            // padded row length (multiple of 128 lanes)
    IS_REQUIRE(ctx, pitch >= lanes && pitch <= 12 * 1024, IS_ERR_UNSUPPORTED, "seam component wider than 12288 lanes");
    IS_TRY(P.alloc(ctx, sizeof(float) * (size_t)pitch * (steps + DP_ROW_PAD) + 64));     // spare rows: k_seam_fwd fetches past the last step
    IS_TRY(Q.alloc(ctx, sizeof(float) * (size_t)pitch * (steps + DP_ROW_PAD) + 64));
    IS_TRY(control.alloc(ctx, (size_t)pitch * steps + 64));
    IS_TRY(seam_lane.alloc(ctx, sizeof(int) * ((size_t)steps + 1)));   // [steps] = "destination reached" flag: one download for both
    const int dx1 = unionTl.x - tl1_.x, dy1 = unionTl.y - tl1_.y, dx2 = unionTl.x - tl2_.x, dy2 = unionTl.y - tl2_.y;
    {
        dim3 block(64, 4), grid(div_up(pitch, 64), div_up(steps, 4));
        ctx->next_bytes = (double)lanes * steps * (is_u8 ? 6 : 24) + (double)lanes * steps * 12;   // two image overlaps + labels read, P and Q written
        const auto v1u = make_view<uint8_t>(*img1, dx1, dy1), v2u = make_view<uint8_t>(*img2, dx2, dy2);
        const auto v1f = make_view<float>(*img1, dx1, dy1), v2f = make_view<float>(*img2, dx2, dy2);
        const int hz = horizontal ? 1 : 0;
        if (cost_fn == IS_COST_COLOR_GRAD) {
            ctx->next_bytes += (double)lanes * steps * 16;                                   // four gradient planes
            if (is_u8)
                IS_LAUNCH(ctx, (k_cost_pq<uint8_t, true>), grid, block, 0, v1u, v2u, labels.as<int>(), frame(), l1, rx, ry, rw, rh, hz, P.as<float>(),
                          Q.as<float>(), pitch, gview);
            else
                IS_LAUNCH(ctx, (k_cost_pq<float, true>), grid, block, 0, v1f, v2f, labels.as<int>(), frame(), l1, rx, ry, rw, rh, hz, P.as<float>(),
                          Q.as<float>(), pitch, gview);
        } else if (is_u8)
            IS_LAUNCH(ctx, (k_cost_pq<uint8_t, false>), grid, block, 0, v1u, v2u, labels.as<int>(), frame(), l1, rx, ry, rw, rh, hz, P.as<float>(),
                      Q.as<float>(), pitch, gview);
        else
            IS_LAUNCH(ctx, (k_cost_pq<float, false>), grid, block, 0, v1f, v2f, labels.as<int>(), frame(), l1, rx, ry, rw, rh, hz, P.as<float>(),
                      Q.as<float>(), pitch, gview);
    }
    DpArgs A;
    A.P = P.as<float>(); A.Q = Q.as<float>(); A.control = control.as<uint8_t>();
    A.lanes = lanes; A.pitch = pitch; A.steps = steps;
    A.s0 = horizontal ? src.x : src.y; A.lane0 = horizontal ? src.y : src.x;
    A.s1 = horizontal ? dst.x : dst.y; A.lane1 = horizontal ? dst.y : dst.x;
    A.seam_lane = seam_lane.as<int>(); A.reached = seam_lane.as<int>() + steps;
    // the forward pass of the batched path (halo windows, register ring, clusters) + parallel back-track, one seam per launch;
    // IS_PAIR_DP_V0=1 keeps round 1's kernel (TMA ring, block barrier per step, back-track inside)
    DpShape shape;
    dp_choose_shape(lanes, steps, A.s0, A.s1, dp_variant_default(), &shape);
    DevBuf bt_map, dp_tab;
    if (shape.v1 && shape.pitch == pitch && !getenv("IS_PAIR_DP_V0")) {
        const int nchunks = div_up(A.s1 - A.s0, BT_CHUNK);
        IS_TRY(bt_map.alloc(ctx, sizeof(short) * (size_t)std::max(nchunks, 1) * pitch + 64));
        A.G = shape.G; A.D = shape.D;
        struct alignas(16) Tables { DpArgs a; alignas(16) BtArgs b; } tabs;
        std::memset(&tabs, 0, sizeof(tabs));
        tabs.a = A;
        tabs.b.A = A; tabs.b.map = bt_map.as<short>(); tabs.b.nchunks = nchunks;
        IS_TRY(dp_tab.alloc(ctx, sizeof(Tables)));
        IS_TRY(upload(ctx, dp_tab.p, &tabs, sizeof(Tables)));
        IS_CUDA(ctx, cudaMemsetAsync(seam_lane.p, 0, sizeof(int) * ((size_t)steps + 1), ctx->stream));
        const Tables* td = dp_tab.as<Tables>();
        IS_TRY(launch_dp_all(ctx, std::vector<DpShape>(1, shape), &td->a, &td->b));
    } else {
        const size_t row_pair = 2 * sizeof(float) * (size_t)pitch;            // one row of P + one of Q
        const size_t budget = 192 * 1024;
        // Two stages of as many rows as fit: the per-stage work (mbarrier wait, refill issued by thread 0 while the other warps
        // wait at the next barrier) costs about as much as a DP step, the depth of the ring does not matter (measured on B200:
        // 1.17 ms with 2 x 7 rows, 1.24 ms with 3 x 5, 1.65 ms with 7 x 2 for a 4000 x 1500 overlap).
        A.D = 2;
        A.G = (int)std::min<size_t>(16, budget / (A.D * row_pair));
        if (A.G < 1) { A.D = 2; A.G = 1; }
        if (const char* e = getenv("IS_DP_G")) A.G = std::max(1, atoi(e));     // tuning knobs: rows per ring stage, stages
        if (const char* e = getenv("IS_DP_D")) A.D = std::max(2, atoi(e));
        const size_t smem = std::max<size_t>((size_t)A.D * A.G * row_pair + 8 * (size_t)A.D + 16, (size_t)32 * 65 + 16);
        IS_REQUIRE(ctx, smem <= (size_t)DP_SMEM_MAX, IS_ERR_INTERNAL, "DP shared-memory budget");
        ctx->next_bytes = (double)(A.s1 - A.s0) * lanes * 9;                    // P, Q read once, control written once
        switch (lpt) {
            case 4:
                IS_CUDA(ctx, cudaFuncSetAttribute(k_seam_dp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX));
                IS_LAUNCH(ctx, k_seam_dp<4>, 1, nt, smem, A);
                break;
            case 8:
                IS_CUDA(ctx, cudaFuncSetAttribute(k_seam_dp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX));
                IS_LAUNCH(ctx, k_seam_dp<8>, 1, nt, smem, A);
                break;
            default:
                IS_CUDA(ctx, cudaFuncSetAttribute(k_seam_dp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX));
                IS_LAUNCH(ctx, k_seam_dp<16>, 1, nt, smem, A);
                break;
        }
    }
    // ---- updateLabelsUsingSeam, device part (launched before knowing `reached`; harmless if unreachable
    //      because nothing is written to labels until the host has decided)
    const int nseam = A.s1 - A.s0 + 1;
    DevBuf klass, sub_parent;
    IS_TRY(klass.alloc(ctx, (size_t)rw * rh));
    IS_TRY(sub_parent.alloc(ctx, sizeof(int) * (size_t)rw * rh));
    DbgTimer dbg;
    std::vector<int> lane_all((size_t)steps + 1);
    IS_TRY(download(ctx, lane_all.data(), seam_lane.p, sizeof(int) * lane_all.size()));
    const int ok = lane_all[steps];
    dbg.lap("  cost+dp (sync)");
    if (!ok) return IS_OK;                                             // [SEAM]:918-919: estimateSeam returned false
    const int* lane_h = lane_all.data();
    std::vector<Pt> seam(nseam);   // union-frame coordinates, ordered p1 -> p2 ([SEAM]:949-954)
    for (int i = 0; i < nseam; ++i) {
        int step = A.s0 + i, lane = lane_h[i];
        Pt q = horizontal ? Pt{step + rx, lane + ry} : Pt{lane + rx, step + ry};
        seam[swapped ? nseam - 1 - i : i] = q;
    }
    if (!(seam.front().x == p1.x && seam.front().y == p1.y && seam.back().x == p2.x && seam.back().y == p2.y))
        return fail(ctx, IS_ERR_ASSERT, "seam end points differ from the seam tips ([SEAM]:953-954)");
    if (trace) {
        size_t need = 5 + 2 * (size_t)nseam;
        if (trace->buf && trace->len + need <= trace->cap) {
            int32_t* t = trace->buf + trace->len;
            t[0] = pair_i; t[1] = pair_j; t[2] = c1; t[3] = horizontal ? 1 : 0; t[4] = nseam;
            for (int i = 0; i < nseam; ++i) { t[5 + 2 * i] = seam[i].x + unionTl.x; t[6 + 2 * i] = seam[i].y + unionTl.y; }
        }
        trace->len += need;
    }
    {
        dim3 block(64, 4), grid(div_up(rw, 64), div_up(rh, 4));
        IS_LAUNCH(ctx, k_uls_class, grid, block, 0, labels.as<int>(), frame(), l1, rx, ry, rw, rh, klass.as<uint8_t>());
        IS_LAUNCH(ctx, k_uls_paint_seam, div_up(nseam, 256), 256, 0, seam_lane.as<int>(), nseam, A.s0, horizontal ? 1 : 0, rw, klass.as<uint8_t>());
    }
    // flood fill of the interior ([SEAM]:978-981): klass == 1 regions; klass 2 pixels form components of
    // their own that are ignored below
    IS_TRY(ccl(klass.as<uint8_t>(), 3, rw, rh, sub_parent.as<int>()));
    std::vector<std::pair<int, int>> roots_all;
    IS_TRY(collect_roots(sub_parent.as<int>(), klass.as<uint8_t>(), (size_t)rw * rh, &roots_all));
    std::vector<int> roots;   // raster order of the first pixel -> component id = rank + 1
    for (auto& r : roots_all) if (r.second == 1) roots.push_back(r.first);
    const int nsub = (int)roots.size();
    auto id_of = [&](int v) -> int {   // gathered value -> reference mask value
        if (v <= 0) return v;          // 0, -1 (outside), -255 (painted)
        auto it = std::lower_bound(roots.begin(), roots.end(), v - 1);
        return (int)(it - roots.begin()) + 1;
    };
    // gather the neighbourhoods the sequential part needs
    const std::vector<ContourRec>& cont = contours[c1];
    const int nc = (int)cont.size();
    std::vector<int2> cpts(nc);
    for (int i = 0; i < nc; ++i) cpts[i] = make_int2(cont[i].x - rx, cont[i].y - ry);
    DevBuf cpts_d, gath_d;                                   // gathered neighbourhoods: 8 ints per contour pixel, then 3 per seam pixel
    IS_TRY(cpts_d.alloc(ctx, sizeof(int2) * (size_t)std::max(nc, 1)));
    std::vector<int> gath(8 * (size_t)nc + 3 * (size_t)nseam);
    IS_TRY(gath_d.alloc(ctx, sizeof(int) * gath.size()));
    int* g8_dp = gath_d.as<int>();
    int* gs_dp = gath_d.as<int>() + 8 * (size_t)nc;
    if (nc) {
        IS_TRY(upload(ctx, cpts_d.p, cpts.data(), sizeof(int2) * (size_t)nc));
        IS_LAUNCH(ctx, k_uls_gather8, div_up(nc, 128), 128, 0, klass.as<uint8_t>(), sub_parent.as<int>(), rw, rh, cpts_d.as<int2>(), nc, g8_dp);
    }
    IS_LAUNCH(ctx, k_uls_gather_seam, div_up(nseam, 128), 128, 0, klass.as<uint8_t>(), sub_parent.as<int>(), rw, rh, seam_lane.as<int>(), nseam,
              A.s0, horizontal ? 1 : 0, gs_dp);
    IS_TRY(download(ctx, gath.data(), gath_d.p, sizeof(int) * gath.size()));   // one round trip for both
    const int* g8 = gath.data();
    const int* gs = gath.data() + 8 * (size_t)nc;
    dbg.lap("  uls device part (sync)");

    // ---- sequential part on the host ([SEAM]:983-1034).  `painted` holds the current mask value of every
    //      contour / seam pixel (255 until assigned); all other pixels are constant (gathered above).
    FlatMap painted((size_t)(nc + nseam));
    auto key = [&](int x, int y) { return (long long)y * rw + x; };
    for (int i = 0; i < nc; ++i) painted[key(cpts[i].x, cpts[i].y)] = 255;
    for (int i = 0; i < nseam; ++i) painted[key(gs[3 * i], gs[3 * i + 1])] = 255;
    static const int ddx[8] = {-1, +1, 0, 0, -1, +1, -1, +1};
    static const int ddy[8] = {0, 0, -1, +1, -1, -1, +1, +1};
    auto value_at = [&](int gathered, int x, int y) -> int {   // current reference mask value at (x, y)
        if (gathered == -1) return -1;                          // outside the mask
        if (gathered == -255) return painted[key(x, y)];
        return id_of(gathered);
    };
    for (int i = 0; i < nc; ++i) {
        const int x = cpts[i].x, y = cpts[i].y;
        bool okc = false;
        int val = 255;
        for (int j = 0; j < 8; ++j) {
            int v = value_at(g8[(size_t)i * 8 + j], x + ddx[j], y + ddy[j]);
            if (v > 0 && v != 255) { okc = true; val = v; }
        }
        painted[key(x, y)] = okc ? val : 0;
    }
    for (int i = 0; i < nseam; ++i) {
        // order-independent: a seam has one pixel per step, the inspected neighbour lies in the same step
        const int k = swapped ? nseam - 1 - i : i;
        const int x = gs[3 * k], y = gs[3 * k + 1];
        int v = horizontal ? value_at(gs[3 * k + 2], x, y + 1) : value_at(gs[3 * k + 2], x + 1, y);
        painted[key(x, y)] = (v > 0 && v != 255) ? v : 0;
    }
    // adjacency vote ([SEAM]:1039-1085).  The reference's std::map<int,int> counters hold the keys 1..ncomps plus key 0
    // when a contour pixel ended up unassigned; plain arrays with a "key 0 present" flag are equivalent.
    std::vector<int> connect2(nsub + 1, 0), connectOther(nsub + 1, 0);
    bool c2_has0 = false, co_has0 = false;
    for (int i = 0; i < nc; ++i) {
        const ContourRec& r = cont[i];
        int mv = painted[key(cpts[i].x, cpts[i].y)];
        if (mv < 0 || mv > nsub) mv = 0;
        if (r.nl[0] == l2 || r.nl[1] == l2 || r.nl[2] == l2 || r.nl[3] == l2) { connect2[mv]++; if (mv == 0) c2_has0 = true; }
        bool other = false;
        for (int k = 0; k < 4; ++k) if (r.nl[k] >= 0 && r.nl[k] != l1 && r.nl[k] != l2) other = true;
        if (other) { connectOther[mv]++; if (mv == 0) co_has0 = true; }
    }
    const int maxKey = nsub;
    std::vector<int> isAdj(maxKey + 1, 0);
    const double len = (double)nc;
    for (int k = c2_has0 ? 0 : 1; k <= nsub; ++k) {
        int res = 0;
        if (connect2[k] / len > 0.05) {
            const bool sub_exists = k >= 1 || co_has0;
            if (sub_exists && (connectOther[k] / len < 0.1)) res = 1;
        }
        isAdj[k] = res;
    }
    dbg.lap("  uls host walk");
    // ---- relabel ([SEAM]:1089-1092)
    std::vector<int> adj_roots;
    for (int i = 1; i <= nsub; ++i) if (isAdj[i]) adj_roots.push_back(roots[i - 1]);
    std::vector<int2> flips;
    for (size_t h = 0; h < painted.keys.size(); ++h) {
        const long long k = painted.keys[h];
        if (k < 0) continue;
        const int v = painted.vals[h];
        if (v > 0 && v <= maxKey && isAdj[v]) flips.push_back(make_int2((int)(k % rw) + rx, (int)(k / rw) + ry));
    }
    if (!adj_roots.empty()) {
        DevBuf ar;
        IS_TRY(ar.alloc(ctx, sizeof(int) * adj_roots.size()));
        IS_TRY(upload(ctx, ar.p, adj_roots.data(), sizeof(int) * adj_roots.size()));
        dim3 block(64, 4), grid(div_up(rw, 64), div_up(rh, 4));
        IS_LAUNCH(ctx, k_uls_apply, grid, block, 0, klass.as<uint8_t>(), sub_parent.as<int>(), rw, rh, ar.as<int>(), (int)adj_roots.size(),
                  labels.as<int>(), frame(), rx, ry, l2);
    }
    if (!flips.empty()) {
        DevBuf fl;
        IS_TRY(fl.alloc(ctx, sizeof(int2) * flips.size()));
        IS_TRY(upload(ctx, fl.p, flips.data(), sizeof(int2) * flips.size()));
        IS_LAUNCH(ctx, k_scatter_label, div_up((int)flips.size(), 256), 256, 0, fl.as<int2>(), (int)flips.size(), labels.as<int>(), frame(), l2);
    }
    return IS_OK;
}

// [SEAM]:395-546
int PairSeam::resolve_conflicts(const DevMat& in1, const DevMat& in2, const DevMat& mask1, const DevMat& mask2) {
    bool hasConflict = true;
    while (hasConflict) {
        int c1 = 0, c2 = 0;
        hasConflict = false;
        for (auto itr = edges.begin(); itr != edges.end(); ++itr) {
            c1 = itr->first;
            c2 = itr->second;
            if ((states[c1] & ST_INTERS) && (states[c1] & (~ST_INTERS)) != states[c2]) { hasConflict = true; break; }
        }
        if (!hasConflict) break;
        const int l1 = c1 + 1, l2 = c2 + 1;
        // The reference recomputes the bbox and the contour of c1 (and of an INTERS c2) after every conflict ([SEAM]:483-513).
        // c1 only ever loses pixels, so its recomputation is deferred until something reads it: a stale bbox is a superset
        // (good enough to relabel the whole component) and scanning the superset later finds the same pixels in the same
        // raster order.  A component that is about to gain pixels is brought up to date first, because the reference scans
        // the bbox it had at that moment.
        if (stale.size() < (size_t)ncomps) stale.resize((size_t)ncomps, 0);
        if ((states[c2] & ST_INTERS) && stale[c2]) { IS_TRY(refresh_component(c2)); stale[c2] = 0; }
        if (has_only_one_neighbor(c1)) {
            const int w = brs[c1].x - tls[c1].x, h = brs[c1].y - tls[c1].y;
            if (w > 0 && h > 0) {
                dim3 block(64, 4), grid(div_up(w, 64), div_up(h, 4));
                IS_LAUNCH(ctx, k_relabel_rect, grid, block, 0, labels.as<int>(), frame(), tls[c1].x, tls[c1].y, w, h, l1, l2);
            }
            states[c1] = states[c2] == ST_FIRST ? ST_SECOND : ST_FIRST;
        } else {
            if (stale[c1]) { IS_TRY(refresh_component(c1)); stale[c1] = 0; }
            Pt p1, p2;
            if (get_seam_tips(c1, c2, &p1, &p2)) IS_TRY(estimate_and_update(c1, c2, p1, p2));
            states[c1] = states[c2] == ST_FIRST ? (ST_INTERS | ST_SECOND) : (ST_INTERS | ST_FIRST);
        }
        stale[c1] = 1;
        if (states[c2] & ST_INTERS) IS_TRY(refresh_component(c2));   // geometry of a non-INTERS component is never read again
        edges.erase({c1, c2});
        edges.erase({c2, c1});
    }
    // update masks ([SEAM]:524-545): mask2 first (reads the original mask1), then mask1 (reads the updated mask2)
    if (mask1.data != in1.data)
        IS_CUDA(ctx, cudaMemcpy2DAsync(mask1.data, mask1.step, in1.data, in1.step, (size_t)in1.cols, in1.rows, cudaMemcpyDeviceToDevice, ctx->stream));
    if (mask2.data != in2.data)
        IS_CUDA(ctx, cudaMemcpy2DAsync(mask2.data, mask2.step, in2.data, in2.step, (size_t)in2.cols, in2.rows, cudaMemcpyDeviceToDevice, ctx->stream));
    DevBuf st;
    IS_TRY(st.alloc(ctx, sizeof(int) * (size_t)std::max(ncomps, 1)));
    if (ncomps) IS_TRY(upload(ctx, st.p, states.data(), sizeof(int) * (size_t)ncomps));
    const int o1x = tl1_.x - unionTl.x, o1y = tl1_.y - unionTl.y, o2x = tl2_.x - unionTl.x, o2y = tl2_.y - unionTl.y;
    MaskView v1{mask1.ptr<uint8_t>(), mask1.step, mask1.rows, mask1.cols, o1x, o1y};
    MaskView v2{mask2.ptr<uint8_t>(), mask2.step, mask2.rows, mask2.cols, o2x, o2y};
    const int ix = std::max(o1x, o2x), iy = std::max(o1y, o2y);
    const int iw = std::min(o1x + mask1.cols, o2x + mask2.cols) - ix, ih = std::min(o1y + mask1.rows, o2y + mask2.rows) - iy;
    dim3 mblock(64, 4), mgrid(div_up(iw, 64), div_up(ih, 4));
    IS_LAUNCH(ctx, k_mask_update, mgrid, mblock, 0, mask2.ptr<uint8_t>(), mask2.step, o2x, o2y, v1, labels.as<int>(), frame(), st.as<int>(),
              (int)ST_FIRST, ix, iy, iw, ih);
    IS_LAUNCH(ctx, k_mask_update, mgrid, mblock, 0, mask1.ptr<uint8_t>(), mask1.step, o1x, o1y, v2, labels.as<int>(), frame(), st.as<int>(),
              (int)ST_SECOND, ix, iy, iw, ih);
    return IS_OK;
}

// Dense fallback (masks with very many runs): per-pixel union-find over the whole frame.
int PairSeam::label_dense() {
    const size_t n = (size_t)uw * uh;
    fr.wx = 0; fr.wy = 0; fr.ww = uw; fr.wh = uh;
    IS_TRY(cls.alloc(ctx, n));
    IS_TRY(parent.alloc(ctx, sizeof(int) * n));
    IS_TRY(labels.alloc(ctx, sizeof(int) * n));
    {
        dim3 block(64, 4), grid(div_up(uw, 64), div_up(uh, 4));
        IS_LAUNCH(ctx, k_classify, grid, block, 0, fr.m1, fr.m2, cls.as<uint8_t>(), uw, uh);
    }
    IS_TRY(ccl(cls.as<uint8_t>(), 3, uw, uh, parent.as<int>()));
    std::vector<std::pair<int, int>> roots;
    IS_TRY(collect_roots(parent.as<int>(), cls.as<uint8_t>(), n, &roots));
    ncomps = (int)roots.size();
    states.assign(ncomps, 0);
    if (ncomps) {
        std::vector<int> root_idx(ncomps);
        for (int k = 0; k < ncomps; ++k) {
            root_idx[k] = roots[k].first;
            const int c = roots[k].second & 3;
            states[k] = c == 3 ? ST_INTERS : (c == 1 ? ST_FIRST : ST_SECOND);
        }
        DevBuf rd;
        IS_TRY(rd.alloc(ctx, sizeof(int) * (size_t)ncomps));
        IS_TRY(upload(ctx, rd.p, root_idx.data(), sizeof(int) * (size_t)ncomps));
        IS_LAUNCH(ctx, k_scatter_ids, div_up(ncomps, 256), 256, 0, rd.as<int>(), ncomps, labels.as<int>());
        IS_LAUNCH(ctx, k_labels_from_roots, (unsigned)((n + 255) / 256), 256, 0, parent.as<int>(), labels.as<int>(), n);
    } else {
        IS_CUDA(ctx, cudaMemsetAsync(labels.p, 0, sizeof(int) * n, ctx->stream));
    }
    parent.release();
    cls.release();
    return IS_OK;
}

// Run-based labelling (see k_row_changes).  *done stays false when the masks have too many runs for it to pay off.
int PairSeam::label_runs(const Pt& iTl, const Pt& iBr, bool* done) {
    *done = false;
    // toggles of both masks (own coordinates), one launch each, one download for all of it
    const MaskView mv[2] = {fr.m1, fr.m2};
    const size_t nrow[2] = {(size_t)mv[0].rows, (size_t)mv[1].rows};
    const size_t ints = 1 + nrow[0] + nrow[1] + (nrow[0] + nrow[1]) * ROW_CAP;   // overflow flag | counts1 | counts2 | xs1 | xs2
    DevBuf tg;
    IS_TRY(tg.alloc(ctx, sizeof(int) * ints));
    int* tg_d = tg.as<int>();
    IS_CUDA(ctx, cudaMemsetAsync(tg_d, 0, sizeof(int), ctx->stream));
    int* cnt_d[2] = {tg_d + 1, tg_d + 1 + nrow[0]};
    int* xs_d[2] = {tg_d + 1 + nrow[0] + nrow[1], tg_d + 1 + nrow[0] + nrow[1] + nrow[0] * ROW_CAP};
    for (int k = 0; k < 2; ++k) {
        ctx->next_bytes = (double)mv[k].rows * mv[k].cols;
        IS_LAUNCH(ctx, k_mask_row_toggles, div_up(mv[k].rows, 8), 256, 0, mv[k].p, mv[k].step, mv[k].rows, mv[k].cols, ROW_CAP, cnt_d[k], xs_d[k], tg_d);
    }
    // counts first (tiny), then only as many slots per row as the fullest row uses
    const size_t head = 1 + nrow[0] + nrow[1];
    std::vector<int> tg_h(head);
    IS_TRY(download(ctx, tg_h.data(), tg_d, sizeof(int) * head));
    if (tg_h[0]) return IS_OK;                                              // noisy masks: dense path
    const int* cnt_h[2] = {tg_h.data() + 1, tg_h.data() + 1 + nrow[0]};
    int slots = 1;
    for (size_t r = 0; r < nrow[0] + nrow[1]; ++r) slots = std::max(slots, tg_h[1 + r]);
    std::vector<int> xs_all((nrow[0] + nrow[1]) * (size_t)slots);
    IS_TRY(download2d(ctx, xs_all.data(), sizeof(int) * (size_t)slots, xs_d[0], sizeof(int) * ROW_CAP, sizeof(int) * (size_t)slots, nrow[0] + nrow[1]));
    const int* xs_h[2] = {xs_all.data(), xs_all.data() + nrow[0] * (size_t)slots};
    // merge per union-frame row: class = (inside mask1) | (inside mask2) << 1, a change point wherever it differs from the pixel
    // on its left (class 0 left of x = 0) -- the run-length form of [SEAM]:205-218
    std::vector<int> row_off((size_t)uh + 1, 0);
    std::vector<ChangePt> cps;
    cps.reserve((size_t)uh * 6);
    for (int y = 0; y < uh; ++y) {
        int na[2] = {0, 0};
        const int* xa[2] = {nullptr, nullptr};
        for (int k = 0; k < 2; ++k) {
            const int my = y - mv[k].oy;
            if (my >= 0 && my < mv[k].rows) { na[k] = cnt_h[k][my]; xa[k] = xs_h[k] + (size_t)my * slots; }
        }
        // position i of mask k: xa[k][i] + ox for i < na[k], then (when the row ends inside the mask) the closing toggle at ox + cols
        auto pos = [&](int k, int i) { return i < na[k] ? xa[k][i] + mv[k].ox : ((na[k] & 1) && i == na[k] ? mv[k].ox + mv[k].cols : INT_MAX); };
        int i0 = 0, i1 = 0, st = 0, prev_cls = 0, emitted = 0;
        for (;;) {
            const int p0 = pos(0, i0), p1 = pos(1, i1);
            const int x = std::min(p0, p1);
            if (x >= uw) break;                     // INT_MAX (both exhausted) or a closing toggle on the frame's right edge
            if (p0 == x) { st ^= 1; ++i0; }
            if (p1 == x) { st ^= 2; ++i1; }
            if (st != prev_cls) { cps.push_back(ChangePt{x, st}); prev_cls = st; ++emitted; }
        }
        if (emitted > ROW_CAP) return IS_OK;        // more runs than the window kernel's per-row table holds: dense path
        row_off[y + 1] = row_off[y] + emitted;
    }
    const int R = row_off[uh];
    if (cps.empty()) cps.push_back(ChangePt{0, 0});
    // union-find over the runs (run k = change point k with cls != 0, spanning [x, next change point or uw))
    std::vector<int> uf((size_t)R);
    for (int k = 0; k < R; ++k) uf[k] = k;
    auto find = [&](int k) { while (uf[k] != k) { uf[k] = uf[uf[k]]; k = uf[k]; } return k; };
    auto run_end = [&](int k, int y) { return k + 1 < row_off[y + 1] ? cps[k + 1].x : uw; };
    for (int y = 1; y < uh; ++y) {
        int a = row_off[y - 1], ae = row_off[y], b = row_off[y], be = row_off[y + 1];
        while (a < ae && b < be) {
            const int ax1 = run_end(a, y - 1), bx1 = run_end(b, y);
            if (cps[a].cls && cps[a].cls == cps[b].cls && cps[a].x < bx1 && cps[b].x < ax1) {
                int ra = find(a), rb = find(b);
                if (ra != rb) { if (ra < rb) uf[rb] = ra; else uf[ra] = rb; }   // the root is the raster-first run
            }
            if (ax1 <= bx1) ++a; else ++b;
        }
    }
    std::vector<int> cp_label((size_t)std::max(R, 1), 0), id_of_root((size_t)std::max(R, 1), 0);
    ncomps = 0;
    states.clear();
    for (int k = 0; k < R; ++k) {               // roots in increasing index = raster order of the first pixel
        if (!cps[k].cls || find(k) != k) continue;
        id_of_root[k] = ++ncomps;
        states.push_back(cps[k].cls == 3 ? ST_INTERS : (cps[k].cls == 1 ? ST_FIRST : ST_SECOND));
    }
    for (int k = 0; k < R; ++k) cp_label[k] = cps[k].cls ? id_of_root[find(k)] : 0;
    // labels only where they are read: the intersection rectangle grown by one pixel
    fr.wx = std::max(0, iTl.x - unionTl.x - 1);
    fr.wy = std::max(0, iTl.y - unionTl.y - 1);
    fr.ww = std::min(uw, iBr.x - unionTl.x + 1) - fr.wx;
    fr.wh = std::min(uh, iBr.y - unionTl.y + 1) - fr.wy;
    IS_TRY(labels.alloc(ctx, sizeof(int) * (size_t)fr.ww * fr.wh));
    // change points + their labels of the window's rows in the window kernel's layout: per row a count, then ROW_CAP slots
    const size_t wrows = (size_t)fr.wh;
    size_t wcap = 1;                                           // slots per row: the fullest window row
    for (int y = fr.wy; y < fr.wy + fr.wh; ++y) wcap = std::max(wcap, (size_t)(row_off[y + 1] - row_off[y]));
    std::vector<int> tab(wrows + wrows * wcap * 3, 0);         // counts | ChangePt (x, cls) | labels
    int* t_cnt = tab.data();
    ChangePt* t_cps = reinterpret_cast<ChangePt*>(tab.data() + wrows);
    int* t_lab = tab.data() + wrows + wrows * wcap * 2;
    for (int y = fr.wy; y < fr.wy + fr.wh; ++y) {
        const size_t r = (size_t)(y - fr.wy);
        t_cnt[r] = row_off[y + 1] - row_off[y];
        std::copy(cps.begin() + row_off[y], cps.begin() + row_off[y + 1], t_cps + r * wcap);
        std::copy(cp_label.begin() + row_off[y], cp_label.begin() + row_off[y + 1], t_lab + r * wcap);
    }
    DevBuf tab_d;
    IS_TRY(tab_d.alloc(ctx, sizeof(int) * tab.size()));
    IS_TRY(upload(ctx, tab_d.p, tab.data(), sizeof(int) * tab.size()));
    {
        dim3 block(64, 4), grid(div_up(fr.ww, 64), div_up(fr.wh, 4));
        const int* d = tab_d.as<int>();
        // the kernel indexes its tables with the frame row: bias the pointers by the window's first row
        IS_LAUNCH(ctx, k_label_window, grid, block, 0, fr, (int)wcap, d - fr.wy, reinterpret_cast<const ChangePt*>(d + wrows) - (size_t)fr.wy * wcap,
                  d + wrows + wrows * wcap * 2 - (size_t)fr.wy * wcap, labels.as<int>());
    }
    *done = true;
    return IS_OK;
}

// [SEAM]:127-193
int PairSeam::process(const DevMat& image1, const DevMat& image2, Pt tl1, Pt tl2, const DevMat& mask1, const DevMat& mask2, const DevMat& out1,
                      const DevMat& out2, int pi, int pj, bool structure_only) {
    pair_i = pi; pair_j = pj;
    DbgTimer dbg;
    IS_REQUIRE(ctx, image1.rows == mask1.rows && image1.cols == mask1.cols, IS_ERR_ASSERT, "image1.size() == mask1.size()");
    IS_REQUIRE(ctx, image2.rows == mask2.rows && image2.cols == mask2.cols, IS_ERR_ASSERT, "image2.size() == mask2.size()");
    Pt iTl{std::max(tl1.x, tl2.x), std::max(tl1.y, tl2.y)};
    Pt iBr{std::min(tl1.x + image1.cols, tl2.x + image2.cols), std::min(tl1.y + image1.rows, tl2.y + image2.rows)};
    if (iTl.x >= iBr.x || iTl.y >= iBr.y) return IS_OK;   // no conflicts
    img1 = &image1; img2 = &image2; tl1_ = tl1; tl2_ = tl2;
    unionTl = {std::min(tl1.x, tl2.x), std::min(tl1.y, tl2.y)};
    Pt unionBr{std::max(tl1.x + image1.cols, tl2.x + image2.cols), std::max(tl1.y + image1.rows, tl2.y + image2.rows)};
    uw = unionBr.x - unionTl.x;
    uh = unionBr.y - unionTl.y;
    const size_t n = (size_t)uw * uh;
    IS_REQUIRE(ctx, n < (size_t)INT_MAX, IS_ERR_UNSUPPORTED, "union frame of an image pair exceeds 2^31 pixels");
    MaskView v1{mask1.ptr<uint8_t>(), mask1.step, mask1.rows, mask1.cols, tl1.x - unionTl.x, tl1.y - unionTl.y};
    MaskView v2{mask2.ptr<uint8_t>(), mask2.step, mask2.rows, mask2.cols, tl2.x - unionTl.x, tl2.y - unionTl.y};
    fr = Frame{uw, uh, 0, 0, uw, uh, v1, v2};
    // findComponents [SEAM]:196-308
    bool done = false;
    const char* dense = getenv("IS_SEAM_DENSE");
    if (!(dense && dense[0] == '1')) IS_TRY(label_runs(iTl, iBr, &done));
    if (!done) IS_TRY(label_dense());
    dbg.lap("component labelling");
    tls.assign(ncomps, Pt{INT_MAX, INT_MAX});
    brs.assign(ncomps, Pt{INT_MIN, INT_MIN});
    contours.assign(ncomps, std::vector<ContourRec>());
    // Contour lists, bounding boxes and edges are only ever consulted for INTERS components (the conflict loop
    // picks edges whose first component is INTERS, [SEAM]:427; getSeamTips / updateLabelsUsingSeam walk
    // contours_[comp1] with comp1 INTERS), and every adjacency of an INTERS component is witnessed by one of
    // its own contour pixels.  INTERS components lie inside the intersection rectangle, so only that window is
    // scanned -- the raster order of the extracted lists is the same as in a full-frame scan.
    std::vector<ContourRec> all;
    IS_TRY(extract_contours(iTl.x - unionTl.x, iTl.y - unionTl.y, iBr.x - iTl.x, iBr.y - iTl.y, CT_FILTER_INTERS, 0, &all));
    for (const ContourRec& r : all) {
        const int c = r.label - 1;
        contours[c].push_back(r);
        tls[c].x = std::min(tls[c].x, r.x); tls[c].y = std::min(tls[c].y, r.y);
        brs[c].x = std::max(brs[c].x, r.x + 1); brs[c].y = std::max(brs[c].y, r.y + 1);
    }
    dbg.lap("labels+contours");
    fp_records = all;
    fp_states = states;
    if (structure_only) { release_device(); return IS_OK; }
    find_edges();
    dbg.lap("find_edges");
    if (cost_fn == IS_COST_COLOR_GRAD) IS_TRY(compute_gradients(iTl, iBr));                 // [SEAM]:398-399
    int rc = resolve_conflicts(mask1, mask2, out1, out2);
    if (dbg.on) cudaStreamSynchronize(ctx->stream);
    dbg.lap("resolve_conflicts+masks");
    release_device();
    return rc;
}

// dst keeps its value where src is non-zero and becomes 0 where src is 0 (intersection of two clear sets)
__global__ void k_mask_and(uint8_t* dst, size_t dstep, const uint8_t* __restrict__ src, size_t sstep, int rows, int cols) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    if (src[(size_t)y * sstep + x] == 0) dst[(size_t)y * dstep + x] = 0;
}

static int mask_and(is_ctx* ctx, const DevMat& dst, const DevMat& src) {
    dim3 block(64, 4), grid(div_up(dst.cols, 64), div_up(dst.rows, 4));
    IS_LAUNCH(ctx, k_mask_and, grid, block, 0, dst.ptr<uint8_t>(), dst.step, src.ptr<uint8_t>(), src.step, dst.rows, dst.cols);
    return IS_OK;
}

// same, restricted to the rectangle (x0, y0, w, h) of the masks: a pair only clears pixels inside its intersection
static int mask_and_rect(is_ctx* ctx, const DevMat& dst, const DevMat& src, int x0, int y0, int w, int h) {
    if (w <= 0 || h <= 0) return IS_OK;
    dim3 block(64, 4), grid(div_up(w, 64), div_up(h, 4));
    IS_LAUNCH(ctx, k_mask_and, grid, block, 0, dst.ptr<uint8_t>() + (size_t)y0 * dst.step + x0, dst.step,
              src.ptr<uint8_t>() + (size_t)y0 * src.step + x0, src.step, h, w);
    return IS_OK;
}

static int mask_copy(is_ctx* ctx, const DevMat& dst, const DevMat& src) {
    IS_CUDA(ctx, cudaMemcpy2DAsync(dst.data, dst.step, src.data, src.step, (size_t)src.cols, src.rows, cudaMemcpyDeviceToDevice, ctx->stream));
    return IS_OK;
}

// The reference's order ([SEAM]:97-121): pairs (i, j), i < j, lexicographic, reversed; masks updated in place.
static int seam_find_sequential(is_ctx* ctx, const std::vector<std::pair<int, int>>& pairs, const DevMat* images, const is_point* corners,
                                const DevMat* masks, TraceSink* trace, int cost_fn = IS_COST_COLOR) {
    for (auto& pr : pairs) {
        PairSeam ps(ctx, images[pr.first].depth == IS_8U, trace, cost_fn);
        IS_TRY(ps.process(images[pr.first], images[pr.second], Pt{corners[pr.first].x, corners[pr.first].y},
                          Pt{corners[pr.second].x, corners[pr.second].y}, masks[pr.first], masks[pr.second], masks[pr.first], masks[pr.second],
                          pr.first, pr.second));
    }
    return IS_OK;
}

// ---- speculative concurrent execution ------------------------------------------------------------------------------
// Every pair modifies its two masks in place, so the reference's loop is a dependency chain: pair p must see the
// clears of every earlier pair that shares an image with it.  In a panorama those clears almost never reach the
// part of the mask pair p looks at, but that is a property of the data, not of the algorithm.  So:
//   1. all overlapping pairs run CONCURRENTLY (one host thread + CUDA stream each) on the masks as they were at
//      entry, each writing its clears into private copies, and each keeping its structural fingerprint
//      (PairSeam::fp_records / fp_states: everything the pair computes after labelling is a function of it);
//   2. every pair with an earlier neighbour recomputes ONLY that fingerprint on the masks it would really have
//      seen (entry masks minus the clears of the earlier pairs) -- also concurrently;
//   3. if all fingerprints agree, the speculative clears are exactly what the sequential loop produces and the
//      final masks are their intersection; otherwise everything is redone sequentially.
// The DP passes (one CTA each, latency bound) and the host/device round trips of different pairs overlap.
struct PairJob {
    int i, j;
    DevMat out_i, out_j;          // private result masks
    PairSeam* spec = nullptr;
    TraceSink trace;
    std::vector<int32_t> trace_buf;
    int status = IS_OK;
    bool needs_check = false, valid = true;
};

template <typename F>
static void run_on_workers(is_ctx* parent, std::vector<is_ctx*>& workers, size_t njobs, F&& fn) {
    std::vector<std::thread> th;
    const size_t T = std::min(workers.size(), njobs);
    for (size_t t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            cudaSetDevice(parent->device);
            for (size_t k = t; k < njobs; k += T) fn(workers[t], k);
        });
    for (auto& x : th) x.join();
}

static int seam_find_concurrent(is_ctx* ctx, const std::vector<std::pair<int, int>>& active, int n, const DevMat* images, const is_point* corners,
                                const DevMat* masks, TraceSink* trace, bool* accepted) {
    *accepted = false;
    const size_t np = active.size();
    std::vector<PairJob> jobs(np);
    size_t T = std::min<size_t>(np, 8);
    if (const char* e = getenv("IS_SEAM_WORKERS")) T = std::max<size_t>(1, std::min<size_t>(np, (size_t)atoi(e)));
    std::vector<is_ctx*> workers(T);
    for (size_t t = 0; t < T; ++t) {
        IS_TRY(child_ctx(ctx, t, &workers[t]));
        workers[t]->ktiming = ctx->ktiming;
        workers[t]->launches = 0;
    }
    for (size_t k = 0; k < np; ++k) {
        jobs[k].i = active[k].first; jobs[k].j = active[k].second;
        IS_TRY(alloc_mat(ctx, masks[jobs[k].i].rows, masks[jobs[k].i].cols, 1, IS_8U, &jobs[k].out_i));
        IS_TRY(alloc_mat(ctx, masks[jobs[k].j].rows, masks[jobs[k].j].cols, 1, IS_8U, &jobs[k].out_j));
        if (trace) { jobs[k].trace_buf.resize(5 + 2 * (size_t)(images[jobs[k].i].rows + images[jobs[k].i].cols + images[jobs[k].j].rows + images[jobs[k].j].cols) * 4);
                     jobs[k].trace.buf = jobs[k].trace_buf.data(); jobs[k].trace.cap = jobs[k].trace_buf.size(); }
    }
    // fork: the workers' streams start after everything queued on the caller's stream
    IS_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    for (size_t t = 0; t < T; ++t) IS_CUDA(ctx, cudaStreamWaitEvent(workers[t]->stream, ctx->ev[4], 0));
    DbgTimer dbg;
    // 1. speculative runs on the entry masks
    run_on_workers(ctx, workers, np, [&](is_ctx* w, size_t k) {
        PairJob& J = jobs[k];
        J.spec = new PairSeam(w, images[J.i].depth == IS_8U, trace ? &J.trace : nullptr);
        J.status = J.spec->process(images[J.i], images[J.j], Pt{corners[J.i].x, corners[J.i].y}, Pt{corners[J.j].x, corners[J.j].y}, masks[J.i],
                                   masks[J.j], J.out_i, J.out_j, J.i, J.j);
        if (J.status == IS_OK && cudaStreamSynchronize(w->stream) != cudaSuccess) J.status = IS_ERR_CUDA;
    });
    dbg.lap("concurrent: speculative runs");
    // 2. validate the pairs that had an earlier neighbour
    auto earlier = [&](size_t k, int img) {   // results of earlier pairs for image img
        std::vector<const DevMat*> v;
        for (size_t q = 0; q < k; ++q) {
            if (jobs[q].i == img) v.push_back(&jobs[q].out_i);
            if (jobs[q].j == img) v.push_back(&jobs[q].out_j);
        }
        return v;
    };
    bool failed = false;
    for (size_t k = 0; k < np; ++k) {
        if (jobs[k].status != IS_OK) failed = true;
        jobs[k].needs_check = !earlier(k, jobs[k].i).empty() || !earlier(k, jobs[k].j).empty();
    }
    if (!failed)
        run_on_workers(ctx, workers, np, [&](is_ctx* w, size_t k) {
            PairJob& J = jobs[k];
            if (!J.needs_check) return;
            DevMat true_i, true_j;
            const DevMat* in[2] = {&masks[J.i], &masks[J.j]};
            DevMat* tmp[2] = {&true_i, &true_j};
            const int img[2] = {J.i, J.j};
            for (int s2 = 0; s2 < 2 && J.status == IS_OK; ++s2) {
                auto ev = earlier(k, img[s2]);
                if (ev.empty()) continue;
                J.status = alloc_mat(w, in[s2]->rows, in[s2]->cols, 1, IS_8U, tmp[s2]);
                if (J.status == IS_OK) J.status = mask_copy(w, *tmp[s2], *ev[0]);
                for (size_t q = 1; q < ev.size() && J.status == IS_OK; ++q) J.status = mask_and(w, *tmp[s2], *ev[q]);
                in[s2] = tmp[s2];
            }
            if (J.status != IS_OK) return;
            PairSeam chk(w, images[J.i].depth == IS_8U, nullptr);
            J.status = chk.process(images[J.i], images[J.j], Pt{corners[J.i].x, corners[J.i].y}, Pt{corners[J.j].x, corners[J.j].y}, *in[0], *in[1],
                                   *in[0], *in[1], J.i, J.j, /*structure_only=*/true);
            if (J.status == IS_OK) J.valid = chk.same_structure(*J.spec);
            cudaStreamSynchronize(w->stream);
        });
    dbg.lap("concurrent: validation");
    int rc = IS_OK;
    bool all_valid = !failed;
    for (size_t k = 0; k < np; ++k) {
        if (jobs[k].status != IS_OK && rc == IS_OK) { rc = jobs[k].status; }
        if (!jobs[k].valid) all_valid = false;
        delete jobs[k].spec;
        jobs[k].spec = nullptr;
    }
    // join: account the workers' launches / timing records to the caller's context
    for (size_t t = 0; t < T; ++t) {
        if (rc != IS_OK && !workers[t]->last_error.empty() && ctx->last_error.empty()) ctx->last_error = workers[t]->last_error;
        ctx->launches += workers[t]->launches;
        for (auto& r : workers[t]->krecs) ctx->krecs.push_back(r);
        workers[t]->krecs.clear();
        IS_CUDA(ctx, cudaEventRecord(workers[t]->ev[0], workers[t]->stream));
        IS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, workers[t]->ev[0], 0));
    }
    if (rc != IS_OK) {
        for (size_t t = 0; t < T; ++t) if (!workers[t]->last_error.empty()) { ctx->last_error = workers[t]->last_error; break; }
        return rc;
    }
    if (!all_valid) return IS_OK;                        // caller falls back to the sequential loop
    // 3. final masks = entry masks minus every pair's clears
    for (size_t k = 0; k < np; ++k) {
        const int i = jobs[k].i, j = jobs[k].j;
        const int x0 = std::max(corners[i].x, corners[j].x), y0 = std::max(corners[i].y, corners[j].y);
        const int x1 = std::min(corners[i].x + images[i].cols, corners[j].x + images[j].cols);
        const int y1 = std::min(corners[i].y + images[i].rows, corners[j].y + images[j].rows);
        IS_TRY(mask_and_rect(ctx, masks[i], jobs[k].out_i, x0 - corners[i].x, y0 - corners[i].y, x1 - x0, y1 - y0));
        IS_TRY(mask_and_rect(ctx, masks[j], jobs[k].out_j, x0 - corners[j].x, y0 - corners[j].y, x1 - x0, y1 - y0));
        if (trace) {
            const size_t len = jobs[k].trace.len;
            if (trace->buf && trace->len + len <= trace->cap && len <= jobs[k].trace.cap) std::memcpy(trace->buf + trace->len, jobs[k].trace_buf.data(), len * sizeof(int32_t));
            trace->len += len;
        }
    }
    (void)n;
    *accepted = true;
    return IS_OK;
}

#include "seam_runs.inl"
#include "seam_batch.inl"

// synthetic cost rows for is_debug_dp_bench: multiples of 0.5 like real colour costs, a curved band of cells outside the
// component (+inf) on either side so that the reachability logic is exercised
__global__ void k_dp_bench_fill(float* P, float* Q, int lanes, int pitch, int steps, unsigned seed) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    const int step = blockIdx.y;
    if (lane >= pitch || step >= steps) return;
    unsigned h = (unsigned)lane * 2654435761u ^ ((unsigned)step * 40503u + seed) * 2246822519u;
    h ^= h >> 15; h *= 2654435761u; h ^= h >> 13;
    const int margin = 3 + (int)(12.f * (0.5f + 0.5f * __sinf(step * 0.01f)));
    const bool inside = lane >= margin && lane < lanes - margin;
    float p = 0.5f * (float)(h % 1500u), q = 0.5f * (float)((h >> 11) % 1500u);
    if ((h >> 28) == 0) { p = 0.f; q = 0.f; }                       // ties
    P[(size_t)step * pitch + lane] = (lane < lanes && inside) ? p : __int_as_float(0x7f800000);
    Q[(size_t)step * pitch + lane] = lane < lanes ? q : 0.f;
}

// One wave of the batched seam path: which of the remaining pairs (reference order) start together.  rects[k] = intersection
// rectangle (x0, y0, x1, y1) of active[k] in panorama coordinates, the only place its clears fall.
// A wave takes the remaining pairs in the reference's order and leaves out (a) a pair that shares an image with an earlier pair of
// the wave whose intersection rectangle comes within a few pixels of its own -- the clears of such a neighbour would very likely
// change what the pair sees, and everything computed for it would be thrown away by the validation -- and (b) every later pair
// that shares an image with a pair left out: two pairs without a common image touch different masks and commute, so a pair may
// overtake the pairs it has nothing in common with, but never one it could depend on or that could depend on it.  In a strip
// every wave is the whole pair list; in a mosaic the waves follow the real conflicts.  The validation still decides what is
// accepted: this only chooses what is worth speculating on.  The first remaining pair is always in the wave.
static void choose_wave(const std::vector<std::pair<int, int>>& active, const std::vector<int4>& rects, const std::vector<char>& finished, int n_images,
                        bool whole, bool in_order, std::vector<size_t>* wave) {
    wave->clear();
    std::vector<char> blocked_img((size_t)n_images, 0);
    for (size_t k = 0; k < active.size(); ++k) {
        if (finished[k]) continue;
        const int i = active[k].first, j = active[k].second;
        bool skip = blocked_img[(size_t)i] || blocked_img[(size_t)j];
        if (!skip && !whole) {
            const int4 a = rects[k];
            for (size_t q : *wave) {
                const std::pair<int, int>& w = active[q];
                if (w.first != i && w.first != j && w.second != i && w.second != j) continue;
                const int4 b = rects[q];
                if (b.x < a.z + 4 && a.x - 4 < b.z && b.y < a.w + 4 && a.y - 4 < b.w) { skip = true; break; }
            }
        }
        if (skip) {
            if (in_order) break;
            blocked_img[(size_t)i] = 1; blocked_img[(size_t)j] = 1;
            continue;
        }
        wave->push_back(k);
    }
}

// device-resident images / masks (masks in-out); used by is_seam_dp_find* and by the pipeline
int seam_find_core(is_ctx* ctx, int n, const DevMat* images, const is_point* corners, const DevMat* masks, TraceSink* trace,
                   int cost_fn = IS_COST_COLOR) {
    std::vector<std::pair<int, int>> pairs;                                  // [SEAM]:97-111 (no sort, reversed)
    for (int i = 0; i + 1 < n; ++i)
        for (int j = i + 1; j < n; ++j) pairs.push_back({i, j});
    std::reverse(pairs.begin(), pairs.end());
    std::vector<std::pair<int, int>> active;                                 // pairs that do not return at [SEAM]:142-143
    for (auto& pr : pairs) {
        const int i = pr.first, j = pr.second;
        const int x0 = std::max(corners[i].x, corners[j].x), y0 = std::max(corners[i].y, corners[j].y);
        const int x1 = std::min(corners[i].x + images[i].cols, corners[j].x + images[j].cols);
        const int y1 = std::min(corners[i].y + images[i].rows, corners[j].y + images[j].rows);
        if (x0 < x1 && y0 < y1) active.push_back(pr);
    }
    const char* seq = getenv("IS_SEAM_SEQUENTIAL");
    const char* mode = getenv("IS_SEAM_PATH");            // tuning / test knob: "pairs" = per-pair concurrent path, "seq" = sequential loop
    ctx->seam_speculation_accepted = -1;
    ctx->seam_path = 0;
    const bool want_seq = (seq && seq[0] == '1') || (mode && !strcmp(mode, "seq"));
    if (!active.empty() && !want_seq && !(mode && !strcmp(mode, "pairs"))) {
        // all pairs through every kernel at once, three host consultations per wave (seam_batch.inl)
        bool all_batched = true;
        int waves = 0;
        std::vector<int4> rects(active.size());
        for (size_t k = 0; k < active.size(); ++k) {
            const int i = active[k].first, j = active[k].second;
            rects[k] = make_int4(std::max(corners[i].x, corners[j].x), std::max(corners[i].y, corners[j].y),
                                 std::min(corners[i].x + images[i].cols, corners[j].x + images[j].cols), std::min(corners[i].y + images[i].rows, corners[j].y + images[j].rows));
        }
        const bool whole = getenv("IS_SEAM_WAVE_ALL") != nullptr;           // tuning knob: every wave takes all remaining pairs, in order
        const bool in_order = getenv("IS_SEAM_WAVE_PREFIX") != nullptr;     // tuning knob: a wave ends at the first conflict (no pair is taken out of order)
        // The seams are reported in the reference's pair order whatever order the waves take the pairs in: with a trace buffer the
        // records are staged and sorted back at the end.
        std::vector<int32_t> staged;
        TraceSink stage_sink;
        TraceSink* tw = trace;
        if (trace && trace->buf) { staged.resize(trace->cap); stage_sink.buf = staged.data(); stage_sink.cap = staged.size(); tw = &stage_sink; }
        std::vector<char> finished(active.size(), 0);
        size_t left = active.size();
        while (left > 0) {
            std::vector<size_t> rest_idx;
            choose_wave(active, rects, finished, n, whole, in_order, &rest_idx);
            std::vector<std::pair<int, int>> rest;
            for (size_t k : rest_idx) rest.push_back(active[k]);
            size_t accepted = 0;
            bool first_unsupported = false;
            IS_TRY(seam_batch_wave(ctx, rest, n, images, corners, masks, tw, cost_fn, &accepted, &first_unsupported));
            if (first_unsupported) {                       // this pair alone through the general path, then on with the waves
                std::vector<std::pair<int, int>> one(1, rest[0]);          // the first pair of a wave is the first remaining pair of the reference's order
                IS_TRY(seam_find_sequential(ctx, one, images, corners, masks, tw, cost_fn));
                all_batched = false;
                finished[rest_idx[0]] = 1;
                --left;
                continue;
            }
            ++waves;
            // A wave member left behind (validation) keeps its place: it comes before every member after it in the reference's order
            // and shares no image with the pairs that were skipped in front of it, so taking it up again later is the same loop.
            // But the members AFTER it must not have been applied either: seam_batch_wave applies a prefix of the wave only.
            for (size_t q = 0; q < accepted; ++q) { finished[rest_idx[q]] = 1; --left; }
        }
        if (tw == &stage_sink) {
            if (stage_sink.len > stage_sink.cap) {
                trace->len += stage_sink.len;                              // too small either way: the caller sees the length it needs
            } else {
                std::map<std::pair<int, int>, std::vector<std::pair<size_t, size_t>>> by_pair;    // (i, j) -> records (offset, length) in the order they were estimated
                for (size_t pos = 0; pos + 5 <= stage_sink.len;) {
                    const size_t len = 5 + 2 * (size_t)staged[pos + 4];
                    by_pair[{staged[pos], staged[pos + 1]}].push_back({pos, len});
                    pos += len;
                }
                for (const auto& pr : active) {
                    auto it = by_pair.find(pr);
                    if (it == by_pair.end()) continue;
                    for (const auto& rec : it->second) {
                        if (trace->len + rec.second <= trace->cap) std::memcpy(trace->buf + trace->len, staged.data() + rec.first, rec.second * sizeof(int32_t));
                        trace->len += rec.second;
                    }
                }
            }
        }
        ctx->seam_speculation_accepted = waves <= 1 ? 1 : 0;
        ctx->seam_path = all_batched ? 2 : 0;
        ctx->seam_waves = waves;
        return IS_OK;
    }
    if (cost_fn != IS_COST_COLOR) return seam_find_sequential(ctx, active, images, corners, masks, trace, cost_fn);   // not on the per-pair concurrent path
    if (active.size() >= 2 && !want_seq) {
        bool accepted = false;
        IS_TRY(seam_find_concurrent(ctx, active, n, images, corners, masks, trace, &accepted));
        ctx->seam_speculation_accepted = accepted ? 1 : 0;
        if (accepted) { ctx->seam_path = 1; return IS_OK; }
    }
    return seam_find_sequential(ctx, active, images, corners, masks, trace);
}

static int seam_find_impl(is_ctx* ctx, int n, const is_mat* images, const is_point* corners, is_mat* masks, int cost_fn, TraceSink* trace) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return IS_OK;                                                // [SEAM]:94-95
    IS_REQUIRE(ctx, images && corners && masks, IS_ERR_BAD_ARG, "null argument");
    IS_REQUIRE(ctx, cost_fn == IS_COST_COLOR || cost_fn == IS_COST_COLOR_GRAD, IS_ERR_BAD_ARG, "unknown cost function");
    const int depth = images[0].depth;
    for (int i = 0; i < n; ++i) {
        IS_TRY(check_mat(ctx, &images[i], "image"));
        IS_TRY(check_mat(ctx, &masks[i], "mask"));
        IS_REQUIRE(ctx, (images[i].channels == 3 || images[i].channels == 4) && images[i].channels == images[0].channels &&
                            (images[i].depth == IS_8U || images[i].depth == IS_32F) && images[i].depth == depth,
                   IS_ERR_BAD_ARG, "both images must have CV_32FC3(4) or CV_8UC3(4) type");     // [SEAM]:741-750
        IS_REQUIRE(ctx, masks[i].depth == IS_8U && masks[i].channels == 1, IS_ERR_BAD_ARG, "masks must be CV_8U");
        IS_REQUIRE(ctx, images[i].rows == masks[i].rows && images[i].cols == masks[i].cols, IS_ERR_ASSERT, "image.size() == mask.size()");
    }
    std::vector<DevMat> dimg(n), dmask(n);
    for (int i = 0; i < n; ++i) {
        IS_TRY(stage_in(ctx, &images[i], &dimg[i]));
        IS_TRY(stage_out(ctx, &masks[i], &dmask[i], true));
    }
    IS_TRY(seam_find_core(ctx, n, dimg.data(), corners, dmask.data(), trace, cost_fn));
    for (int i = 0; i < n; ++i) IS_TRY(commit(ctx, &dmask[i]));
    return IS_OK;
}

// used by the pipeline with device-resident mats
int seam_find_device(is_ctx* ctx, int n, const DevMat* images, const is_point* corners, const DevMat* masks, int cost_fn) {
    return seam_find_core(ctx, n, images, corners, masks, nullptr, cost_fn);
}

}  // namespace is

using namespace is;

extern "C" {

int is_seam_dp_find(is_ctx* ctx, int n, const is_mat* images, const is_point* corners, is_mat* masks, int cost_fn) {
    return seam_find_impl(ctx, n, images, corners, masks, cost_fn, nullptr);
}

int is_seam_dp_find_trace(is_ctx* ctx, int n, const is_mat* images, const is_point* corners, is_mat* masks, int cost_fn,
                          int32_t* trace, size_t trace_cap, size_t* trace_len) {
    TraceSink sink;
    sink.buf = trace;
    sink.cap = trace ? trace_cap : 0;
    int rc = seam_find_impl(ctx, n, images, corners, masks, cost_fn, &sink);
    if (trace_len) *trace_len = sink.len;
    return rc;
}

// ---- single-pair primitives for sharded execution (one strip of the panorama per GPU) ---------------------------
struct is_seam_pair_impl { std::vector<ContourRec> records; std::vector<int> states; };

static int pair_common(is_ctx* ctx, const is_mat* image_i, const is_mat* image_j, const is_mat* mask_i, const is_mat* mask_j) {
    IS_TRY(check_mat(ctx, image_i, "image_i"));
    IS_TRY(check_mat(ctx, image_j, "image_j"));
    IS_TRY(check_mat(ctx, mask_i, "mask_i"));
    IS_TRY(check_mat(ctx, mask_j, "mask_j"));
    IS_REQUIRE(ctx, (image_i->channels == 3 || image_i->channels == 4) && image_j->channels == image_i->channels && image_i->depth == image_j->depth &&
                        (image_i->depth == IS_8U || image_i->depth == IS_32F), IS_ERR_BAD_ARG, "both images must have CV_32FC3(4) or CV_8UC3(4) type");
    IS_REQUIRE(ctx, mask_i->depth == IS_8U && mask_i->channels == 1 && mask_j->depth == IS_8U && mask_j->channels == 1, IS_ERR_BAD_ARG, "masks must be CV_8U");
    return IS_OK;
}

int is_seam_pair_run(is_ctx* ctx, const is_mat* image_i, const is_mat* image_j, is_point tl_i, is_point tl_j, const is_mat* mask_i,
                     const is_mat* mask_j, is_mat* out_i, is_mat* out_j, is_seam_pair** result) {
    if (!ctx || !result) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(pair_common(ctx, image_i, image_j, mask_i, mask_j));
    IS_TRY(check_mat(ctx, out_i, "out_i"));
    IS_TRY(check_mat(ctx, out_j, "out_j"));
    IS_REQUIRE(ctx, out_i->rows == mask_i->rows && out_i->cols == mask_i->cols && out_j->rows == mask_j->rows && out_j->cols == mask_j->cols &&
                        out_i->depth == IS_8U && out_j->depth == IS_8U, IS_ERR_BAD_ARG, "out masks must match the input masks");
    DevMat di, dj, mi, mj, oi, oj;
    IS_TRY(stage_in(ctx, image_i, &di));
    IS_TRY(stage_in(ctx, image_j, &dj));
    IS_TRY(stage_in(ctx, mask_i, &mi));
    IS_TRY(stage_in(ctx, mask_j, &mj));
    IS_TRY(stage_out(ctx, out_i, &oi, false));
    IS_TRY(stage_out(ctx, out_j, &oj, false));
    PairSeam ps(ctx, image_i->depth == IS_8U, nullptr);
    const bool overlap = std::max(tl_i.x, tl_j.x) < std::min(tl_i.x + image_i->cols, tl_j.x + image_j->cols) &&
                         std::max(tl_i.y, tl_j.y) < std::min(tl_i.y + image_i->rows, tl_j.y + image_j->rows);
    if (overlap) IS_TRY(ps.process(di, dj, Pt{tl_i.x, tl_i.y}, Pt{tl_j.x, tl_j.y}, mi, mj, oi, oj, 0, 1));
    else { IS_TRY(mask_copy(ctx, oi, mi)); IS_TRY(mask_copy(ctx, oj, mj)); }
    IS_TRY(commit(ctx, &oi));
    IS_TRY(commit(ctx, &oj));
    is_seam_pair_impl* r = new is_seam_pair_impl();
    r->records = ps.fp_records;
    r->states = ps.fp_states;
    *result = reinterpret_cast<is_seam_pair*>(r);
    return IS_OK;
}

int is_seam_pair_check(is_ctx* ctx, const is_mat* image_i, const is_mat* image_j, is_point tl_i, is_point tl_j, const is_mat* mask_i,
                       const is_mat* mask_j, const is_seam_pair* spec, int* same) {
    if (!ctx || !spec || !same) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(pair_common(ctx, image_i, image_j, mask_i, mask_j));
    DevMat di, dj, mi, mj;
    IS_TRY(stage_in(ctx, image_i, &di));
    IS_TRY(stage_in(ctx, image_j, &dj));
    IS_TRY(stage_in(ctx, mask_i, &mi));
    IS_TRY(stage_in(ctx, mask_j, &mj));
    PairSeam chk(ctx, image_i->depth == IS_8U, nullptr);
    const bool overlap = std::max(tl_i.x, tl_j.x) < std::min(tl_i.x + image_i->cols, tl_j.x + image_j->cols) &&
                         std::max(tl_i.y, tl_j.y) < std::min(tl_i.y + image_i->rows, tl_j.y + image_j->rows);
    if (overlap) IS_TRY(chk.process(di, dj, Pt{tl_i.x, tl_i.y}, Pt{tl_j.x, tl_j.y}, mi, mj, mi, mj, 0, 1, /*structure_only=*/true));
    const is_seam_pair_impl* sp = reinterpret_cast<const is_seam_pair_impl*>(spec);
    PairSeam ref(ctx, true, nullptr);
    ref.fp_records = sp->records;
    ref.fp_states = sp->states;
    *same = chk.same_structure(ref) ? 1 : 0;
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IS_OK;
}

// Would pair (i, j) take the same decisions on (mask_i, mask_j_b) as on (mask_i, mask_j_a)?  Structure and plan of both
// settings through the batched path's machinery (row toggles + special points on the device, PairRuns on the host).
// *same = 1: identical; 0: different, or outside what the run tables cover (the caller then takes its general path).
int is_seam_pair_same_structure(is_ctx* ctx, const is_mat* mask_i, const is_mat* mask_j_a, const is_mat* mask_j_b, is_point tl_i, is_point tl_j, int* same) {
    if (!ctx || !same) return IS_ERR_BAD_ARG;
    *same = 0;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, mask_i, "mask_i"));
    IS_TRY(check_mat(ctx, mask_j_a, "mask_j_a"));
    IS_TRY(check_mat(ctx, mask_j_b, "mask_j_b"));
    IS_REQUIRE(ctx, mask_i->depth == IS_8U && mask_i->channels == 1 && mask_j_a->depth == IS_8U && mask_j_a->channels == 1 && mask_j_b->depth == IS_8U &&
                        mask_j_b->channels == 1 && mask_j_a->rows == mask_j_b->rows && mask_j_a->cols == mask_j_b->cols, IS_ERR_BAD_ARG,
               "masks must be CV_8U, both settings of mask j of equal size");
    DevMat mi, ma, mb;
    IS_TRY(stage_in(ctx, mask_i, &mi));
    IS_TRY(stage_in(ctx, mask_j_a, &ma));
    IS_TRY(stage_in(ctx, mask_j_b, &mb));
    const bool overlap = std::max(tl_i.x, tl_j.x) < std::min(tl_i.x + mi.cols, tl_j.x + ma.cols) && std::max(tl_i.y, tl_j.y) < std::min(tl_i.y + mi.rows, tl_j.y + ma.rows);
    if (!overlap) { *same = 1; return IS_OK; }
    StructureQuery Q;
    Q.masks.push_back(plain_mask(mi));
    Q.masks.push_back(plain_mask(ma));
    Q.masks.push_back(plain_mask(mb));
    Q.pairs.push_back(StructureQuery::PairQ{0, 1, Pt{tl_i.x, tl_i.y}, Pt{tl_j.x, tl_j.y}});
    Q.pairs.push_back(StructureQuery::PairQ{0, 2, Pt{tl_i.x, tl_i.y}, Pt{tl_j.x, tl_j.y}});
    IS_TRY(run_structure_query(ctx, Q));
    if (Q.pair_overflow[0] || Q.pair_overflow[1]) return IS_OK;
    PairRuns P[2];
    host_pool(ctx)->run(2, [&](size_t k) {
        P[k].setup(0, 1, Q.pairs[k].tl1, Q.pairs[k].tl2, &Q.runs[(size_t)Q.pairs[k].m1], &Q.runs[(size_t)Q.pairs[k].m2]);
        P[k].specials = Q.specials[k];
        P[k].build();
        if (!P[k].too_many_runs) P[k].plan();
    });
    if (P[0].too_many_runs || P[1].too_many_runs || P[0].unsupported || P[1].unsupported) return IS_OK;
    *same = P[0].same_structure(P[1]) ? 1 : 0;
    return IS_OK;
}

int is_seam_pair_destroy(is_seam_pair* p) {
    delete reinterpret_cast<is_seam_pair_impl*>(p);
    return IS_OK;
}

int is_mask_and(is_ctx* ctx, is_mat* dst, const is_mat* src) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, dst, "dst"));
    IS_TRY(check_mat(ctx, src, "src"));
    IS_REQUIRE(ctx, dst->depth == IS_8U && src->depth == IS_8U && dst->channels == 1 && src->channels == 1 && dst->rows == src->rows &&
                        dst->cols == src->cols, IS_ERR_BAD_ARG, "masks must be CV_8U of equal size");
    DevMat d, sm;
    IS_TRY(stage_out(ctx, dst, &d, true));
    IS_TRY(stage_in(ctx, src, &sm));
    IS_TRY(mask_and(ctx, d, sm));
    IS_TRY(commit(ctx, &d));
    return IS_OK;
}

// Host-only diagnostic (no device needed): the run-domain structure and plan of ONE pair, computed by the same code the
// batched path runs between its kernels (PairRuns), with the device-side inputs (row toggles, special points) produced by
// host loops.  out (int32): [too_many_runs, unsupported (1; 2 = staged plan: only its first round is listed), ncomps, nops, nrecords, union_tl.x, union_tl.y, states[ncomps],
// ops[nops][11] = (kind, c1, c2, p1.x, p1.y, p2.x, p2.y, rx, ry, rw, rh), records[nrecords][7] = (x, y, label, nl[4])];
// coordinates are union-frame.
int is_debug_seam_pair_plan(const uint8_t* mask1, int rows1, int cols1, size_t step1, int tl1x, int tl1y, const uint8_t* mask2, int rows2, int cols2,
                            size_t step2, int tl2x, int tl2y, int32_t* out, size_t cap, size_t* len) {
    if (!mask1 || !mask2 || !len) return IS_ERR_BAD_ARG;
    MaskRuns r1, r2;
    mask_runs_from_host(mask1, step1, rows1, cols1, &r1);
    mask_runs_from_host(mask2, step2, rows2, cols2, &r2);
    PairRuns P;
    P.setup(0, 1, Pt{tl1x, tl1y}, Pt{tl2x, tl2y}, &r1, &r2);
    if (P.iTl.x < P.iBr.x && P.iTl.y < P.iBr.y) {
        // special points: host restatement of k_special_points_batch
        auto at = [&](int k, int x, int y) { return k == 0 ? r1.inside(x - P.o1x, y - P.o1y) : r2.inside(x - P.o2x, y - P.o2y); };
        auto contour = [&](int k, int x, int y) { return at(k, x, y) && !(at(k, x - 1, y) && at(k, x + 1, y) && at(k, x, y - 1) && at(k, x, y + 1)); };
        auto close_to = [&](int k, int x, int y) {
            for (int dy = -2; dy <= 2; ++dy)
                for (int dx = -2; dx <= 2; ++dx) {
                    const int xx = x + dx, yy = y + dy;
                    if (xx >= 0 && xx < P.uw && yy >= 0 && yy < P.uh && contour(k, xx, yy)) return true;
                }
            return false;
        };
        for (int y = P.iTl.y - P.unionTl.y; y < P.iBr.y - P.unionTl.y; ++y)
            for (int x = P.iTl.x - P.unionTl.x; x < P.iBr.x - P.unionTl.x; ++x) {
                if (!at(0, x, y) || !at(1, x, y)) continue;
                const int nx[4] = {x - 1, x, x + 1, x}, ny[4] = {y, y - 1, y, y + 1};
                bool touches = false;
                for (int k = 0; k < 4; ++k) touches = touches || (at(0, nx[k], ny[k]) != at(1, nx[k], ny[k]));
                if (touches && close_to(0, x, y) && close_to(1, x, y)) P.specials.push_back(Pt{x, y});
            }
        const bool timing = getenv("IS_DEBUG_PLAN_TIMING") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        P.build();
        auto t1 = std::chrono::steady_clock::now();
        if (!P.too_many_runs) P.plan();
        auto t2 = std::chrono::steady_clock::now();
        if (timing) fprintf(stderr, "[plan timing] build %.3f ms, plan %.3f ms (uh=%d, runs=%d)\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                            std::chrono::duration<double, std::milli>(t2 - t1).count(), P.uh, P.row_off.empty() ? 0 : P.row_off.back());
    }
    std::vector<int32_t> v;
    size_t nrec = 0;
    for (auto& c : P.contours) nrec += c.size();
    v.push_back(P.too_many_runs); v.push_back(P.unsupported ? 1 : (P.blocked ? 2 : 0)); v.push_back(P.ncomps); v.push_back((int)P.ops.size()); v.push_back((int)nrec);   // 2: staged plan, only the first round is listed
    v.push_back(P.unionTl.x); v.push_back(P.unionTl.y);
    for (int st : P.states) v.push_back(st);
    for (auto& op : P.ops) for (int q : {op.kind, op.c1, op.c2, op.p1.x, op.p1.y, op.p2.x, op.p2.y, op.rx, op.ry, op.rw, op.rh}) v.push_back(q);
    for (auto& c : P.contours) for (auto& r : c) for (int q : {r.x, r.y, r.label, r.nl[0], r.nl[1], r.nl[2], r.nl[3]}) v.push_back(q);
    *len = v.size();
    if (out && cap >= v.size()) std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
    return IS_OK;
}

// Host-only diagnostic: finishes ONE pair on the CPU with the batched path's host code, given the seams (as the oracle / the
// reference trace them): plan -> run-domain updateLabelsUsingSeam (UlsRuns) -> clear intervals -> masks.  seams: for every
// kind-1 operation of the plan, in order, [npts, x0, y0, x1, y1, ...] in panorama coordinates ordered tip 1 -> tip 2
// (npts = 0: estimateSeam failed).  mask1 / mask2 are updated in place.  Returns IS_ERR_UNSUPPORTED when the plan does not
// cover the pair.
int is_debug_seam_pair_finish(uint8_t* mask1, int rows1, int cols1, size_t step1, int tl1x, int tl1y, uint8_t* mask2, int rows2, int cols2,
                              size_t step2, int tl2x, int tl2y, const int32_t* seams, size_t seams_len) {
    if (!mask1 || !mask2) return IS_ERR_BAD_ARG;
    MaskRuns r1, r2;
    mask_runs_from_host(mask1, step1, rows1, cols1, &r1);
    mask_runs_from_host(mask2, step2, rows2, cols2, &r2);
    PairRuns P;
    P.setup(0, 1, Pt{tl1x, tl1y}, Pt{tl2x, tl2y}, &r1, &r2);
    if (!(P.iTl.x < P.iBr.x && P.iTl.y < P.iBr.y)) return IS_OK;
    {
        auto at = [&](int k, int x, int y) { return k == 0 ? r1.inside(x - P.o1x, y - P.o1y) : r2.inside(x - P.o2x, y - P.o2y); };
        auto contour = [&](int k, int x, int y) { return at(k, x, y) && !(at(k, x - 1, y) && at(k, x + 1, y) && at(k, x, y - 1) && at(k, x, y + 1)); };
        auto close_to = [&](int k, int x, int y) {
            for (int dy = -2; dy <= 2; ++dy)
                for (int dx = -2; dx <= 2; ++dx) {
                    const int xx = x + dx, yy = y + dy;
                    if (xx >= 0 && xx < P.uw && yy >= 0 && yy < P.uh && contour(k, xx, yy)) return true;
                }
            return false;
        };
        for (int y = P.iTl.y - P.unionTl.y; y < P.iBr.y - P.unionTl.y; ++y)
            for (int x = P.iTl.x - P.unionTl.x; x < P.iBr.x - P.unionTl.x; ++x) {
                if (!at(0, x, y) || !at(1, x, y)) continue;
                const int nx[4] = {x - 1, x, x + 1, x}, ny[4] = {y, y - 1, y, y + 1};
                bool touches = false;
                for (int k = 0; k < 4; ++k) touches = touches || (at(0, nx[k], ny[k]) != at(1, nx[k], ny[k]));
                if (touches && close_to(0, x, y) && close_to(1, x, y)) P.specials.push_back(Pt{x, y});
            }
    }
    P.build();
    if (P.too_many_runs) return IS_ERR_UNSUPPORTED;
    P.plan();
    if (P.unsupported) return IS_ERR_UNSUPPORTED;
    // rounds of the staged plan: the seams are consumed in the order the operations are planned
    std::vector<std::unique_ptr<UlsRuns>> uls;
    std::vector<std::vector<int>> lanes;
    lanes.reserve(256);
    std::vector<const std::vector<Interval>*> flips, round_flips;
    size_t pos = 0;
    for (;;) {
        round_flips.clear();
        for (size_t q = P.round_begin; q < P.ops.size(); ++q) {
            const SeamOp op = P.ops[q];
            if (op.kind != 1) continue;
            if (pos >= seams_len) return IS_ERR_BAD_ARG;
            const int npts = seams[pos++];
            if (npts == 0) { flips.push_back(nullptr); round_flips.push_back(nullptr); continue; }
            if (pos + 2 * (size_t)npts > seams_len) return IS_ERR_BAD_ARG;
            Pt src{op.p1.x - op.rx, op.p1.y - op.ry}, dst{op.p2.x - op.rx, op.p2.y - op.ry};
            const bool horizontal = std::abs(dst.x - src.x) > std::abs(dst.y - src.y);
            bool swapped = false;
            if (horizontal) { if (src.x > dst.x) { std::swap(src, dst); swapped = true; } }
            else if (src.y > dst.y) { std::swap(src, dst); swapped = true; }
            const int s0 = horizontal ? src.x : src.y, s1 = horizontal ? dst.x : dst.y;
            if (npts != s1 - s0 + 1) return IS_ERR_BAD_ARG;
            if (lanes.size() == lanes.capacity()) return IS_ERR_UNSUPPORTED;          // the lane arrays must not move
            lanes.emplace_back((size_t)npts);
            for (int i = 0; i < npts; ++i) {
                const int k = swapped ? npts - 1 - i : i;                          // trace order tip 1 -> tip 2; steps ascend from the (swapped) source
                const int x = seams[pos + 2 * (size_t)k] - P.unionTl.x - op.rx, y = seams[pos + 2 * (size_t)k + 1] - P.unionTl.y - op.ry;
                lanes.back()[(size_t)i] = horizontal ? y : x;
                if ((horizontal ? x : y) != s0 + i) return IS_ERR_BAD_ARG;
            }
            pos += 2 * (size_t)npts;
            uls.emplace_back(new UlsRuns());
            UlsRuns& U = *uls.back();
            U.P = &P; U.op = op; U.horizontal = horizontal; U.s0 = s0; U.nseam = npts; U.lane = lanes.back().data();
            auto t0 = std::chrono::steady_clock::now();
            U.run();
            if (getenv("IS_DEBUG_PLAN_TIMING")) fprintf(stderr, "[finish timing] UlsRuns::run %.3f ms (nc=%zu, nseam=%d, flips=%zu)\n",
                                                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), P.contours[(size_t)op.c1].size(), npts, U.flips.size());
            if (U.too_many_regions) return IS_ERR_UNSUPPORTED;
            flips.push_back(&U.flips);
            round_flips.push_back(&U.flips);
        }
        if (!P.blocked) break;
        if (!P.apply_round(round_flips)) return IS_ERR_UNSUPPORTED;
        P.plan_resume();
        if (P.unsupported) return IS_ERR_UNSUPPORTED;
    }
    std::vector<ClearIv> clears;
    auto t1 = std::chrono::steady_clock::now();
    if (P.staged) {
        if (!P.apply_round(round_flips)) return IS_ERR_UNSUPPORTED;
        P.final_clears(&clears);
    } else {
        pair_clear_intervals(P, flips, &clears);
    }
    if (getenv("IS_DEBUG_PLAN_TIMING")) fprintf(stderr, "[finish timing] pair_clear_intervals %.3f ms (%zu intervals)\n",
                                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), clears.size());
    for (const ClearIv& c : clears)
        for (int x = c.x0; x < c.x1; ++x) {
            if (c.bits & 1) mask1[(size_t)(c.y - P.o1y) * step1 + (x - P.o1x)] = 0;
            if (c.bits & 2) mask2[(size_t)(c.y - P.o2y) * step2 + (x - P.o2x)] = 0;
        }
    return IS_OK;
}

// Host-only diagnostic: the waves the batched seam path would form for a set of warped image rectangles if every wave were
// accepted in full.  out (int32): [nwaves, then per wave: count, (i, j) x count].
int is_debug_seam_wave_schedule(int n, const is_point* corners, const is_size* sizes, int32_t* out, size_t cap, size_t* len) {
    if (n < 0 || !corners || !sizes || !len) return IS_ERR_BAD_ARG;
    std::vector<std::pair<int, int>> pairs, active;
    for (int i = 0; i + 1 < n; ++i)
        for (int j = i + 1; j < n; ++j) pairs.push_back({i, j});
    std::reverse(pairs.begin(), pairs.end());
    std::vector<int4> rects;
    for (auto& pr : pairs) {
        const int i = pr.first, j = pr.second;
        const int4 r = make_int4(std::max(corners[i].x, corners[j].x), std::max(corners[i].y, corners[j].y),
                                 std::min(corners[i].x + sizes[i].width, corners[j].x + sizes[j].width), std::min(corners[i].y + sizes[i].height, corners[j].y + sizes[j].height));
        if (r.x < r.z && r.y < r.w) { active.push_back(pr); rects.push_back(r); }
    }
    std::vector<char> finished(active.size(), 0);
    std::vector<int32_t> v(1, 0);
    size_t left = active.size();
    while (left > 0) {
        std::vector<size_t> wave;
        choose_wave(active, rects, finished, n, false, false, &wave);
        if (wave.empty()) return IS_ERR_INTERNAL;
        v[0]++;
        v.push_back((int32_t)wave.size());
        for (size_t k : wave) { v.push_back(active[k].first); v.push_back(active[k].second); finished[k] = 1; --left; }
    }
    *len = v.size();
    if (out && cap >= v.size()) std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
    return IS_OK;
}

// Diagnostic / tuning entry: `njobs` synthetic seams of lanes x steps through the DP kernels of formulation `variant`
// (0: barrier per step + in-kernel back-track, 1: halo windows + parallel back-track), `iters` timed launches.
// seam_out[njobs][steps] receives the seam lanes, ms[0] = mean milliseconds per launch group (CUDA events).
int is_debug_dp_bench(is_ctx* ctx, int lanes, int steps, int njobs, int variant, unsigned seed, int iters, int32_t* seam_out, float* ms) {
    if (!ctx || lanes < 8 || steps < 2 || njobs < 1 || iters < 1) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    DpShape S;
    dp_choose_shape(lanes, steps, 0, steps - 1, variant, &S);
    IS_REQUIRE(ctx, S.pitch >= lanes && S.pitch <= 12 * 1024, IS_ERR_UNSUPPORTED, "too many lanes");
    const size_t cells = (size_t)S.pitch * (steps + DP_ROW_PAD);
    const int nchunks = div_up(steps - 1, BT_CHUNK);
    DevBuf P, Q, ctl, res, map, tab;
    IS_TRY(P.alloc(ctx, sizeof(float) * cells * njobs + 64));
    IS_TRY(Q.alloc(ctx, sizeof(float) * cells * njobs + 64));
    IS_TRY(ctl.alloc(ctx, cells * njobs + 64));
    IS_TRY(res.alloc(ctx, sizeof(int) * (size_t)(steps + 4) * njobs));
    IS_TRY(map.alloc(ctx, sizeof(short) * (size_t)nchunks * S.pitch * njobs + 64));
    std::vector<DpArgs> da((size_t)njobs);
    std::vector<BtArgs> ba((size_t)njobs);
    for (int j = 0; j < njobs; ++j) {
        IS_LAUNCH(ctx, k_dp_bench_fill, dim3(div_up(S.pitch, 256), steps), 256, 0, P.as<float>() + cells * j, Q.as<float>() + cells * j, lanes, S.pitch, steps, seed + 977u * j);
        DpArgs& A = da[(size_t)j];
        A.P = P.as<float>() + cells * j; A.Q = Q.as<float>() + cells * j; A.control = ctl.as<uint8_t>() + cells * j;
        A.lanes = lanes; A.pitch = S.pitch; A.steps = steps;
        A.s0 = 0; A.lane0 = lanes / 2 - 7 * j; A.s1 = steps - 1;
        A.lane1 = std::min(lanes - 1, std::max(0, A.lane0 + ((j & 1) ? -1 : 1) * std::min(steps / 3, lanes / 5) + 11 * j));   // inside the source's cone
        A.seam_lane = res.as<int>() + (size_t)(steps + 4) * j + 4; A.reached = res.as<int>() + (size_t)(steps + 4) * j;
        A.G = S.G; A.D = S.D;
        ba[(size_t)j].A = A; ba[(size_t)j].map = map.as<short>() + (size_t)nchunks * S.pitch * j; ba[(size_t)j].nchunks = nchunks;
    }
    IS_TRY(tab.alloc(ctx, sizeof(DpArgs) * njobs + sizeof(BtArgs) * njobs + 64));
    IS_TRY(upload(ctx, tab.p, da.data(), sizeof(DpArgs) * njobs));
    BtArgs* bt_d = reinterpret_cast<BtArgs*>(tab.as<unsigned char>() + align_up(sizeof(DpArgs) * njobs, 16));
    IS_TRY(upload(ctx, bt_d, ba.data(), sizeof(BtArgs) * njobs));
    std::vector<DpShape> shapes((size_t)njobs, S);
    IS_CUDA(ctx, cudaMemsetAsync(res.p, 0, sizeof(int) * (size_t)(steps + 4) * njobs, ctx->stream));
    IS_TRY(launch_dp_all(ctx, shapes, tab.as<DpArgs>(), bt_d));                  // warm-up
    cudaEvent_t e0, e1;
    IS_CUDA(ctx, cudaEventCreate(&e0));
    IS_CUDA(ctx, cudaEventCreate(&e1));
    IS_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    for (int it = 0; it < iters; ++it) IS_TRY(launch_dp_all(ctx, shapes, tab.as<DpArgs>(), bt_d));
    IS_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms) ms[0] = t / iters;
    if (seam_out) {
        std::vector<int> h((size_t)(steps + 4) * njobs);
        IS_TRY(download(ctx, h.data(), res.p, sizeof(int) * h.size()));
        for (int j = 0; j < njobs; ++j)
            for (int k = 0; k < steps; ++k) seam_out[(size_t)j * steps + k] = h[(size_t)(steps + 4) * j] ? h[(size_t)(steps + 4) * j + 4 + k] : -1;
    }
    return IS_OK;
}

int is_seam_cost_maps(is_ctx* ctx, const is_mat* image1, const is_mat* image2, is_point tl1, is_point tl2, const is_mat* labels,
                      is_point union_tl, int label, is_rect roi, is_mat* costV, is_mat* costH) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaSetDevice(ctx->device));
    IS_TRY(check_mat(ctx, image1, "image1"));
    IS_TRY(check_mat(ctx, image2, "image2"));
    IS_TRY(check_mat(ctx, labels, "labels"));
    IS_TRY(check_mat(ctx, costV, "costV"));
    IS_TRY(check_mat(ctx, costH, "costH"));
    IS_REQUIRE(ctx, (image1->channels == 3 || image1->channels == 4) && image2->channels == image1->channels && image1->depth == image2->depth &&
                        (image1->depth == IS_8U || image1->depth == IS_32F), IS_ERR_BAD_ARG, "both images must have CV_32FC3 or CV_8UC3 type");
    IS_REQUIRE(ctx, labels->depth == IS_32S && labels->channels == 1, IS_ERR_BAD_ARG, "labels must be CV_32S");
    IS_REQUIRE(ctx, labels->device >= 0 || labels->step == (size_t)labels->cols * 4, IS_ERR_BAD_ARG, "host labels must be dense");
    IS_REQUIRE(ctx, costV->depth == IS_32F && costV->rows == roi.height && costV->cols == roi.width + 1, IS_ERR_BAD_ARG, "costV must be h x (w+1) CV_32F");
    IS_REQUIRE(ctx, costH->depth == IS_32F && costH->rows == roi.height + 1 && costH->cols == roi.width, IS_ERR_BAD_ARG, "costH must be (h+1) x w CV_32F");
    IS_REQUIRE(ctx, roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= labels->cols && roi.y + roi.height <= labels->rows, IS_ERR_BAD_ARG, "roi outside the label image");
    DevMat a, b, cv, ch;
    IS_TRY(stage_in(ctx, image1, &a));
    IS_TRY(stage_in(ctx, image2, &b));
    // labels: dense int32 on the device
    DevBuf lab;
    const int* lab_d = nullptr;
    if (labels->device >= 0 && labels->step == (size_t)labels->cols * 4) lab_d = (const int*)labels->data;
    else {
        IS_TRY(lab.alloc(ctx, sizeof(int) * (size_t)labels->rows * labels->cols));
        IS_CUDA(ctx, cudaMemcpy2DAsync(lab.p, (size_t)labels->cols * 4, labels->data, labels->step, (size_t)labels->cols * 4, labels->rows,
                                       labels->device >= 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
        lab_d = lab.as<int>();
    }
    IS_TRY(stage_out(ctx, costV, &cv, false));
    IS_TRY(stage_out(ctx, costH, &ch, false));
    Frame f{labels->cols, labels->rows, 0, 0, labels->cols, labels->rows, MaskView{nullptr, 0, 0, 0, 0, 0}, MaskView{nullptr, 0, 0, 0, 0, 0}};   // dense labels: window = frame
    const int dx1 = union_tl.x - tl1.x, dy1 = union_tl.y - tl1.y, dx2 = union_tl.x - tl2.x, dy2 = union_tl.y - tl2.y;
    dim3 block(32, 8), grid(div_up(roi.width + 1, 32), div_up(roi.height + 1, 8));
    if (image1->depth == IS_8U)
        IS_LAUNCH(ctx, k_cost_maps<uint8_t>, grid, block, 0, make_view<uint8_t>(a, dx1, dy1), make_view<uint8_t>(b, dx2, dy2), lab_d, f, label,
                  roi.x, roi.y, roi.width, roi.height, cv.ptr<float>(), cv.step, ch.ptr<float>(), ch.step);
    else
        IS_LAUNCH(ctx, k_cost_maps<float>, grid, block, 0, make_view<float>(a, dx1, dy1), make_view<float>(b, dx2, dy2), lab_d, f, label,
                  roi.x, roi.y, roi.width, roi.height, cv.ptr<float>(), cv.step, ch.ptr<float>(), ch.step);
    IS_TRY(commit(ctx, &cv));
    IS_TRY(commit(ctx, &ch));
    return IS_OK;
}

}  // extern "C"
