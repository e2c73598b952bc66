// seam_batch.inl -- the batched seam path: ALL image pairs of a call go through each kernel in one launch, the host is
// consulted three times per call instead of ~15 times per pair.  (included by seam.cu inside namespace is)
//
//   A  device  k_row_toggles_batch (every mask -> per-row toggle positions), k_special_points_batch (per pair: the handful of
//              pixels that can become seam tips)                                            -> one download
//      host    PairRuns::build / plan per pair (thread pool): components, contours, edges, conflict loop -> operations
//   C  device  k_label_window_batch, k_relabel_batch, k_cost_pq_batch, k_seam_dp_batch (one CTA per seam), the device part of
//              updateLabelsUsingSeam (class, paint, flood fill as CCL, neighbourhood gathers)  -> one download
//      host    the order-dependent walk + adjacency vote of updateLabelsUsingSeam per seam (thread pool)
//   E  device  k_uls_apply_batch, k_scatter_label_batch, k_pair_clears_batch (each pair's mask clears as bits, private)
//   F  device  toggles + special points of the masks every pair WOULD have seen in the reference's sequential loop (entry
//              masks minus the clears of the earlier pairs)                                   -> one download
//      host    PairRuns of those masks; identical structure and plan <=> the speculative result is the sequential loop's
//   G  device  k_apply_clears_batch: the clears go into the real masks
// Every pair of a wave starts from the same masks; the longest prefix (in the reference's order) proven independent of the
// earlier pairs' clears is accepted, the rest forms the next wave (a strip needs one wave, a mosaic whose images overlap
// mutually a few).  A pair the plan cannot cover (noisy masks, a component cut twice) takes the general path (PairSeam) alone.

constexpr int TG_CAP = 8;              // toggles per mask row the batched path handles (rows are a few runs; more -> general path)
constexpr int MAX_LAYERS = 8;          // earlier pairs whose clears a validation mask can carry
constexpr int SPECIAL_CAP = 512;       // candidate seam tips per pair (a panorama pair has a few dozen; more -> general path)

struct ClearLayer { const uint8_t* p; int pitch; int x0, y0, w, h; int bit; };   // rectangle in the mask's own coordinates

struct LayeredMask {                   // a mask minus the clears of some pairs
    const uint8_t* p; size_t step; int rows, cols;
    int nlayers;
    ClearLayer layer[MAX_LAYERS];
};

__device__ __forceinline__ bool lm_cleared(const LayeredMask& m, int x, int y) {
    for (int k = 0; k < m.nlayers; ++k) {
        const ClearLayer& L = m.layer[k];
        const int lx = x - L.x0, ly = y - L.y0;
        if ((unsigned)lx < (unsigned)L.w && (unsigned)ly < (unsigned)L.h && (L.p[(size_t)ly * L.pitch + lx] & L.bit)) return true;
    }
    return false;
}

__device__ __forceinline__ bool lm_at(const LayeredMask& m, int x, int y) {       // own coordinates
    if ((unsigned)x >= (unsigned)m.cols || (unsigned)y >= (unsigned)m.rows) return false;
    if (!m.p[(size_t)y * m.step + x]) return false;
    return m.nlayers == 0 || !lm_cleared(m, x, y);
}

struct ToggleJob { LayeredMask m; unsigned char* counts; unsigned short* xs; };   // counts[rows], xs[rows][TG_CAP]

// One warp per mask row (blockIdx.y = job): the x positions where (mask != 0) toggles, in increasing x; the state left of
// x = 0 is "outside".  16 pixels per lane and trip.  Rows with more than TG_CAP toggles set *overflow.
__global__ void __launch_bounds__(256) k_row_toggles_batch(const ToggleJob* __restrict__ jobs, int* __restrict__ overflow) {   // overflow[job]
    const ToggleJob& J = jobs[blockIdx.y];
    const int rows = J.m.rows, cols = J.m.cols;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* row = J.m.p + (size_t)y * J.m.step;
    const bool aligned = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
    unsigned short* out = J.xs + (size_t)y * TG_CAP;
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
        if (J.m.nlayers && bits) {                       // validation masks: drop the pixels an earlier pair has cleared
            for (int i = 0; i < 16; ++i)
                if ((bits >> i & 1u) && lm_cleared(J.m, x0 + i, y)) bits &= ~(1u << i);
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
            if (k < TG_CAP) out[k] = (unsigned short)(x0 + i);
            ++k;
        }
        n += __shfl_sync(0xffffffffu, incl, 31);
        carry = __shfl_sync(0xffffffffu, bits >> 15, 31);
    }
    if (lane == 0) {
        J.counts[y] = (unsigned char)min(n, 255);
        if (n > TG_CAP) overflow[blockIdx.y] = 1;
    }
}

struct SpecialJob {
    LayeredMask m1, m2;
    int o1x, o1y, o2x, o2y;            // mask origins in the union frame
    int uw, uh;
    int ix, iy, iw, ih;                // intersection rectangle (frame coordinates)
    const unsigned char* cnt1; const unsigned short* xs1;   // the two masks' row toggles (k_row_toggles_batch)
    const unsigned char* cnt2; const unsigned short* xs2;
    int2* out; int* count;
};

__device__ __forceinline__ bool sp_at(const LayeredMask& m, int ox, int oy, int x, int y) { return lm_at(m, x - ox, y - oy); }
__device__ __forceinline__ bool sp_contour(const LayeredMask& m, int ox, int oy, int x, int y) {   // contour{1,2}mask_ [SEAM]:165-186
    return sp_at(m, ox, oy, x, y) && !(sp_at(m, ox, oy, x - 1, y) && sp_at(m, ox, oy, x + 1, y) && sp_at(m, ox, oy, x, y - 1) && sp_at(m, ox, oy, x, y + 1));
}
__device__ __forceinline__ bool sp_close(const LayeredMask& m, int ox, int oy, int uw, int uh, int x, int y) {   // closeToContour [SEAM]:584-604
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= uh) continue;
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = x + dx;
            if (xx >= 0 && xx < uw && sp_contour(m, ox, oy, xx, yy)) return true;
        }
    }
    return false;
}

// bit i of the result: pixel mx0 + i (own coordinates) of mask row my is set, i < 18 -- from the row's toggles
__device__ __forceinline__ unsigned row_bits18(const unsigned char* __restrict__ counts, const unsigned short* __restrict__ xs, int rows, int cols, int mx0, int my) {
    if ((unsigned)my >= (unsigned)rows) return 0u;
    const int n = min((int)counts[my], TG_CAP);
    const uint4 v = *reinterpret_cast<const uint4*>(xs + (size_t)my * TG_CAP);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    unsigned bits = 0;
    for (int k = 0; k < n; ++k) {
        const int t = (int)((w[k >> 1] >> (16 * (k & 1))) & 0xffffu) - mx0;      // toggle position inside the window
        bits ^= t <= 0 ? 0x3ffffu : (t >= 18 ? 0u : (0x3ffffu << t) & 0x3ffffu);
    }
    if (mx0 + 18 > cols) bits &= mx0 < cols ? (1u << (cols - mx0)) - 1u : 0u;      // a row that ends inside the mask has no closing toggle
    return bits;
}

// The only pixels getSeamTips can ever pick ([SEAM]:621-629): pixels of both masks with a 4-neighbour inside exactly one
// mask (a contour pixel of an INTERS component touching a FIRST / SECOND component), close to both masks' contours.
// The scan works on the row toggles k_row_toggles_batch has just produced (16 pixels of a row per step as bit sets: two
// 16-byte loads per row instead of ten byte loads per pixel); only the few candidates look at the mask bytes themselves.
constexpr int SP_ROWS = 8;             // rows per thread
__global__ void __launch_bounds__(128) k_special_points_batch(const SpecialJob* __restrict__ jobs) {
    const SpecialJob& J = jobs[blockIdx.z];
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ly0 = blockIdx.y * SP_ROWS;
    if (16 * cx >= J.iw || ly0 >= J.ih) return;
    const int x0 = J.ix + 16 * cx;                                               // frame column of bit 1
    const int m1x = x0 - 1 - J.o1x, m2x = x0 - 1 - J.o2x;
    auto bits = [&](int y, unsigned* b1, unsigned* b2) {                         // frame row y
        *b1 = row_bits18(J.cnt1, J.xs1, J.m1.rows, J.m1.cols, m1x, y - J.o1y);
        *b2 = row_bits18(J.cnt2, J.xs2, J.m2.rows, J.m2.cols, m2x, y - J.o2y);
    };
    unsigned p1, p2, c1, c2, n1, n2;
    bits(J.iy + ly0 - 1, &p1, &p2);
    bits(J.iy + ly0, &c1, &c2);
    const int rows = min(SP_ROWS, J.ih - ly0);
    for (int r = 0; r < rows; ++r) {
        const int y = J.iy + ly0 + r;
        bits(y + 1, &n1, &n2);
        const unsigned both = c1 & c2, x_cur = c1 ^ c2, x_up = p1 ^ p2, x_dn = n1 ^ n2;
        unsigned cand = both & ((x_cur << 1) | (x_cur >> 1) | x_up | x_dn) & 0x1fffeu;   // bits 1..16: this chunk's pixels
        while (cand) {
            const int i = __ffs(cand) - 1;
            cand &= cand - 1;
            const int x = x0 - 1 + i;
            if (x >= J.ix + J.iw) break;
            if (sp_close(J.m1, J.o1x, J.o1y, J.uw, J.uh, x, y) && sp_close(J.m2, J.o2x, J.o2y, J.uw, J.uh, x, y)) {
                const int pos = atomicAdd(J.count, 1);
                if (pos < SPECIAL_CAP) J.out[pos] = make_int2(x, y);
            }
        }
        p1 = c1; p2 = c2; c1 = n1; c2 = n2;
    }
}

// ---- per pair / per seam tables ----------------------------------------------------------------------------------------
struct PairDev {
    Frame fr;                          // union frame, label window, the ENTRY masks placed in the frame
    int* labels;
    const int* tab_cnt; const ChangePt* tab_cps; const int* tab_lab; int wcap;   // biased by the window's first row (k_label_window)
    const void* img1; const void* img2; size_t step1, step2; int rows1, cols1, rows2, cols2; int dx1, dy1, dx2, dy2;
    uint8_t* clear; int cpitch; int ix, iy, iw, ih;     // clear bits over the intersection rectangle: 1 = first mask, 2 = second mask
    const int* states;                 // final states of the components
    uint8_t* mask1; size_t mstep1; uint8_t* mask2; size_t mstep2;   // the real masks (written by k_apply_clears_batch only)
    GradView g;
};

struct JobDev {
    int pair;
    int l1, l2;
    int rx, ry, rw, rh;
    int horizontal, lanes, steps, pitch;
    float* P; float* Q;
    int s0, s1;
    int* res;                          // [0] destination reached, [1] interior components, [2 .. 2 + nseam) seam lanes, gathers of the contour (8 nc), of the seam (3 nseam)
    uint8_t* klass; int* sub_parent;
    const int2* cpts; int nc;
};

struct JobDevE { const int* adj_roots; int nadj; const int2* flips; int nflips; };
struct RelabelDev { int pair, x0, y0, w, h, from, to; };

__global__ void k_label_window_batch(const PairDev* __restrict__ pairs) {
    const PairDev& D = pairs[blockIdx.z];
    const Frame& f = D.fr;
    const int x = f.wx + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = f.wy + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= f.wx + f.ww || y >= f.wy + f.wh) return;
    int lo = y * D.wcap, hi = y * D.wcap + D.tab_cnt[y] - 1, best = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (D.tab_cps[mid].x <= x) { best = mid; lo = mid + 1; } else hi = mid - 1;
    }
    D.labels[lidx(f, x, y)] = best >= 0 ? D.tab_lab[best] : 0;
}

__global__ void k_relabel_batch(const PairDev* __restrict__ pairs, const RelabelDev* __restrict__ ops) {
    const RelabelDev R = ops[blockIdx.z];
    const PairDev& D = pairs[R.pair];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= R.w || y >= R.h) return;
    int* p = D.labels + lidx(D.fr, R.x0 + x, R.y0 + y);
    if (*p == R.from) *p = R.to;
}

template <typename T, bool GRAD>
__global__ void k_cost_pq_batch(const PairDev* __restrict__ pairs, const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.z];
    const PairDev& D = pairs[J.pair];
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    const int step = blockIdx.y * blockDim.y + threadIdx.y;
    if (lane >= J.pitch || step >= J.steps) return;
    float* P = J.P;
    float* Q = J.Q;
    if (lane >= J.lanes) {   // padding lanes: outside the component
        P[(size_t)step * J.pitch + lane] = __int_as_float(0x7f800000);
        Q[(size_t)step * J.pitch + lane] = 0.f;
        return;
    }
    const ImgView<T> a{reinterpret_cast<const T*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1};
    const ImgView<T> b{reinterpret_cast<const T*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2};
    const int x = J.rx + (J.horizontal ? step : lane), y = J.ry + (J.horizontal ? lane : step);
    float p, q;
    if (J.horizontal) { p = cost_h<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); q = cost_v<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); }
    else { p = cost_v<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); q = cost_h<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); }
    if (lab(D.labels, D.fr, x, y) != J.l1) p = __int_as_float(0x7f800000);   // +inf: the cell can never be on a path
    P[(size_t)step * J.pitch + lane] = p;
    Q[(size_t)step * J.pitch + lane] = q;
}

// ---- updateLabelsUsingSeam, device part, all seams at once (same arithmetic as the k_uls_* kernels) -----------------------
__global__ void k_uls_class_batch(const PairDev* __restrict__ pairs, const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.z];
    if (!J.res[0]) return;                                                       // estimateSeam failed: nothing to update
    const PairDev& D = pairs[J.pair];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= J.rw || y >= J.rh) return;
    const int ux = J.rx + x, uy = J.ry + y;
    int k = 0;
    if (lab(D.labels, D.fr, ux, uy) == J.l1) k = is_contour(D.labels, D.fr, ux, uy, J.l1) ? 2 : 1;
    J.klass[(size_t)y * J.rw + x] = (uint8_t)k;
}

__global__ void k_uls_paint_seam_batch(const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    if (!J.res[0]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > J.s1 - J.s0) return;
    const int step = J.s0 + i, lane = J.res[2 + i];
    const int x = J.horizontal ? step : lane, y = J.horizontal ? lane : step;
    J.klass[(size_t)y * J.rw + x] = 2;
}

__global__ void k_ccl_rows_batch(const JobDev* __restrict__ jobs) {   // k_ccl_rows, kmask 3
    const JobDev& J = jobs[blockIdx.y];
    if (!J.res[0]) return;
    const int y = blockIdx.x, w = J.rw;
    if (y >= J.rh) return;
    const uint8_t* row = J.klass + (size_t)y * w;
    int* prow = J.sub_parent + (size_t)y * w;
    __shared__ int warp_max[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < w; base += blockDim.x) {
        const int x = base + threadIdx.x;
        int k = 0, start = -1;
        if (x < w) {
            k = row[x] & 3;
            const int kprev = x > 0 ? (row[x - 1] & 3) : -1;
            if (k != kprev) start = x;
        }
        int v = start;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = max(v, t);
        }
        if (lane == 31) warp_max[wid] = v;
        __syncthreads();
        if (wid == 0) {
            int t = lane < nw ? warp_max[lane] : -1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t = max(t, u);
            }
            warp_max[lane] = t;
        }
        __syncthreads();
        const int prefix = wid > 0 ? warp_max[wid - 1] : -1;
        v = max(max(v, prefix), carry_s);
        if (x < w) prow[x] = k ? y * w + v : -1;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = v;
        __syncthreads();
    }
}

__global__ void k_ccl_merge_batch(const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.z];
    if (!J.res[0]) return;
    const int w = J.rw, h = J.rh;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (x >= w || y >= h) return;
    const uint8_t* r1 = J.klass + (size_t)y * w;
    const uint8_t* r0 = r1 - w;
    const int k = r1[x] & 3;
    if (!k || (r0[x] & 3) != k) return;
    const bool start1 = x == 0 || (r1[x - 1] & 3) != k;
    const bool start0 = x == 0 || (r0[x - 1] & 3) != k;
    if (start1 || start0) uf_union(J.sub_parent, y * w + x, (y - 1) * w + x);
}

// flatten + count the interior components (klass 1 roots)
__global__ void k_ccl_flatten_batch(const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    if (!J.res[0]) return;
    const size_t n = (size_t)J.rw * J.rh;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = J.sub_parent[i];
    if (p < 0) return;
    volatile int* vp = J.sub_parent;
    while (true) { const int q = vp[p]; if (q == p) break; p = q; }
    J.sub_parent[i] = p;
    if (p == (int)i && J.klass[i] == 1) atomicAdd(J.res + 1, 1);
}

// the neighbourhoods the host walk needs: 8 values per contour pixel ([SEAM]:989-990 order), then (x, y, value) per seam pixel
__global__ void k_uls_gather_batch(const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    if (!J.res[0]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nseam = J.s1 - J.s0 + 1;
    int* g8 = J.res + 2 + nseam;
    int* gs = g8 + 8 * (size_t)J.nc;
    if (i < J.nc) {
        const int dx[8] = {-1, +1, 0, 0, -1, +1, -1, +1};
        const int dy[8] = {0, 0, -1, +1, -1, -1, +1, +1};
        const int2 p = J.cpts[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) g8[(size_t)i * 8 + j] = uls_value(J.klass, J.sub_parent, J.rw, J.rh, p.x + dx[j], p.y + dy[j]);
    } else if (i < J.nc + nseam) {
        const int k = i - J.nc;
        const int step = J.s0 + k, lane = J.res[2 + k];
        const int x = J.horizontal ? step : lane, y = J.horizontal ? lane : step;
        gs[3 * k] = x;
        gs[3 * k + 1] = y;
        gs[3 * k + 2] = J.horizontal ? uls_value(J.klass, J.sub_parent, J.rw, J.rh, x, y + 1) : uls_value(J.klass, J.sub_parent, J.rw, J.rh, x + 1, y);
    }
}

__global__ void k_uls_apply_batch(const PairDev* __restrict__ pairs, const JobDev* __restrict__ jobs, const JobDevE* __restrict__ ext) {
    const JobDev& J = jobs[blockIdx.z];
    const JobDevE E = ext[blockIdx.z];
    if (!J.res[0] || E.nadj == 0) return;
    const PairDev& D = pairs[J.pair];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= J.rw || y >= J.rh) return;
    const size_t i = (size_t)y * J.rw + x;
    if (J.klass[i] != 1) return;
    const int r = J.sub_parent[i];
    int lo = 0, hi = E.nadj - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = E.adj_roots[mid];
        if (v == r) { D.labels[lidx(D.fr, J.rx + x, J.ry + y)] = J.l2; return; }
        if (v < r) lo = mid + 1; else hi = mid - 1;
    }
}

__global__ void k_scatter_label_batch(const PairDev* __restrict__ pairs, const JobDev* __restrict__ jobs, const JobDevE* __restrict__ ext) {
    const JobDev& J = jobs[blockIdx.y];
    const JobDevE E = ext[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.nflips) return;
    const PairDev& D = pairs[J.pair];
    D.labels[lidx(D.fr, E.flips[i].x, E.flips[i].y)] = J.l2;
}

// The final mask update [SEAM]:527-545 of a pair as clear bits over its intersection rectangle: mask2 loses the pixels whose
// label's state has FIRST where mask1 is set, then mask1 loses those whose state has SECOND where the UPDATED mask2 is set.
__global__ void k_pair_clears_batch(const PairDev* __restrict__ pairs) {
    const PairDev& D = pairs[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= D.iw || y >= D.ih) return;
    const int ux = D.ix + x, uy = D.iy + y;
    const int l = lab(D.labels, D.fr, ux, uy);
    const int st = l > 0 ? D.states[l - 1] : 0;
    const int m1 = D.fr.m1.at(ux, uy), m2 = D.fr.m2.at(ux, uy);
    const bool c2 = (st & ST_FIRST) && m1;
    const int m2n = c2 ? 0 : m2;
    const bool c1 = (st & ST_SECOND) && m2n;
    D.clear[(size_t)y * D.cpitch + x] = (uint8_t)((c1 && m1 ? 1 : 0) | (c2 && m2 ? 2 : 0));
}

__global__ void k_apply_clears_batch(const PairDev* __restrict__ pairs) {
    const PairDev& D = pairs[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= D.iw || y >= D.ih) return;
    const int c = D.clear[(size_t)y * D.cpitch + x];
    if (!c) return;
    const int ux = D.ix + x, uy = D.iy + y;
    if (c & 1) D.mask1[(size_t)(uy - D.fr.m1.oy) * D.mstep1 + (ux - D.fr.m1.ox)] = 0;
    if (c & 2) D.mask2[(size_t)(uy - D.fr.m2.oy) * D.mstep2 + (ux - D.fr.m2.ox)] = 0;
}

// ---- DP launches ----------------------------------------------------------------------------------------------------------
// Two formulations of the forward pass (seam.cu): 0 = one __syncthreads per step over a TMA-fed shared-memory ring, back-track
// inside the kernel; 1 = warp-private halo windows, one __syncthreads per H steps, parallel back-track kernels.
struct DpShape {
    int v1 = 0;                        // formulation
    int tmpl = 0, nwarps = 0;          // v1: template choice, warps per CTA
    int lpt = 4, nt = 32;              // v0
    int pitch = 128, G = 1, D = 2;
    int lanes = 0, s0 = 0, s1 = 0;
    bool same(const DpShape& o) const { return v1 == o.v1 && (v1 ? (tmpl == o.tmpl && nwarps == o.nwarps) : (lpt == o.lpt && nt == o.nt)); }
};

static int dp_variant_default() {
    if (const char* e = getenv("IS_DP_VARIANT")) return atoi(e) ? 1 : 0;
    return 1;
}

// window shapes of k_seam_fwd<LPT, H, R>: {LPT, H, owned lanes per warp}
static const int V1_TMPL[4][3] = {{4, 8, 112}, {8, 16, 224}, {16, 16, 480}, {4, 16, 96}};

static void dp_choose_shape(int lanes, int steps, int s0, int s1, int variant, DpShape* S) {
    S->lanes = lanes; S->s0 = s0; S->s1 = s1;
    S->lpt = lanes <= 4096 ? 4 : (lanes <= 8192 ? 8 : 16);
    if (const char* e = getenv("IS_DP_LPT")) {
        const int v = atoi(e);
        if ((v == 4 || v == 8 || v == 16) && lanes <= 1024 * v) S->lpt = v;
    }
    S->nt = std::min(1024, div_up(div_up(lanes, S->lpt), 32) * 32);
    S->pitch = S->nt * S->lpt;
    const size_t row_pair = 2 * sizeof(float) * (size_t)S->pitch;
    S->D = 2;
    S->G = (int)std::min<size_t>(16, (192 * 1024) / (S->D * row_pair));
    if (S->G < 1) S->G = 1;
    if (const char* e = getenv("IS_DP_G")) S->G = std::max(1, atoi(e));
    if (const char* e = getenv("IS_DP_D")) S->D = std::max(2, atoi(e));
    S->v1 = 0;
    if (variant == 1) {
        int forced = -1;
        if (const char* e = getenv("IS_DP_V1_TMPL")) forced = atoi(e);
        for (int t = 0; t < 4; ++t) {
            if (forced >= 0 ? t != forced : t == 3) continue;
            if (lanes <= 16 * V1_TMPL[t][2]) { S->v1 = 1; S->tmpl = t; S->nwarps = div_up(lanes, V1_TMPL[t][2]); break; }
        }
    }
    (void)steps;
}

template <int LPT>
static int launch_dp_v0(is_ctx* ctx, const DpArgs* table_d, int njobs, int nt, size_t smem, double bytes) {
    IS_CUDA(ctx, cudaFuncSetAttribute(k_seam_dp_batch<LPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->next_bytes = bytes;
    IS_LAUNCH(ctx, k_seam_dp_batch<LPT>, njobs, nt, smem, table_d);
    return IS_OK;
}

// shapes[q] describes table entry q; entries of equal shape are adjacent.  bt_d: BtArgs of the v1 entries (same indices).
static int launch_dp_all(is_ctx* ctx, const std::vector<DpShape>& shapes, const DpArgs* dp_d, const BtArgs* bt_d) {
    const size_t nj = shapes.size();
    int v1_first = -1, v1_count = 0, max_lanes = 0, max_chunks = 0;
    for (size_t q = 0; q < nj;) {
        size_t e = q;
        const DpShape& S0 = shapes[q];
        double bytes = 0;
        while (e < nj && shapes[e].same(S0)) { bytes += (double)(shapes[e].s1 - shapes[e].s0) * shapes[e].lanes * 9; ++e; }
        const int cnt = (int)(e - q);
        if (!S0.v1) {
            const size_t row_pair = 2 * sizeof(float) * (size_t)S0.pitch;
            const size_t smem = std::max<size_t>((size_t)S0.D * S0.G * row_pair + 8 * (size_t)S0.D + 16, (size_t)32 * 65 + 16);
            IS_REQUIRE(ctx, smem <= 200 * 1024, IS_ERR_INTERNAL, "DP shared-memory budget");
            switch (S0.lpt) {
                case 4: IS_TRY(launch_dp_v0<4>(ctx, dp_d + q, cnt, S0.nt, smem, bytes)); break;
                case 8: IS_TRY(launch_dp_v0<8>(ctx, dp_d + q, cnt, S0.nt, smem, bytes)); break;
                default: IS_TRY(launch_dp_v0<16>(ctx, dp_d + q, cnt, S0.nt, smem, bytes)); break;
            }
        } else {
            const int H = V1_TMPL[S0.tmpl][1];
            const size_t smem = 2 * sizeof(float) * (size_t)(S0.pitch + 2 * H);
            ctx->next_bytes = bytes;
            switch (S0.tmpl) {
                case 0: IS_LAUNCH(ctx, (k_seam_fwd<4, 8, 4>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
                case 1: IS_LAUNCH(ctx, (k_seam_fwd<8, 16, 2>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
                case 2: IS_LAUNCH(ctx, (k_seam_fwd<16, 16, 1>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
                default: IS_LAUNCH(ctx, (k_seam_fwd<4, 16, 4>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
            }
            if (v1_first < 0) v1_first = (int)q;
            v1_count = (int)e - v1_first;                               // v1 entries are contiguous (sorted by formulation first)
            for (size_t k = q; k < e; ++k) { max_lanes = std::max(max_lanes, shapes[k].lanes); max_chunks = std::max(max_chunks, div_up(shapes[k].s1 - shapes[k].s0, BT_CHUNK)); }
        }
        q = e;
    }
    if (v1_count > 0 && max_chunks > 0) {
        IS_LAUNCH(ctx, k_bt_compose, dim3(div_up(max_lanes, 256), max_chunks, v1_count), 256, 0, bt_d + v1_first);
        IS_LAUNCH(ctx, k_bt_walk, v1_count, 256, sizeof(int) * (size_t)max_chunks, bt_d + v1_first);
    }
    return IS_OK;
}

// ---- host ------------------------------------------------------------------------------------------------------------------

// a device arena + its pinned host mirror: everything a phase uploads goes up in ONE copy
struct Blob {
    std::vector<unsigned char> host;
    size_t put(const void* src, size_t bytes, size_t align = 16) {
        const size_t off = align_up(host.size(), align);
        host.resize(off + bytes);
        if (src && bytes) std::memcpy(host.data() + off, src, bytes);
        return off;
    }
};

struct SeamJobHost {
    int pair;                          // index into the active pair list
    SeamOp op;
    bool horizontal = false, swapped = false;
    int lanes = 0, steps = 0, lpt = 4, nt = 32, pitch = 128;
    int s0 = 0, lane0 = 0, s1 = 0, lane1 = 0, nseam = 0, nc = 0;
    DpShape shape;
    size_t off_map = 0;
    size_t off_P = 0, off_Q = 0, off_ctl = 0, off_klass = 0, off_parent = 0, off_res = 0;   // device arena offsets
    size_t res_ints = 0;
    std::vector<int> adj_roots;
    std::vector<int2> flips;
    std::vector<int32_t> trace;
    int status = IS_OK;
};

static HostPool* host_pool(is_ctx* ctx) {
    if (!ctx->hpool) {
        unsigned hc = std::thread::hardware_concurrency();
        size_t w = hc > 2 ? std::min<size_t>(hc - 1, 7) : 0;
        if (const char* e = getenv("IS_SEAM_HOST_THREADS")) w = (size_t)std::max(0, atoi(e) - 1);
        ctx->hpool = new HostPool(w);
    }
    return ctx->hpool;
}

// the order-dependent part of updateLabelsUsingSeam ([SEAM]:983-1085) for one seam, from the gathered neighbourhoods
static int uls_host_walk(is_ctx* ctx, const PairRuns& PR, SeamJobHost& J, const int* res, TraceSink* trace_on) {
    const SeamOp& op = J.op;
    const int l1 = op.c1 + 1, l2 = op.c2 + 1;
    const int rx = op.rx, ry = op.ry;
    const int nseam = J.nseam, nc = J.nc;
    J.adj_roots.clear(); J.flips.clear(); J.trace.clear();
    if (!res[0]) return IS_OK;                                        // [SEAM]:918-919: estimateSeam returned false
    const int nsub_total = res[1];
    const int* lane_h = res + 2;
    const int* g8 = res + 2 + nseam;
    const int* gs = g8 + 8 * (size_t)nc;
    const bool horizontal = J.horizontal, swapped = J.swapped;
    // seam in union-frame coordinates, ordered p1 -> p2 ([SEAM]:949-954)
    {
        const int first = swapped ? nseam - 1 : 0, last = swapped ? 0 : nseam - 1;
        auto pt = [&](int i) { const int step = J.s0 + i, lane = lane_h[i]; return horizontal ? Pt{step + rx, lane + ry} : Pt{lane + rx, step + ry}; };
        const Pt a = pt(first), b = pt(last);
        if (!(a.x == op.p1.x && a.y == op.p1.y && b.x == op.p2.x && b.y == op.p2.y))
            return fail(ctx, IS_ERR_ASSERT, "seam end points differ from the seam tips ([SEAM]:953-954)");
        if (trace_on) {
            J.trace.resize(5 + 2 * (size_t)nseam);
            int32_t* t = J.trace.data();
            t[0] = PR.pi; t[1] = PR.pj; t[2] = op.c1; t[3] = horizontal ? 1 : 0; t[4] = nseam;
            for (int i = 0; i < nseam; ++i) {
                const Pt q = pt(swapped ? nseam - 1 - i : i);
                t[5 + 2 * i] = q.x + PR.unionTl.x; t[6 + 2 * i] = q.y + PR.unionTl.y;
            }
        }
    }
    if (nsub_total >= 255) return IS_ERR_UNSUPPORTED;                 // the reference's mask value 255 would collide with a component id: general path
    // interior components are identified by their root index; ids 1.. in order of first appearance (only equality, "> 0" and
    // "!= 255" are ever asked of them, and there are fewer than 255)
    int roots[256];
    int nroots = 0;
    auto id_of = [&](int v) -> int {   // gathered value (> 0: root index + 1) -> reference mask value
        for (int k = 0; k < nroots; ++k) if (roots[k] == v - 1) return k + 1;
        roots[nroots] = v - 1;
        return ++nroots;
    };
    const std::vector<ContourRec>& cont = PR.contours[(size_t)op.c1];
    // The reference paints contour and seam pixels 255 and then assigns them one by one.  Contour pixels go first, in raster
    // order: a painted neighbour counts only once it has been assigned, i.e. when it is a contour pixel EARLIER in raster order
    // (neighbours 0, 2, 4, 5 of the reference's list: left, up, up-left, up-right); later contour pixels and all seam pixels
    // still hold 255.  The records are raster ordered, so a neighbour is found by a search inside its row's slice.
    const int rh = op.rh;
    std::vector<int> row_first((size_t)rh + 1, 0);
    for (int i = 0; i < nc; ++i) row_first[(size_t)(cont[(size_t)i].y - ry) + 1]++;
    for (int y = 0; y < rh; ++y) row_first[(size_t)y + 1] += row_first[(size_t)y];
    auto find_contour = [&](int x, int y) -> int {                     // index of the contour record at bbox position (x, y), -1 if none
        if ((unsigned)y >= (unsigned)rh) return -1;
        int lo = row_first[(size_t)y], hi = row_first[(size_t)y + 1] - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int cx = cont[(size_t)mid].x - rx;
            if (cx == x) return mid;
            if (cx < x) lo = mid + 1; else hi = mid - 1;
        }
        return -1;
    };
    static const int ddx[8] = {-1, +1, 0, 0, -1, +1, -1, +1};
    static const int ddy[8] = {0, 0, -1, +1, -1, -1, +1, +1};
    std::vector<int> val((size_t)nc, 255);                             // current mask value of every contour pixel
    for (int i = 0; i < nc; ++i) {
        const int x = cont[(size_t)i].x - rx, y = cont[(size_t)i].y - ry;
        int v = 0;
        // the last neighbour (in the reference's order) with an assigned value wins: scan backwards, stop at the first hit
        for (int j = 7; j >= 0; --j) {
            const int g = g8[(size_t)i * 8 + j];
            if (g > 0) { v = id_of(g); break; }                        // interior pixel of a flood-filled component
            if (g == -255 && (j == 0 || j == 2 || j == 4 || j == 5)) { // painted, earlier in raster order: assigned if it is a contour pixel
                const int k = find_contour(x + ddx[j], y + ddy[j]);
                if (k >= 0 && k < i && val[(size_t)k] > 0 && val[(size_t)k] != 255) { v = val[(size_t)k]; break; }
            }
        }
        val[(size_t)i] = v;
    }
    // then the seam pixels: each looks at one neighbour of its own step (never a seam pixel itself), so their order is irrelevant;
    // a seam pixel that is also a contour pixel is assigned a second time
    std::vector<int> sval((size_t)nseam, 0);
    std::vector<int> seam_contour((size_t)nseam, -1);
    for (int i = 0; i < nseam; ++i) {
        const int x = gs[3 * i], y = gs[3 * i + 1], g = gs[3 * i + 2];
        int v = 0;
        if (g > 0) v = id_of(g);
        else if (g == -255) {
            const int k = horizontal ? find_contour(x, y + 1) : find_contour(x + 1, y);
            if (k >= 0 && val[(size_t)k] > 0 && val[(size_t)k] != 255) v = val[(size_t)k];
        }
        sval[(size_t)i] = v;
        seam_contour[(size_t)i] = find_contour(x, y);
    }
    for (int i = 0; i < nseam; ++i)
        if (seam_contour[(size_t)i] >= 0) val[(size_t)seam_contour[(size_t)i]] = sval[(size_t)i];
    // adjacency vote ([SEAM]:1039-1085)
    const int nsub = nroots;
    std::vector<int> connect2((size_t)nsub + 1, 0), connectOther((size_t)nsub + 1, 0);
    bool c2_has0 = false, co_has0 = false;
    for (int i = 0; i < nc; ++i) {
        const ContourRec& r = cont[(size_t)i];
        int mv = val[(size_t)i];
        if (mv < 0 || mv > nsub) mv = 0;
        if (r.nl[0] == l2 || r.nl[1] == l2 || r.nl[2] == l2 || r.nl[3] == l2) { connect2[(size_t)mv]++; if (mv == 0) c2_has0 = true; }
        bool other = false;
        for (int k = 0; k < 4; ++k) if (r.nl[k] >= 0 && r.nl[k] != l1 && r.nl[k] != l2) other = true;
        if (other) { connectOther[(size_t)mv]++; if (mv == 0) co_has0 = true; }
    }
    std::vector<int> isAdj((size_t)nsub + 1, 0);
    const double len = (double)nc;
    for (int k = c2_has0 ? 0 : 1; k <= nsub; ++k) {
        int r = 0;
        if (connect2[(size_t)k] / len > 0.05) {
            const bool sub_exists = k >= 1 || co_has0;
            if (sub_exists && (connectOther[(size_t)k] / len < 0.1)) r = 1;
        }
        isAdj[(size_t)k] = r;
    }
    for (int i = 1; i <= nsub; ++i) if (isAdj[(size_t)i]) J.adj_roots.push_back(roots[i - 1]);
    std::sort(J.adj_roots.begin(), J.adj_roots.end());
    // painted pixels that ended up in an adjacent component are relabelled one by one ([SEAM]:1089-1092 covers them with the rest)
    for (int i = 0; i < nc; ++i) {
        const int v = val[(size_t)i];
        if (v > 0 && v <= nsub && isAdj[(size_t)v]) J.flips.push_back(make_int2(cont[(size_t)i].x, cont[(size_t)i].y));
    }
    for (int i = 0; i < nseam; ++i) {
        const int v = sval[(size_t)i];
        if (seam_contour[(size_t)i] < 0 && v > 0 && v <= nsub && isAdj[(size_t)v]) J.flips.push_back(make_int2(gs[3 * i] + rx, gs[3 * i + 1] + ry));
    }
    return IS_OK;
}

struct BatchTimer {
    bool on; std::chrono::steady_clock::time_point t;
    BatchTimer() : on(getenv("IS_SEAM_DEBUG") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[seam batch] %-34s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// toggles + special points for a set of (possibly layered) masks / pairs: launch, one download, parse
struct StructureQuery {
    std::vector<LayeredMask> masks;                    // toggle jobs
    struct PairQ { int m1, m2; Pt tl1, tl2; };         // indices into masks
    std::vector<PairQ> pairs;
    // results
    std::vector<MaskRuns> runs;
    std::vector<std::vector<Pt>> specials;
    std::vector<char> mask_overflow, pair_overflow;    // more toggles per row / more candidate tips than the batched path carries
};

static int run_structure_query(is_ctx* ctx, StructureQuery& Q) {
    const size_t nm = Q.masks.size(), np = Q.pairs.size();
    Q.runs.assign(nm, MaskRuns());
    Q.specials.assign(np, std::vector<Pt>());
    Q.mask_overflow.assign(nm, 0);
    Q.pair_overflow.assign(np, 0);
    if (nm == 0) return IS_OK;
    // device layout: [hdr: overflow flags nm, special counts np][special points np x SPECIAL_CAP][counts (bytes)][xs (u16)]   tables at the end
    size_t total_rows = 0;
    int max_rows = 0;
    for (auto& m : Q.masks) {
        total_rows += (size_t)m.rows; max_rows = std::max(max_rows, m.rows);
    }
    const size_t off_hdr = 0, hdr_bytes = align_up(sizeof(int) * (nm + np), 16);
    const size_t off_sp_dl = off_hdr + hdr_bytes, sp_dl_bytes = align_up(sizeof(int2) * SPECIAL_CAP * np, 16);
    const size_t off_cnt = off_sp_dl + sp_dl_bytes, cnt_bytes = align_up(total_rows, 16);
    const size_t off_xs = off_cnt + cnt_bytes, xs_bytes = align_up(total_rows * TG_CAP * sizeof(unsigned short), 16);
    const size_t dl_bytes = off_xs + xs_bytes;
    const size_t off_tab = align_up(dl_bytes, 16);
    const size_t tab_bytes = sizeof(ToggleJob) * nm + sizeof(SpecialJob) * np;
    DevBuf dev;
    IS_TRY(dev.alloc(ctx, off_tab + tab_bytes));
    unsigned char* base = dev.as<unsigned char>();
    std::vector<unsigned char> tabs(tab_bytes);
    ToggleJob* tj = reinterpret_cast<ToggleJob*>(tabs.data());
    SpecialJob* sj = reinterpret_cast<SpecialJob*>(tabs.data() + sizeof(ToggleJob) * nm);
    std::vector<size_t> row0(nm);
    size_t r = 0;
    for (size_t k = 0; k < nm; ++k) {
        row0[k] = r;
        tj[k].m = Q.masks[k];
        tj[k].counts = base + off_cnt + r;
        tj[k].xs = reinterpret_cast<unsigned short*>(base + off_xs) + r * TG_CAP;
        r += (size_t)Q.masks[k].rows;
    }
    int max_iw = 0, max_ih = 0;
    for (size_t k = 0; k < np; ++k) {
        const auto& pq = Q.pairs[k];
        const LayeredMask& a = Q.masks[(size_t)pq.m1];
        const LayeredMask& b = Q.masks[(size_t)pq.m2];
        SpecialJob& S = sj[k];
        S.m1 = a; S.m2 = b;
        const Pt utl{std::min(pq.tl1.x, pq.tl2.x), std::min(pq.tl1.y, pq.tl2.y)};
        const Pt ubr{std::max(pq.tl1.x + a.cols, pq.tl2.x + b.cols), std::max(pq.tl1.y + a.rows, pq.tl2.y + b.rows)};
        S.o1x = pq.tl1.x - utl.x; S.o1y = pq.tl1.y - utl.y; S.o2x = pq.tl2.x - utl.x; S.o2y = pq.tl2.y - utl.y;
        S.uw = ubr.x - utl.x; S.uh = ubr.y - utl.y;
        S.ix = std::max(S.o1x, S.o2x); S.iy = std::max(S.o1y, S.o2y);
        S.iw = std::min(S.o1x + a.cols, S.o2x + b.cols) - S.ix; S.ih = std::min(S.o1y + a.rows, S.o2y + b.rows) - S.iy;
        S.cnt1 = tj[(size_t)pq.m1].counts; S.xs1 = tj[(size_t)pq.m1].xs; S.cnt2 = tj[(size_t)pq.m2].counts; S.xs2 = tj[(size_t)pq.m2].xs;
        S.out = reinterpret_cast<int2*>(base + off_sp_dl) + k * SPECIAL_CAP;
        S.count = reinterpret_cast<int*>(base + off_hdr) + nm + k;
        max_iw = std::max(max_iw, S.iw); max_ih = std::max(max_ih, S.ih);
    }
    IS_CUDA(ctx, cudaMemsetAsync(base + off_hdr, 0, hdr_bytes, ctx->stream));
    IS_TRY(upload(ctx, base + off_tab, tabs.data(), tab_bytes));
    const ToggleJob* tj_d = reinterpret_cast<const ToggleJob*>(base + off_tab);
    const SpecialJob* sj_d = reinterpret_cast<const SpecialJob*>(base + off_tab + sizeof(ToggleJob) * nm);
    {
        double bytes = 0;
        for (auto& m : Q.masks) bytes += (double)m.rows * m.cols;
        ctx->next_bytes = bytes;
        dim3 grid(div_up(max_rows, 8), (unsigned)nm);
        IS_LAUNCH(ctx, k_row_toggles_batch, grid, 256, 0, tj_d, reinterpret_cast<int*>(base + off_hdr));
    }
    if (np && max_iw > 0 && max_ih > 0) {
        dim3 grid(div_up(div_up(max_iw, 16), 128), div_up(max_ih, SP_ROWS), (unsigned)np);
        IS_LAUNCH(ctx, k_special_points_batch, grid, 128, 0, sj_d);
    }
    const unsigned char* h = nullptr;                                   // view of the pinned bounce buffer, valid until the next download
    IS_TRY(download_view(ctx, base, dl_bytes, reinterpret_cast<const void**>(&h)));
    const int* hdr = reinterpret_cast<const int*>(h + off_hdr);
    for (size_t k = 0; k < nm; ++k) Q.mask_overflow[k] = hdr[k] != 0 || Q.masks[k].cols >= 65535;
    for (size_t k = 0; k < np; ++k)   // more candidate tips than any panorama mask produces, or a mask the run tables cannot hold: general path
        Q.pair_overflow[k] = hdr[nm + k] > SPECIAL_CAP || Q.mask_overflow[(size_t)Q.pairs[k].m1] || Q.mask_overflow[(size_t)Q.pairs[k].m2];
    for (size_t k = 0; k < nm; ++k) {
        if (Q.mask_overflow[k]) continue;
        MaskRuns& R = Q.runs[k];
        R.rows = Q.masks[k].rows; R.cols = Q.masks[k].cols; R.slots = TG_CAP;
        R.counts.assign(h + off_cnt + row0[k], h + off_cnt + row0[k] + (size_t)R.rows);
        const unsigned short* x = reinterpret_cast<const unsigned short*>(h + off_xs) + row0[k] * TG_CAP;
        R.xs.assign(x, x + (size_t)R.rows * TG_CAP);
    }
    for (size_t k = 0; k < np; ++k) {
        if (Q.pair_overflow[k]) continue;
        const int n = hdr[nm + k];
        const int2* p = reinterpret_cast<const int2*>(h + off_sp_dl) + k * SPECIAL_CAP;
        std::vector<Pt>& out = Q.specials[k];
        out.resize((size_t)n);
        for (int i = 0; i < n; ++i) out[(size_t)i] = Pt{p[i].x, p[i].y};
        std::sort(out.begin(), out.end(), [](const Pt& a, const Pt& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });   // raster order
    }
    return IS_OK;
}

static LayeredMask plain_mask(const DevMat& m) {
    LayeredMask L;
    std::memset(&L, 0, sizeof(L));
    L.p = m.ptr<uint8_t>(); L.step = m.step; L.rows = m.rows; L.cols = m.cols; L.nlayers = 0;
    return L;
}

// One wave: the pairs `active_in` (the reference's order) all start from the masks as they are now.  The longest prefix of them
// whose speculative results are proven to be the sequential loop's is applied to the masks; *accepted_n is its length (at
// least 1, unless the FIRST pair is outside what the batched path covers: then *first_unsupported is set and nothing is done).
static int seam_batch_wave(is_ctx* ctx, const std::vector<std::pair<int, int>>& active_in, int n, const DevMat* images, const is_point* corners,
                           const DevMat* masks, TraceSink* trace, int cost_fn, size_t* accepted_n, bool* first_unsupported) {
    *accepted_n = 0;
    *first_unsupported = false;
    std::vector<std::pair<int, int>> active = active_in;
    size_t np = active.size();
    if (np == 0) return IS_OK;
    const bool is_u8 = images[active[0].first].depth == IS_8U;
    BatchTimer tm;
    HostPool* pool = host_pool(ctx);
    // ---- A: structure of every pair on the entry masks
    std::vector<int> mask_slot((size_t)n, -1);
    StructureQuery Q;
    for (auto& pr : active)
        for (int img : {pr.first, pr.second})
            if (mask_slot[(size_t)img] < 0) { mask_slot[(size_t)img] = (int)Q.masks.size(); Q.masks.push_back(plain_mask(masks[img])); }
    for (auto& pr : active)
        Q.pairs.push_back(StructureQuery::PairQ{mask_slot[(size_t)pr.first], mask_slot[(size_t)pr.second], Pt{corners[pr.first].x, corners[pr.first].y},
                                                Pt{corners[pr.second].x, corners[pr.second].y}});
    IS_TRY(run_structure_query(ctx, Q));
    tm.lap("A toggles + special points");
    std::vector<PairRuns> PR(np);
    pool->run(np, [&](size_t k) {
        PairRuns& P = PR[k];
        if (Q.pair_overflow[k]) { P.unsupported = true; return; }
        P.setup(active[k].first, active[k].second, Q.pairs[k].tl1, Q.pairs[k].tl2, &Q.runs[(size_t)Q.pairs[k].m1], &Q.runs[(size_t)Q.pairs[k].m2]);
        P.specials = Q.specials[k];
        if ((size_t)P.uw * P.uh >= (size_t)INT_MAX) { P.unsupported = true; return; }
        P.build();
        if (P.too_many_runs) return;
        P.plan();
        for (const SeamOp& op : P.ops) {                               // seams wider than the DP kernels' tables: general path (it reports the limit)
            if (op.kind != 1) continue;
            const bool horizontal = std::abs(op.p2.x - op.p1.x) > std::abs(op.p2.y - op.p1.y);
            if ((horizontal ? op.rh : op.rw) > 12 * 1024 - 128) P.unsupported = true;
        }
    });
    tm.lap("A host: runs, contours, plan");
    // a pair the plan does not cover ends the wave in front of it (it goes through the general path, PairSeam, on its own)
    for (size_t k = 0; k < np; ++k)
        if (PR[k].too_many_runs || PR[k].unsupported) {
            if (k == 0) { *first_unsupported = true; return IS_OK; }
            np = k;
            active.resize(np);
            PR.resize(np);
            break;
        }
    // ---- C: plan -> device tables
    std::vector<SeamJobHost> jobs;
    std::vector<RelabelDev> pre_relabel, post_relabel;
    const int dp_variant = dp_variant_default();
    for (size_t k = 0; k < np; ++k) {
        std::vector<char> cut((size_t)PR[k].ncomps, 0);
        for (const SeamOp& op : PR[k].ops) {
            if (op.kind == 0) {
                RelabelDev R{(int)k, op.rx, op.ry, op.rw, op.rh, op.c1 + 1, op.c2 + 1};
                (cut[(size_t)op.c1] ? post_relabel : pre_relabel).push_back(R);
                continue;
            }
            cut[(size_t)op.c1] = 1;
            SeamJobHost J;
            J.pair = (int)k; J.op = op;
            Pt src{op.p1.x - op.rx, op.p1.y - op.ry}, dst{op.p2.x - op.rx, op.p2.y - op.ry};
            J.horizontal = std::abs(dst.x - src.x) > std::abs(dst.y - src.y);                     // [SEAM]:828
            if (J.horizontal) { if (src.x > dst.x) { std::swap(src, dst); J.swapped = true; } }
            else if (src.y > dst.y) { std::swap(src, dst); J.swapped = true; }
            J.lanes = J.horizontal ? op.rh : op.rw; J.steps = J.horizontal ? op.rw : op.rh;
            J.s0 = J.horizontal ? src.x : src.y; J.lane0 = J.horizontal ? src.y : src.x;
            J.s1 = J.horizontal ? dst.x : dst.y; J.lane1 = J.horizontal ? dst.y : dst.x;
            dp_choose_shape(J.lanes, J.steps, J.s0, J.s1, dp_variant, &J.shape);
            J.lpt = J.shape.lpt; J.nt = J.shape.nt; J.pitch = J.shape.pitch;
            IS_REQUIRE(ctx, J.pitch >= J.lanes && J.pitch <= 12 * 1024, IS_ERR_INTERNAL, "seam wider than planned");
            J.nseam = J.s1 - J.s0 + 1;
            J.nc = (int)PR[k].contours[(size_t)op.c1].size();
            jobs.push_back(std::move(J));
        }
    }
    const size_t nj = jobs.size();
    // device arena: per pair labels / clear bits, per job P, Q, control, klass, parent, results; blob 1 = tables + small inputs
    size_t arena = 0;
    auto take = [&](size_t bytes) { const size_t o = arena; arena = align_up(arena + bytes, 256); return o; };
    std::vector<size_t> off_labels(np), off_clear(np), off_grad(np, 0);
    std::vector<int> cpitch(np), gpitch(np, 0);
    for (size_t k = 0; k < np; ++k) {
        const PairRuns& P = PR[k];
        off_labels[k] = take(sizeof(int) * (size_t)P.ww * P.wh);
        const int iw = P.iBr.x - P.iTl.x, ih = P.iBr.y - P.iTl.y;
        cpitch[k] = (iw + 15) & ~15;
        off_clear[k] = take((size_t)cpitch[k] * ih);
        if (cost_fn == IS_COST_COLOR_GRAD) {
            gpitch[k] = (iw + 31) & ~31;
            off_grad[k] = take(sizeof(float) * (size_t)gpitch[k] * ih * 4);
        }
    }
    size_t res_total_ints = 0;
    for (auto& J : jobs) {
        J.off_P = take(sizeof(float) * (size_t)J.pitch * (J.steps + DP_ROW_PAD) + 64);
        J.off_Q = take(sizeof(float) * (size_t)J.pitch * (J.steps + DP_ROW_PAD) + 64);
        J.off_ctl = take((size_t)J.pitch * J.steps + 64);
        J.off_klass = take((size_t)J.op.rw * J.op.rh);
        J.off_parent = take(sizeof(int) * (size_t)J.op.rw * J.op.rh);
        if (J.shape.v1) J.off_map = take(sizeof(short) * (size_t)div_up(J.s1 - J.s0, BT_CHUNK) * J.pitch + 64);
        J.res_ints = 2 + (size_t)J.nseam + 8 * (size_t)J.nc + 3 * (size_t)J.nseam;
        J.off_res = res_total_ints;
        res_total_ints += (J.res_ints + 3) & ~(size_t)3;
    }
    const size_t off_res_all = take(sizeof(int) * std::max<size_t>(res_total_ints, 4));
    // blob 1
    Blob B1;
    std::vector<size_t> off_tab(np), off_cpts(nj);
    for (size_t k = 0; k < np; ++k) off_tab[k] = B1.put(PR[k].tab.data(), sizeof(int) * PR[k].tab.size());
    for (size_t j = 0; j < nj; ++j) {
        const SeamJobHost& J = jobs[j];
        const auto& cont = PR[(size_t)J.pair].contours[(size_t)J.op.c1];
        std::vector<int2> cpts((size_t)J.nc);
        for (int i = 0; i < J.nc; ++i) cpts[(size_t)i] = make_int2(cont[(size_t)i].x - J.op.rx, cont[(size_t)i].y - J.op.ry);
        off_cpts[j] = B1.put(cpts.data(), sizeof(int2) * cpts.size());
    }
    const size_t off_pairs = B1.put(nullptr, sizeof(PairDev) * np);
    const size_t off_jobs = B1.put(nullptr, sizeof(JobDev) * std::max<size_t>(nj, 1));
    const size_t off_dp = B1.put(nullptr, sizeof(DpArgs) * std::max<size_t>(nj, 1));
    const size_t off_bt = B1.put(nullptr, sizeof(BtArgs) * std::max<size_t>(nj, 1));
    const size_t off_pre = B1.put(pre_relabel.data(), sizeof(RelabelDev) * pre_relabel.size());
    const size_t off_b1 = take(B1.host.size());
    // blob 2 (after the host walk): states, adjacency roots, flips, post relabels -- sized now, filled later
    size_t b2_bytes = 0;
    std::vector<size_t> off_states(np);
    for (size_t k = 0; k < np; ++k) { off_states[k] = b2_bytes; b2_bytes = align_up(b2_bytes + sizeof(int) * (size_t)std::max(PR[k].ncomps, 1), 16); }
    const size_t b2_fixed = b2_bytes;
    size_t b2_cap = b2_fixed + sizeof(JobDevE) * std::max<size_t>(nj, 1) + sizeof(RelabelDev) * post_relabel.size() + 64;
    for (auto& J : jobs) b2_cap += sizeof(int) * 256 + sizeof(int2) * ((size_t)J.nc + (size_t)J.nseam) + 64;
    const size_t off_b2 = take(b2_cap);
    DevBuf dev;
    IS_TRY(dev.alloc(ctx, arena));
    unsigned char* base = dev.as<unsigned char>();
    unsigned char* b1d = base + off_b1;
    // group the seams by DP launch shape
    std::vector<size_t> order(nj);
    for (size_t j = 0; j < nj; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        const DpShape& x = jobs[a].shape;
        const DpShape& y = jobs[b].shape;
        return std::make_tuple(x.v1, x.v1 ? x.tmpl : x.lpt, x.v1 ? x.nwarps : x.nt) < std::make_tuple(y.v1, y.v1 ? y.tmpl : y.lpt, y.v1 ? y.nwarps : y.nt);
    });
    {
        PairDev* pd = reinterpret_cast<PairDev*>(B1.host.data() + off_pairs);
        for (size_t k = 0; k < np; ++k) {
            const PairRuns& P = PR[k];
            const DevMat& m1 = masks[P.pi];
            const DevMat& m2 = masks[P.pj];
            const DevMat& i1 = images[P.pi];
            const DevMat& i2 = images[P.pj];
            PairDev& D = pd[k];
            std::memset(&D, 0, sizeof(D));
            D.fr = Frame{P.uw, P.uh, P.wx, P.wy, P.ww, P.wh, MaskView{m1.ptr<uint8_t>(), m1.step, m1.rows, m1.cols, P.o1x, P.o1y},
                         MaskView{m2.ptr<uint8_t>(), m2.step, m2.rows, m2.cols, P.o2x, P.o2y}};
            D.labels = reinterpret_cast<int*>(base + off_labels[k]);
            const int* d = reinterpret_cast<const int*>(b1d + off_tab[k]);
            const size_t wrows = (size_t)P.wh;
            D.wcap = (int)P.wcap;
            D.tab_cnt = d - P.wy;
            D.tab_cps = reinterpret_cast<const ChangePt*>(d + wrows) - (size_t)P.wy * P.wcap;
            D.tab_lab = d + wrows + wrows * P.wcap * 2 - (size_t)P.wy * P.wcap;
            D.img1 = i1.data; D.img2 = i2.data; D.step1 = i1.step; D.step2 = i2.step;
            D.rows1 = i1.rows; D.cols1 = i1.cols; D.rows2 = i2.rows; D.cols2 = i2.cols;
            D.dx1 = P.unionTl.x - P.tl1.x; D.dy1 = P.unionTl.y - P.tl1.y; D.dx2 = P.unionTl.x - P.tl2.x; D.dy2 = P.unionTl.y - P.tl2.y;
            D.clear = base + off_clear[k]; D.cpitch = cpitch[k];
            D.ix = P.iTl.x - P.unionTl.x; D.iy = P.iTl.y - P.unionTl.y; D.iw = P.iBr.x - P.iTl.x; D.ih = P.iBr.y - P.iTl.y;
            D.states = reinterpret_cast<const int*>(base + off_b2 + off_states[k]);
            D.mask1 = m1.ptr<uint8_t>(); D.mstep1 = m1.step; D.mask2 = m2.ptr<uint8_t>(); D.mstep2 = m2.step;
            if (cost_fn == IS_COST_COLOR_GRAD) {
                float* g = reinterpret_cast<float*>(base + off_grad[k]);
                const size_t plane = (size_t)gpitch[k] * D.ih;
                D.g = GradView{g, g + plane, g + 2 * plane, g + 3 * plane, gpitch[k], D.ix, D.iy};
            }
        }
        JobDev* jd = reinterpret_cast<JobDev*>(B1.host.data() + off_jobs);
        DpArgs* da = reinterpret_cast<DpArgs*>(B1.host.data() + off_dp);
        BtArgs* ba = reinterpret_cast<BtArgs*>(B1.host.data() + off_bt);
        for (size_t q = 0; q < nj; ++q) {
            const SeamJobHost& J = jobs[order[q]];
            JobDev& D = jd[q];
            std::memset(&D, 0, sizeof(D));
            D.pair = J.pair; D.l1 = J.op.c1 + 1; D.l2 = J.op.c2 + 1;
            D.rx = J.op.rx; D.ry = J.op.ry; D.rw = J.op.rw; D.rh = J.op.rh;
            D.horizontal = J.horizontal ? 1 : 0; D.lanes = J.lanes; D.steps = J.steps; D.pitch = J.pitch;
            D.P = reinterpret_cast<float*>(base + J.off_P); D.Q = reinterpret_cast<float*>(base + J.off_Q);
            D.s0 = J.s0; D.s1 = J.s1;
            D.res = reinterpret_cast<int*>(base + off_res_all) + J.off_res;
            D.klass = base + J.off_klass; D.sub_parent = reinterpret_cast<int*>(base + J.off_parent);
            D.cpts = reinterpret_cast<const int2*>(b1d + off_cpts[order[q]]); D.nc = J.nc;
            DpArgs& A = da[q];
            A.P = D.P; A.Q = D.Q; A.control = base + J.off_ctl;
            A.lanes = J.lanes; A.pitch = J.pitch; A.steps = J.steps;
            A.s0 = J.s0; A.lane0 = J.lane0; A.s1 = J.s1; A.lane1 = J.lane1;
            A.seam_lane = D.res + 2; A.reached = D.res;
            A.G = J.shape.G; A.D = J.shape.D;
            ba[q].A = A;
            ba[q].map = reinterpret_cast<short*>(base + J.off_map);
            ba[q].nchunks = div_up(J.s1 - J.s0, BT_CHUNK);
        }
    }
    IS_TRY(upload(ctx, b1d, B1.host.data(), B1.host.size()));
    IS_CUDA(ctx, cudaMemsetAsync(base + off_res_all, 0, sizeof(int) * std::max<size_t>(res_total_ints, 4), ctx->stream));
    const PairDev* pairs_d = reinterpret_cast<const PairDev*>(b1d + off_pairs);
    const JobDev* jobs_d = reinterpret_cast<const JobDev*>(b1d + off_jobs);
    const DpArgs* dp_d = reinterpret_cast<const DpArgs*>(b1d + off_dp);
    int max_ww = 0, max_wh = 0, max_iw = 0, max_ih = 0;
    for (auto& P : PR) {
        max_ww = std::max(max_ww, P.ww); max_wh = std::max(max_wh, P.wh);
        max_iw = std::max(max_iw, P.iBr.x - P.iTl.x); max_ih = std::max(max_ih, P.iBr.y - P.iTl.y);
    }
    {
        dim3 block(64, 4), grid(div_up(max_ww, 64), div_up(max_wh, 4), (unsigned)np);
        IS_LAUNCH(ctx, k_label_window_batch, grid, block, 0, pairs_d);
    }
    if (!pre_relabel.empty()) {
        int mw = 0, mh = 0;
        for (auto& R : pre_relabel) { mw = std::max(mw, R.w); mh = std::max(mh, R.h); }
        dim3 block(64, 4), grid(div_up(mw, 64), div_up(mh, 4), (unsigned)pre_relabel.size());
        IS_LAUNCH(ctx, k_relabel_batch, grid, block, 0, pairs_d, reinterpret_cast<const RelabelDev*>(b1d + off_pre));
    }
    if (nj) {
        if (cost_fn == IS_COST_COLOR_GRAD) {                                                      // computeGradients [SEAM]:549-572 over the intersection rectangles
            const PairDev* pd = reinterpret_cast<const PairDev*>(B1.host.data() + off_pairs);
            for (size_t k = 0; k < np; ++k) {
                const PairDev& D = pd[k];
                float* g = reinterpret_cast<float*>(base + off_grad[k]);
                const size_t plane = (size_t)gpitch[k] * D.ih;
                dim3 block(64, 4), grid(div_up(D.iw, 64), div_up(D.ih, 4));
                if (is_u8) {
                    IS_LAUNCH(ctx, k_sobel_window<uint8_t>, grid, block, 0, ImgView<uint8_t>{reinterpret_cast<const uint8_t*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1},
                              D.ix, D.iy, D.iw, D.ih, g, g + plane, gpitch[k]);
                    IS_LAUNCH(ctx, k_sobel_window<uint8_t>, grid, block, 0, ImgView<uint8_t>{reinterpret_cast<const uint8_t*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2},
                              D.ix, D.iy, D.iw, D.ih, g + 2 * plane, g + 3 * plane, gpitch[k]);
                } else {
                    IS_LAUNCH(ctx, k_sobel_window<float>, grid, block, 0, ImgView<float>{reinterpret_cast<const float*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1},
                              D.ix, D.iy, D.iw, D.ih, g, g + plane, gpitch[k]);
                    IS_LAUNCH(ctx, k_sobel_window<float>, grid, block, 0, ImgView<float>{reinterpret_cast<const float*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2},
                              D.ix, D.iy, D.iw, D.ih, g + 2 * plane, g + 3 * plane, gpitch[k]);
                }
            }
        }
        int max_pitch = 0, max_steps = 0, max_rw = 0, max_rh = 0, max_items = 0, max_nseam = 0;
        double cost_bytes = 0;
        for (auto& J : jobs) {
            max_pitch = std::max(max_pitch, J.pitch); max_steps = std::max(max_steps, J.steps);
            max_rw = std::max(max_rw, J.op.rw); max_rh = std::max(max_rh, J.op.rh);
            max_items = std::max(max_items, J.nc + J.nseam); max_nseam = std::max(max_nseam, J.nseam);
            cost_bytes += (double)J.lanes * J.steps * ((is_u8 ? 6 : 24) + 12 + (cost_fn == IS_COST_COLOR_GRAD ? 16 : 0));
        }
        {
            dim3 block(64, 4), grid(div_up(max_pitch, 64), div_up(max_steps, 4), (unsigned)nj);
            ctx->next_bytes = cost_bytes;
            if (cost_fn == IS_COST_COLOR_GRAD) {
                if (is_u8) IS_LAUNCH(ctx, (k_cost_pq_batch<uint8_t, true>), grid, block, 0, pairs_d, jobs_d);
                else IS_LAUNCH(ctx, (k_cost_pq_batch<float, true>), grid, block, 0, pairs_d, jobs_d);
            } else {
                if (is_u8) IS_LAUNCH(ctx, (k_cost_pq_batch<uint8_t, false>), grid, block, 0, pairs_d, jobs_d);
                else IS_LAUNCH(ctx, (k_cost_pq_batch<float, false>), grid, block, 0, pairs_d, jobs_d);
            }
        }
        {                                                                                        // one DP launch per shape, one CTA per seam
            std::vector<DpShape> shapes(nj);
            for (size_t q = 0; q < nj; ++q) shapes[q] = jobs[order[q]].shape;
            IS_TRY(launch_dp_all(ctx, shapes, dp_d, reinterpret_cast<const BtArgs*>(b1d + off_bt)));
        }
        {
            dim3 block(64, 4), grid(div_up(max_rw, 64), div_up(max_rh, 4), (unsigned)nj);
            IS_LAUNCH(ctx, k_uls_class_batch, grid, block, 0, pairs_d, jobs_d);
            IS_LAUNCH(ctx, k_uls_paint_seam_batch, dim3(div_up(max_nseam, 256), (unsigned)nj), 256, 0, jobs_d);
            IS_LAUNCH(ctx, k_ccl_rows_batch, dim3(max_rh, (unsigned)nj), 256, 0, jobs_d);
            if (max_rh > 1) {
                dim3 mgrid(div_up(max_rw, 64), div_up(max_rh - 1, 4), (unsigned)nj);
                IS_LAUNCH(ctx, k_ccl_merge_batch, mgrid, block, 0, jobs_d);
            }
            const size_t nmax = (size_t)max_rw * max_rh;
            IS_LAUNCH(ctx, k_ccl_flatten_batch, dim3((unsigned)((nmax + 255) / 256), (unsigned)nj), 256, 0, jobs_d);
            IS_LAUNCH(ctx, k_uls_gather_batch, dim3(div_up(max_items, 128), (unsigned)nj), 128, 0, jobs_d);
        }
    }
    const int* res_h = nullptr;                                        // view of the pinned bounce buffer, valid until the next download
    if (nj) IS_TRY(download_view(ctx, base + off_res_all, sizeof(int) * res_total_ints, reinterpret_cast<const void**>(&res_h)));
    tm.lap("C labels, costs, DP, uls device part");
    // ---- host walks
    pool->run(nj, [&](size_t j) {
        SeamJobHost& J = jobs[j];
        J.status = uls_host_walk(ctx, PR[(size_t)J.pair], J, res_h + J.off_res, trace);
    });
    size_t limit = np;                                                 // pairs [0, limit) can still be accepted
    for (auto& J : jobs) {
        if (J.status == IS_ERR_UNSUPPORTED) { limit = std::min(limit, (size_t)J.pair); continue; }   // that pair takes the general path
        if (J.status != IS_OK) return J.status;
    }
    if (limit == 0) { *first_unsupported = true; return IS_OK; }
    tm.lap("D host: uls walk + vote");
    // ---- E: relabel by the seams, clears per pair (private)
    {
        Blob B2;
        B2.host.resize(b2_fixed);
        for (size_t k = 0; k < np; ++k)
            if (PR[k].ncomps) std::memcpy(B2.host.data() + off_states[k], PR[k].final_states.data(), sizeof(int) * (size_t)PR[k].ncomps);
        std::vector<JobDevE> ext(std::max<size_t>(nj, 1));
        int max_flips = 0;
        bool any_adj = false;
        for (size_t q = 0; q < nj; ++q) {
            SeamJobHost& J = jobs[order[q]];
            const size_t oa = B2.put(J.adj_roots.data(), sizeof(int) * J.adj_roots.size());
            const size_t of = B2.put(J.flips.data(), sizeof(int2) * J.flips.size());
            ext[q] = JobDevE{reinterpret_cast<const int*>(base + off_b2 + oa), (int)J.adj_roots.size(), reinterpret_cast<const int2*>(base + off_b2 + of), (int)J.flips.size()};
            max_flips = std::max(max_flips, (int)J.flips.size());
            any_adj = any_adj || !J.adj_roots.empty();
        }
        const size_t off_ext = B2.put(ext.data(), sizeof(JobDevE) * ext.size());
        const size_t off_post = B2.put(post_relabel.data(), sizeof(RelabelDev) * post_relabel.size());
        IS_REQUIRE(ctx, B2.host.size() <= b2_cap, IS_ERR_INTERNAL, "seam batch: second upload larger than planned");
        IS_TRY(upload(ctx, base + off_b2, B2.host.data(), B2.host.size()));
        const JobDevE* ext_d = reinterpret_cast<const JobDevE*>(base + off_b2 + off_ext);
        if (nj && any_adj) {
            int max_rw = 0, max_rh = 0;
            for (auto& J : jobs) { max_rw = std::max(max_rw, J.op.rw); max_rh = std::max(max_rh, J.op.rh); }
            dim3 block(64, 4), grid(div_up(max_rw, 64), div_up(max_rh, 4), (unsigned)nj);
            IS_LAUNCH(ctx, k_uls_apply_batch, grid, block, 0, pairs_d, jobs_d, ext_d);
        }
        if (nj && max_flips) IS_LAUNCH(ctx, k_scatter_label_batch, dim3(div_up(max_flips, 256), (unsigned)nj), 256, 0, pairs_d, jobs_d, ext_d);
        if (!post_relabel.empty()) {
            int mw = 0, mh = 0;
            for (auto& R : post_relabel) { mw = std::max(mw, R.w); mh = std::max(mh, R.h); }
            dim3 block(64, 4), grid(div_up(mw, 64), div_up(mh, 4), (unsigned)post_relabel.size());
            IS_LAUNCH(ctx, k_relabel_batch, grid, block, 0, pairs_d, reinterpret_cast<const RelabelDev*>(base + off_b2 + off_post));
        }
        dim3 block(64, 4), grid(div_up(max_iw, 64), div_up(max_ih, 4), (unsigned)np);
        IS_LAUNCH(ctx, k_pair_clears_batch, grid, block, 0, pairs_d);
    }
    // ---- F: validation -- the masks every pair would have seen in the sequential loop
    {
        StructureQuery V;
        std::vector<int> vpair;                                        // active index of every validation pair
        for (size_t k = 0; k < limit; ++k) {
            const int img[2] = {active[k].first, active[k].second};
            LayeredMask lm[2] = {plain_mask(masks[img[0]]), plain_mask(masks[img[1]])};
            bool any = false, too_many = false;
            for (int s = 0; s < 2 && !too_many; ++s)
                for (size_t q = 0; q < k; ++q) {
                    int bit = 0;
                    if (active[q].first == img[s]) bit = 1; else if (active[q].second == img[s]) bit = 2;
                    if (!bit) continue;
                    if (lm[s].nlayers == MAX_LAYERS) { too_many = true; break; }       // more earlier neighbours than a mask carries
                    const PairRuns& E = PR[q];
                    ClearLayer& L = lm[s].layer[lm[s].nlayers++];
                    L.p = base + off_clear[q]; L.pitch = cpitch[q];
                    L.x0 = E.iTl.x - corners[img[s]].x; L.y0 = E.iTl.y - corners[img[s]].y;
                    L.w = E.iBr.x - E.iTl.x; L.h = E.iBr.y - E.iTl.y; L.bit = bit;
                    any = true;
                }
            if (too_many) { limit = k; break; }                        // ends the wave here: the pair starts the next one with fewer layers
            if (!any) continue;
            V.pairs.push_back(StructureQuery::PairQ{(int)V.masks.size(), (int)V.masks.size() + 1, PR[k].tl1, PR[k].tl2});
            V.masks.push_back(lm[0]);
            V.masks.push_back(lm[1]);
            vpair.push_back((int)k);
        }
        if (!V.pairs.empty()) {
            IS_TRY(run_structure_query(ctx, V));
            tm.lap("F toggles + special points (check)");
            std::vector<char> ok(vpair.size(), 0);
            pool->run(vpair.size(), [&](size_t v) {
                if (V.pair_overflow[v]) return;
                PairRuns C;
                const size_t k = (size_t)vpair[v];
                C.setup(active[k].first, active[k].second, PR[k].tl1, PR[k].tl2, &V.runs[(size_t)V.pairs[v].m1], &V.runs[(size_t)V.pairs[v].m2]);
                C.specials = V.specials[v];
                C.build();
                if (C.too_many_runs) return;
                C.plan();
                ok[v] = C.same_structure(PR[k]) ? 1 : 0;
            });
            for (size_t v = 0; v < vpair.size(); ++v)
                if (!ok[v]) { limit = std::min(limit, (size_t)vpair[v]); break; }   // the first pair the earlier clears change starts the next wave
            tm.lap("F host: check structures");
        }
    }
    // ---- G: the clears of the accepted prefix go into the masks
    IS_REQUIRE(ctx, limit >= 1, IS_ERR_INTERNAL, "seam wave without progress");
    {
        dim3 block(64, 4), grid(div_up(max_iw, 64), div_up(max_ih, 4), (unsigned)limit);
        IS_LAUNCH(ctx, k_apply_clears_batch, grid, block, 0, pairs_d);
    }
    if (trace)
        for (size_t k = 0; k < limit; ++k)
            for (auto& J : jobs) {
                if ((size_t)J.pair != k || J.trace.empty()) continue;
                const size_t len = J.trace.size();
                if (trace->buf && trace->len + len <= trace->cap) std::memcpy(trace->buf + trace->len, J.trace.data(), len * sizeof(int32_t));
                trace->len += len;
            }
    if (tm.on) cudaStreamSynchronize(ctx->stream);
    tm.lap("G apply");
    (void)n;
    *accepted_n = limit;
    return IS_OK;
}
