// seam_batch.inl -- the batched seam path: ALL image pairs of a call go through each kernel in one launch, and everything the
// reference derives from masks and labels is done in the RUN domain on the host (seam_runs.inl); the device keeps what is
// per-pixel arithmetic: cost maps and the DP.  (included by seam.cu inside namespace is)
//
//   A  device  k_row_toggles_batch (every mask -> per-row toggle positions), k_special_points_batch (per pair: the handful of
//              pixels that can become seam tips)                                            -> one download (~1 MB)
//      host    PairRuns::build / plan per pair (thread pool): components, contours, edges, conflict loop -> operations
//   C  device  k_label_window_batch (labels of the overlap for the cost kernel), k_cost_pq_batch, DP forward + back-track (one
//              CTA per seam)                                                                  -> one download (the seam lanes)
//      host    UlsRuns per seam (thread pool): updateLabelsUsingSeam on runs -> the pixels that change sides;
//              pair_clear_intervals per pair: the final mask update as clear intervals per row -> one upload
//   F  device  toggles + special points of the masks every pair WOULD have seen in the reference's sequential loop (entry
//              masks minus the clear intervals of the earlier pairs)                          -> one download
//      host    PairRuns of those masks; identical structure and plan <=> the speculative result is the sequential loop's
//   G  device  k_apply_clear_runs_batch: the clear intervals go into the real masks
// Every pair of a wave starts from the same masks; the longest prefix (in the reference's order) proven independent of the
// earlier pairs' clears is accepted, the rest forms the next wave (a strip needs one wave, a mosaic whose images overlap
// mutually a few).  A pair the plan cannot cover (noisy masks, a component cut twice) takes the general path (PairSeam) alone.

constexpr int TG_CAP = 16;             // toggles per mask row the batched path handles (rows are a few runs, up to a dozen in a mosaic after several
                                       // pairs have cut their seams into a mask; more -> general path)
constexpr int MAX_LAYERS = 8;          // earlier pairs whose clears a validation mask can carry
constexpr int SPECIAL_CAP = 512;       // candidate seam tips per pair (a panorama pair has a few dozen; more -> general path)

// clear intervals of one pair over the rows of its intersection rectangle: ivs[row_start[r] .. row_start[r + 1]) = (x0, x1, bits, -)
// in the pair's FRAME coordinates, row r = frame y - iy; bits: 1 = the first image's mask loses the pixels, 2 = the second's
struct ClearLayer {
    const int* row_start; const int4* ivs;
    int dx, dy;                        // mask coordinates = frame coordinates + (dx, dy)
    int iy, ih;
    int bit;
};

struct LayeredMask {                   // a mask minus the clears of some pairs
    const uint8_t* p; size_t step; int rows, cols;
    int nlayers;
    ClearLayer layer[MAX_LAYERS];
};

__device__ __forceinline__ bool lm_cleared(const LayeredMask& m, int x, int y) {
    for (int k = 0; k < m.nlayers; ++k) {
        const ClearLayer& L = m.layer[k];
        const int r = y - L.dy - L.iy, fx = x - L.dx;
        if ((unsigned)r >= (unsigned)L.ih) continue;
        for (int q = L.row_start[r]; q < L.row_start[r + 1]; ++q) {
            const int4 iv = L.ivs[q];
            if (fx >= iv.x && fx < iv.y && (iv.z & L.bit)) return true;
        }
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
// x = 0 is "outside".  16 pixels per lane and trip.  Rows with more than TG_CAP toggles set overflow[job].
__global__ void __launch_bounds__(256) k_row_toggles_batch(const ToggleJob* __restrict__ jobs, int* __restrict__ overflow) {
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
        if (J.m.nlayers && bits) {                       // validation masks: drop the pixels an earlier pair has cleared (a few intervals per row)
            for (int k = 0; k < J.m.nlayers; ++k) {
                const ClearLayer& L = J.m.layer[k];
                const int r = y - L.dy - L.iy;
                if ((unsigned)r >= (unsigned)L.ih) continue;
                for (int q = L.row_start[r]; q < L.row_start[r + 1]; ++q) {
                    const int4 iv = L.ivs[q];
                    if (!(iv.z & L.bit)) continue;
                    const int a = max(iv.x + L.dx - x0, 0), b = min(iv.y + L.dx - x0, 16);   // cleared bits [a, b) of this chunk
                    if (a < b) bits &= ~(((1u << (b - a)) - 1u) << a);
                }
            }
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

// bit i of the result: pixel mx0 + i (own coordinates) of a mask row is set, i < 18 -- from the row's toggles.  Every toggle XORs a
// suffix of the window, so the toggles can be taken eight at a time (n of them in v) and the partial results XORed.
__device__ __forceinline__ unsigned toggles_xor18(int n, const uint4& v, int mx0) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    unsigned bits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k >= n) break;
        const int t = (int)((w[k >> 1] >> (16 * (k & 1))) & 0xffffu) - mx0;      // toggle position inside the window
        bits ^= t <= 0 ? 0x3ffffu : (t >= 18 ? 0u : (0x3ffffu << t) & 0x3ffffu);
    }
    return bits;
}
__device__ __forceinline__ unsigned clamp_bits18(unsigned bits, int cols, int mx0) {
    if (mx0 + 18 > cols) bits &= mx0 < cols ? (1u << (cols - mx0)) - 1u : 0u;      // a row that ends inside the mask has no closing toggle
    return bits;
}

// The only pixels getSeamTips can ever pick ([SEAM]:621-629): pixels of both masks with a 4-neighbour inside exactly one
// mask (a contour pixel of an INTERS component touching a FIRST / SECOND component), close to both masks' contours.
// The scan works on the row toggles k_row_toggles_batch has just produced (16 pixels of a row per step as bit sets: two
// 16-byte loads per row instead of ten byte loads per pixel); only the few candidates look at the mask bytes themselves.
constexpr int SP_ROWS = 4;             // rows per thread; a block of 128 threads covers 32 chunks (512 pixels) x 16 rows
constexpr int SP_QUEUE = 1024;         // candidates a block can hold (its 8192 pixels could all be candidates: beyond the queue the pair overflows)
struct SpMask { const uint8_t* p; size_t step; int rows, cols, ox, oy, nlayers; };
__device__ __forceinline__ bool spm_at(const SpMask& M, const LayeredMask& full, int fx, int fy) {   // frame coordinates
    const int x = fx - M.ox, y = fy - M.oy;
    if ((unsigned)x >= (unsigned)M.cols || (unsigned)y >= (unsigned)M.rows) return false;
    if (!M.p[(size_t)y * M.step + x]) return false;
    return M.nlayers == 0 || !lm_cleared(full, x, y);
}
__device__ __forceinline__ bool spm_contour(const SpMask& M, const LayeredMask& full, int x, int y) {   // contour{1,2}mask_ [SEAM]:165-186
    return spm_at(M, full, x, y) && !(spm_at(M, full, x - 1, y) && spm_at(M, full, x + 1, y) && spm_at(M, full, x, y - 1) && spm_at(M, full, x, y + 1));
}

__global__ void __launch_bounds__(128) k_special_points_batch(const SpecialJob* __restrict__ jobs) {
    __shared__ int2 queue[SP_QUEUE];
    __shared__ int qn;
    __shared__ SpMask M1, M2;
    __shared__ int geo[6];                                                       // uw, uh, ix, iy, iw, ih
    const SpecialJob& J = jobs[blockIdx.z];
    if (threadIdx.x == 0) {
        qn = 0;
        M1 = SpMask{J.m1.p, J.m1.step, J.m1.rows, J.m1.cols, J.o1x, J.o1y, J.m1.nlayers};
        M2 = SpMask{J.m2.p, J.m2.step, J.m2.rows, J.m2.cols, J.o2x, J.o2y, J.m2.nlayers};
        geo[0] = J.uw; geo[1] = J.uh; geo[2] = J.ix; geo[3] = J.iy; geo[4] = J.iw; geo[5] = J.ih;
    }
    __syncthreads();
    const int ix = geo[2], iy = geo[3], iw = geo[4], ih = geo[5];
    const int cx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ly0 = (blockIdx.y * 4 + (threadIdx.x >> 5)) * SP_ROWS;
    if ((int)(blockIdx.y * 4 * SP_ROWS) >= ih || (int)(blockIdx.x * 512) >= iw) return;   // uniform over the block
    if (16 * cx < iw && ly0 < ih) {
        const int x0 = ix + 16 * cx;                                             // frame column of bit 1
        const int m1x = x0 - 1 - M1.ox, m2x = x0 - 1 - M2.ox;
        // the toggles of the SP_ROWS + 2 rows this thread looks at, both masks: all loads first (they are independent), then the bits
        int cn1[SP_ROWS + 2], cn2[SP_ROWS + 2];
        uint4 tv1[SP_ROWS + 2], tv2[SP_ROWS + 2];
#pragma unroll
        for (int r = 0; r < SP_ROWS + 2; ++r) {
            const int y1 = iy + ly0 - 1 + r - M1.oy, y2 = iy + ly0 - 1 + r - M2.oy;
            const bool in1 = (unsigned)y1 < (unsigned)M1.rows, in2 = (unsigned)y2 < (unsigned)M2.rows;
            cn1[r] = in1 ? min((int)J.cnt1[y1], TG_CAP) : 0;
            cn2[r] = in2 ? min((int)J.cnt2[y2], TG_CAP) : 0;
            tv1[r] = in1 ? *reinterpret_cast<const uint4*>(J.xs1 + (size_t)y1 * TG_CAP) : make_uint4(0, 0, 0, 0);
            tv2[r] = in2 ? *reinterpret_cast<const uint4*>(J.xs2 + (size_t)y2 * TG_CAP) : make_uint4(0, 0, 0, 0);
        }
        unsigned b1[SP_ROWS + 2], b2[SP_ROWS + 2];
#pragma unroll
        for (int r = 0; r < SP_ROWS + 2; ++r) {
            b1[r] = toggles_xor18(cn1[r], tv1[r], m1x);
            b2[r] = toggles_xor18(cn2[r], tv2[r], m2x);
            if (cn1[r] > 8) {                                                      // rare: the second eight toggles of the row
                const int y1 = iy + ly0 - 1 + r - M1.oy;
                b1[r] ^= toggles_xor18(cn1[r] - 8, *reinterpret_cast<const uint4*>(J.xs1 + (size_t)y1 * TG_CAP + 8), m1x);
            }
            if (cn2[r] > 8) {
                const int y2 = iy + ly0 - 1 + r - M2.oy;
                b2[r] ^= toggles_xor18(cn2[r] - 8, *reinterpret_cast<const uint4*>(J.xs2 + (size_t)y2 * TG_CAP + 8), m2x);
            }
            b1[r] = clamp_bits18(b1[r], M1.cols, m1x);
            b2[r] = clamp_bits18(b2[r], M2.cols, m2x);
        }
        const int rows = min(SP_ROWS, ih - ly0);
#pragma unroll
        for (int r = 0; r < SP_ROWS; ++r) {
            if (r >= rows) break;
            const int y = iy + ly0 + r;
            const unsigned both = b1[r + 1] & b2[r + 1], x_cur = b1[r + 1] ^ b2[r + 1], x_up = b1[r] ^ b2[r], x_dn = b1[r + 2] ^ b2[r + 2];
            unsigned cand = both & ((x_cur << 1) | (x_cur >> 1) | x_up | x_dn) & 0x1fffeu;   // bits 1..16: this chunk's pixels
            if (cand) {
                int k = atomicAdd(&qn, __popc(cand));
                while (cand) {
                    const int i = __ffs(cand) - 1;
                    cand &= cand - 1;
                    if (k < SP_QUEUE) queue[k] = make_int2(x0 - 1 + i, y);
                    ++k;
                }
            }
        }
    }
    __syncthreads();
    const int n = qn;
    if (n > SP_QUEUE) { if (threadIdx.x == 0) atomicAdd(J.count, SPECIAL_CAP + 1); return; }   // overflow: the pair takes the general path
    // closeToContour of both masks for every candidate ([SEAM]:584-604): one warp per candidate, one window position per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int uw = geo[0], uh = geo[1];
    for (int q = warp; q < n; q += nw) {
        const int2 c = queue[q];
        const int x = c.x + (lane % 5) - 2, y = c.y + (lane / 5) - 2;
        const bool in = lane < 25 && x >= 0 && x < uw && y >= 0 && y < uh;
        const unsigned b1 = __ballot_sync(0xffffffffu, in && spm_contour(M1, J.m1, x, y));
        if (!b1) continue;                                                       // uniform over the warp
        const unsigned b2 = __ballot_sync(0xffffffffu, in && spm_contour(M2, J.m2, x, y));
        if (lane == 0 && b2 && c.x < ix + iw) {
            const int pos = atomicAdd(J.count, 1);
            if (pos < SPECIAL_CAP) J.out[pos] = c;
        }
    }
}

// ---- per pair / per seam tables ----------------------------------------------------------------------------------------
struct PairDev {
    Frame fr;                          // union frame, label window, the ENTRY masks placed in the frame
    int* labels;
    const int* tab_cnt; const ChangePt* tab_cps; const int* tab_lab; int wcap;   // biased by the window's first row (k_label_window)
    const void* img1; const void* img2; size_t step1, step2; int rows1, cols1, rows2, cols2; int dx1, dy1, dx2, dy2;
    int cn;                            // elements per pixel of both images (3 or 4)
    GradView g;
};

struct JobDev {
    int pair;
    int l1;
    int rx, ry, rw, rh;
    int horizontal, lanes, steps, pitch;
    float* P; float* Q;
    const int2* runs;                  // the component's runs per bounding-box row: runs[(y - ry) * COST_RUNS + k] = (x0, x1) in frame coordinates
};
constexpr int COST_RUNS = 4;

struct ClearTabDev {                   // the clear intervals of a pair and the two masks they go into
    const int* row_start; const int4* ivs;
    int iy, ih;
    uint8_t* mask1; size_t mstep1; int o1x, o1y;
    uint8_t* mask2; size_t mstep2; int o2x, o2y;
};

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
    const ImgView<T> a{reinterpret_cast<const T*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1, D.cn};
    const ImgView<T> b{reinterpret_cast<const T*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2, D.cn};
    const int x = J.rx + (J.horizontal ? step : lane), y = J.ry + (J.horizontal ? lane : step);
    float p, q;
    if (J.horizontal) { p = cost_h<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); q = cost_v<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); }
    else { p = cost_v<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); q = cost_h<T, GRAD>(a, b, D.labels, D.fr, J.l1, x, y, D.g); }
    if (lab(D.labels, D.fr, x, y) != J.l1) p = __int_as_float(0x7f800000);   // +inf: the cell can never be on a path
    P[(size_t)step * J.pitch + lane] = p;
    Q[(size_t)step * J.pitch + lane] = q;
}

// COLOR costs of all seams, second formulation: a thread owns one lane and walks COST_WALK steps.  Both costs of a cell pair
// the cell with one neighbour -- the neighbouring lane of the same step (P) and the same lane of the previous step (Q):
//   vertical seam    lane = x, step = y:   P = costV(x, y) pairs (x - 1, y),  Q = costH(x, y) pairs (x, y - 1)
//   horizontal seam  lane = y, step = x:   P = costH(x, y) pairs (x, y - 1),  Q = costV(x, y) pairs (x - 1, y)
// so the cell's own pixels and label are loaded once, the previous step's stay in registers and the neighbouring lane's come
// by warp shuffle: 6 pixel bytes + one label per cell instead of ~24 + 5.  Same arithmetic as cost_v / cost_h ([SEAM]:756-802).
constexpr int COST_WALK = 32;
// Membership of a cell in the component comes from the component's runs of that row (a few intervals per row, uploaded with the
// plan), not from a label image: no k_label_window pass, no 4-byte label per cell.
template <typename T>
__global__ void __launch_bounds__(128) k_cost_pq_walk(const PairDev* __restrict__ pairs, const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.z];
    const int s_begin = blockIdx.y * COST_WALK;
    if (s_begin >= J.steps || (int)(blockIdx.x * blockDim.x) >= J.pitch) return;      // uniform over the block
    const PairDev& D = pairs[J.pair];
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane_id = threadIdx.x & 31;
    const bool hz = J.horizontal != 0;
    const int l1 = J.l1;
    const ImgView<T> A{reinterpret_cast<const T*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1, D.cn};
    const ImgView<T> B{reinterpret_cast<const T*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2, D.cn};
    struct Cell { float a[3], b[3]; int lab; };
    auto load = [&](int ln, int st) {                                               // pixels only where the label says both images cover the cell
        Cell c;
        const int x = J.rx + (hz ? st : ln), y = J.ry + (hz ? ln : st);
        c.lab = -3;
        const int by = y - J.ry;
        if (ln < J.lanes && (unsigned)by < (unsigned)J.rh) {
            const int4* rr = reinterpret_cast<const int4*>(J.runs + (size_t)by * COST_RUNS);
            const int4 r01 = __ldg(rr), r23 = __ldg(rr + 1);
            if ((x >= r01.x && x < r01.y) || (x >= r01.z && x < r01.w) || (x >= r23.x && x < r23.y) || (x >= r23.z && x < r23.w)) c.lab = l1;
        }
        if (c.lab == l1) {
            const T* pa = A.px(x, y);
            const T* pb = B.px(x, y);
            c.a[0] = (float)pa[0]; c.a[1] = (float)pa[1]; c.a[2] = (float)pa[2];
            c.b[0] = (float)pb[0]; c.b[1] = (float)pb[1]; c.b[2] = (float)pb[2];
        } else {
            c.a[0] = c.a[1] = c.a[2] = c.b[0] = c.b[1] = c.b[2] = 0.f;
        }
        return c;
    };
    auto pair_cost = [&](const Cell& n, const Cell& c) {
        if (n.lab != l1 || c.lab != l1) return IS_BAD_REGION_COST;
        return __fmul_rn(__fadd_rn(diff3(n.a, c.b), diff3(c.a, n.b)), 0.5f);
    };
    const float INF = __int_as_float(0x7f800000);
    Cell prev = load(lane, s_begin - 1);                                             // step -1 is the frame row / column in front of the bounding box
    const int s_end = min(s_begin + COST_WALK, J.steps);
    for (int st = s_begin; st < s_end; ++st) {
        const Cell cur = load(lane, st);
        Cell nb;
        nb.lab = __shfl_up_sync(0xffffffffu, cur.lab, 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) { nb.a[k] = __shfl_up_sync(0xffffffffu, cur.a[k], 1); nb.b[k] = __shfl_up_sync(0xffffffffu, cur.b[k], 1); }
        if (lane_id == 0) nb = load(lane - 1, st);                                   // the neighbouring lane belongs to another warp (or lies in front of the box)
        if (lane < J.pitch) {
            float p, q;
            if (lane >= J.lanes) { p = INF; q = 0.f; }                               // padding lanes: outside the component
            else {
                p = pair_cost(nb, cur);
                q = pair_cost(prev, cur);
                if (cur.lab != l1) p = INF;                                          // +inf: the cell can never be on a path
            }
            J.P[(size_t)st * J.pitch + lane] = p;
            J.Q[(size_t)st * J.pitch + lane] = q;
        }
        prev = cur;
    }
}

// the final mask update: one warp per row of a pair's intersection rectangle writes zeros over its clear intervals
__global__ void __launch_bounds__(256) k_apply_clear_runs_batch(const ClearTabDev* __restrict__ tabs) {
    const ClearTabDev& T = tabs[blockIdx.y];
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= T.ih) return;
    const int lane = threadIdx.x & 31;
    const int y = T.iy + r;
    for (int q = T.row_start[r]; q < T.row_start[r + 1]; ++q) {
        const int4 iv = T.ivs[q];
#pragma unroll
        for (int bit = 1; bit <= 2; ++bit) {
            if (!(iv.z & bit)) continue;
            uint8_t* m = bit == 1 ? T.mask1 + (size_t)(y - T.o1y) * T.mstep1 - T.o1x : T.mask2 + (size_t)(y - T.o2y) * T.mstep2 - T.o2x;   // indexed by frame x
            // byte stores up to a 4-byte boundary, then words, then the tail
            const int body = min(iv.y, (int)(iv.x + ((4 - (int)((uintptr_t)(m + iv.x) & 3)) & 3)));
            if (iv.x + lane < body) m[iv.x + lane] = 0;
            const int nwords = (iv.y - body) >> 2;
            uint32_t* w = reinterpret_cast<uint32_t*>(m + body);
            for (int k = lane; k < nwords; k += 32) w[k] = 0u;
            for (int x = body + 4 * nwords + lane; x < iv.y; x += 32) m[x] = 0;
        }
    }
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
    int pair;                          // index into the wave's pair list
    SeamOp op;
    bool horizontal = false, swapped = false;
    int lanes = 0, steps = 0, pitch = 128;
    int s0 = 0, lane0 = 0, s1 = 0, lane1 = 0, nseam = 0;
    DpShape shape;
    size_t off_map = 0, off_P = 0, off_Q = 0, off_ctl = 0, off_res = 0;   // device arena offsets (off_res in ints)
    std::vector<int2> runs;            // component_runs of c1 over its bounding box
    UlsRuns uls;
    bool reached = false;
    std::vector<int32_t> trace;
    int status = IS_OK;
};

// one seam back on the host: end points, trace record, updateLabelsUsingSeam on runs.  res: [reached, -, lanes ...]
static int seam_job_finish(is_ctx* ctx, const PairRuns& PR, SeamJobHost& J, const int* res, bool want_trace) {
    const SeamOp& op = J.op;
    J.trace.clear();
    J.reached = res[0] != 0;
    if (!J.reached) return IS_OK;                                      // [SEAM]:918-919: estimateSeam returned false
    const int* lane_h = res + 2;
    const int nseam = J.nseam;
    auto pt = [&](int i) { const int step = J.s0 + i, lane = lane_h[i]; return J.horizontal ? Pt{step + op.rx, lane + op.ry} : Pt{lane + op.rx, step + op.ry}; };
    const Pt a = pt(J.swapped ? nseam - 1 : 0), b = pt(J.swapped ? 0 : nseam - 1);   // ordered p1 -> p2 ([SEAM]:949-954)
    if (!(a.x == op.p1.x && a.y == op.p1.y && b.x == op.p2.x && b.y == op.p2.y))
        return fail(ctx, IS_ERR_ASSERT, "seam end points differ from the seam tips ([SEAM]:953-954)");
    if (want_trace) {
        J.trace.resize(5 + 2 * (size_t)nseam);
        int32_t* t = J.trace.data();
        t[0] = PR.pi; t[1] = PR.pj; t[2] = op.c1; t[3] = J.horizontal ? 1 : 0; t[4] = nseam;
        for (int i = 0; i < nseam; ++i) {
            const Pt q = pt(J.swapped ? nseam - 1 - i : i);
            t[5 + 2 * i] = q.x + PR.unionTl.x; t[6 + 2 * i] = q.y + PR.unionTl.y;
        }
    }
    J.uls.P = &PR; J.uls.op = op; J.uls.horizontal = J.horizontal; J.uls.s0 = J.s0; J.uls.nseam = nseam; J.uls.lane = lane_h;
    J.uls.run();
    return J.uls.too_many_regions ? IS_ERR_UNSUPPORTED : IS_OK;
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
    for (auto& m : Q.masks) { total_rows += (size_t)m.rows; max_rows = std::max(max_rows, m.rows); }
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
        dim3 grid(div_up(max_iw, 512), div_up(max_ih, 4 * SP_ROWS), (unsigned)np);
        IS_LAUNCH(ctx, k_special_points_batch, grid, 128, 0, sj_d);
    }
    if (ctx->deferred_side_work) {                                      // see common.cuh: the side stream's work goes behind these kernels
        std::function<int()> f = std::move(ctx->deferred_side_work);
        ctx->deferred_side_work = nullptr;
        IS_TRY(f());
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
            if (tm.on)
                fprintf(stderr, "[seam batch] pair (%d, %d) leaves the batched path: %s%s%s (toggle overflow %d / %d, special points %d)\n", active[k].first, active[k].second,
                        Q.pair_overflow[k] ? "toggle / special-point tables " : "", PR[k].too_many_runs ? "too many runs in a row " : "",
                        PR[k].unsupported && !Q.pair_overflow[k] ? "plan (INTERS neighbour / DP width)" : "", (int)Q.mask_overflow[(size_t)Q.pairs[k].m1],
                        (int)Q.mask_overflow[(size_t)Q.pairs[k].m2], (int)Q.specials[k].size());
            if (k == 0) { *first_unsupported = true; return IS_OK; }
            np = k;
            active.resize(np);
            PR.resize(np);
            break;
        }
    // ---- C + D, in rounds: normally one; a pair whose plan stopped in front of a second seam on a component already cut (staged
    //      plan, seam_runs.inl) goes on in the next round with the labels its first seams have left
    const int dp_variant = dp_variant_default();
    std::deque<SeamJobHost> all_jobs;                                  // the finished seams of all rounds: per pair in the order they were estimated
    std::vector<std::vector<const std::vector<Interval>*>> flips_of(np), round_flips(np);
    std::vector<char> in_round(np, 1);
    size_t limit = np;                                                 // pairs [0, limit) can still be accepted
    for (int round = 0;; ++round) {
    std::vector<SeamJobHost> jobs;
    for (size_t k = 0; k < limit; ++k) {
        if (!in_round[k]) continue;
        round_flips[k].clear();
        for (size_t q = PR[k].round_begin; q < PR[k].ops.size(); ++q) {
            const SeamOp& op = PR[k].ops[q];
            if (op.kind != 1) continue;                                // wholesale relabels never reach the device: the labels only feed the cost kernel
            if ((std::abs(op.p2.x - op.p1.x) > std::abs(op.p2.y - op.p1.y) ? op.rh : op.rw) > 12 * 1024 - 128) { limit = std::min(limit, k); break; }   // wider than the DP tables: general path
            jobs.emplace_back();
            SeamJobHost& J = jobs.back();
            J.pair = (int)k; J.op = op;
            Pt src{op.p1.x - op.rx, op.p1.y - op.ry}, dst{op.p2.x - op.rx, op.p2.y - op.ry};
            J.horizontal = std::abs(dst.x - src.x) > std::abs(dst.y - src.y);                     // [SEAM]:828
            if (J.horizontal) { if (src.x > dst.x) { std::swap(src, dst); J.swapped = true; } }
            else if (src.y > dst.y) { std::swap(src, dst); J.swapped = true; }
            J.lanes = J.horizontal ? op.rh : op.rw; J.steps = J.horizontal ? op.rw : op.rh;
            J.s0 = J.horizontal ? src.x : src.y; J.lane0 = J.horizontal ? src.y : src.x;
            J.s1 = J.horizontal ? dst.x : dst.y; J.lane1 = J.horizontal ? dst.y : dst.x;
            dp_choose_shape(J.lanes, J.steps, J.s0, J.s1, dp_variant, &J.shape);
            J.pitch = J.shape.pitch;
            IS_REQUIRE(ctx, J.pitch >= J.lanes && J.pitch <= 12 * 1024, IS_ERR_INTERNAL, "seam wider than planned");
            J.nseam = J.s1 - J.s0 + 1;
        }
    }
    if (limit == 0) { *first_unsupported = true; return IS_OK; }
    while (!jobs.empty() && (size_t)jobs.back().pair >= limit) jobs.pop_back();   // pairs are visited in order: the dropped ones are at the end
    const size_t nj = jobs.size();
    // COLOR costs read the component's runs; the label image (k_label_window_batch) is only built for the cost kernels that still
    // want it: COLOR_GRAD, or a component with more than COST_RUNS runs in a row
    std::vector<char> runs_ok(std::max<size_t>(nj, 1), 0);
    pool->run(nj, [&](size_t j) { runs_ok[j] = jobs[j].runs.empty() && PR[(size_t)jobs[j].pair].component_runs(jobs[j].op.c1, jobs[j].op.ry, jobs[j].op.rh, COST_RUNS, &jobs[j].runs); });
    bool need_labels = cost_fn == IS_COST_COLOR_GRAD || getenv("IS_COST_KERNEL_CELL") != nullptr;
    for (size_t j = 0; j < nj; ++j) need_labels = need_labels || !runs_ok[j];
    if (need_labels) pool->run(np, [&](size_t k) { PR[k].build_tab(); });
    // device arena: per pair the label window, per seam P, Q, control (+ back-track maps), results; blob 1 = tables + label tables
    size_t arena = 0;
    auto take = [&](size_t bytes) { const size_t o = arena; arena = align_up(arena + bytes, 256); return o; };
    std::vector<size_t> off_labels(np), off_grad(np, 0);
    std::vector<int> gpitch(np, 0);
    for (size_t k = 0; k < np; ++k) {
        const PairRuns& P = PR[k];
        off_labels[k] = need_labels ? take(sizeof(int) * (size_t)P.ww * P.wh) : 0;
        if (cost_fn == IS_COST_COLOR_GRAD) {
            const int iw = P.iBr.x - P.iTl.x, ih = P.iBr.y - P.iTl.y;
            gpitch[k] = (iw + 31) & ~31;
            off_grad[k] = take(sizeof(float) * (size_t)gpitch[k] * ih * 4);
        }
    }
    size_t res_total_ints = 0;
    for (auto& J : jobs) {
        J.off_P = take(sizeof(float) * (size_t)J.pitch * (J.steps + DP_ROW_PAD) + 64);
        J.off_Q = take(sizeof(float) * (size_t)J.pitch * (J.steps + DP_ROW_PAD) + 64);
        J.off_ctl = take((size_t)J.pitch * J.steps + 64);
        if (J.shape.v1) J.off_map = take(sizeof(short) * (size_t)div_up(J.s1 - J.s0, BT_CHUNK) * J.pitch + 64);
        J.off_res = res_total_ints;
        res_total_ints += (2 + (size_t)J.nseam + 3) & ~(size_t)3;
    }
    const size_t off_res_all = take(sizeof(int) * std::max<size_t>(res_total_ints, 4));
    Blob B1;
    std::vector<size_t> off_tab(np);
    for (size_t k = 0; k < np; ++k) off_tab[k] = need_labels ? B1.put(PR[k].tab.data(), sizeof(int) * PR[k].tab.size()) : 0;
    std::vector<size_t> off_runs(std::max<size_t>(nj, 1), 0);
    if (!need_labels)
        for (size_t j = 0; j < nj; ++j) off_runs[j] = B1.put(jobs[j].runs.data(), sizeof(int2) * jobs[j].runs.size());
    const size_t off_pairs = B1.put(nullptr, sizeof(PairDev) * np);
    const size_t off_jobs = B1.put(nullptr, sizeof(JobDev) * std::max<size_t>(nj, 1));
    const size_t off_dp = B1.put(nullptr, sizeof(DpArgs) * std::max<size_t>(nj, 1));
    const size_t off_bt = B1.put(nullptr, sizeof(BtArgs) * std::max<size_t>(nj, 1));
    const size_t off_b1 = take(B1.host.size());
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
            if (need_labels) {
                D.labels = reinterpret_cast<int*>(base + off_labels[k]);
                const int* d = reinterpret_cast<const int*>(b1d + off_tab[k]);
                const size_t wrows = (size_t)P.wh;
                D.wcap = (int)P.wcap;
                D.tab_cnt = d - P.wy;
                D.tab_cps = reinterpret_cast<const ChangePt*>(d + wrows) - (size_t)P.wy * P.wcap;
                D.tab_lab = d + wrows + wrows * P.wcap * 2 - (size_t)P.wy * P.wcap;
            }
            D.img1 = i1.data; D.img2 = i2.data; D.step1 = i1.step; D.step2 = i2.step;
            D.rows1 = i1.rows; D.cols1 = i1.cols; D.rows2 = i2.rows; D.cols2 = i2.cols; D.cn = i1.channels;
            D.dx1 = P.unionTl.x - P.tl1.x; D.dy1 = P.unionTl.y - P.tl1.y; D.dx2 = P.unionTl.x - P.tl2.x; D.dy2 = P.unionTl.y - P.tl2.y;
            if (cost_fn == IS_COST_COLOR_GRAD) {
                float* g = reinterpret_cast<float*>(base + off_grad[k]);
                const size_t plane = (size_t)gpitch[k] * (size_t)(P.iBr.y - P.iTl.y);
                D.g = GradView{g, g + plane, g + 2 * plane, g + 3 * plane, gpitch[k], P.iTl.x - P.unionTl.x, P.iTl.y - P.unionTl.y};
            }
        }
        JobDev* jd = reinterpret_cast<JobDev*>(B1.host.data() + off_jobs);
        DpArgs* da = reinterpret_cast<DpArgs*>(B1.host.data() + off_dp);
        BtArgs* ba = reinterpret_cast<BtArgs*>(B1.host.data() + off_bt);
        for (size_t q = 0; q < nj; ++q) {
            const SeamJobHost& J = jobs[order[q]];
            JobDev& D = jd[q];
            std::memset(&D, 0, sizeof(D));
            D.pair = J.pair; D.l1 = J.op.c1 + 1;
            D.rx = J.op.rx; D.ry = J.op.ry; D.rw = J.op.rw; D.rh = J.op.rh;
            D.horizontal = J.horizontal ? 1 : 0; D.lanes = J.lanes; D.steps = J.steps; D.pitch = J.pitch;
            D.P = reinterpret_cast<float*>(base + J.off_P); D.Q = reinterpret_cast<float*>(base + J.off_Q);
            D.runs = need_labels ? nullptr : reinterpret_cast<const int2*>(b1d + off_runs[order[q]]);
            int* res = reinterpret_cast<int*>(base + off_res_all) + J.off_res;
            DpArgs& A = da[q];
            A.P = D.P; A.Q = D.Q; A.control = base + J.off_ctl;
            A.lanes = J.lanes; A.pitch = J.pitch; A.steps = J.steps;
            A.s0 = J.s0; A.lane0 = J.lane0; A.s1 = J.s1; A.lane1 = J.lane1;
            A.seam_lane = res + 2; A.reached = res;
            A.G = J.shape.G; A.D = J.shape.D;
            ba[q].A = A;
            ba[q].map = reinterpret_cast<short*>(base + J.off_map);
            ba[q].nchunks = div_up(J.s1 - J.s0, BT_CHUNK);
        }
    }
    IS_TRY(upload(ctx, b1d, B1.host.data(), B1.host.size()));
    const PairDev* pairs_d = reinterpret_cast<const PairDev*>(b1d + off_pairs);
    const JobDev* jobs_d = reinterpret_cast<const JobDev*>(b1d + off_jobs);
    const DpArgs* dp_d = reinterpret_cast<const DpArgs*>(b1d + off_dp);
    const int* res_h = nullptr;                                        // view of the pinned bounce buffer, valid until the next download
    if (nj) {
        IS_CUDA(ctx, cudaMemsetAsync(base + off_res_all, 0, sizeof(int) * std::max<size_t>(res_total_ints, 4), ctx->stream));
        if (need_labels) {                                                                       // labels only where seams are estimated
            int max_ww = 0, max_wh = 0;
            for (auto& J : jobs) { max_ww = std::max(max_ww, PR[(size_t)J.pair].ww); max_wh = std::max(max_wh, PR[(size_t)J.pair].wh); }
            dim3 block(64, 4), grid(div_up(max_ww, 64), div_up(max_wh, 4), (unsigned)np);
            IS_LAUNCH(ctx, k_label_window_batch, grid, block, 0, pairs_d);
        }
        if (cost_fn == IS_COST_COLOR_GRAD) {                                                      // computeGradients [SEAM]:549-572 over the intersection rectangles
            const PairDev* pd = reinterpret_cast<const PairDev*>(B1.host.data() + off_pairs);
            for (size_t k = 0; k < np; ++k) {
                const PairDev& D = pd[k];
                const PairRuns& P = PR[k];
                const int ix = P.iTl.x - P.unionTl.x, iy = P.iTl.y - P.unionTl.y, iw = P.iBr.x - P.iTl.x, ih = P.iBr.y - P.iTl.y;
                float* g = reinterpret_cast<float*>(base + off_grad[k]);
                const size_t plane = (size_t)gpitch[k] * ih;
                dim3 block(64, 4), grid(div_up(iw, 64), div_up(ih, 4));
                if (is_u8) {
                    IS_LAUNCH(ctx, k_sobel_window<uint8_t>, grid, block, 0, ImgView<uint8_t>{reinterpret_cast<const uint8_t*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1, D.cn},
                              ix, iy, iw, ih, g, g + plane, gpitch[k]);
                    IS_LAUNCH(ctx, k_sobel_window<uint8_t>, grid, block, 0, ImgView<uint8_t>{reinterpret_cast<const uint8_t*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2, D.cn},
                              ix, iy, iw, ih, g + 2 * plane, g + 3 * plane, gpitch[k]);
                } else {
                    IS_LAUNCH(ctx, k_sobel_window<float>, grid, block, 0, ImgView<float>{reinterpret_cast<const float*>(D.img1), D.step1, D.rows1, D.cols1, D.dx1, D.dy1, D.cn},
                              ix, iy, iw, ih, g, g + plane, gpitch[k]);
                    IS_LAUNCH(ctx, k_sobel_window<float>, grid, block, 0, ImgView<float>{reinterpret_cast<const float*>(D.img2), D.step2, D.rows2, D.cols2, D.dx2, D.dy2, D.cn},
                              ix, iy, iw, ih, g + 2 * plane, g + 3 * plane, gpitch[k]);
                }
            }
        }
        int max_pitch = 0, max_steps = 0;
        double cost_bytes = 0;
        for (auto& J : jobs) {
            max_pitch = std::max(max_pitch, J.pitch); max_steps = std::max(max_steps, J.steps);
            cost_bytes += (double)J.lanes * J.steps * ((is_u8 ? 6 : 24) + 12 + (cost_fn == IS_COST_COLOR_GRAD ? 16 : 0));
        }
        {
            dim3 block(64, 4), grid(div_up(max_pitch, 64), div_up(max_steps, 4), (unsigned)nj);
            ctx->next_bytes = cost_bytes;
            if (cost_fn == IS_COST_COLOR_GRAD) {
                if (is_u8) IS_LAUNCH(ctx, (k_cost_pq_batch<uint8_t, true>), grid, block, 0, pairs_d, jobs_d);
                else IS_LAUNCH(ctx, (k_cost_pq_batch<float, true>), grid, block, 0, pairs_d, jobs_d);
            } else if (need_labels) {                                                            // one thread per cell, label image
                if (is_u8) IS_LAUNCH(ctx, (k_cost_pq_batch<uint8_t, false>), grid, block, 0, pairs_d, jobs_d);
                else IS_LAUNCH(ctx, (k_cost_pq_batch<float, false>), grid, block, 0, pairs_d, jobs_d);
            } else {
                dim3 wgrid(div_up(max_pitch, 128), div_up(max_steps, COST_WALK), (unsigned)nj);
                if (is_u8) IS_LAUNCH(ctx, k_cost_pq_walk<uint8_t>, wgrid, 128, 0, pairs_d, jobs_d);
                else IS_LAUNCH(ctx, k_cost_pq_walk<float>, wgrid, 128, 0, pairs_d, jobs_d);
            }
        }
        {                                                                                        // one DP launch per shape, one CTA per seam
            std::vector<DpShape> shapes(nj);
            for (size_t q = 0; q < nj; ++q) shapes[q] = jobs[order[q]].shape;
            IS_TRY(launch_dp_all(ctx, shapes, dp_d, reinterpret_cast<const BtArgs*>(b1d + off_bt)));
        }
        IS_TRY(download_view(ctx, base + off_res_all, sizeof(int) * res_total_ints, reinterpret_cast<const void**>(&res_h)));
    }
    tm.lap("C labels, costs, DP");
    // ---- host: updateLabelsUsingSeam per seam -- one task per pair
    std::vector<std::vector<size_t>> jobs_of(np);
    for (size_t j = 0; j < nj; ++j) jobs_of[(size_t)jobs[j].pair].push_back(j);   // plan order
    pool->run(limit, [&](size_t k) {
        for (size_t j : jobs_of[k]) {
            SeamJobHost& J = jobs[j];
            J.status = seam_job_finish(ctx, PR[k], J, res_h + J.off_res, trace != nullptr);
            if (J.status != IS_OK) return;
        }
    });
    for (auto& J : jobs) {
        if (J.status == IS_ERR_UNSUPPORTED && tm.on) fprintf(stderr, "[seam batch] pair (%d, %d) leaves the batched path: 255 or more flood-fill regions\n", active[(size_t)J.pair].first, active[(size_t)J.pair].second);
        if (J.status == IS_ERR_UNSUPPORTED) { limit = std::min(limit, (size_t)J.pair); continue; }   // that pair takes the general path
        if (J.status != IS_OK) return J.status;
    }
    if (limit == 0) { *first_unsupported = true; return IS_OK; }
    for (auto& J : jobs) {                                             // keep the finished seams (their flips must not move any more)
        if ((size_t)J.pair >= limit) continue;
        all_jobs.emplace_back(std::move(J));
        SeamJobHost& K = all_jobs.back();
        const std::vector<Interval>* f = K.reached ? &K.uls.flips : nullptr;
        flips_of[(size_t)K.pair].push_back(f);
        round_flips[(size_t)K.pair].push_back(f);
    }
    // staged pairs: the round's relabels into the runs, then on with the conflict loop
    bool more = false;
    std::vector<char> next_round(np, 0);
    pool->run(limit, [&](size_t k) {
        if (!in_round[k] || !PR[k].blocked) return;
        if (round >= 32 || !PR[k].apply_round(round_flips[k])) { PR[k].unsupported = true; return; }
        PR[k].plan_resume();
        next_round[k] = 1;
    });
    for (size_t k = 0; k < limit; ++k) {
        if ((PR[k].unsupported || PR[k].too_many_runs) && tm.on)
            fprintf(stderr, "[seam batch] pair (%d, %d) leaves the batched path in round %d of its staged plan\n", active[k].first, active[k].second, round);
        if (PR[k].unsupported || PR[k].too_many_runs) { limit = k; break; }   // the plan does not cover what came up: general path from this pair on
        more = more || next_round[k];
    }
    if (limit == 0) { *first_unsupported = true; return IS_OK; }
    in_round = next_round;
    if (!more) break;
    }   // rounds
    // ---- host: the final mask update of every pair as clear intervals + the upload tables -- one task per pair
    std::vector<std::vector<ClearIv>> clears(np);
    std::vector<std::vector<int>> rs_all(np);
    std::vector<std::vector<int4>> iv_all(np);
    std::vector<char> clears_failed(np, 0);
    pool->run(limit, [&](size_t k) {
        if (PR[k].staged) {
            if (!PR[k].apply_round(round_flips[k])) { clears_failed[k] = 1; return; }     // the last round's relabels
            PR[k].final_clears(&clears[k]);
        } else {
            pair_clear_intervals(PR[k], flips_of[k], &clears[k]);
        }
        // the pair's upload tables while the task is at it: row_start[ih + 1] and the intervals
        const PairRuns& P = PR[k];
        const int iy = P.iTl.y - P.unionTl.y, ih = P.iBr.y - P.iTl.y;
        std::vector<int>& rs = rs_all[k];
        std::vector<int4>& iv = iv_all[k];
        rs.assign((size_t)ih + 1, 0);
        iv.resize(clears[k].size());
        for (size_t q = 0; q < clears[k].size(); ++q) {
            const ClearIv& c = clears[k][q];
            rs[(size_t)(c.y - iy) + 1]++;
            iv[q] = make_int4(c.x0, c.x1, c.bits, 0);
        }
        for (int r = 0; r < ih; ++r) rs[(size_t)r + 1] += rs[(size_t)r];
    });
    for (size_t k = 0; k < limit; ++k) if (clears_failed[k]) { limit = k; break; }
    if (limit == 0) { *first_unsupported = true; return IS_OK; }
    clears.resize(limit);
    tm.lap("D host: updateLabelsUsingSeam on runs, clear intervals");
    // one upload: per pair row_start[ih + 1] and the intervals, then the table
    Blob B2;
    std::vector<size_t> off_rs(limit), off_iv(limit);
    {
        size_t total = 64 + sizeof(ClearTabDev) * limit;
        for (size_t k = 0; k < limit; ++k) total += sizeof(int) * rs_all[k].size() + sizeof(int4) * iv_all[k].size() + 32;
        B2.host.reserve(total);
    }
    for (size_t k = 0; k < limit; ++k) {
        off_rs[k] = B2.put(rs_all[k].data(), sizeof(int) * rs_all[k].size());
        off_iv[k] = B2.put(iv_all[k].data(), sizeof(int4) * iv_all[k].size());
    }
    const size_t off_ct = B2.put(nullptr, sizeof(ClearTabDev) * limit);
    DevBuf dev2;
    IS_TRY(dev2.alloc(ctx, B2.host.size()));
    unsigned char* b2d = dev2.as<unsigned char>();
    int max_ih = 0;
    {
        ClearTabDev* ct = reinterpret_cast<ClearTabDev*>(B2.host.data() + off_ct);
        for (size_t k = 0; k < limit; ++k) {
            const PairRuns& P = PR[k];
            const DevMat& m1 = masks[P.pi];
            const DevMat& m2 = masks[P.pj];
            ct[k].row_start = reinterpret_cast<const int*>(b2d + off_rs[k]);
            ct[k].ivs = reinterpret_cast<const int4*>(b2d + off_iv[k]);
            ct[k].iy = P.iTl.y - P.unionTl.y; ct[k].ih = P.iBr.y - P.iTl.y;
            ct[k].mask1 = m1.ptr<uint8_t>(); ct[k].mstep1 = m1.step; ct[k].o1x = P.o1x; ct[k].o1y = P.o1y;
            ct[k].mask2 = m2.ptr<uint8_t>(); ct[k].mstep2 = m2.step; ct[k].o2x = P.o2x; ct[k].o2y = P.o2y;
            max_ih = std::max(max_ih, ct[k].ih);
        }
    }
    IS_TRY(upload(ctx, b2d, B2.host.data(), B2.host.size()));
    tm.lap("E host: clear intervals + upload");
    // ---- F: validation -- the masks every pair would have seen in the sequential loop
    //      Where the earlier pairs' clears stay away from a pair's intersection rectangle (every pair of a strip), its special points
    //      cannot change and the masks it would have seen follow from the entry toggles minus the clear intervals: such a pair is
    //      validated on the host alone (runs_minus_clears, the first half of build(), same_window); the others ask the device again.
    std::vector<char> host_checked(limit, 0);
    std::vector<int> host_decision(limit, -1);
    const bool check_both = getenv("IS_SEAM_CHECK_BOTH") != nullptr;   // test knob: validate on the device as well and insist on the same verdicts
    if (!getenv("IS_SEAM_CHECK_DEVICE")) {
        std::vector<std::vector<RunLayer>> lay(2 * limit);
        std::vector<size_t> hk;
        for (size_t k = 0; k < limit; ++k) {
            const int img[2] = {active[k].first, active[k].second};
            bool any = false, near = false;
            for (int s = 0; s < 2; ++s)
                for (size_t q = 0; q < k; ++q) {
                    int bit = 0;
                    if (active[q].first == img[s]) bit = 1; else if (active[q].second == img[s]) bit = 2;
                    if (!bit || clears[q].empty()) continue;
                    const PairRuns& E = PR[q];
                    lay[2 * k + (size_t)s].push_back(RunLayer{&clears[q], bit, E.unionTl.x - corners[img[s]].x, E.unionTl.y - corners[img[s]].y});
                    any = true;
                    // the clears of pair q lie inside its intersection rectangle; special points look three pixels around theirs
                    const PairRuns& P = PR[k];
                    if (E.iTl.x < P.iBr.x + 3 && P.iTl.x - 3 < E.iBr.x && E.iTl.y < P.iBr.y + 3 && P.iTl.y - 3 < E.iBr.y) near = true;
                }
            if (any && !near) { hk.push_back(k); host_checked[k] = 1; }
        }
        if (!hk.empty()) {
            std::vector<char> ok(hk.size(), 0);
            pool->run(hk.size(), [&](size_t v) {
                const size_t k = hk[v];
                MaskRuns nr[2];
                const MaskRuns* use[2];
                for (int s = 0; s < 2; ++s) {
                    const MaskRuns& base = Q.runs[(size_t)(s == 0 ? Q.pairs[k].m1 : Q.pairs[k].m2)];
                    use[s] = &base;
                    if (lay[2 * k + (size_t)s].empty()) continue;
                    if (!runs_minus_clears(base, lay[2 * k + (size_t)s], TG_CAP, &nr[s])) return;   // too many toggles: the pair starts the next wave
                    use[s] = &nr[s];
                }
                PairRuns C;
                C.setup(active[k].first, active[k].second, PR[k].tl1, PR[k].tl2, use[0], use[1]);
                C.specials = PR[k].specials;
                C.build_runs();
                if (C.too_many_runs) return;
                if (C.same_window(PR[k])) { ok[v] = 1; return; }
                C.build_contours();
                C.plan();
                ok[v] = C.same_structure(PR[k]) ? 1 : 0;
            });
            for (size_t v = 0; v < hk.size(); ++v) host_decision[hk[v]] = ok[v];
            if (check_both) std::fill(host_checked.begin(), host_checked.end(), 0);
            else
                for (size_t v = 0; v < hk.size(); ++v)
                    if (!ok[v]) { limit = std::min(limit, hk[v]); break; }
            tm.lap("F host: check structures (runs)");
        }
    }
    {
        StructureQuery V;
        std::vector<int> vpair;                                        // wave index of every validation pair
        for (size_t k = 0; k < limit; ++k) {
            if (host_checked[k]) continue;
            const int img[2] = {active[k].first, active[k].second};
            LayeredMask lm[2] = {plain_mask(masks[img[0]]), plain_mask(masks[img[1]])};
            bool any = false, too_many = false;
            for (int s = 0; s < 2 && !too_many; ++s)
                for (size_t q = 0; q < k; ++q) {
                    int bit = 0;
                    if (active[q].first == img[s]) bit = 1; else if (active[q].second == img[s]) bit = 2;
                    if (!bit) continue;
                    if (clears[q].empty()) continue;                   // that pair clears nothing
                    if (lm[s].nlayers == MAX_LAYERS) { too_many = true; break; }       // more earlier neighbours than a mask carries
                    const PairRuns& E = PR[q];
                    ClearLayer& L = lm[s].layer[lm[s].nlayers++];
                    L.row_start = reinterpret_cast<const int*>(b2d + off_rs[q]);
                    L.ivs = reinterpret_cast<const int4*>(b2d + off_iv[q]);
                    L.dx = E.unionTl.x - corners[img[s]].x; L.dy = E.unionTl.y - corners[img[s]].y;
                    L.iy = E.iTl.y - E.unionTl.y; L.ih = E.iBr.y - E.iTl.y; L.bit = bit;
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
                C.build_runs();
                if (C.too_many_runs) return;
                // the same test as on the host path (a staged pair is compared as build() left it), plus the special points, which
                // were recomputed here; only then the full comparison of contours and plans
                bool same_specials = C.specials.size() == PR[k].specials.size();
                for (size_t q = 0; same_specials && q < C.specials.size(); ++q)
                    same_specials = C.specials[q].x == PR[k].specials[q].x && C.specials[q].y == PR[k].specials[q].y;
                if (same_specials && C.same_window(PR[k])) { ok[v] = 1; return; }
                C.build_contours();
                C.plan();
                ok[v] = C.same_structure(PR[k]) ? 1 : 0;
            });
            if (check_both)
                for (size_t v = 0; v < vpair.size(); ++v)
                    if (host_decision[(size_t)vpair[v]] >= 0 && host_decision[(size_t)vpair[v]] != ok[v])
                        return fail(ctx, IS_ERR_ASSERT, "seam validation: host verdict %d, device verdict %d for pair %d of the wave", host_decision[(size_t)vpair[v]], (int)ok[v], vpair[v]);
            for (size_t v = 0; v < vpair.size(); ++v)
                if (!ok[v]) { limit = std::min(limit, (size_t)vpair[v]); break; }   // the first pair the earlier clears change starts the next wave
            tm.lap("F host: check structures");
        }
    }
    // ---- G: the clear intervals of the accepted prefix go into the masks
    IS_REQUIRE(ctx, limit >= 1, IS_ERR_INTERNAL, "seam wave without progress");
    if (max_ih > 0) {
        double bytes = 0;
        for (size_t k = 0; k < limit; ++k) for (const ClearIv& c : clears[k]) bytes += c.x1 - c.x0;
        ctx->next_bytes = bytes;
        IS_LAUNCH(ctx, k_apply_clear_runs_batch, dim3(div_up(max_ih, 8), (unsigned)limit), 256, 0, reinterpret_cast<const ClearTabDev*>(b2d + off_ct));
    }
    if (trace)
        for (size_t k = 0; k < limit; ++k)
            for (auto& J : all_jobs) {
                if ((size_t)J.pair != k || J.trace.empty()) continue;
                const size_t len = J.trace.size();
                if (trace->buf && trace->len + len <= trace->cap) std::memcpy(trace->buf + trace->len, J.trace.data(), len * sizeof(int32_t));
                trace->len += len;
            }
    if (tm.on) cudaStreamSynchronize(ctx->stream);
    tm.lap("G apply");
    *accepted_n = limit;
    return IS_OK;
}
