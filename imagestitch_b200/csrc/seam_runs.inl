// seam_runs.inl -- host side of the batched seam path: the structure of an image pair in the RUN domain.
// (included by seam.cu inside namespace is, after PairSeam)
//
// Warped masks are a few horizontal runs per row.  k_row_toggles_batch reduces every mask to its per-row toggle positions once
// per call; everything the reference derives from the masks before the first seam estimation -- the class image and its
// connected components ([SEAM]:196-308), bounding boxes and raster-ordered contour lists of the INTERS components, the edge
// set ([SEAM]:311-392), the seam tips ([SEAM]:607-706) and the order of the conflict loop ([SEAM]:395-546) -- is computed
// here from those runs, without a labels image and without a device round trip.  What it produces is a PLAN: the list of
// relabel operations and seam estimations (component, orientation, tips) the device then executes for all pairs at once.

struct MaskRuns {                     // toggles of one mask in its own coordinates
    int rows = 0, cols = 0, slots = 0;
    std::vector<unsigned char> counts;   // [rows]
    std::vector<unsigned short> xs;      // [rows][slots]
    int count(int y) const { return counts[(size_t)y]; }
    const unsigned short* row(int y) const { return xs.data() + (size_t)y * slots; }
    // number of intervals of row y and the i-th one [a, b)
    int nint(int y) const { return (count(y) + 1) >> 1; }
    void interval(int y, int i, int* a, int* b) const {
        const unsigned short* r = row(y);
        *a = r[2 * i];
        *b = 2 * i + 1 < count(y) ? r[2 * i + 1] : cols;
    }
    bool inside(int x, int y) const {
        if ((unsigned)y >= (unsigned)rows || (unsigned)x >= (unsigned)cols) return false;
        const unsigned short* r = row(y);
        const int n = count(y);
        int k = 0;
        while (k < n && r[k] <= x) ++k;      // rows hold a handful of toggles
        return (k & 1) != 0;
    }
};

// host reference of k_row_toggles_batch (used by the CPU-side structure check, tests/test_seam_runs_host.py)
static void mask_runs_from_host(const uint8_t* mask, size_t step, int rows, int cols, MaskRuns* out) {
    out->rows = rows; out->cols = cols;
    out->counts.assign((size_t)rows, 0);
    std::vector<std::vector<int>> tmp((size_t)rows);
    int slots = 1;
    for (int y = 0; y < rows; ++y) {
        int prev = 0;
        for (int x = 0; x < cols; ++x) {
            const int v = mask[(size_t)y * step + x] != 0;
            if (v != prev) { tmp[(size_t)y].push_back(x); prev = v; }
        }
        out->counts[(size_t)y] = (unsigned char)std::min<size_t>(tmp[(size_t)y].size(), 255);
        slots = std::max(slots, (int)tmp[(size_t)y].size());
    }
    out->slots = slots;
    out->xs.assign((size_t)rows * slots, 0);
    for (int y = 0; y < rows; ++y) std::copy(tmp[(size_t)y].begin(), tmp[(size_t)y].end(), out->xs.begin() + (size_t)y * slots);
}

struct SeamOp {                       // one step of the conflict loop ([SEAM]:423-520)
    int kind;                         // 0: relabel the whole component c1 -> label of c2 ([SEAM]:440-450); 1: estimateSeam + updateLabelsUsingSeam
    int c1, c2;
    Pt p1, p2;                        // seam tips (kind 1), union-frame coordinates
    int rx, ry, rw, rh;               // bounding box of c1 at that moment
};

class PairRuns {
public:
    int pi = 0, pj = 0;
    Pt tl1{}, tl2{}, unionTl{}, iTl{}, iBr{};
    int uw = 0, uh = 0;
    int rows1 = 0, cols1 = 0, rows2 = 0, cols2 = 0;
    int o1x = 0, o1y = 0, o2x = 0, o2y = 0;          // mask origins in the union frame
    const MaskRuns* mr[2] = {nullptr, nullptr};
    // class change points of every union-frame row, their labels
    std::vector<int> row_off;
    std::vector<ChangePt> cps;
    std::vector<int> cp_label;
    int ncomps = 0;
    std::vector<int> states;                          // initial states ([SEAM]:240-250)
    std::vector<int> final_states;                    // after the conflict loop
    int wx = 0, wy = 0, ww = 0, wh = 0;               // label window: intersection rectangle grown by one pixel
    size_t wcap = 1;
    std::vector<int> tab;                             // counts | ChangePt (x, cls) | labels of the window rows (k_label_window layout)
    std::vector<Pt> tls, brs;
    std::vector<std::vector<ContourRec>> contours;    // INTERS components only, raster order; flags are NOT filled here
    std::set<std::pair<int, int>> edges;
    std::vector<SeamOp> ops;
    std::vector<Pt> specials;                         // see get_seam_tips
    bool too_many_runs = false;                       // a row has more change points than the window kernel's table holds
    bool unsupported = false;                         // the conflict loop needs device results the plan cannot anticipate

    void setup(int i, int j, Pt t1, Pt t2, const MaskRuns* m1, const MaskRuns* m2);
    void build();                                     // merge, components, contours, edges
    void build_runs();                                // the first half: merged runs, components, labels, the label window
    void build_contours();                            // the second half: contour records of the INTERS components, edges
    // true when the runs and labels of `o` (same geometry) agree with these on everything the second half, the plan and the seam
    // tips read: the label window (intersection rectangle grown by one pixel) and the states of the labels in it
    bool same_window(const PairRuns& o) const;
    void build_tab();                                 // the label-window table (only the label-image path of the cost kernels needs it)
    // the runs of component c (0-based) in the rows [ry, ry + rh) of the frame, at most `cap` per row: out[(y - ry) * cap + k] = (x0, x1),
    // unused slots (0, 0); false when a row has more runs than cap
    bool component_runs(int c, int ry, int rh, int cap, std::vector<int2>* out) const;
    void plan();                                      // conflict loop -> ops, final_states
    // Staged plan.  The conflict loop can decide everything before any seam is known -- except a SECOND seam estimation on a
    // component a seam of this pair has already cut (curved mosaic masks leave an INTERS component two neighbours of the same
    // image): its tips and its cost region depend on the first seam.  plan() then stops in front of that operation (`blocked`);
    // once the seams of the operations planned so far are back, apply_round() writes their relabels into the runs and refreshes
    // the contours, plan_resume() carries on.  A pair that has been through apply_round() is `staged`: its final mask update comes
    // from the relabelled runs (final_clears) instead of pair_clear_intervals.
    bool blocked = false, staged = false;
    size_t round_begin = 0;                           // first operation of the current round in `ops`
    std::vector<int> snap_row_off, snap_label;        // staged pairs: the runs as build() left them (what a validation compares with)
    std::vector<ChangePt> snap_cps;
    void plan_resume();
    // seam_flips: UlsRuns::flips of the round's kind-1 operations in order (nullptr: estimateSeam failed); false: a row now has more runs than the tables hold
    bool apply_round(const std::vector<const std::vector<struct Interval>*>& seam_flips);
    void final_clears(std::vector<struct ClearIv>* out) const;
    bool same_structure(const PairRuns& o) const;

private:
    std::set<std::pair<int, int>> ed_;                // conflict-loop state between rounds: remaining edges, components cut by a seam of this round
    std::vector<char> cut_;
    void plan_loop();
    bool relabel(int l_from, int l_to, const std::vector<struct Interval>* fl);   // fl == nullptr: every pixel of l_from
    int label_at(int x, int y) const;                 // label of a frame pixel from the runs (0 background, -1 outside the frame)
    void contour_rows();
    void find_edges();
    bool has_only_one_neighbor(int comp) const;
    bool get_seam_tips(int c1, int c2, Pt* p1, Pt* p2) const;
};

void PairRuns::setup(int i, int j, Pt t1, Pt t2, const MaskRuns* m1, const MaskRuns* m2) {
    pi = i; pj = j; tl1 = t1; tl2 = t2; mr[0] = m1; mr[1] = m2;
    rows1 = m1->rows; cols1 = m1->cols; rows2 = m2->rows; cols2 = m2->cols;
    iTl = {std::max(tl1.x, tl2.x), std::max(tl1.y, tl2.y)};
    iBr = {std::min(tl1.x + cols1, tl2.x + cols2), std::min(tl1.y + rows1, tl2.y + rows2)};
    unionTl = {std::min(tl1.x, tl2.x), std::min(tl1.y, tl2.y)};
    const Pt unionBr{std::max(tl1.x + cols1, tl2.x + cols2), std::max(tl1.y + rows1, tl2.y + rows2)};
    uw = unionBr.x - unionTl.x;
    uh = unionBr.y - unionTl.y;
    o1x = tl1.x - unionTl.x; o1y = tl1.y - unionTl.y; o2x = tl2.x - unionTl.x; o2y = tl2.y - unionTl.y;
}

int PairRuns::label_at(int x, int y) const {
    if ((unsigned)x >= (unsigned)uw || (unsigned)y >= (unsigned)uh) return -1;
    const int b = row_off[(size_t)y], e = row_off[(size_t)y + 1];
    int k = b;
    while (k < e && cps[(size_t)k].x <= x) ++k;
    return k == b ? 0 : cp_label[(size_t)k - 1];
}

// [SEAM]:196-308 in run-length form
void PairRuns::build() {
    build_runs();
    build_contours();
}

void PairRuns::build_runs() {
    const bool tim = getenv("IS_DEBUG_PLAN_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lapt = [&](const char* w) { if (!tim) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "   [build] %-16s %.3f ms\n", w, std::chrono::duration<double, std::milli>(n - T0).count()); T0 = n; };
    const int ox[2] = {o1x, o2x}, oy[2] = {o1y, o2y};
    row_off.assign((size_t)uh + 1, 0);
    cps.clear();
    cps.reserve((size_t)uh * 6);
    too_many_runs = false;
    for (int y = 0; y < uh; ++y) {
        int na[2] = {0, 0};
        const unsigned short* xa[2] = {nullptr, nullptr};
        for (int k = 0; k < 2; ++k) {
            const int my = y - oy[k];
            if (my >= 0 && my < mr[k]->rows) { na[k] = mr[k]->count(my); xa[k] = mr[k]->row(my); }
        }
        // position i of mask k: xa[k][i] + ox for i < na[k], then (when the row ends inside the mask) the closing toggle at ox + cols
        auto pos = [&](int k, int i) { return i < na[k] ? xa[k][i] + ox[k] : ((na[k] & 1) && i == na[k] ? ox[k] + mr[k]->cols : INT_MAX); };
        int i0 = 0, i1 = 0, st = 0, prev_cls = 0, emitted = 0;
        for (;;) {
            const int p0 = pos(0, i0), p1 = pos(1, i1);
            const int x = std::min(p0, p1);
            if (x >= uw) break;                     // INT_MAX (both exhausted) or a closing toggle on the frame's right edge
            if (p0 == x) { st ^= 1; ++i0; }
            if (p1 == x) { st ^= 2; ++i1; }
            if (st != prev_cls) { cps.push_back(ChangePt{x, st}); prev_cls = st; ++emitted; }
        }
        if (emitted > ROW_CAP) too_many_runs = true;
        row_off[(size_t)y + 1] = row_off[(size_t)y] + emitted;
    }
    lapt("merge");
    const int R = row_off[(size_t)uh];
    // union-find over the runs (run k = change point k with cls != 0, spanning [x, next change point or uw))
    std::vector<int> uf((size_t)std::max(R, 1));
    for (int k = 0; k < R; ++k) uf[(size_t)k] = k;
    auto find = [&](int k) { while (uf[(size_t)k] != k) { uf[(size_t)k] = uf[(size_t)uf[(size_t)k]]; k = uf[(size_t)k]; } return k; };
    auto run_end = [&](int k, int y) { return k + 1 < row_off[(size_t)y + 1] ? cps[(size_t)k + 1].x : uw; };
    for (int y = 1; y < uh; ++y) {
        int a = row_off[(size_t)y - 1], ae = row_off[(size_t)y], b = row_off[(size_t)y], be = row_off[(size_t)y + 1];
        while (a < ae && b < be) {
            const int ax1 = run_end(a, y - 1), bx1 = run_end(b, y);
            if (cps[(size_t)a].cls && cps[(size_t)a].cls == cps[(size_t)b].cls && cps[(size_t)a].x < bx1 && cps[(size_t)b].x < ax1) {
                const int ra = find(a), rb = find(b);
                if (ra != rb) { if (ra < rb) uf[(size_t)rb] = ra; else uf[(size_t)ra] = rb; }   // the root is the raster-first run
            }
            if (ax1 <= bx1) ++a; else ++b;
        }
    }
    lapt("union-find");
    cp_label.assign((size_t)std::max(R, 1), 0);
    std::vector<int> id_of_root((size_t)std::max(R, 1), 0);
    ncomps = 0;
    states.clear();
    for (int k = 0; k < R; ++k) {               // roots in increasing index = raster order of the first pixel
        if (!cps[(size_t)k].cls || find(k) != k) continue;
        id_of_root[(size_t)k] = ++ncomps;
        states.push_back(cps[(size_t)k].cls == 3 ? ST_INTERS : (cps[(size_t)k].cls == 1 ? ST_FIRST : ST_SECOND));
    }
    for (int k = 0; k < R; ++k) cp_label[(size_t)k] = cps[(size_t)k].cls ? id_of_root[(size_t)find(k)] : 0;
    lapt("labels");
    // labels are materialised on the device only where they are read: the intersection rectangle grown by one pixel
    wx = std::max(0, iTl.x - unionTl.x - 1);
    wy = std::max(0, iTl.y - unionTl.y - 1);
    ww = std::min(uw, iBr.x - unionTl.x + 1) - wx;
    wh = std::min(uh, iBr.y - unionTl.y + 1) - wy;
    lapt("tab");
}

void PairRuns::build_contours() {
    const bool tim = getenv("IS_DEBUG_PLAN_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lapt = [&](const char* w) { if (!tim) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "   [build] %-16s %.3f ms\n", w, std::chrono::duration<double, std::milli>(n - T0).count()); T0 = n; };
    tls.assign((size_t)ncomps, Pt{INT_MAX, INT_MAX});
    brs.assign((size_t)ncomps, Pt{INT_MIN, INT_MIN});
    contours.assign((size_t)ncomps, std::vector<ContourRec>());
    contour_rows();
    lapt("contours");
    find_edges();
    lapt("edges");
}

bool PairRuns::same_window(const PairRuns& o) const {
    if (uw != o.uw || uh != o.uh || wx != o.wx || wy != o.wy || ww != o.ww || wh != o.wh || too_many_runs != o.too_many_runs) return false;
    auto state_of = [](const std::vector<int>& st, int label) { return (label >= 1 && label <= (int)st.size()) ? st[(size_t)label - 1] : -1; };
    const int x_lo = wx, x_hi = wx + ww;
    // a staged pair is compared as build() left it, before its own seams were written into its runs
    const std::vector<int>& o_row_off = o.staged ? o.snap_row_off : o.row_off;
    const std::vector<ChangePt>& o_cps = o.staged ? o.snap_cps : o.cps;
    const std::vector<int>& o_label = o.staged ? o.snap_label : o.cp_label;
    for (int y = wy; y < wy + wh; ++y) {
        // the label function of the row over [x_lo, x_hi) as (start, label) pieces, from either structure, walked in step
        const int ab = row_off[(size_t)y], ae = row_off[(size_t)y + 1], bb = o_row_off[(size_t)y], be = o_row_off[(size_t)y + 1];
        int ia = ab, ib = bb;
        while (ia < ae && cps[(size_t)ia].x <= x_lo) ++ia;                  // ia: first change point right of x_lo
        while (ib < be && o_cps[(size_t)ib].x <= x_lo) ++ib;
        int x = x_lo;
        for (;;) {
            const int la = ia == ab ? 0 : cp_label[(size_t)ia - 1], lb = ib == bb ? 0 : o_label[(size_t)ib - 1];
            if (la != lb) return false;
            if (la > 0 && state_of(states, la) != state_of(o.states, la)) return false;
            const int na = ia < ae ? std::min(cps[(size_t)ia].x, x_hi) : x_hi, nb = ib < be ? std::min(o_cps[(size_t)ib].x, x_hi) : x_hi;
            if (na != nb) return false;
            x = na;
            if (x >= x_hi) break;
            ++ia; ++ib;
        }
    }
    return true;
}

// The toggles of a mask after the clear intervals of earlier pairs have been applied to it (what k_row_toggles_batch reports for a
// layered mask), on the host.  `layers`: per layer the clear list (sorted by row, frame coordinates of the pair that produced
// it), the bit that selects this mask's intervals and the offset (dx, dy) from those frame coordinates to this mask's own.
// false: a row ends up with more toggles than the tables hold.
struct RunLayer { const std::vector<struct ClearIv>* clears; int bit, dx, dy; };
static bool runs_minus_clears(const MaskRuns& in, const std::vector<RunLayer>& layers, int cap, MaskRuns* out);

void PairRuns::build_tab() {
    wcap = 1;
    for (int y = wy; y < wy + wh; ++y) wcap = std::max(wcap, (size_t)(row_off[(size_t)y + 1] - row_off[(size_t)y]));
    const size_t wrows = (size_t)wh;
    tab.assign(wrows + wrows * wcap * 3, 0);
    int* t_cnt = tab.data();
    ChangePt* t_cps = reinterpret_cast<ChangePt*>(tab.data() + wrows);
    int* t_lab = tab.data() + wrows + wrows * wcap * 2;
    for (int y = wy; y < wy + wh; ++y) {
        const size_t r = (size_t)(y - wy);
        t_cnt[r] = row_off[(size_t)y + 1] - row_off[(size_t)y];
        std::copy(cps.begin() + row_off[(size_t)y], cps.begin() + row_off[(size_t)y + 1], t_cps + r * wcap);
        std::copy(cp_label.begin() + row_off[(size_t)y], cp_label.begin() + row_off[(size_t)y + 1], t_lab + r * wcap);
    }
}

bool PairRuns::component_runs(int c, int ry, int rh, int cap, std::vector<int2>* out) const {
    out->assign((size_t)rh * cap, make_int2(0, 0));
    const int l = c + 1;
    for (int y = 0; y < rh; ++y) {
        const int fy = ry + y;
        const int rb = row_off[(size_t)fy], re = row_off[(size_t)fy + 1];
        int k = 0;
        for (int q = rb; q < re; ++q) {
            if (cp_label[(size_t)q] != l) continue;
            if (k == cap) return false;
            (*out)[(size_t)y * cap + k++] = make_int2(cps[(size_t)q].x, q + 1 < re ? cps[(size_t)q + 1].x : uw);
        }
    }
    return true;
}

// Raster-ordered contour records of the INTERS components: a pixel of an INTERS run is a contour pixel when one of its
// 4-neighbours lies outside the frame or carries another label ([SEAM]:262-273).  Inside a run only the two end pixels can
// have a different left / right neighbour; the up / down neighbours are constant over the segments into which the runs of
// the rows above and below cut the run.
void PairRuns::contour_rows() {
    struct Seg { int x0, x1, lab; };      // [x0, x1): label of row r over that range
    std::vector<Seg> up, dn;
    auto row_segs = [&](int r, int a, int b, std::vector<Seg>& out) {      // labels of row r over [a, b)
        out.clear();
        if (r < 0 || r >= uh) { out.push_back(Seg{a, b, -1}); return; }
        const int rb = row_off[(size_t)r], re = row_off[(size_t)r + 1];
        int k = rb;
        while (k < re && cps[(size_t)k].x <= a) ++k;                       // k: first change point right of a
        int x = a;
        while (x < b) {
            const int lab = k == rb ? 0 : cp_label[(size_t)k - 1];
            const int nx = k < re ? std::min(cps[(size_t)k].x, b) : b;
            out.push_back(Seg{x, nx, lab});
            x = nx;
            ++k;
        }
    };
    for (int y = std::max(0, iTl.y - unionTl.y); y < std::min(uh, iBr.y - unionTl.y); ++y) {
        const int rb = row_off[(size_t)y], re = row_off[(size_t)y + 1];
        for (int k = rb; k < re; ++k) {
            if (cps[(size_t)k].cls != 3) continue;
            const int a = cps[(size_t)k].x, b = k + 1 < re ? cps[(size_t)k + 1].x : uw;
            const int l = cp_label[(size_t)k];
            if (!(states[(size_t)l - 1] & ST_INTERS)) continue;               // staged pairs: pixels a seam has handed to a FIRST / SECOND component
            const int left = a == 0 ? -1 : (k == rb ? 0 : cp_label[(size_t)k - 1]);
            const int right = b == uw ? -1 : (k + 1 < re ? cp_label[(size_t)k + 1] : 0);   // b < uw implies a following change point
            row_segs(y - 1, a, b, up);
            row_segs(y + 1, a, b, dn);
            std::vector<ContourRec>& out = contours[(size_t)l - 1];
            if (out.capacity() == 0) out.reserve(4 * (size_t)((iBr.x - iTl.x) + (iBr.y - iTl.y)) + 64);
            Pt& tl = tls[(size_t)l - 1];
            Pt& br = brs[(size_t)l - 1];
            size_t iu = 0, id = 0;
            int x = a;
            while (x < b) {
                while (up[iu].x1 <= x) ++iu;
                while (dn[id].x1 <= x) ++id;
                const int e = std::min(up[iu].x1, dn[id].x1);          // [x, e): constant up / down labels
                const int ul = up[iu].lab, dl = dn[id].lab;
                auto emit = [&](int px) {
                    ContourRec r;
                    r.x = px; r.y = y; r.label = l;
                    r.nl[0] = px == a ? left : l;
                    r.nl[1] = ul;
                    r.nl[2] = px == b - 1 ? right : l;
                    r.nl[3] = dl;
                    r.flags = 0;
                    out.push_back(r);
                    tl.x = std::min(tl.x, px); tl.y = std::min(tl.y, y);
                    br.x = std::max(br.x, px + 1); br.y = std::max(br.y, y + 1);
                };
                if (ul != l || dl != l) {
                    for (int px = x; px < e; ++px) emit(px);
                } else {
                    if (x == a) emit(a);                                   // the end pixels of a run always differ from their outer neighbour
                    if (b - 1 != a && b - 1 >= x && b - 1 < e) emit(b - 1);
                }
                x = e;
            }
        }
    }
}

// [SEAM]:311-392 (edges whose first component is INTERS, both directions; the only ones the conflict loop reads)
void PairRuns::find_edges() {
    edges.clear();
    std::vector<char> seen((size_t)ncomps + 1, 0);
    std::vector<int> touched;
    for (int ci = 0; ci < ncomps; ++ci) {
        if (contours[(size_t)ci].empty()) continue;
        const int l = ci + 1;
        touched.clear();
        for (const ContourRec& r : contours[(size_t)ci])
            for (int k = 0; k < 4; ++k) {
                const int nl = r.nl[k];
                if (nl > 0 && nl != l && !seen[(size_t)nl]) { seen[(size_t)nl] = 1; touched.push_back(nl); }
            }
        for (int nl : touched) {
            edges.insert({ci, nl - 1});
            edges.insert({nl - 1, ci});
            seen[(size_t)nl] = 0;
        }
    }
}

bool PairRuns::has_only_one_neighbor(int comp) const {   // [SEAM]:575-581
    auto begin = edges.lower_bound({comp, INT_MIN});
    auto end = edges.upper_bound({comp, INT_MAX});
    return ++begin == end;
}

// [SEAM]:607-706.  `specials`: the pixels of the pair's INTERS class that touch a pixel of exactly one mask and lie within
// two pixels of both masks' contours (closeToContour, [SEAM]:584-604), in raster order -- found on the device by
// k_special_points_batch straight from the masks; the labels come from the runs.
bool PairRuns::get_seam_tips(int c1, int c2, Pt* p1, Pt* p2) const {
    const int l1 = c1 + 1, l2 = c2 + 1;
    std::vector<Pt> special;
    for (const Pt& p : specials) {
        if (label_at(p.x, p.y) != l1) continue;
        if (label_at(p.x - 1, p.y) == l2 || label_at(p.x, p.y - 1) == l2 || label_at(p.x + 1, p.y) == l2 || label_at(p.x, p.y + 1) == l2) special.push_back(p);
    }
    if (special.size() < 2) return false;
    // cv::partition with ClosePoints(10): connected components of "dist^2 < 100", classes numbered by first member
    const int n = (int)special.size();
    std::vector<int> uf((size_t)n);
    for (int i = 0; i < n; ++i) uf[(size_t)i] = i;
    auto find = [&](int i) { while (uf[(size_t)i] != i) { uf[(size_t)i] = uf[(size_t)uf[(size_t)i]]; i = uf[(size_t)i]; } return i; };
    for (int i = 0; i < n; ++i)                                   // raster order: only rows within 10 of each other can be close
        for (int j = i + 1; j < n && special[(size_t)j].y - special[(size_t)i].y < 10; ++j) {
            const int dx = special[(size_t)i].x - special[(size_t)j].x, dy = special[(size_t)i].y - special[(size_t)j].y;
            if (dx * dx + dy * dy < 100) { const int a = find(i), b = find(j); if (a != b) uf[(size_t)b] = a; }
        }
    std::vector<int> cls_of_root((size_t)n, -1), lab((size_t)n);
    int nlabels = 0;
    for (int i = 0; i < n; ++i) { const int r = find(i); if (cls_of_root[(size_t)r] < 0) cls_of_root[(size_t)r] = nlabels++; lab[(size_t)i] = cls_of_root[(size_t)r]; }
    if (nlabels < 2) return false;
    std::vector<long long> sumx((size_t)nlabels, 0), sumy((size_t)nlabels, 0);
    std::vector<std::vector<Pt>> points((size_t)nlabels);
    for (int i = 0; i < n; ++i) { sumx[(size_t)lab[(size_t)i]] += special[(size_t)i].x; sumy[(size_t)lab[(size_t)i]] += special[(size_t)i].y; points[(size_t)lab[(size_t)i]].push_back(special[(size_t)i]); }
    int idx[2] = {-1, -1};
    double maxDist = -std::numeric_limits<double>::max();
    for (int i = 0; i < nlabels - 1; ++i)
        for (int j = i + 1; j < nlabels; ++j) {
            const double s1 = (double)points[(size_t)i].size(), s2 = (double)points[(size_t)j].size();
            const double cx1 = round_half_even((int)sumx[(size_t)i] / s1), cy1 = round_half_even((int)sumy[(size_t)i] / s1);
            const double cx2 = round_half_even((int)sumx[(size_t)j] / s2), cy2 = round_half_even((int)sumy[(size_t)j] / s2);
            const double dist = (cx1 - cx2) * (cx1 - cx2) + (cy1 - cy2) * (cy1 - cy2);
            if (dist > maxDist) { maxDist = dist; idx[0] = i; idx[1] = j; }
        }
    Pt p[2];
    for (int i = 0; i < 2; ++i) {
        const std::vector<Pt>& pts = points[(size_t)idx[i]];
        const double size = (double)pts.size();
        const double cx = round_half_even((int)sumx[(size_t)idx[i]] / size), cy = round_half_even((int)sumy[(size_t)idx[i]] / size);
        size_t closest = 0;
        double minDist = std::numeric_limits<double>::max();
        for (size_t j = 0; j < pts.size(); ++j) {
            const double dist = (pts[j].x - cx) * (pts[j].x - cx) + (pts[j].y - cy) * (pts[j].y - cy);
            if (dist < minDist) { minDist = dist; closest = j; }
        }
        p[i] = pts[closest];
    }
    *p1 = p[0];
    *p2 = p[1];
    return true;
}

// The conflict loop [SEAM]:395-546 as far as it can be decided before any seam is known.  Everything it decides -- which
// component meets which, who is relabelled wholesale, which seams are estimated between which tips, the final states -- depends
// on the run structure only; what a seam estimation changes (part of c1 becomes l2) is never read again by the loop unless the
// same component enters a second estimation, or an INTERS component is on the receiving side (neither happens with warped
// panorama masks; both are reported as `unsupported` and take the general path of PairSeam).
void PairRuns::plan() {
    ops.clear();
    unsupported = false;
    blocked = false;
    staged = false;
    round_begin = 0;
    final_states = states;
    ed_ = edges;
    cut_.assign((size_t)ncomps, 0);
    plan_loop();
}

void PairRuns::plan_resume() {
    blocked = false;
    std::fill(cut_.begin(), cut_.end(), 0);           // the runs and contours are current again
    plan_loop();
}

void PairRuns::plan_loop() {
    const bool no_resume = getenv("IS_SEAM_NO_RESUME") != nullptr;   // test knob: such pairs go to the general path as before
    std::vector<int>& st = final_states;
    std::set<std::pair<int, int>>& ed = ed_;
    auto only_one = [&](int comp) {
        auto begin = ed.lower_bound({comp, INT_MIN});
        auto end = ed.upper_bound({comp, INT_MAX});
        return ++begin == end;
    };
    for (;;) {
        int c1 = 0, c2 = 0;
        bool hasConflict = false;
        for (auto itr = ed.begin(); itr != ed.end(); ++itr) {
            c1 = itr->first;
            c2 = itr->second;
            if ((st[(size_t)c1] & ST_INTERS) && (st[(size_t)c1] & (~ST_INTERS)) != st[(size_t)c2]) { hasConflict = true; break; }
        }
        if (!hasConflict) break;
        if (st[(size_t)c2] & ST_INTERS) { unsupported = true; return; }   // c2's geometry would have to be refreshed ([SEAM]:499-513)
        SeamOp op;
        op.c1 = c1; op.c2 = c2;
        const bool empty_box = tls[(size_t)c1].x >= brs[(size_t)c1].x || tls[(size_t)c1].y >= brs[(size_t)c1].y;   // no pixel of c1 is left
        op.rx = empty_box ? 0 : tls[(size_t)c1].x; op.ry = empty_box ? 0 : tls[(size_t)c1].y;
        op.rw = empty_box ? 0 : brs[(size_t)c1].x - tls[(size_t)c1].x; op.rh = empty_box ? 0 : brs[(size_t)c1].y - tls[(size_t)c1].y;
        op.p1 = op.p2 = Pt{0, 0};
        if (only_one(c1)) {
            op.kind = 0;                                                     // a stale bounding box is a superset: good enough
            if (op.rw > 0 && op.rh > 0) ops.push_back(op);
            st[(size_t)c1] = st[(size_t)c2] == ST_FIRST ? ST_SECOND : ST_FIRST;
        } else {
            if (cut_[(size_t)c1]) {                                          // tips of a component a seam of this round has already cut
                if (no_resume) { unsupported = true; return; }
                blocked = true;                                              // the same conflict is found again by plan_resume()
                return;
            }
            op.kind = 1;
            if (get_seam_tips(c1, c2, &op.p1, &op.p2)) ops.push_back(op);
            st[(size_t)c1] = st[(size_t)c2] == ST_FIRST ? (ST_INTERS | ST_SECOND) : (ST_INTERS | ST_FIRST);
        }
        cut_[(size_t)c1] = 1;
        ed.erase({c1, c2});
        ed.erase({c2, c1});
    }
}

// Everything a pair computes after labelling is a function of the contour records of its INTERS components (position, label,
// neighbour labels), the states of the labels they mention, and -- through the seam tips -- the contour-proximity flags of the
// candidate points; the latter are compared through the planned operations themselves.
bool PairRuns::same_structure(const PairRuns& o) const {
    auto state_of = [](const std::vector<int>& st, int label) { return (label >= 1 && label <= (int)st.size()) ? st[(size_t)label - 1] : -1; };
    // INTERS components in label order, records in raster order
    std::vector<const std::vector<ContourRec>*> a, b;
    std::vector<int> la, lb;
    for (size_t c = 0; c < contours.size(); ++c) if (!contours[c].empty()) { a.push_back(&contours[c]); la.push_back((int)c + 1); }
    for (size_t c = 0; c < o.contours.size(); ++c) if (!o.contours[c].empty()) { b.push_back(&o.contours[c]); lb.push_back((int)c + 1); }
    if (la != lb) return false;
    for (size_t k = 0; k < a.size(); ++k) {
        const auto& ra = *a[k];
        const auto& rb = *b[k];
        if (ra.size() != rb.size()) return false;
        for (size_t i = 0; i < ra.size(); ++i) {
            const ContourRec& x = ra[i];
            const ContourRec& y = rb[i];
            if (x.x != y.x || x.y != y.y || x.label != y.label || x.nl[0] != y.nl[0] || x.nl[1] != y.nl[1] || x.nl[2] != y.nl[2] || x.nl[3] != y.nl[3]) return false;
            if (state_of(states, x.label) != state_of(o.states, x.label)) return false;
            for (int q = 0; q < 4; ++q)
                if (x.nl[q] > 0 && state_of(states, x.nl[q]) != state_of(o.states, x.nl[q])) return false;
        }
    }
    if (unsupported != o.unsupported || ops.size() != o.ops.size()) return false;
    for (size_t k = 0; k < ops.size(); ++k) {
        const SeamOp& x = ops[k];
        const SeamOp& y = o.ops[k];
        if (x.kind != y.kind || x.c1 != y.c1 || x.c2 != y.c2 || x.p1.x != y.p1.x || x.p1.y != y.p1.y || x.p2.x != y.p2.x || x.p2.y != y.p2.y ||
            x.rx != y.rx || x.ry != y.ry || x.rw != y.rw || x.rh != y.rh) return false;
    }
    return true;
}

// ---- updateLabelsUsingSeam ([SEAM]:960-1093) in the run domain ---------------------------------------------------------------
// The reference paints the contour of comp1 and the seam into a mask of comp1's bounding box, flood-fills what is left of the
// component (4-connected) and lets every painted pixel join the region of one of its neighbours; the regions that mostly touch
// comp2 are then relabelled.  Here the bounding box is never rasterised: a row of comp1 is a few runs, the painted pixels cut
// them into interior runs, a union-find over vertically overlapping runs is the flood fill, and "the value of pixel (x, y)" is
// a search in its row.
struct Interval { int y, x0, x1; };                   // pixels [x0, x1) of frame row y

struct UlsRuns {
    // inputs
    const PairRuns* P = nullptr;
    SeamOp op{};
    bool horizontal = false;
    int s0 = 0, nseam = 0;
    const int* lane = nullptr;                        // lane of the seam at step s0 + i (bbox coordinates)
    // output: the pixels of comp1 that take comp2's label, sorted by (y, x0), disjoint
    std::vector<Interval> flips;
    bool too_many_regions = false;                    // 255 or more flood-filled regions: the reference's mask value 255 collides with an id

    void run();
};

void UlsRuns::run() {
    const bool tim = getenv("IS_DEBUG_PLAN_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lapt = [&](const char* w) { if (!tim) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "   [uls] %-16s %.3f ms\n", w, std::chrono::duration<double, std::milli>(n - T0).count()); T0 = n; };
    const PairRuns& R = *P;
    const int l1 = op.c1 + 1, l2 = op.c2 + 1;
    const int rx = op.rx, ry = op.ry, rw = op.rw, rh = op.rh;
    const std::vector<ContourRec>& cont = R.contours[(size_t)op.c1];
    const int nc = (int)cont.size();
    flips.clear();
    too_many_regions = false;
    // contour records per bbox row
    std::vector<int> row_first((size_t)rh + 1, 0);
    for (int i = 0; i < nc; ++i) row_first[(size_t)(cont[(size_t)i].y - ry) + 1]++;
    for (int y = 0; y < rh; ++y) row_first[(size_t)y + 1] += row_first[(size_t)y];
    auto find_contour = [&](int x, int y) -> int {                   // bbox coordinates -> record index, -1 if none
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
    // seam pixels per bbox row (a vertical seam has one per row of its steps, a horizontal seam any number)
    std::vector<int> seam_first((size_t)rh + 1, 0), seam_x((size_t)nseam);
    auto seam_px = [&](int i, int* x, int* y) { const int step = s0 + i; if (horizontal) { *x = step; *y = lane[i]; } else { *x = lane[i]; *y = step; } };
    for (int i = 0; i < nseam; ++i) { int x, y; seam_px(i, &x, &y); seam_first[(size_t)y + 1]++; }
    for (int y = 0; y < rh; ++y) seam_first[(size_t)y + 1] += seam_first[(size_t)y];
    {
        std::vector<int> fill(seam_first.begin(), seam_first.end() - 1);
        for (int i = 0; i < nseam; ++i) { int x, y; seam_px(i, &x, &y); seam_x[(size_t)fill[(size_t)y]++] = x; }   // increasing i = increasing x within a row of a horizontal seam
    }
    auto is_seam = [&](int x, int y) -> bool {
        if ((unsigned)y >= (unsigned)rh) return false;
        for (int k = seam_first[(size_t)y]; k < seam_first[(size_t)y + 1]; ++k) if (seam_x[(size_t)k] == x) return true;
        return false;
    };
    // interior runs of every bbox row: comp1's runs minus the painted pixels
    struct IRun { int x0, x1, parent; };
    std::vector<IRun> ir;
    std::vector<int> ir_first((size_t)rh + 1, 0);
    std::vector<int> painted;                                         // painted x of the current row, sorted
    ir.reserve((size_t)rh * 3);
    for (int y = 0; y < rh; ++y) {
        const int fy = ry + y;
        painted.clear();
        {
            int a = row_first[(size_t)y], ae = row_first[(size_t)y + 1], b = seam_first[(size_t)y], be = seam_first[(size_t)y + 1];
            // seam xs of a row are increasing for a horizontal seam; a vertical seam has at most one: merge two sorted lists
            while (a < ae || b < be) {
                const int xa = a < ae ? cont[(size_t)a].x - rx : INT_MAX, xb = b < be ? seam_x[(size_t)b] : INT_MAX;
                const int x = std::min(xa, xb);
                painted.push_back(x);
                if (xa == x) ++a;
                if (xb == x) ++b;
            }
        }
        size_t pi = 0;
        const int rb = R.row_off[(size_t)fy], re = R.row_off[(size_t)fy + 1];
        for (int k = rb; k < re; ++k) {
            if (R.cp_label[(size_t)k] != l1) continue;
            int a = R.cps[(size_t)k].x - rx, b = (k + 1 < re ? R.cps[(size_t)k + 1].x : R.uw) - rx;
            a = std::max(a, 0); b = std::min(b, rw);
            int x = a;
            while (pi < painted.size() && painted[pi] < a) ++pi;
            while (x < b) {
                const int nextp = pi < painted.size() && painted[pi] < b ? painted[pi] : b;
                if (nextp > x) ir.push_back(IRun{x, nextp, (int)ir.size()});
                x = nextp + 1;
                if (nextp < b) ++pi;
            }
        }
        ir_first[(size_t)y + 1] = (int)ir.size();
    }
    lapt("interior runs");
    auto find = [&](int k) { while (ir[(size_t)k].parent != k) { ir[(size_t)k].parent = ir[(size_t)ir[(size_t)k].parent].parent; k = ir[(size_t)k].parent; } return k; };
    for (int y = 1; y < rh; ++y) {
        int a = ir_first[(size_t)y - 1], ae = ir_first[(size_t)y], b = ir_first[(size_t)y], be = ir_first[(size_t)y + 1];
        while (a < ae && b < be) {
            if (ir[(size_t)a].x0 < ir[(size_t)b].x1 && ir[(size_t)b].x0 < ir[(size_t)a].x1) {
                const int ra = find(a), rb2 = find(b);
                if (ra != rb2) { if (ra < rb2) ir[(size_t)rb2].parent = ra; else ir[(size_t)ra].parent = rb2; }
            }
            if (ir[(size_t)a].x1 <= ir[(size_t)b].x1) ++a; else ++b;
        }
    }
    lapt("union-find");
    // ids 1.. of the regions in order of first use (only equality, "> 0" and "!= 255" are ever asked of them)
    int nregions = 0;
    for (size_t k = 0; k < ir.size(); ++k) if (find((int)k) == (int)k) ++nregions;
    if (nregions >= 255) { too_many_regions = true; return; }
    std::vector<int> region_id(ir.size(), 0);
    int nsub = 0;
    auto id_of_run = [&](int k) { const int r = find(k); if (!region_id[(size_t)r]) region_id[(size_t)r] = ++nsub; return region_id[(size_t)r]; };
    // value of the reference's mask at bbox pixel (x, y) BEFORE any painted pixel is assigned: -1 outside, 0 not comp1, -255 painted, else region id
    auto value = [&](int x, int y) -> int {
        if ((unsigned)x >= (unsigned)rw || (unsigned)y >= (unsigned)rh) return -1;
        for (int k = ir_first[(size_t)y]; k < ir_first[(size_t)y + 1]; ++k) {
            if (x < ir[(size_t)k].x0) break;
            if (x < ir[(size_t)k].x1) return id_of_run(k);
        }
        if (find_contour(x, y) >= 0 || is_seam(x, y)) return -255;
        return 0;
    };
    static const int ddx[8] = {-1, +1, 0, 0, -1, +1, -1, +1};
    static const int ddy[8] = {0, 0, -1, +1, -1, -1, +1, +1};
    // contour pixels in raster order ([SEAM]:983-1010): the last neighbour, in the reference's order, holding an assigned value wins.
    // A painted neighbour is assigned only if it is a contour pixel earlier in raster order (neighbours 0, 2, 4, 5).
    std::vector<int> val((size_t)nc, 255);
    for (int i = 0; i < nc; ++i) {
        const int x = cont[(size_t)i].x - rx, y = cont[(size_t)i].y - ry;
        int v = 0;
        for (int j = 7; j >= 0; --j) {
            const int g = value(x + ddx[j], y + ddy[j]);
            if (g > 0) { v = g; break; }
            if (g == -255 && (j == 0 || j == 2 || j == 4 || j == 5)) {
                const int k = find_contour(x + ddx[j], y + ddy[j]);
                if (k >= 0 && k < i && val[(size_t)k] > 0 && val[(size_t)k] != 255) { v = val[(size_t)k]; break; }
            }
        }
        val[(size_t)i] = v;
    }
    lapt("contour walk");
    // seam pixels ([SEAM]:1012-1034): each takes the value of one neighbour of its own step
    std::vector<int> sval((size_t)nseam, 0), seam_contour((size_t)nseam, -1);
    for (int i = 0; i < nseam; ++i) {
        int x, y;
        seam_px(i, &x, &y);
        const int nx = horizontal ? x : x + 1, ny = horizontal ? y + 1 : y;
        const int g = value(nx, ny);
        int v = 0;
        if (g > 0) v = g;
        else if (g == -255) {
            const int k = find_contour(nx, ny);
            if (k >= 0 && val[(size_t)k] > 0 && val[(size_t)k] != 255) v = val[(size_t)k];
        }
        sval[(size_t)i] = v;
        seam_contour[(size_t)i] = find_contour(x, y);
    }
    for (int i = 0; i < nseam; ++i)
        if (seam_contour[(size_t)i] >= 0) val[(size_t)seam_contour[(size_t)i]] = sval[(size_t)i];
    lapt("seam walk");
    // adjacency vote ([SEAM]:1039-1085)
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
    std::vector<char> isAdj((size_t)nsub + 1, 0);
    const double len = (double)nc;
    for (int k = c2_has0 ? 0 : 1; k <= nsub; ++k) {
        if (connect2[(size_t)k] / len > 0.05) {
            const bool sub_exists = k >= 1 || co_has0;
            if (sub_exists && (connectOther[(size_t)k] / len < 0.1)) isAdj[(size_t)k] = 1;
        }
    }
    lapt("vote");
    // relabel ([SEAM]:1089-1092): the regions voted adjacent and the painted pixels that joined them, as intervals per frame row
    // per row: three lists sorted by x (interior runs, contour pixels, seam-only pixels) -> merged, touching intervals joined
    std::vector<int> seam_idx((size_t)nseam);                         // seam pixels in the order of seam_x (per row)
    {
        std::vector<int> fill(seam_first.begin(), seam_first.end() - 1);
        for (int i = 0; i < nseam; ++i) { int x, y; seam_px(i, &x, &y); seam_idx[(size_t)fill[(size_t)y]++] = i; }
    }
    auto adj = [&](int v) { return v > 0 && v <= nsub && isAdj[(size_t)v]; };
    std::vector<std::pair<int, int>> row;                             // (x0, x1) of the current row
    for (int y = 0; y < rh; ++y) {
        row.clear();
        for (int k = ir_first[(size_t)y]; k < ir_first[(size_t)y + 1]; ++k)
            if (adj(region_id[(size_t)find(k)])) row.push_back({ir[(size_t)k].x0, ir[(size_t)k].x1});
        const size_t n_runs = row.size();
        for (int i = row_first[(size_t)y]; i < row_first[(size_t)y + 1]; ++i)
            if (adj(val[(size_t)i])) row.push_back({cont[(size_t)i].x - rx, cont[(size_t)i].x - rx + 1});
        const size_t n_cont = row.size();
        for (int k = seam_first[(size_t)y]; k < seam_first[(size_t)y + 1]; ++k) {
            const int i = seam_idx[(size_t)k];
            if (seam_contour[(size_t)i] < 0 && adj(sval[(size_t)i])) row.push_back({seam_x[(size_t)k], seam_x[(size_t)k] + 1});   // contour pixels are covered above
        }
        if (row.empty()) continue;
        // three sorted lists of a handful of intervals: one sort (insertion sort at this size, and unlike std::inplace_merge it
        // does not ask the allocator for a scratch buffer twice per row)
        if (row.size() > n_runs || n_cont > n_runs) std::sort(row.begin(), row.end());
        (void)n_cont;
        for (auto& iv : row) {
            if (!flips.empty() && flips.back().y == ry + y && flips.back().x1 >= iv.first + rx) flips.back().x1 = std::max(flips.back().x1, iv.second + rx);
            else flips.push_back(Interval{ry + y, iv.first + rx, iv.second + rx});
        }
    }
    lapt("flips");
}

// The pair's final mask update ([SEAM]:524-545) as clear intervals: for every pixel of both masks, the state of its FINAL label
// decides which mask loses it -- bit 1: the first image's mask, bit 2: the second's (mask2 is updated first and mask1 then looks
// at the updated mask2, so a pixel is never cleared in both).  `seam_flips[k]`: UlsRuns::flips of the k-th kind-1 operation of
// the plan that succeeded (nullptr when estimateSeam failed or was not run).
struct ClearIv { int y, x0, x1, bits; };              // frame coordinates

// ---- staged plan: the relabels of a round written into the runs -----------------------------------------------------------------
bool PairRuns::relabel(int l_from, int l_to, const std::vector<Interval>* fl) {
    if (!fl) {                                                            // wholesale relabel ([SEAM]:440-450)
        for (int& l : cp_label) if (l == l_from) l = l_to;
        return true;
    }
    if (fl->empty()) return true;
    std::vector<ChangePt> ncps;
    std::vector<int> nlab, nro((size_t)uh + 1, 0);
    ncps.reserve(cps.size() + 2 * fl->size());
    nlab.reserve(cps.size() + 2 * fl->size());
    size_t fi = 0;
    for (int y = 0; y < uh; ++y) {
        const int rb = row_off[(size_t)y], re = row_off[(size_t)y + 1];
        while (fi < fl->size() && (*fl)[fi].y < y) ++fi;
        size_t fe = fi;
        while (fe < fl->size() && (*fl)[fe].y == y) ++fe;
        const size_t row_begin = ncps.size();
        if (fi == fe) {
            ncps.insert(ncps.end(), cps.begin() + rb, cps.begin() + re);
            nlab.insert(nlab.end(), cp_label.begin() + rb, cp_label.begin() + re);
        } else {
            size_t f = fi;                                                // flips of a row are sorted by x0 and disjoint
            for (int k = rb; k < re; ++k) {
                const int a = cps[(size_t)k].x, b = k + 1 < re ? cps[(size_t)k + 1].x : uw;
                const int lab = cp_label[(size_t)k], cls = cps[(size_t)k].cls;
                if (lab != l_from) { ncps.push_back(ChangePt{a, cls}); nlab.push_back(lab); continue; }
                int x = a;
                while (f < fe && (*fl)[f].x1 <= a) ++f;
                for (size_t g = f; g < fe && (*fl)[g].x0 < b; ++g) {
                    const int s0 = std::max((*fl)[g].x0, a), e0 = std::min((*fl)[g].x1, b);
                    if (s0 >= e0) continue;
                    if (s0 > x) { ncps.push_back(ChangePt{x, cls}); nlab.push_back(l_from); }
                    ncps.push_back(ChangePt{s0, cls}); nlab.push_back(l_to);
                    x = e0;
                }
                if (x < b) { ncps.push_back(ChangePt{x, cls}); nlab.push_back(l_from); }
            }
        }
        if ((int)(ncps.size() - row_begin) > ROW_CAP) { too_many_runs = true; return false; }
        nro[(size_t)y + 1] = (int)ncps.size();
        fi = fe;
    }
    cps.swap(ncps);
    cp_label.swap(nlab);
    row_off.swap(nro);
    return true;
}

bool PairRuns::apply_round(const std::vector<const std::vector<Interval>*>& seam_flips) {
    if (!staged) { snap_row_off = row_off; snap_cps = cps; snap_label = cp_label; }
    size_t k1 = 0;
    for (size_t q = round_begin; q < ops.size(); ++q) {
        const SeamOp& op = ops[q];
        if (op.kind == 0) { relabel(op.c1 + 1, op.c2 + 1, nullptr); continue; }
        const std::vector<Interval>* f = k1 < seam_flips.size() ? seam_flips[k1] : nullptr;
        ++k1;
        if (f && !relabel(op.c1 + 1, op.c2 + 1, f)) return false;         // nullptr: estimateSeam failed, the labels stay
    }
    round_begin = ops.size();
    staged = true;
    // what the reference refreshes after an operation ([SEAM]:484-520) for the components the loop can still ask about: bounding
    // boxes and contours of the INTERS components, from the labels as they are now.  The edge set is NOT recomputed (the reference
    // only erases from it).
    tls.assign((size_t)ncomps, Pt{INT_MAX, INT_MAX});
    brs.assign((size_t)ncomps, Pt{INT_MIN, INT_MIN});
    contours.assign((size_t)ncomps, std::vector<ContourRec>());
    contour_rows();
    return true;
}

// the final mask update [SEAM]:524-545 of a staged pair: every pixel of both masks takes the bits of the state its label ends up in
void PairRuns::final_clears(std::vector<ClearIv>* out) const {
    out->clear();
    const int y_lo = std::max(0, iTl.y - unionTl.y), y_hi = std::min(uh, iBr.y - unionTl.y);
    for (int y = y_lo; y < y_hi; ++y) {
        const int rb = row_off[(size_t)y], re = row_off[(size_t)y + 1];
        for (int k = rb; k < re; ++k) {
            if (cps[(size_t)k].cls != 3) continue;
            const int l = cp_label[(size_t)k];
            if (l <= 0) continue;
            const int st = final_states[(size_t)l - 1];
            const int bits = (st & ST_FIRST) ? 2 : ((st & ST_SECOND) ? 1 : 0);
            const int a = cps[(size_t)k].x, b = k + 1 < re ? cps[(size_t)k + 1].x : uw;
            if (!bits || a >= b) continue;
            if (!out->empty() && out->back().y == y && out->back().x1 == a && out->back().bits == bits) out->back().x1 = b;
            else out->push_back(ClearIv{y, a, b, bits});
        }
    }
}

static bool runs_minus_clears(const MaskRuns& in, const std::vector<RunLayer>& layers, int cap, MaskRuns* out) {
    *out = in;
    if (out->slots < cap) return false;
    const int cols = in.cols, slots = out->slots;
    for (const RunLayer& L : layers) {
        const std::vector<ClearIv>& cl = *L.clears;
        size_t i = 0;
        while (i < cl.size()) {
            const int fy = cl[i].y;
            size_t j = i;
            while (j < cl.size() && cl[j].y == fy) ++j;
            const int my = fy + L.dy;
            if (my >= 0 && my < in.rows) {
                unsigned short* r = out->xs.data() + (size_t)my * slots;
                int n = out->counts[(size_t)my];
                if (n > slots) return false;
                // intervals of the row
                int iv[16][2], ni = 0;
                for (int k = 0; k < n; k += 2) { iv[ni][0] = r[k]; iv[ni][1] = k + 1 < n ? r[k + 1] : cols; ++ni; }
                bool changed = false;
                for (size_t q = i; q < j; ++q) {
                    if (!(cl[q].bits & L.bit)) continue;
                    const int a = std::max(cl[q].x0 + L.dx, 0), b = std::min(cl[q].x1 + L.dx, cols);
                    if (a >= b) continue;
                    for (int k = 0; k < ni; ++k) {
                        const int s0 = iv[k][0], s1 = iv[k][1];
                        if (b <= s0 || a >= s1) continue;
                        changed = true;
                        if (a <= s0 && b >= s1) { for (int m = k; m + 1 < ni; ++m) { iv[m][0] = iv[m + 1][0]; iv[m][1] = iv[m + 1][1]; } --ni; --k; }
                        else if (a <= s0) iv[k][0] = b;
                        else if (b >= s1) iv[k][1] = a;
                        else {                                                     // a hole: the interval splits
                            if (ni >= 15) return false;
                            for (int m = ni; m > k + 1; --m) { iv[m][0] = iv[m - 1][0]; iv[m][1] = iv[m - 1][1]; }
                            iv[k + 1][0] = b; iv[k + 1][1] = s1; iv[k][1] = a;
                            ++ni; ++k;
                        }
                    }
                }
                if (changed) {
                    n = 0;
                    for (int k = 0; k < ni; ++k) {
                        if (n < slots) r[n] = (unsigned short)iv[k][0];
                        ++n;
                        if (iv[k][1] < cols) { if (n < slots) r[n] = (unsigned short)iv[k][1]; ++n; }
                    }
                    if (n > cap) return false;
                    out->counts[(size_t)my] = (unsigned char)n;
                }
            }
            i = j;
        }
    }
    return true;
}

static void pair_clear_intervals(const PairRuns& R, const std::vector<const std::vector<Interval>*>& seam_flips, std::vector<ClearIv>* out) {
    out->clear();
    auto bits_of = [&](int label) {
        const int st = R.final_states[(size_t)label - 1];
        if (st & ST_FIRST) return 2;
        if (st & ST_SECOND) return 1;
        return 0;
    };
    // per INTERS component: label the pixels a seam did not move end up with (a later wholesale relabel), and its seam flips
    std::vector<int> rest_label((size_t)R.ncomps, 0);
    std::vector<const std::vector<Interval>*> flips_of((size_t)R.ncomps, nullptr);
    std::vector<int> flip_label((size_t)R.ncomps, 0);
    for (int c = 0; c < R.ncomps; ++c) rest_label[(size_t)c] = c + 1;
    size_t k1 = 0;
    for (const SeamOp& op : R.ops) {
        if (op.kind == 0) rest_label[(size_t)op.c1] = op.c2 + 1;
        else { flips_of[(size_t)op.c1] = seam_flips[k1++]; flip_label[(size_t)op.c1] = op.c2 + 1; }
    }
    std::vector<size_t> cursor((size_t)R.ncomps, 0);                  // position in each component's flip list (sorted by row)
    const int y_lo = std::max(0, R.iTl.y - R.unionTl.y), y_hi = std::min(R.uh, R.iBr.y - R.unionTl.y);
    auto emit = [&](int y, int x0, int x1, int bits) {
        if (!bits || x0 >= x1) return;
        if (!out->empty() && out->back().y == y && out->back().x1 == x0 && out->back().bits == bits) out->back().x1 = x1;
        else out->push_back(ClearIv{y, x0, x1, bits});
    };
    for (int y = y_lo; y < y_hi; ++y) {
        const int rb = R.row_off[(size_t)y], re = R.row_off[(size_t)y + 1];
        for (int k = rb; k < re; ++k) {
            if (R.cps[(size_t)k].cls != 3) continue;
            const int c = R.cp_label[(size_t)k] - 1;
            const int a = R.cps[(size_t)k].x, b = k + 1 < re ? R.cps[(size_t)k + 1].x : R.uw;
            const int rest_bits = bits_of(rest_label[(size_t)c]);
            const std::vector<Interval>* fl = flips_of[(size_t)c];
            int x = a;
            if (fl) {
                size_t& cu = cursor[(size_t)c];
                while (cu < fl->size() && ((*fl)[cu].y < y || ((*fl)[cu].y == y && (*fl)[cu].x1 <= a))) ++cu;
                const int fbits = bits_of(flip_label[(size_t)c]);
                size_t q = cu;
                while (q < fl->size() && (*fl)[q].y == y && (*fl)[q].x0 < b) {
                    const int f0 = std::max((*fl)[q].x0, a), f1 = std::min((*fl)[q].x1, b);
                    emit(y, x, f0, rest_bits);
                    emit(y, f0, f1, fbits);
                    x = f1;
                    ++q;
                }
            }
            emit(y, x, b, rest_bits);
        }
    }
}
