// oracle/seam.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the reference's dynamic-programming seam finder ([SEAM]:87-1093, itself a
// free-function copy of cv::detail::DpSeamFinder).  Function-by-function citations below.
// COLOR_GRAD ([SEAM]:549-572, :769-771, :794-796) is restated with cvtColor(BGR2GRAY) + Sobel(3x3, CV_32F) written out:
// OpenCV's float results for these two depend on the SIMD dispatch (FMA or not, vector body or scalar tail), so the
// gradients are pinned against cv2 to a few ulp, the seam masks exactly (tests/test_oracle_cv2.py).
//
// Parity pin: cv2.detail_DpSeamFinder (tests/golden/seam_blend_cases.npz, color_grad_cases.npz, tests/test_oracle_cv2.py)
// and the reference's own find() compiled into oracle/_ref/libref_seam.so (tests/test_oracle_reference_build.py,
// tests/golden/seam_ref_cases.npz): seam masks bit for bit.
//
// Compile with -ffp-contract=off.
#include "oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace {

struct Pt { int x, y; };
inline bool operator==(Pt a, Pt b) { return a.x == b.x && a.y == b.y; }

enum { FIRST = 1, SECOND = 2, INTERS = 4, INTERS_FIRST = 5, INTERS_SECOND = 6 };   // [SEAM]:73-78

struct Img {
    const void* data; int rows, cols; bool u8;
    int cn = 3;                       // 3, or 4 (CV_8UC4 / CV_32FC4: the fourth channel is skipped, [SEAM]:722-730 diffL2Square4)
    inline float at(int y, int x, int c) const {
        size_t i = ((size_t)y * cols + x) * cn + c;
        return u8 ? (float)((const uint8_t*)data)[i] : ((const float*)data)[i];
    }
};

template <typename T> struct Grid {
    int rows = 0, cols = 0;
    std::vector<T> v;
    void create(int r, int c, T init = T()) { rows = r; cols = c; v.assign((size_t)r * c, init); }
    inline T& operator()(int y, int x) { return v[(size_t)y * cols + x]; }
    inline const T& operator()(int y, int x) const { return v[(size_t)y * cols + x]; }
};

// cv::floodFill on a CV_32S image, 4-connectivity, zero tolerance: fills the pixels connected to the
// seed whose value equals the seed's value.
void floodFill(Grid<int>& g, int sx, int sy, int newVal) {
    const int old = g(sy, sx);
    if (old == newVal) return;
    std::vector<Pt> stack;
    stack.push_back({sx, sy});
    g(sy, sx) = newVal;
    while (!stack.empty()) {
        Pt p = stack.back();
        stack.pop_back();
        static const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
        for (int k = 0; k < 4; ++k) {
            int x = p.x + dx[k], y = p.y + dy[k];
            if (x >= 0 && x < g.cols && y >= 0 && y < g.rows && g(y, x) == old) {
                g(y, x) = newVal;
                stack.push_back({x, y});
            }
        }
    }
}

// cv::partition with the ClosePoints(minDist) predicate ([SEAM]:50-63,638): connected components of
// the "dist^2 < minDist^2" graph, classes numbered by first member index.
int partitionClose(const std::vector<Pt>& pts, int minDist, std::vector<int>& labels) {
    const int n = (int)pts.size();
    std::vector<int> parent(n);
    for (int i = 0; i < n; ++i) parent[i] = i;
    auto find = [&](int i) {
        while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; }
        return i;
    };
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            int ddx = pts[i].x - pts[j].x, ddy = pts[i].y - pts[j].y;
            if (ddx * ddx + ddy * ddy < minDist * minDist) {
                int a = find(i), b = find(j);
                if (a != b) parent[b] = a;
            }
        }
    labels.assign(n, -1);
    std::vector<int> rootClass(n, -1);
    int ncls = 0;
    for (int i = 0; i < n; ++i) {
        int r = find(i);
        if (rootClass[r] < 0) rootClass[r] = ncls++;
        labels[i] = rootClass[r];
    }
    return ncls;
}

inline double cvRoundD(double v) { return (double)lrint(v); }

struct DpSeam {
    int costFunc = ORC_COST_COLOR;
    Pt unionTl{}, unionBr{};
    int uw = 0, uh = 0;
    Grid<uint8_t> mask1_, mask2_, contour1mask_, contour2mask_;
    int ncomps_ = 0;
    Grid<int> labels_;
    std::vector<int> states_;
    std::vector<Pt> tls_, brs_;
    std::vector<std::vector<Pt>> contours_;
    std::set<std::pair<int, int>> edges_;

    // trace
    int32_t* trace = nullptr; size_t trace_cap = 0; size_t trace_len = 0;
    int cur_i = 0, cur_j = 0;

    inline int label(int y, int x) const {   // out-of-frame reads count as "no label" 
        if (x < 0 || y < 0 || x >= uw || y >= uh) return -1;
        return labels_(y, x);
    }

    // [SEAM]:127-193
    void process(const Img& image1, const Img& image2, Pt tl1, Pt tl2, uint8_t* mask1, uint8_t* mask2) {
        Pt iTl{std::max(tl1.x, tl2.x), std::max(tl1.y, tl2.y)};
        Pt iBr{std::min(tl1.x + image1.cols, tl2.x + image2.cols), std::min(tl1.y + image1.rows, tl2.y + image2.rows)};
        if (iTl.x >= iBr.x || iTl.y >= iBr.y) return;   // :142-143
        unionTl = {std::min(tl1.x, tl2.x), std::min(tl1.y, tl2.y)};
        unionBr = {std::max(tl1.x + image1.cols, tl2.x + image2.cols), std::max(tl1.y + image1.rows, tl2.y + image2.rows)};
        uw = unionBr.x - unionTl.x;
        uh = unionBr.y - unionTl.y;
        mask1_.create(uh, uw, 0);
        mask2_.create(uh, uw, 0);
        for (int y = 0; y < image1.rows; ++y)            // :157-158
            std::memcpy(&mask1_(tl1.y - unionTl.y + y, tl1.x - unionTl.x), mask1 + (size_t)y * image1.cols, image1.cols);
        for (int y = 0; y < image2.rows; ++y)            // :160-161
            std::memcpy(&mask2_(tl2.y - unionTl.y + y, tl2.x - unionTl.x), mask2 + (size_t)y * image2.cols, image2.cols);
        contour1mask_.create(uh, uw, 0);
        contour2mask_.create(uh, uw, 0);
        for (int y = 0; y < uh; ++y)                     // :168-186
            for (int x = 0; x < uw; ++x) {
                if (mask1_(y, x) &&
                    ((x == 0 || !mask1_(y, x - 1)) || (x == uw - 1 || !mask1_(y, x + 1)) ||
                     (y == 0 || !mask1_(y - 1, x)) || (y == uh - 1 || !mask1_(y + 1, x))))
                    contour1mask_(y, x) = 255;
                if (mask2_(y, x) &&
                    ((x == 0 || !mask2_(y, x - 1)) || (x == uw - 1 || !mask2_(y, x + 1)) ||
                     (y == 0 || !mask2_(y - 1, x)) || (y == uh - 1 || !mask2_(y + 1, x))))
                    contour2mask_(y, x) = 255;
            }
        findComponents();
        findEdges();
        resolveConflicts(image1, image2, tl1, tl2, mask1, mask2);
    }

    // [SEAM]:196-308
    void findComponents() {
        ncomps_ = 0;
        labels_.create(uh, uw, 0);
        states_.clear(); tls_.clear(); brs_.clear(); contours_.clear();
        for (int y = 0; y < uh; ++y)
            for (int x = 0; x < uw; ++x) {
                if (mask1_(y, x) && mask2_(y, x)) labels_(y, x) = INT_MAX;
                else if (mask1_(y, x)) labels_(y, x) = INT_MAX - 1;
                else if (mask2_(y, x)) labels_(y, x) = INT_MAX - 2;
                else labels_(y, x) = 0;
            }
        for (int y = 0; y < uh; ++y)
            for (int x = 0; x < uw; ++x) {
                if (labels_(y, x) >= INT_MAX - 2) {
                    if (labels_(y, x) == INT_MAX) states_.push_back(INTERS);
                    else if (labels_(y, x) == INT_MAX - 1) states_.push_back(FIRST);
                    else states_.push_back(SECOND);
                    floodFill(labels_, x, y, ++ncomps_);
                    tls_.push_back({x, y});
                    brs_.push_back({x + 1, y + 1});
                    contours_.push_back(std::vector<Pt>());
                }
                if (labels_(y, x)) {
                    int l = labels_(y, x);
                    int ci = l - 1;
                    tls_[ci].x = std::min(tls_[ci].x, x);
                    tls_[ci].y = std::min(tls_[ci].y, y);
                    brs_[ci].x = std::max(brs_[ci].x, x + 1);
                    brs_[ci].y = std::max(brs_[ci].y, y + 1);
                    if ((x == 0 || labels_(y, x - 1) != l) || (x == uw - 1 || labels_(y, x + 1) != l) ||
                        (y == 0 || labels_(y - 1, x) != l) || (y == uh - 1 || labels_(y + 1, x) != l))
                        contours_[ci].push_back({x, y});
                }
            }
    }

    // [SEAM]:311-392
    void findEdges() {
        std::map<std::pair<int, int>, int> wedges;
        for (int ci = 0; ci < ncomps_ - 1; ++ci)
            for (int cj = ci + 1; cj < ncomps_; ++cj) {
                wedges[{ci, cj}] = 0;
                wedges[{cj, ci}] = 0;
            }
        for (int ci = 0; ci < ncomps_; ++ci)
            for (size_t i = 0; i < contours_[ci].size(); ++i) {
                int x = contours_[ci][i].x, y = contours_[ci][i].y, l = ci + 1;
                if (x > 0 && labels_(y, x - 1) && labels_(y, x - 1) != l) {
                    wedges[{ci, labels_(y, x - 1) - 1}]++;
                    wedges[{labels_(y, x - 1) - 1, ci}]++;
                }
                if (y > 0 && labels_(y - 1, x) && labels_(y - 1, x) != l) {
                    wedges[{ci, labels_(y - 1, x) - 1}]++;
                    wedges[{labels_(y - 1, x) - 1, ci}]++;
                }
                if (x < uw - 1 && labels_(y, x + 1) && labels_(y, x + 1) != l) {
                    wedges[{ci, labels_(y, x + 1) - 1}]++;
                    wedges[{labels_(y, x + 1) - 1, ci}]++;
                }
                if (y < uh - 1 && labels_(y + 1, x) && labels_(y + 1, x) != l) {
                    wedges[{ci, labels_(y + 1, x) - 1}]++;
                    wedges[{labels_(y + 1, x) - 1, ci}]++;
                }
            }
        edges_.clear();
        for (int ci = 0; ci < ncomps_ - 1; ++ci)
            for (int cj = ci + 1; cj < ncomps_; ++cj) {
                auto itr = wedges.find({ci, cj});
                if (itr != wedges.end() && itr->second > 0) edges_.insert(itr->first);
                itr = wedges.find({cj, ci});
                if (itr != wedges.end() && itr->second > 0) edges_.insert(itr->first);
            }
    }

    // [SEAM]:575-581
    bool hasOnlyOneNeighbor(int comp) {
        auto begin = edges_.lower_bound({comp, INT_MIN});
        auto end = edges_.upper_bound({comp, INT_MAX});
        return ++begin == end;
    }

    // [SEAM]:584-604
    bool closeToContour(int y, int x, const Grid<uint8_t>& cm) {
        const int rad = 2;
        for (int dy = -rad; dy <= rad; ++dy)
            if (y + dy >= 0 && y + dy < uh)
                for (int dx = -rad; dx <= rad; ++dx)
                    if (x + dx >= 0 && x + dx < uw && cm(y + dy, x + dx)) return true;
        return false;
    }

    // [SEAM]:607-706
    bool getSeamTips(int comp1, int comp2, Pt& p1, Pt& p2) {
        std::vector<Pt> special;
        int l2 = comp2 + 1;
        for (size_t i = 0; i < contours_[comp1].size(); ++i) {
            int x = contours_[comp1][i].x, y = contours_[comp1][i].y;
            if (closeToContour(y, x, contour1mask_) && closeToContour(y, x, contour2mask_) &&
                ((x > 0 && labels_(y, x - 1) == l2) || (y > 0 && labels_(y - 1, x) == l2) ||
                 (x < uw - 1 && labels_(y, x + 1) == l2) || (y < uh - 1 && labels_(y + 1, x) == l2)))
                special.push_back({x, y});
        }
        if (special.size() < 2) return false;
        std::vector<int> labels;
        int nlabels = partitionClose(special, 10, labels);
        if (nlabels < 2) return false;
        std::vector<Pt> sum(nlabels, Pt{0, 0});
        std::vector<std::vector<Pt>> points(nlabels);
        for (size_t i = 0; i < special.size(); ++i) {
            sum[labels[i]].x += special[i].x;
            sum[labels[i]].y += special[i].y;
            points[labels[i]].push_back(special[i]);
        }
        int idx[2] = {-1, -1};
        double maxDist = -std::numeric_limits<double>::max();
        for (int i = 0; i < nlabels - 1; ++i)
            for (int j = i + 1; j < nlabels; ++j) {
                double size1 = (double)points[i].size(), size2 = (double)points[j].size();
                double cx1 = cvRoundD(sum[i].x / size1), cy1 = cvRoundD(sum[i].y / size1);
                double cx2 = cvRoundD(sum[j].x / size2), cy2 = cvRoundD(sum[j].y / size2);
                double dist = (cx1 - cx2) * (cx1 - cx2) + (cy1 - cy2) * (cy1 - cy2);
                if (dist > maxDist) { maxDist = dist; idx[0] = i; idx[1] = j; }
            }
        Pt p[2];
        for (int i = 0; i < 2; ++i) {
            double size = (double)points[idx[i]].size();
            double cx = cvRoundD(sum[idx[i]].x / size), cy = cvRoundD(sum[idx[i]].y / size);
            size_t closest = points[idx[i]].size();
            double minDist = std::numeric_limits<double>::max();
            for (size_t j = 0; j < points[idx[i]].size(); ++j) {
                double dist = (points[idx[i]][j].x - cx) * (points[idx[i]][j].x - cx) +
                              (points[idx[i]][j].y - cy) * (points[idx[i]][j].y - cy);
                if (dist < minDist) { minDist = dist; closest = j; }
            }
            p[i] = points[idx[i]][closest];
        }
        p1 = p[0];
        p2 = p[1];
        return true;
    }

    // [SEAM]:713-730 diffL2Square3<T> / diffL2Square4<T> (pixel stride 4, the same three channels); float: ((d0^2 + d1^2) + d2^2)
    // with d in float. uchar: the
    // same sum in int then converted -- identical values for 8-bit data.
    static inline float diff3(const Img& a, int y1, int x1, const Img& b, int y2, int x2) {
        if (a.u8) {
            const uint8_t* r1 = (const uint8_t*)a.data + ((size_t)y1 * a.cols + x1) * a.cn;
            const uint8_t* r2 = (const uint8_t*)b.data + ((size_t)y2 * b.cols + x2) * b.cn;
            int d0 = r1[0] - r2[0], d1 = r1[1] - r2[1], d2 = r1[2] - r2[2];
            return static_cast<float>(d0 * d0 + d1 * d1 + d2 * d2);
        }
        const float* r1 = (const float*)a.data + ((size_t)y1 * a.cols + x1) * a.cn;
        const float* r2 = (const float*)b.data + ((size_t)y2 * b.cols + x2) * b.cn;
        float d0 = r1[0] - r2[0], d1 = r1[1] - r2[1], d2 = r1[2] - r2[2];
        return d0 * d0 + d1 * d1 + d2 * d2;
    }

    // [SEAM]:549-572 computeGradients: gray = cvtColor(BGR2GRAY) of the CV_32F image (coefficients 0.114/0.587/0.299,
    // association of OpenCV's FMA vector body), gradx/grady = Sobel(gray, CV_32F, 1,0 / 0,1), ksize 3, BORDER_REFLECT_101:
    // separable, row filter first; [1 2 1] is evaluated as (a + c) + 2b, [-1 0 1] as c - a.
    // 8-bit images are taken as their CV_32F conversion (what the mains pass to find(): images_warped_f, [SEAM]:1188-1190).
    Grid<float> gradx1_, grady1_, gradx2_, grady2_;
    static void sobelPair(const Img& im, Grid<float>& gx, Grid<float>& gy) {
        const int H = im.rows, W = im.cols;
        Grid<float> gray, dxr, smr;
        gray.create(H, W);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                gray(y, x) = std::fmaf(im.at(y, x, 2), 0.299f, std::fmaf(im.at(y, x, 0), 0.114f, im.at(y, x, 1) * 0.587f));
        auto refl = [](int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); };
        dxr.create(H, W);
        smr.create(H, W);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float a = gray(y, refl(x - 1, W)), b = gray(y, x), c = gray(y, refl(x + 1, W));
                dxr(y, x) = c - a;
                smr(y, x) = (a + c) + b * 2.f;
            }
        gx.create(H, W);
        gy.create(H, W);
        for (int y = 0; y < H; ++y) {
            int ya = refl(y - 1, H), yc = refl(y + 1, H);
            for (int x = 0; x < W; ++x) {
                gx(y, x) = (dxr(ya, x) + dxr(yc, x)) + dxr(y, x) * 2.f;
                gy(y, x) = smr(yc, x) - smr(ya, x);
            }
        }
    }
    void computeGradients(const Img& image1, const Img& image2) {
        sobelPair(image1, gradx1_, grady1_);
        sobelPair(image2, gradx2_, grady2_);
    }

    // [SEAM]:733-803
    void computeCosts(const Img& image1, const Img& image2, Pt tl1, Pt tl2, int comp,
                      Grid<float>& costV, Grid<float>& costH) {
        int l = comp + 1;
        int rx = tls_[comp].x, ry = tls_[comp].y;
        int rw = brs_[comp].x - rx, rh = brs_[comp].y - ry;
        int dx1 = unionTl.x - tl1.x, dy1 = unionTl.y - tl1.y;
        int dx2 = unionTl.x - tl2.x, dy2 = unionTl.y - tl2.y;
        const float badRegionCost = 195075.f;   // detail::normL2(Point3f(255,255,255), 0): squared norm
        costV.create(rh, rw + 1);
        for (int y = ry; y < ry + rh; ++y)
            for (int x = rx; x < rx + rw + 1; ++x) {
                if (label(y, x) == l && x > 0 && label(y, x - 1) == l) {
                    float costColor = (diff3(image1, y + dy1, x + dx1 - 1, image2, y + dy2, x + dx2) +
                                       diff3(image1, y + dy1, x + dx1, image2, y + dy2, x + dx2 - 1)) / 2;
                    if (costFunc == ORC_COST_COLOR_GRAD) {                                   // :767-772
                        float costGrad = std::fabs(gradx1_(y + dy1, x + dx1)) + std::fabs(gradx1_(y + dy1, x + dx1 - 1)) +
                                         std::fabs(gradx2_(y + dy2, x + dx2)) + std::fabs(gradx2_(y + dy2, x + dx2 - 1)) + 1.f;
                        costColor = costColor / costGrad;
                    }
                    costV(y - ry, x - rx) = costColor;
                } else
                    costV(y - ry, x - rx) = badRegionCost;
            }
        costH.create(rh + 1, rw);
        for (int y = ry; y < ry + rh + 1; ++y)
            for (int x = rx; x < rx + rw; ++x) {
                if (label(y, x) == l && y > 0 && label(y - 1, x) == l) {
                    float costColor = (diff3(image1, y + dy1 - 1, x + dx1, image2, y + dy2, x + dx2) +
                                       diff3(image1, y + dy1, x + dx1, image2, y + dy2 - 1, x + dx2)) / 2;
                    if (costFunc == ORC_COST_COLOR_GRAD) {                                   // :792-797
                        float costGrad = std::fabs(grady1_(y + dy1, x + dx1)) + std::fabs(grady1_(y + dy1 - 1, x + dx1)) +
                                         std::fabs(grady2_(y + dy2, x + dx2)) + std::fabs(grady2_(y + dy2 - 1, x + dx2)) + 1.f;
                        costColor = costColor / costGrad;
                    }
                    costH(y - ry, x - rx) = costColor;
                } else
                    costH(y - ry, x - rx) = badRegionCost;
            }
    }

    // [SEAM]:806-957
    bool estimateSeam(const Img& image1, const Img& image2, Pt tl1, Pt tl2, int comp,
                      Pt p1, Pt p2, std::vector<Pt>& seam, bool& isHorizontal) {
        Grid<float> costV, costH;
        computeCosts(image1, image2, tl1, tl2, comp, costV, costH);
        int rx = tls_[comp].x, ry = tls_[comp].y;
        int rw = brs_[comp].x - rx, rh = brs_[comp].y - ry;
        Pt src{p1.x - rx, p1.y - ry};
        Pt dst{p2.x - rx, p2.y - ry};
        int l = comp + 1;
        bool swapped = false;
        isHorizontal = std::abs(dst.x - src.x) > std::abs(dst.y - src.y);
        if (isHorizontal) {
            if (src.x > dst.x) { std::swap(src, dst); swapped = true; }
        } else if (src.y > dst.y) {
            swapped = true;
            std::swap(src, dst);
        }
        Grid<uint8_t> control, reachable;
        Grid<float> cost;
        control.create(rh, rw, 0);
        reachable.create(rh, rw, 0);
        cost.create(rh, rw, 0.f);
        reachable(src.y, src.x) = 1;
        cost(src.y, src.x) = 0.f;
        int nsteps;
        std::pair<float, int> steps[3];
        if (isHorizontal) {                                      // :859-886
            for (int x = src.x + 1; x <= dst.x; ++x)
                for (int y = 0; y < rh; ++y) {
                    nsteps = 0;
                    if (labels_(y + ry, x + rx) == l) {
                        if (reachable(y, x - 1))
                            steps[nsteps++] = std::make_pair(cost(y, x - 1) + costH(y, x - 1), 1);
                        if (y > 0 && reachable(y - 1, x - 1))
                            steps[nsteps++] = std::make_pair(cost(y - 1, x - 1) + costH(y - 1, x - 1) + costV(y - 1, x), 2);
                        if (y < rh - 1 && reachable(y + 1, x - 1))
                            steps[nsteps++] = std::make_pair(cost(y + 1, x - 1) + costH(y + 1, x - 1) + costV(y, x), 3);
                    }
                    if (nsteps) {
                        std::pair<float, int> opt = *std::min_element(steps, steps + nsteps);
                        cost(y, x) = opt.first;
                        control(y, x) = (uint8_t)opt.second;
                        reachable(y, x) = 255;
                    }
                }
        } else {                                                 // :889-916
            for (int y = src.y + 1; y <= dst.y; ++y)
                for (int x = 0; x < rw; ++x) {
                    nsteps = 0;
                    if (labels_(y + ry, x + rx) == l) {
                        if (reachable(y - 1, x))
                            steps[nsteps++] = std::make_pair(cost(y - 1, x) + costV(y - 1, x), 1);
                        if (x > 0 && reachable(y - 1, x - 1))
                            steps[nsteps++] = std::make_pair(cost(y - 1, x - 1) + costV(y - 1, x - 1) + costH(y, x - 1), 2);
                        if (x < rw - 1 && reachable(y - 1, x + 1))
                            steps[nsteps++] = std::make_pair(cost(y - 1, x + 1) + costV(y - 1, x + 1) + costH(y, x), 3);
                    }
                    if (nsteps) {
                        std::pair<float, int> opt = *std::min_element(steps, steps + nsteps);
                        cost(y, x) = opt.first;
                        control(y, x) = (uint8_t)opt.second;
                        reachable(y, x) = 255;
                    }
                }
        }
        if (!reachable(dst.y, dst.x)) return false;
        Pt p = dst;                                              // :923-947
        seam.clear();
        seam.push_back({p.x + rx, p.y + ry});
        if (isHorizontal) {
            for (; p.x != src.x; seam.push_back({p.x + rx, p.y + ry})) {
                if (control(p.y, p.x) == 2) p.y--;
                else if (control(p.y, p.x) == 3) p.y++;
                p.x--;
            }
        } else {
            for (; p.y != src.y; seam.push_back({p.x + rx, p.y + ry})) {
                if (control(p.y, p.x) == 2) p.x--;
                else if (control(p.y, p.x) == 3) p.x++;
                p.y--;
            }
        }
        if (!swapped) std::reverse(seam.begin(), seam.end());
        // :953-954 CV_Assert(seam.front() == p1 && seam.back() == p2)
        return seam.front() == p1 && seam.back() == p2;
    }

    // [SEAM]:960-1093
    void updateLabelsUsingSeam(int comp1, int comp2, const std::vector<Pt>& seam, bool isHorizontalSeam) {
        Grid<int> mask;
        const Pt tl = tls_[comp1];
        mask.create(brs_[comp1].y - tl.y, brs_[comp1].x - tl.x, 0);
        for (size_t i = 0; i < contours_[comp1].size(); ++i)
            mask(contours_[comp1][i].y - tl.y, contours_[comp1][i].x - tl.x) = 255;
        for (size_t i = 0; i < seam.size(); ++i) mask(seam[i].y - tl.y, seam[i].x - tl.x) = 255;
        int l1 = comp1 + 1, l2 = comp2 + 1;
        int ncomps = 0;
        for (int y = 0; y < mask.rows; ++y)
            for (int x = 0; x < mask.cols; ++x)
                if (!mask(y, x) && labels_(y + tl.y, x + tl.x) == l1) floodFill(mask, x, y, ++ncomps);
        for (size_t i = 0; i < contours_[comp1].size(); ++i) {
            int x = contours_[comp1][i].x - tl.x, y = contours_[comp1][i].y - tl.y;
            bool ok = false;
            static const int dx[] = {-1, +1, 0, 0, -1, +1, -1, +1};
            static const int dy[] = {0, 0, -1, +1, -1, -1, +1, +1};
            for (int j = 0; j < 8; ++j) {
                int c = x + dx[j], r = y + dy[j];
                if (c >= 0 && c < mask.cols && r >= 0 && r < mask.rows && mask(r, c) && mask(r, c) != 255) {
                    ok = true;
                    mask(y, x) = mask(r, c);
                }
            }
            if (!ok) mask(y, x) = 0;
        }
        if (isHorizontalSeam) {
            for (size_t i = 0; i < seam.size(); ++i) {
                int x = seam[i].x - tl.x, y = seam[i].y - tl.y;
                if (y < mask.rows - 1 && mask(y + 1, x) && mask(y + 1, x) != 255) mask(y, x) = mask(y + 1, x);
                else mask(y, x) = 0;
            }
        } else {
            for (size_t i = 0; i < seam.size(); ++i) {
                int x = seam[i].x - tl.x, y = seam[i].y - tl.y;
                if (x < mask.cols - 1 && mask(y, x + 1) && mask(y, x + 1) != 255) mask(y, x) = mask(y, x + 1);
                else mask(y, x) = 0;
            }
        }
        std::map<int, int> connect2, connectOther;
        for (int i = 1; i <= ncomps; ++i) {
            connect2.insert({i, 0});
            connectOther.insert({i, 0});
        }
        for (size_t i = 0; i < contours_[comp1].size(); ++i) {
            int x = contours_[comp1][i].x, y = contours_[comp1][i].y;
            if ((x > 0 && labels_(y, x - 1) == l2) || (y > 0 && labels_(y - 1, x) == l2) ||
                (x < uw - 1 && labels_(y, x + 1) == l2) || (y < uh - 1 && labels_(y + 1, x) == l2))
                connect2[mask(y - tl.y, x - tl.x)]++;
            if ((x > 0 && labels_(y, x - 1) != l1 && labels_(y, x - 1) != l2) ||
                (y > 0 && labels_(y - 1, x) != l1 && labels_(y - 1, x) != l2) ||
                (x < uw - 1 && labels_(y, x + 1) != l1 && labels_(y, x + 1) != l2) ||
                (y < uh - 1 && labels_(y + 1, x) != l1 && labels_(y + 1, x) != l2))
                connectOther[mask(y - tl.y, x - tl.x)]++;
        }
        // isAdjComp is indexed by mask value; size it to hold every key (0 and, degenerately, 255)
        int maxKey = ncomps;
        for (auto& kv : connect2) maxKey = std::max(maxKey, kv.first);
        std::vector<int> isAdjComp(maxKey + 1, 0);
        for (auto itr = connect2.begin(); itr != connect2.end(); ++itr) {
            double len = static_cast<double>(contours_[comp1].size());
            int res = 0;
            if (itr->second / len > 0.05) {
                auto sub = connectOther.find(itr->first);
                if (sub != connectOther.end() && (sub->second / len < 0.1)) res = 1;
            }
            isAdjComp[itr->first] = res;
        }
        for (int y = 0; y < mask.rows; ++y)
            for (int x = 0; x < mask.cols; ++x)
                if (mask(y, x) && mask(y, x) <= maxKey && isAdjComp[mask(y, x)]) labels_(y + tl.y, x + tl.x) = l2;
    }

    // [SEAM]:395-546
    void resolveConflicts(const Img& image1, const Img& image2, Pt tl1, Pt tl2, uint8_t* mask1, uint8_t* mask2) {
        if (costFunc == ORC_COST_COLOR_GRAD) computeGradients(image1, image2);           // :398-399
        bool hasConflict = true;
        while (hasConflict) {
            int c1 = 0, c2 = 0;
            hasConflict = false;
            for (auto itr = edges_.begin(); itr != edges_.end(); ++itr) {
                c1 = itr->first;
                c2 = itr->second;
                if ((states_[c1] & INTERS) && (states_[c1] & (~INTERS)) != states_[c2]) {
                    hasConflict = true;
                    break;
                }
            }
            if (hasConflict) {
                int l1 = c1 + 1, l2 = c2 + 1;
                if (hasOnlyOneNeighbor(c1)) {
                    for (int y = tls_[c1].y; y < brs_[c1].y; ++y)
                        for (int x = tls_[c1].x; x < brs_[c1].x; ++x)
                            if (labels_(y, x) == l1) labels_(y, x) = l2;
                    states_[c1] = states_[c2] == FIRST ? SECOND : FIRST;
                } else {
                    Pt p1, p2;
                    if (getSeamTips(c1, c2, p1, p2)) {
                        std::vector<Pt> seam;
                        bool isHorizontalSeam;
                        if (estimateSeam(image1, image2, tl1, tl2, c1, p1, p2, seam, isHorizontalSeam)) {
                            record(c1, isHorizontalSeam, seam);
                            updateLabelsUsingSeam(c1, c2, seam, isHorizontalSeam);
                        }
                    }
                    states_[c1] = states_[c2] == FIRST ? INTERS_SECOND : INTERS_FIRST;
                }
                const int c[] = {c1, c2};
                const int l[] = {l1, l2};
                for (int i = 0; i < 2; ++i) {
                    int x0 = tls_[c[i]].x, x1 = brs_[c[i]].x;
                    int y0 = tls_[c[i]].y, y1 = brs_[c[i]].y;
                    tls_[c[i]] = {INT_MAX, INT_MAX};
                    brs_[c[i]] = {INT_MIN, INT_MIN};
                    contours_[c[i]].clear();
                    for (int y = y0; y < y1; ++y)
                        for (int x = x0; x < x1; ++x)
                            if (labels_(y, x) == l[i]) {
                                tls_[c[i]].x = std::min(tls_[c[i]].x, x);
                                tls_[c[i]].y = std::min(tls_[c[i]].y, y);
                                brs_[c[i]].x = std::max(brs_[c[i]].x, x + 1);
                                brs_[c[i]].y = std::max(brs_[c[i]].y, y + 1);
                                if ((x == 0 || labels_(y, x - 1) != l[i]) || (x == uw - 1 || labels_(y, x + 1) != l[i]) ||
                                    (y == 0 || labels_(y - 1, x) != l[i]) || (y == uh - 1 || labels_(y + 1, x) != l[i]))
                                    contours_[c[i]].push_back({x, y});
                            }
                }
                edges_.erase({c1, c2});
                edges_.erase({c2, c1});
            }
        }
        // update masks  :524-545
        int dx1 = unionTl.x - tl1.x, dy1 = unionTl.y - tl1.y;
        int dx2 = unionTl.x - tl2.x, dy2 = unionTl.y - tl2.y;
        for (int y = 0; y < image2.rows; ++y)
            for (int x = 0; x < image2.cols; ++x) {
                int l = labels_(y - dy2, x - dx2);
                if (l > 0 && (states_[l - 1] & FIRST) && mask1[(size_t)(y - dy2 + dy1) * image1.cols + (x - dx2 + dx1)])
                    mask2[(size_t)y * image2.cols + x] = 0;
            }
        for (int y = 0; y < image1.rows; ++y)
            for (int x = 0; x < image1.cols; ++x) {
                int l = labels_(y - dy1, x - dx1);
                if (l > 0 && (states_[l - 1] & SECOND) && mask2[(size_t)(y - dy1 + dy2) * image2.cols + (x - dx1 + dx2)])
                    mask1[(size_t)y * image1.cols + x] = 0;
            }
    }

    void record(int comp, bool horiz, const std::vector<Pt>& seam) {
        if (!trace) return;
        size_t need = 5 + 2 * seam.size();
        if (trace_len + need <= trace_cap) {
            int32_t* t = trace + trace_len;
            t[0] = cur_i; t[1] = cur_j; t[2] = comp; t[3] = horiz ? 1 : 0; t[4] = (int32_t)seam.size();
            for (size_t i = 0; i < seam.size(); ++i) {
                t[5 + 2 * i] = seam[i].x + unionTl.x;     // pano coordinates
                t[6 + 2 * i] = seam[i].y + unionTl.y;
            }
        }
        trace_len += need;
    }
};

}  // namespace

extern "C" {

int orc_dp_seam_find(int n, const void* const* images, int is_u8, const int* rows, const int* cols,
                     const int* corners_xy, uint8_t* const* masks, int cost_fn,
                     int32_t* trace, size_t trace_cap, size_t* trace_len) {
    if (trace_len) *trace_len = 0;
    if (n == 0) return 0;                                    // [SEAM]:94-95
    if (cost_fn != ORC_COST_COLOR && cost_fn != ORC_COST_COLOR_GRAD) return -5;   // StsBadArg
    std::vector<std::pair<int, int>> pairs;                  // [SEAM]:97-111
    for (int i = 0; i + 1 < n; ++i)
        for (int j = i + 1; j < n; ++j) pairs.push_back({i, j});
    std::reverse(pairs.begin(), pairs.end());
    DpSeam f;
    f.costFunc = cost_fn;
    f.trace = trace;
    f.trace_cap = trace_cap;
    for (size_t k = 0; k < pairs.size(); ++k) {              // [SEAM]:115-121
        int i0 = pairs[k].first, i1 = pairs[k].second;
        Img a{images[i0], rows[i0], cols[i0], (is_u8 & 1) != 0, (is_u8 & 2) ? 4 : 3};     // is_u8: bit 0 = 8-bit, bit 1 = four channels
        Img b{images[i1], rows[i1], cols[i1], (is_u8 & 1) != 0, (is_u8 & 2) ? 4 : 3};
        f.cur_i = i0;
        f.cur_j = i1;
        f.process(a, b, Pt{corners_xy[2 * i0], corners_xy[2 * i0 + 1]}, Pt{corners_xy[2 * i1], corners_xy[2 * i1 + 1]},
                  masks[i0], masks[i1]);
    }
    if (trace_len) *trace_len = f.trace_len;
    return 0;
}

void orc_seam_costs(const void* img1, const void* img2, int is_u8,
                    int rows1, int cols1, int rows2, int cols2,
                    int tl1x, int tl1y, int tl2x, int tl2y,
                    const int32_t* labels, int H, int W, int union_tlx, int union_tly,
                    int l, const int roi[4], float* costV, float* costH) {
    orc_seam_costs_ex(img1, img2, is_u8, rows1, cols1, rows2, cols2, tl1x, tl1y, tl2x, tl2y, labels, H, W, union_tlx, union_tly, l, roi,
                      ORC_COST_COLOR, costV, costH);
}

void orc_seam_gradients(const void* img, int is_u8, int rows, int cols, float* gradx, float* grady) {
    Img a{img, rows, cols, (is_u8 & 1) != 0, (is_u8 & 2) ? 4 : 3};
    Grid<float> gx, gy;
    DpSeam::sobelPair(a, gx, gy);
    std::memcpy(gradx, gx.v.data(), sizeof(float) * gx.v.size());
    std::memcpy(grady, gy.v.data(), sizeof(float) * gy.v.size());
}

void orc_seam_costs_ex(const void* img1, const void* img2, int is_u8,
                       int rows1, int cols1, int rows2, int cols2,
                       int tl1x, int tl1y, int tl2x, int tl2y,
                       const int32_t* labels, int H, int W, int union_tlx, int union_tly,
                       int l, const int roi[4], int cost_fn, float* costV, float* costH) {
    DpSeam f;
    f.costFunc = cost_fn;
    f.uw = W; f.uh = H;
    f.unionTl = {union_tlx, union_tly};
    f.labels_.create(H, W, 0);
    std::memcpy(f.labels_.v.data(), labels, sizeof(int32_t) * (size_t)H * W);
    int comp = l - 1;
    f.tls_.assign(comp + 1, Pt{0, 0});
    f.brs_.assign(comp + 1, Pt{0, 0});
    f.states_.assign(comp + 1, INTERS);
    f.tls_[comp] = {roi[0], roi[1]};
    f.brs_[comp] = {roi[0] + roi[2], roi[1] + roi[3]};
    Img a{img1, rows1, cols1, (is_u8 & 1) != 0, (is_u8 & 2) ? 4 : 3}, b{img2, rows2, cols2, (is_u8 & 1) != 0, (is_u8 & 2) ? 4 : 3};
    if (cost_fn == ORC_COST_COLOR_GRAD) f.computeGradients(a, b);
    Grid<float> cv, ch;
    f.computeCosts(a, b, Pt{tl1x, tl1y}, Pt{tl2x, tl2y}, comp, cv, ch);
    std::memcpy(costV, cv.v.data(), sizeof(float) * cv.v.size());
    std::memcpy(costH, ch.v.data(), sizeof(float) * ch.v.size());
}

}  // extern "C"
