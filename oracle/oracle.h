/*
 * oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain CPU restatement of the reference's
 * composite path (warp -> DP seam -> blend) used as the parity checker and as the
 * timed CPU baseline.  Nothing in imagestitch_b200/ (the product) may include,
 * link or call it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity pin: the reference ships no golden vectors.  The arithmetic it
 * delegates to is OpenCV (pinned 3.4.2, un-vendored); this oracle is pinned against
 * OpenCV 4.13 (python cv2) outputs stored under tests/golden/ (generator:
 * tests/golden/make_golden.py) and, when cv2 is importable, against cv2 directly.
 * The hand-written linear blend (oracle/linblend.cpp) is pinned against the reference's own
 * code: [BLEND]:141-717 compiled from /root/reference with a small cv::Mat shim into
 * oracle/_ref/libref_linblend.so (`make ref`), golden outputs in tests/golden/linblend_ref_cases.npz;
 * ROI + backward maps of the cylindrical warp likewise against the reference's own detectResultRoi /
 * mapBackward ([WARP]:47-88, oracle/_ref/libref_warp.so, tests/golden/warp_ref_cases.npz), and the DP
 * seam finder against the reference's own find() ... updateLabelsUsingSeam ([SEAM]:29-1093,
 * oracle/_ref/libref_seam.so, tests/golden/seam_ref_cases.npz).
 *
 * Reference aliases ([WARP], [SEAM], [BLEND]) are defined in SURVEY.md section 0.
 */
#ifndef IMAGESTITCH_ORACLE_H
#define IMAGESTITCH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PROJ_CYLINDRICAL = 0, ORC_PROJ_SPHERICAL = 1, ORC_PROJ_PLANE = 2, ORC_PROJ_FISHEYE = 3, ORC_PROJ_STEREOGRAPHIC = 4 };
enum { ORC_INTER_NEAREST = 0, ORC_INTER_LINEAR = 1 };
enum { ORC_BORDER_CONSTANT = 0, ORC_BORDER_REFLECT = 2 };
enum { ORC_COST_COLOR = 0, ORC_COST_COLOR_GRAD = 1 };
enum { ORC_WEIGHT_32F = 5, ORC_WEIGHT_16S = 3 };

/* ---- warp ([WARP]:36-161, cv::remap) ---- */

/* [WARP]:90-120 setCameraParams.  out: k_rinv[9], r_kinv[9] */
void orc_camera_params(const float K[9], const float R[9], float k_rinv[9], float r_kinv[9]);

/* [WARP]:64-88 detectResultRoi (full scan, as the reference does).  full_scan=0 scans the
 * image border only (what the product does on the host).  out: tl_x,tl_y,br_x,br_y */
void orc_detect_roi(int proj, int src_w, int src_h, const float K[9], const float R[9],
                    float scale, int full_scan, int roi[4]);

/* [WARP]:122-144 buildMaps.  maps are (br_y-tl_y+1) x (br_x-tl_x+1) float, dense. */
void orc_build_maps(int proj, const float K[9], const float R[9], float scale,
                    const int roi[4], float* xmap, float* ymap);

/* cv::remap for 8-bit images, channels = 1 or 3 ([WARP]:157). dst is dense h x w x ch. */
void orc_remap_u8(const uint8_t* src, int src_h, int src_w, int ch, size_t src_step,
                  const float* xmap, const float* ymap, int h, int w,
                  int interp, int border, uint8_t* dst);

/* ---- DP seam finder ([SEAM]:87-1093) ---- */

/* images: n pointers to dense (rows x cols x 3) float32 or uint8 data (is_u8 selects).
 * masks: n pointers to dense rows x cols uint8, modified in place.
 * trace (optional, may be NULL): receives for every estimateSeam call that succeeded
 *   [pair_i, pair_j, comp, isHorizontal, npoints, x0,y0, x1,y1, ...] appended as int32; trace_cap
 *   is the capacity in int32, *trace_len the used length (calls that do not fit are dropped
 *   but still counted).
 * returns 0, or a negative cv::Error-style code. */
int orc_dp_seam_find(int n, const void* const* images, int is_u8, const int* rows, const int* cols,
                     const int* corners_xy, uint8_t* const* masks, int cost_fn,
                     int32_t* trace, size_t trace_cap, size_t* trace_len);

/* [SEAM]:733-803 for label image `labels` (H x W int32, union frame) and component label l with
 * bbox roi (x,y,w,h).  costV is h x (w+1), costH is (h+1) x w. */
void orc_seam_costs(const void* img1, const void* img2, int is_u8,
                    int rows1, int cols1, int rows2, int cols2,
                    int tl1x, int tl1y, int tl2x, int tl2y,
                    const int32_t* labels, int H, int W, int union_tlx, int union_tly,
                    int l, const int roi[4], float* costV, float* costH);
/* ... with the cost function selectable (ORC_COST_COLOR_GRAD: [SEAM]:767-772, :792-797) */
void orc_seam_costs_ex(const void* img1, const void* img2, int is_u8,
                       int rows1, int cols1, int rows2, int cols2,
                       int tl1x, int tl1y, int tl2x, int tl2y,
                       const int32_t* labels, int H, int W, int union_tlx, int union_tly,
                       int l, const int roi[4], int cost_fn, float* costV, float* costH);
/* [SEAM]:549-572 computeGradients for one image: Sobel x / y of its gray conversion (rows x cols f32 each) */
void orc_seam_gradients(const void* img, int is_u8, int rows, int cols, float* gradx, float* grady);

/* ---- pyramids (cv::pyrDown / cv::pyrUp) ---- */
void orc_pyr_down_s16(const int16_t* src, int h, int w, int ch, int16_t* dst);      /* dst ((h+1)/2,(w+1)/2) */
void orc_pyr_up_s16(const int16_t* src, int h, int w, int ch, int dh, int dw, int16_t* dst);
void orc_pyr_down_f32(const float* src, int h, int w, float* dst);

/* ---- multi-band blender (cv::detail::MultiBandBlender) ---- */
typedef struct orc_mb orc_mb;
orc_mb* orc_mb_create(int num_bands, int weight_type);
void orc_mb_destroy(orc_mb*);
/* dst_roi = x,y,w,h */
void orc_mb_prepare(orc_mb*, const int dst_roi[4]);
int orc_mb_num_bands(const orc_mb*);
/* img: dense rows x cols x 3 int16; mask: dense rows x cols uint8 */
void orc_mb_feed(orc_mb*, const int16_t* img, const uint8_t* mask, int rows, int cols, int tlx, int tly);
/* dst: dense roi.h x roi.w x 3 int16 ; dst_mask: roi.h x roi.w uint8 */
void orc_mb_blend(orc_mb*, int16_t* dst, uint8_t* dst_mask);

/* ---- reference's hand-written linear blend ([BLEND]:141-717) ---- */
/* img1/img2: float32 x3 dense.  pano: panoHe x panoBr x 3 float32 (sizes via orc_lin_geometry).
 * seam_x: panoHe ints.  returns 0, 1 when the two images do not overlap ([BLEND]:182). */
void orc_lin_geometry(int rows1, int cols1, int rows2, int cols2, int tl1x, int tl1y, int tl2x, int tl2y,
                      int* panoHe, int* panoBr);
int orc_lin_blend(const float* img1, int rows1, int cols1, const float* img2, int rows2, int cols2,
                  int tl1x, int tl1y, int tl2x, int tl2y, float* pano, int* seam_x,
                  float* costV_out /* panoHe x (interSectBr+2), may be NULL */);

/* ---- feather blend + mask preparation (oracle/feather.cpp): dilate / distanceTransform / cv::detail::FeatherBlender ---- */
void orc_dilate_rect(const uint8_t* src, int rows, int cols, int kw, int kh, uint8_t* dst);
void orc_distance_l1(const uint8_t* mask, int rows, int cols, float* dist);
void orc_feather_weight(const uint8_t* mask, int rows, int cols, float sharpness, float* weight);
typedef struct orc_fb orc_fb;
orc_fb* orc_fb_create(float sharpness);
void orc_fb_destroy(orc_fb* f);
void orc_fb_prepare(orc_fb* f, const int* roi_xywh);
void orc_fb_feed(orc_fb* f, const int16_t* img, const uint8_t* mask, int rows, int cols, int tlx, int tly);
void orc_fb_blend(orc_fb* f, int16_t* pano, uint8_t* pano_mask);

/* ---- gain exposure compensation (oracle/exposure.cpp): cv::detail::GainCompensator feed / apply ---- */
int orc_gain_feed(int n, const uint8_t* const* images, const uint8_t* const* masks, const int* rows, const int* cols, const int* corners_xy,
                  double* gains);
void orc_gain_apply(uint8_t* image, size_t count, double gain);

/* ---- whole composite path (oracle/pipeline.cpp): warp -> [gain exposure] -> [DP seam] -> multi-band blend ---- */
void orc_set_threads(int n);
int orc_get_max_threads(void);
int orc_pipeline_plan(int n, int proj, const int* src_rows, const int* src_cols, const float* K, const float* R,
                      float scale, int* corners_xy, int* sizes_wh, int* pano_roi);
int orc_pipeline_run(int n, int proj, const uint8_t* const* srcs, const int* src_rows, const int* src_cols,
                     const float* K, const float* R, float scale, int seam, int num_bands, int weight_type,
                     const int* corners_xy, const int* sizes_wh, const int* pano_roi,
                     uint8_t* const* warped_out, uint8_t* const* masks_out,
                     int16_t* pano, uint8_t* pano_mask, double* stage_seconds);
/* the same with the gain exposure compensator of the mains: gains from the warped images and masks ([BLEND]:117-123), the
 * seam finder sees the uncompensated images, apply() before the blender's feed ([SEAM]:1165-1171); gains_out[n] optional */
int orc_pipeline_run_ex(int n, int proj, const uint8_t* const* srcs, const int* src_rows, const int* src_cols,
                        const float* K, const float* R, float scale, int seam, int num_bands, int weight_type, int exposure_gain,
                        const int* corners_xy, const int* sizes_wh, const int* pano_roi,
                        uint8_t* const* warped_out, uint8_t* const* masks_out,
                        int16_t* pano, uint8_t* pano_mask, double* stage_seconds, double* gains_out);
/* ... and with the blender of the mains' live path: blender 0 = multi-band, 1 = feather (sharpness); seam_dilate > 0:
 * masks = dilate(masks, d x d) & warped masks before the blender's feed ([SEAM]:1257-1270) */
/* seam: 0 = keep the warped masks, 1 = DP seam finder with COLOR, 2 = with COLOR_GRAD */
int orc_pipeline_run_ex2(int n, int proj, const uint8_t* const* srcs, const int* src_rows, const int* src_cols,
                         const float* K, const float* R, float scale, int seam, int num_bands, int weight_type, int exposure_gain,
                         int blender, float sharpness, int seam_dilate,
                         const int* corners_xy, const int* sizes_wh, const int* pano_roi,
                         uint8_t* const* warped_out, uint8_t* const* masks_out,
                         int16_t* pano, uint8_t* pano_mask, double* stage_seconds, double* gains_out);

/* ---- ORB features finder ([FEAT]:56-418, 727-1021; cvtColor, resize(INTER_LINEAR_EXACT), FAST, fastAtan2, GaussianBlur) ---- */
void orc_bgr2gray(const uint8_t* src, int rows, int cols, int ch, size_t step, uint8_t* dst);
void orc_resize_linear_exact_u8(const uint8_t* src, int rows, int cols, uint8_t* dst, int drows, int dcols);
int orc_fast(const uint8_t* src, int rows, int cols, int threshold, int* out, int cap);
void orc_gaussian7_u8(const uint8_t* src, int rows, int cols, uint8_t* dst);
float orc_fast_atan2(float y, float x);
int orc_orb_find(const uint8_t* img, int rows, int cols, int channels, size_t step, int grid_w, int grid_h, int nfeatures, float scale_factor,
                 int nlevels, float* kps, uint8_t* desc, int cap);

#ifdef __cplusplus
}
#endif
#endif
