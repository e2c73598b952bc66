/*
 * imagestitch.h -- C ABI of libimagestitch_b200.so: the per-pixel composite stage of the stitching
 * pipeline (cylindrical/spherical backward warp -> DP seam -> multi-band / linear blend) on B200.
 *
 * The reference (mhhai/ImageStitch) has no FFI of its own; its hot path is the set of cv::detail
 * stitching interfaces it both calls and re-implements as free functions.  Each entry point below
 * names the reference interface it replaces (aliases [WARP], [SEAM], [BLEND]: SURVEY.md section 0).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary: every call returns an is_status (0 = OK, negative
 *     values follow cv::Error codes where the reference would CV_Assert / CV_Error, plus CUDA codes).
 *     is_ctx_last_error() gives the message.
 *   - images are described by is_mat, a mirror of cv::Mat: row-major, channels interleaved (BGR),
 *     `step` bytes per row, `depth` uses OpenCV's depth codes.  `device` = -1 for host memory or the
 *     CUDA ordinal that owns `data`; host buffers are staged through the context's stream, device
 *     buffers are used in place.  All buffers are caller-owned.
 *   - an is_ctx owns a CUDA stream, a stream-ordered workspace pool and the launch counter; one
 *     context per host thread (the reference keeps this state in globals and is not re-entrant,
 *     [WARP]:30-35, [SEAM]:65-85).  Calls are asynchronous with respect to device buffers; results in
 *     host buffers are complete on return.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     IS_ERR_CUDA.
 */
#ifndef IMAGESTITCH_H
#define IMAGESTITCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IS_VERSION_MAJOR 0
#define IS_VERSION_MINOR 1

typedef enum is_status {
    IS_OK = 0,
    IS_ERR_NO_MEM = -4,          /* cv::Error::StsNoMem */
    IS_ERR_BAD_ARG = -5,         /* cv::Error::StsBadArg   ([SEAM]:749) */
    IS_ERR_UNSUPPORTED = -213,   /* cv::Error::StsNotImplemented */
    IS_ERR_ASSERT = -215,        /* cv::Error::StsAssert   ([WARP]:94-96, [SEAM]:133-134) */
    IS_ERR_CUDA = -1000,         /* CUDA runtime / driver failure, or no device */
    IS_ERR_INTERNAL = -1001
} is_status;

/* depth codes = OpenCV's */
enum { IS_8U = 0, IS_16S = 3, IS_32S = 4, IS_32F = 5 };

typedef struct is_mat {
    void* data;
    int rows, cols, channels, depth;
    size_t step;      /* bytes per row */
    int device;       /* -1: host memory, >= 0: CUDA device ordinal */
} is_mat;

typedef struct is_point { int x, y; } is_point;
typedef struct is_size { int width, height; } is_size;
typedef struct is_rect { int x, y, width, height; } is_rect;

/* cv::CylindricalWarper / SphericalWarper / PlaneWarper / FisheyeWarper / StereographicWarper: the warper creators of
 * [BLEND]:91-95.  Fisheye and stereographic need libm calls per PIXEL, so their maps are built on the host (as
 * buildMaps [WARP]:122-144 does) and sampled on the device; the other three are evaluated on the device from O(W + H)
 * host tables and take the fused warp path of is_pipeline_run. */
typedef enum is_projection {
    IS_PROJ_CYLINDRICAL = 0, IS_PROJ_SPHERICAL = 1, IS_PROJ_PLANE = 2, IS_PROJ_FISHEYE = 3, IS_PROJ_STEREOGRAPHIC = 4
} is_projection;
typedef enum is_interp { IS_INTER_NEAREST = 0, IS_INTER_LINEAR = 1 } is_interp;           /* cv::INTER_* */
typedef enum is_border { IS_BORDER_CONSTANT = 0, IS_BORDER_REFLECT = 2 } is_border;       /* cv::BORDER_* */
typedef enum is_seam_cost { IS_COST_COLOR = 0, IS_COST_COLOR_GRAD = 1 } is_seam_cost;     /* [SEAM]:71 */
typedef enum is_weight_type { IS_WEIGHT_32F = 5, IS_WEIGHT_16S = 3 } is_weight_type;      /* CV_32F / CV_16S */

typedef struct is_ctx is_ctx;
typedef struct is_blender is_blender;

/* ------------------------------------------------------------------ context */
const char* is_version(void);
const char* is_status_string(int status);
int is_ctx_create(int device, is_ctx** out);
int is_ctx_destroy(is_ctx* ctx);
int is_ctx_synchronize(is_ctx* ctx);
const char* is_ctx_last_error(const is_ctx* ctx);
void* is_ctx_stream(is_ctx* ctx);                     /* cudaStream_t all work of this context runs on */
int is_ctx_set_stream(is_ctx* ctx, void* stream);     /* adopt a caller-owned cudaStream_t; NULL = the legacy default stream */
int is_ctx_reset_stream(is_ctx* ctx);                  /* back to the context's own stream */
uint64_t is_ctx_kernel_launches(const is_ctx* ctx);   /* kernels of this library launched so far */
int is_ctx_device(const is_ctx* ctx);
/* Per-launch CUDA-event timing for measurement (bench.py's roofline figure): when enabled every kernel launch
 * of this context is bracketed by two events on the context's stream.  The report is a JSON array
 * [{"name", "launches", "ms", "bytes"}] aggregated per kernel since the last report; it synchronises the
 * stream, clears the records and returns the buffer length needed. */
int is_ctx_kernel_timing(is_ctx* ctx, int enable);
int is_ctx_kernel_timing_report(is_ctx* ctx, char* buf, size_t cap);

/* ------------------------------------------------------------------ warp
 * Replaces the free functions of [WARP] (== cv::detail::RotationWarper created by
 * WarperCreator::create(scale), [BLEND]:99):
 *   Rect  buildMaps(Size src_size, InputArray K, InputArray R, OutputArray xmap, OutputArray ymap)  [WARP]:122
 *   Point warp(InputArray src, InputArray K, InputArray R, int interp, int border, OutputArray dst) [WARP]:145
 * K and R are 3x3 row-major float (CV_32F is asserted at [WARP]:94-96).
 */

/* detectResultRoi [WARP]:64-88 -> top-left corner in panorama coordinates and the size of the
 * destination the reference allocates (dst.create(roi.height + 1, roi.width + 1), [WARP]:150). */
int is_warp_roi(is_ctx* ctx, int projection, is_size src_size, const float K[9], const float R[9],
                float scale, is_point* dst_tl, is_size* dst_size);

/* buildMaps [WARP]:122-144.  xmap / ymap: dst_size, 1 channel IS_32F, caller-allocated. */
int is_build_maps(is_ctx* ctx, int projection, is_size src_size, const float K[9], const float R[9],
                  float scale, is_mat* xmap, is_mat* ymap, is_rect* dst_roi);

/* warp [WARP]:145-161 (buildMaps + cv::remap fused; no maps are materialised).  src: IS_8U, 1 or 3
 * channels.  dst: caller-allocated, dst_size from is_warp_roi, same type as src. */
int is_warp(is_ctx* ctx, int projection, const is_mat* src, const float K[9], const float R[9], float scale,
            int interp, int border, is_mat* dst, is_point* dst_tl);

/* cv::remap(src, dst, xmap, ymap, interp, border) -- the call warp() ends in, [WARP]:157 -- on its own: src IS_8U with 1 or
 * 3 channels, maps 1-channel IS_32F of the destination's size, dst caller-allocated with the type of src.  NaN and
 * out-of-range map values sample what cv::remap samples on x86 (cvRound's INT_MIN). */
int is_remap(is_ctx* ctx, const is_mat* src, const is_mat* xmap, const is_mat* ymap, int interp, int border, is_mat* dst);

/* The two warp calls every main() makes per image ([BLEND]:105,109): image with INTER_LINEAR +
 * BORDER_REFLECT and an all-255 mask with INTER_NEAREST + BORDER_CONSTANT, in one pass over the
 * destination.  dst: 3 channels IS_8U; dst_mask: 1 channel IS_8U. */
int is_warp_with_mask(is_ctx* ctx, int projection, const is_mat* src, const float K[9], const float R[9],
                      float scale, is_mat* dst, is_mat* dst_mask, is_point* dst_tl);

/* ------------------------------------------------------------------ image files
 * The on-disk format either side of the path: cv::imread(".bmp") [BLEND]:31-34, [SEAM]:1098-1100 and
 * cv::imwrite(".bmp", mat) [BLEND]:717, [SEAM]:1195-1206.  The file's bytes go to the device as they are and a kernel
 * turns them into cv::Mat rows (top-down, BGR interleaved, palette looked up, padding / alpha dropped); writing packs the
 * rows bottom-up on the device with imwrite's saturating conversion to 8 bit fused in.  Uncompressed 1 / 4 / 8 / 24 /
 * 32-bit bitmaps; byte-identical to cv2.imread / cv2.imwrite.  dst / src may be host or device memory. */
int is_bmp_info(is_ctx* ctx, const char* path, is_size* size, int* bits_per_pixel);
/* imread(path) with the default IMREAD_COLOR: dst is IS_8U with 3 channels of the size is_bmp_info reports */
int is_imread_bmp(is_ctx* ctx, const char* path, is_mat* dst);
/* imwrite(path, src): 1 channel (8-bit file with a gray palette) or 3 channels (24-bit file); IS_16S and IS_32F are converted
 * like cv::Mat::convertTo(CV_8U) -- what imwrite does with images_warped_f, pano and result in the mains */
int is_imwrite_bmp(is_ctx* ctx, const char* path, const is_mat* src);

/* ------------------------------------------------------------------ seam
 * Replaces  void find(const std::vector<UMat>& src, const std::vector<Point>& corners,
 *                     std::vector<UMat>& masks)                                        [SEAM]:87
 * (== cv::detail::DpSeamFinder::find).  images: n mats, 3 channels -- or 4 with the fourth ignored, diffL2Square4 [SEAM]:722-730,
 * 745-748 -- IS_32F or IS_8U (all the same);
 * masks: n mats IS_8U, same sizes as the images ([SEAM]:133-134), modified in place.
 * cost_fn: IS_COST_COLOR or IS_COST_COLOR_GRAD ([SEAM]:549-572, :767-772, :792-797; 8-bit images are taken as their
 * CV_32F conversion, which is what the mains pass).
 */
int is_seam_dp_find(is_ctx* ctx, int n, const is_mat* images, const is_point* corners, is_mat* masks, int cost_fn);

/* Same, additionally reporting every seam the DP estimated ([SEAM]:806-957) as int32 records
 * [pair_i, pair_j, comp, isHorizontal, npoints, x0, y0, x1, y1, ...] (panorama coordinates) into the
 * host array `trace` of capacity trace_cap; *trace_len receives the length needed. */
int is_seam_dp_find_trace(is_ctx* ctx, int n, const is_mat* images, const is_point* corners, is_mat* masks,
                          int cost_fn, int32_t* trace, size_t trace_cap, size_t* trace_len);

/* Single-pair primitives for sharded execution (one strip of the panorama per GPU, pairs of a strip boundary run
 * on the GPU that owns the left image).  is_seam_pair_run = process() [SEAM]:127-193 for ONE pair on the given
 * input masks, writing the masks with this pair's clears to out_i / out_j and returning a handle on the pair's
 * structural fingerprint.  is_seam_pair_check recomputes only that fingerprint on the masks the pair would have
 * seen in the reference's sequential loop (entry masks minus the clears of earlier pairs); *same = 1 proves that
 * the speculative result is the sequential one.  is_mask_and intersects clear sets: dst = 0 where src == 0. */
typedef struct is_seam_pair is_seam_pair;
int is_seam_pair_run(is_ctx* ctx, const is_mat* image_i, const is_mat* image_j, is_point tl_i, is_point tl_j, const is_mat* mask_i,
                     const is_mat* mask_j, is_mat* out_i, is_mat* out_j, is_seam_pair** result);
int is_seam_pair_check(is_ctx* ctx, const is_mat* image_i, const is_mat* image_j, is_point tl_i, is_point tl_j, const is_mat* mask_i,
                       const is_mat* mask_j, const is_seam_pair* spec, int* same);
int is_seam_pair_destroy(is_seam_pair* p);
/* Sharded strips: would pair (i, j) decide the same (components, contours, conflict loop, seam tips) with mask_j_b in place of
 * mask_j_a?  Used at a strip boundary: mask_j_a is image j's mask as the pair saw it, mask_j_b the one it would have seen in the
 * reference's sequential loop (after the owner's own pairs).  *same = 0 also when the masks are outside what the check covers. */
int is_seam_pair_same_structure(is_ctx* ctx, const is_mat* mask_i, const is_mat* mask_j_a, const is_mat* mask_j_b,
                                is_point tl_i, is_point tl_j, int* same);
int is_mask_and(is_ctx* ctx, is_mat* dst, const is_mat* src);

/* How the last is_seam_dp_find / is_pipeline_run on this context executed the pair loop: 1 = the pairs ran
 * concurrently and their results were proven equal to the reference's sequential loop, 0 = the proof failed and
 * the sequential loop was run, -1 = sequential (fewer than two overlapping pairs, or IS_SEAM_SEQUENTIAL=1). */
int is_ctx_seam_speculation(const is_ctx* ctx);

/* Which implementation of the pair loop the last call took: 2 = batched (all pairs through each kernel in one launch,
 * csrc/seam_batch.inl), 1 = one host thread + stream per pair, 0 = the reference's sequential loop [SEAM]:100-121.
 * The results are identical; the batched path hands inputs it does not cover (noisy masks with more than 8 toggles
 * per row, a component cut by two seams) to the other two. */
int is_ctx_seam_path(const is_ctx* ctx);

/* Host-only diagnostic, no device needed: finishes one pair on the CPU with the batched path's host code, given the seams
 * of its plan ([npts, x0, y0, ...] per seam operation, panorama coordinates, tip 1 -> tip 2; npts = 0 when the estimation
 * failed): run-domain updateLabelsUsingSeam, clear intervals, masks updated in place. */
int is_debug_seam_pair_finish(uint8_t* mask1, int rows1, int cols1, size_t step1, int tl1x, int tl1y,
                              uint8_t* mask2, int rows2, int cols2, size_t step2, int tl2x, int tl2y,
                              const int32_t* seams, size_t seams_len);

/* Tuning diagnostic: njobs synthetic seams (lanes x steps cost tables generated on the device) through the DP forward /
 * back-track kernels of formulation `variant` (0 or 1, see csrc/seam.cu); seam_out[njobs][steps] = seam lanes (or -1 when the
 * destination is unreachable), ms[0] = mean milliseconds per launch over `iters` launches. */
int is_debug_dp_bench(is_ctx* ctx, int lanes, int steps, int njobs, int variant, unsigned seed, int iters, int32_t* seam_out, float* ms);

/* Waves the batched pair loop needed in the last call (1 for a strip; more when pairs depend on each other's clears). */
int is_ctx_seam_waves(const is_ctx* ctx);

/* Drops the context's memo of warp plans (camera -> result ROI, detectResultRoi [WARP]:64-88).  The memo lets the
 * plan -> run sequence of ONE panorama scan the image borders once; a benchmark that repeats the same panorama clears
 * it at the start of every step so that each step pays for its own scan, as a stream of different panoramas would. */
int is_ctx_clear_plan_cache(is_ctx* ctx);

/* Host-only diagnostic, no device needed: the waves of the batched seam path for a set of warped image rectangles (which pairs
 * start together; pairs without a common image may overtake each other), assuming every wave is accepted in full.
 * out (int32): [nwaves, then per wave: count, (i, j) x count]; *len = values needed. */
int is_debug_seam_wave_schedule(int n, const is_point* corners, const is_size* sizes, int32_t* out, size_t cap, size_t* len);

/* Host-only diagnostic, no device needed: structure and plan of one image pair as the batched path computes them
 * between its kernels (components, states, conflict-loop operations with seam tips, contour records).  See seam.cu. */
int is_debug_seam_pair_plan(const uint8_t* mask1, int rows1, int cols1, size_t step1, int tl1x, int tl1y,
                            const uint8_t* mask2, int rows2, int cols2, size_t step2, int tl2x, int tl2y,
                            int32_t* out, size_t cap, size_t* len);

/* computeCosts [SEAM]:733-803 for the component labelled `label` of a host/device label image
 * (IS_32S, union frame of the pair, top-left union_tl in panorama coordinates) over `roi` (union-frame
 * coordinates).  costV: roi.height x (roi.width + 1), costH: (roi.height + 1) x roi.width, IS_32F. */
int is_seam_cost_maps(is_ctx* ctx, const is_mat* image1, const is_mat* image2, is_point tl1, is_point tl2,
                      const is_mat* labels, is_point union_tl, int label, is_rect roi,
                      is_mat* costV, is_mat* costH);

/* ------------------------------------------------------------------ multi-band blend
 * Replaces the blender calls of every main() ([SEAM]:1244-1252,1271,1280; == cv::detail::MultiBandBlender):
 *   Blender::createDefault(MULTI_BAND) + setNumBands   -> is_blender_create
 *   prepare(corners, sizes) / prepare(Rect)            -> is_blender_prepare / is_blender_prepare_roi
 *   feed(InputArray img, InputArray mask, Point tl)    -> is_blender_feed
 *   blend(InputOutputArray dst, InputOutputArray mask) -> is_blender_blend
 */
int is_blender_create(is_ctx* ctx, int num_bands, int weight_type, is_blender** out);
int is_blender_destroy(is_blender* b);
int is_blender_prepare(is_blender* b, int n, const is_point* corners, const is_size* sizes);
int is_blender_prepare_roi(is_blender* b, is_rect dst_roi);
int is_blender_num_bands(const is_blender* b);        /* after prepare: min(num_bands, ceil(log2(max side))) */
int is_blender_dst_size(const is_blender* b, is_size* size);

/* img: 3 channels IS_16S or IS_8U; mask: IS_8U, same size.  flags: IS_FEED_COPY keeps a private
 * copy (OpenCV semantics); IS_FEED_BORROW uses device buffers in place -- they must stay valid and
 * unchanged until is_blender_blend returns (host buffers are always copied). */
enum { IS_FEED_COPY = 0, IS_FEED_BORROW = 1, IS_FEED_DEFER_WEIGHTS = 2 };
int is_blender_feed(is_blender* b, const is_mat* img, const is_mat* mask, is_point tl, int flags);
/* feed() with an explicit position in the feed order (ascending key > 0; the float weight sums of the blender depend on the
 * order).  IS_FEED_DEFER_WEIGHTS (device buffers, borrowed): the image's Gaussian pyramid is built now on a side stream; the
 * mask may still change and is only read when is_blender_blend / is_blender_blend_strip is called -- lets the caller feed
 * the images before its seam stage has produced the final masks. */
int is_blender_feed_ex(is_blender* b, const is_mat* img, const is_mat* mask, is_point tl, int flags, long long key);

/* dst: 3 channels IS_16S, dst_mask: IS_8U, both is_blender_dst_size, caller-allocated. */
int is_blender_blend(is_blender* b, is_mat* dst, is_mat* dst_mask);

/* Column-strip sharding (one strip of the panorama per GPU): blend only the columns [x0, x1) of the destination
 * ROI.  dst / dst_mask: ROI height x (x1 - x0).  The result equals the same columns of is_blender_blend provided
 * every image for which is_blender_strip_needs reports 1 was fed, in the same order as for the full blend. */
int is_blender_strip_needs(const is_blender* b, is_size img_size, is_point tl, int x0, int x1, int* needed);
int is_blender_blend_strip(is_blender* b, int x0, int x1, is_mat* dst, is_mat* dst_mask);

/* ------------------------------------------------------------------ linear blend
 * Replaces the hand-written pair blend of [BLEND]:141-717 (cost map, greedy seam, seam-guided linear
 * weights, composite).  img1 / img2: 3 channels IS_32F; image 1 must be the left image.
 * pano: 3 channels IS_32F of is_linear_blend_size; seam_x: host array of pano rows ints (may be NULL).
 * Returns 1 (not an error) when the two images do not overlap ([BLEND]:182-183).
 */
int is_linear_blend_size(is_size size1, is_size size2, is_point tl1, is_point tl2, is_size* pano_size);
int is_linear_blend_pair(is_ctx* ctx, const is_mat* img1, const is_mat* img2, is_point tl1, is_point tl2,
                         is_mat* pano, int* seam_x);

/* ------------------------------------------------------------------ composite pipeline
 * The call sequence of every main(): detect -> match -> homography (host control flow, supplied by
 * the caller as cameras or through the hooks below) -> warp -> seam -> blend.
 */
typedef struct is_camera { float K[9]; float R[9]; } is_camera;

/* Host-side registration hooks mirroring (*finder)(img, features) [FEAT]:948, matcher(features,
 * matches) [MATCH]:123 and estimator(features, matches, cameras) [CAM]:118.  They run on the host
 * before the GPU stages; `estimate` must fill n cameras and the warp scale.  All may be NULL when
 * cameras are passed to is_pipeline_run directly. */
typedef struct is_registration_hooks {
    void* user;
    int (*detect)(void* user, int image_index, const is_mat* image);
    int (*match)(void* user, int n_images);
    int (*estimate)(void* user, int n_images, is_camera* cameras, float* scale);
} is_registration_hooks;

/* ---- mask preparation + feather blend: the blend path the reference's mains execute ([SEAM]:1249-1280)
 *   Mat element = getStructuringElement(MORPH_RECT, Size(20, 20)); dilate(masks_seam[k], masks_seam[k], element);
 *   masks_seam[k] = masks_seam[k] & masks_warped[k];                       -> is_mask_dilate_and(mask, 20, 20, masks_warped)
 *   Blender::createDefault(Blender::FEATHER) + setSharpness(0.1)           -> is_feather_create
 *   prepare(corners, sizes) / feed(img_s, mask, corner) / blend(dst, mask) -> is_feather_prepare / _feed / _blend
 * img: 3 channels IS_16S or IS_8U; masks IS_8U; dst: 3 channels IS_16S, dst_mask: IS_8U, of is_feather_dst_size. */
typedef struct is_feather_blender is_feather_blender;
int is_mask_dilate_and(is_ctx* ctx, is_mat* mask /* in-out */, int kw, int kh, const is_mat* and_mask /* may be NULL */);
int is_feather_weight_map(is_ctx* ctx, const is_mat* mask, float sharpness, is_mat* weight /* IS_32F: createWeightMap */);
int is_feather_create(is_ctx* ctx, float sharpness, is_feather_blender** out);
int is_feather_destroy(is_feather_blender* b);
int is_feather_prepare(is_feather_blender* b, int n, const is_point* corners, const is_size* sizes);
int is_feather_prepare_roi(is_feather_blender* b, is_rect dst_roi);
int is_feather_dst_size(const is_feather_blender* b, is_size* size);
int is_feather_feed(is_feather_blender* b, const is_mat* img, const is_mat* mask, is_point tl);
int is_feather_blend(is_feather_blender* b, is_mat* dst, is_mat* dst_mask);

/* ---- ORB features finder: (*finder)(img, features) of every main ([BLEND]:36-41), restated by the reference as
 *      void find(InputArray image, ImageFeatures& features)   [FEAT]:948-1021
 * = gray conversion, grid_width x grid_height cells, per cell ORB::detectAndCompute [FEAT]:727-946 (FAST + Harris + orientation +
 * rBRIEF, wta_k = 2, edgeThreshold = patchSize = 31, fastThreshold = 20).  params == NULL takes the reference's values
 * (nfeatures 510, scaleFactor 1.3f, nlevels 5, grid 3 x 1; [FEAT]:39-55).  image: IS_8U with 1, 3 or 4 channels, host or device.
 * keypoints / descriptors (32 bytes each) receive at most `capacity` entries in the reference's order; *count the number found
 * (at most about nfeatures per cell, ties included).  What a detect hook of is_registration_hooks would call. */
typedef struct is_keypoint { float x, y, size, angle, response; int octave, class_id; } is_keypoint;   /* cv::KeyPoint */
typedef struct is_orb_params { int nfeatures; float scale_factor; int nlevels; int grid_width, grid_height; } is_orb_params;
int is_orb_find(is_ctx* ctx, const is_mat* image, const is_orb_params* params, is_keypoint* keypoints, uint8_t* descriptors,
                int capacity, int* count);

typedef enum is_seam_mode { IS_SEAM_NONE = 0, IS_SEAM_DP = 1 } is_seam_mode;
typedef enum is_exposure_mode { IS_EXPOSURE_NONE = 0, IS_EXPOSURE_GAIN = 1 } is_exposure_mode;
typedef enum is_blender_type { IS_BLEND_MULTI_BAND = 0, IS_BLEND_FEATHER = 1 } is_blender_type;

/* ---- gain exposure compensation: cv::detail::GainCompensator (ExposureCompensator::GAIN), the compensator of every main:
 *      compensator->feed(corners, images_warped, masks_warped)   [BLEND]:117-123 / [SEAM]:1165-1171
 *      compensator->apply(img_idx, corner, img_warped, mask)     compositing loop, before Blender::feed
 * images: n CV_8UC3, masks: n CV_8U (255 = inside), host or device.  gains[n] receives what getMatGains() would return. */
int is_gain_feed(is_ctx* ctx, int n, const is_point* corners, const is_mat* images, const is_mat* masks, double* gains);
/* image (CV_8U, any channel count, in place) = saturate_cast<uchar>(image * gain) */
int is_gain_apply(is_ctx* ctx, is_mat* image, double gain);

typedef struct is_pipeline_config {
    int projection;      /* is_projection */
    int seam;            /* is_seam_mode */
    int seam_cost;       /* is_seam_cost */
    int num_bands;       /* MultiBandBlender::setNumBands, default 5 */
    int weight_type;     /* is_weight_type */
    float scale;         /* warper scale = cameras[0].focal ([BLEND]:99); ignored when hooks->estimate is set */
    int exposure;        /* is_exposure_mode: gains from the warped images, applied before the blender's feed */
    int blender;         /* is_blender_type: IS_BLEND_MULTI_BAND (num_bands, weight_type) or IS_BLEND_FEATHER (sharpness) */
    float sharpness;     /* FeatherBlender::setSharpness ([SEAM]:1251: 0.1) */
    int seam_dilate;     /* > 0: masks_seam = dilate(masks_seam, seam_dilate x seam_dilate rectangle) & masks_warped before the
                            blender's feed ([SEAM]:1257-1270: 20); 0: the seam masks are fed as they are */
} is_pipeline_config;

typedef struct is_pipeline_plan_t {
    is_rect pano_roi;    /* resultRoi(corners, sizes) */
} is_pipeline_plan_t;

/* Registration only (host): validates the images, runs hooks->detect / match / estimate ((*finder)(img, features) [FEAT]:948,
 * matcher [MATCH]:123, estimator [CAM]:118) and hands back the n cameras and the warper scale they produce.  The two-phase form
 * of the call sequence: is_pipeline_estimate -> is_pipeline_plan (sizes the panorama) -> is_pipeline_run with those cameras,
 * cfg.scale = the scale and hooks = NULL.  (is_pipeline_run also accepts hooks directly; the output mats must then already
 * have the size the estimated geometry leads to.) */
int is_pipeline_estimate(is_ctx* ctx, int n, const is_mat* images, const is_registration_hooks* hooks,
                         is_camera* cameras, float* scale);

/* Geometry only (host): corners[n], sizes[n] of the warped images and the panorama ROI. */
int is_pipeline_plan(is_ctx* ctx, int n, const is_size* src_sizes, const is_camera* cameras,
                     const is_pipeline_config* cfg, is_point* corners, is_size* sizes, is_rect* pano_roi);

/* images: n source images, 3 channels IS_8U, host or device.  pano: 3 channels IS_16S, pano_mask:
 * IS_8U, both of pano_roi size.  seam_masks (optional, may be NULL): n caller-allocated IS_8U mats of
 * sizes[i] receiving the final seam masks. */
int is_pipeline_run(is_ctx* ctx, int n, const is_mat* images, const is_camera* cameras,
                    const is_registration_hooks* hooks, const is_pipeline_config* cfg,
                    is_mat* pano, is_mat* pano_mask, is_mat* seam_masks);

/* Exposure gains of the last is_pipeline_run with cfg.exposure == IS_EXPOSURE_GAIN (compensator->getMatGains()). */
int is_pipeline_last_gains(is_ctx* ctx, int n, double* gains);

/* Per-stage device time (ms) of the last is_pipeline_run on this context: warp, seam, blend, total. */
int is_pipeline_last_timings(is_ctx* ctx, float ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* IMAGESTITCH_H */
