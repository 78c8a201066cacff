/*
 * lvt_kernels.h -- the internal seam: a thin C ABI (plain pointers and sizes, no C++ / torch
 * types) that sits exactly where the reference's components call their third-party
 * libraries (OpenCV AGAST / BRIEF / BFMatcher, g2o).  A maintainer of the reference would
 * bind these from lvt_image_features_handler.cpp, lvt_image_features_struct.cpp,
 * lvt_local_map.cpp and lvt_pnp_solver.cpp.
 *
 * All pointers are HOST pointers; the CUDA library stages them to the device, runs the
 * kernels on its own stream and copies the results back before returning.  The CPU oracle
 * exports the identical symbols so that a test can call both and memcmp the outputs.
 *
 * Return value: 0 = ok, <0 = error (LVTK_ERR_*).  Nothing throws.
 */
#ifndef LVT_B200_KERNELS_H__
#define LVT_B200_KERNELS_H__

#include "lvt_c.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVTK_OK 0
#define LVTK_ERR_ARG (-1)      /* bad argument */
#define LVTK_ERR_CUDA (-2)     /* CUDA runtime / driver error */
#define LVTK_ERR_CAPACITY (-3) /* an output or internal list overflowed its capacity */
#define LVTK_ERR_NO_DEVICE (-4)

/* one feature location; response is the AGAST score (integer valued) or 0 for external corners */
typedef struct lvtk_keypoint
{
    float x, y, response;
} lvtk_keypoint;

typedef struct lvtk_ctx lvtk_ctx;

/* One context per sequence: owns the device buffers, stream and tensor maps sized for
 * p->img_width x p->img_height.  device < 0 keeps the current device. */
LVT_API lvtk_ctx *lvtk_ctx_create(const lvt_params_c *p, int device);
LVT_API void lvtk_ctx_destroy(lvtk_ctx *ctx);
/* 1 if this library computes on a GPU, 0 for the CPU oracle */
LVT_API int lvtk_is_gpu(void);

/* cv::AgastFeatureDetector(threshold, nonmax, OAST_9_16)::detect on one image treated as a
 * single tile -- replaces the call at lvt/src/lvt_image_features_handler.cpp:139.
 * Keypoints come out in raster order with integer coordinates. */
LVT_API int lvtk_agast(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, int threshold,
                       int nonmax, lvtk_keypoint *out, int cap, int *n_out);

/* perform_detect_corners + low-corner retry: tile grid, per-tile AGAST + ANMS, tile offset,
 * retry at th*0.5+0.5 when fewer than 200 corners
 * (lvt/src/lvt_image_features_handler.cpp:34-83,95-114,131-154,161-169). */
LVT_API int lvtk_detect(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, lvtk_keypoint *out,
                        int cap, int *n_out);

/* cv::xfeatures2d::BriefDescriptorExtractor::create()->compute -- 32 bytes, no orientation
 * (lvt/src/lvt_image_features_handler.cpp:172,190,247).  Drops keypoints within 28 px of the
 * border (order kept); out_kps/out_desc need room for n_in entries. */
LVT_API int lvtk_brief(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride,
                       const lvtk_keypoint *in, int n_in, lvtk_keypoint *out_kps, uint8_t *out_desc, int *n_out);

/* detect + brief in one call, the device never leaving the GPU in between
 * (perform_compute_features, lvt/src/lvt_image_features_handler.cpp:156-176). */
LVT_API int lvtk_extract(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, lvtk_keypoint *out_kps,
                         uint8_t *out_desc, int cap, int *n_out);

/* Projection matching of m 3-D points (in order, greedy) against one frame's features:
 * is_point_visible + find_match_index + mark_as_matched, and the "< retry_below matches:
 * reset marks, radius x2, redo" pass when retry_below > 0
 * (lvt/src/lvt_local_map.cpp:62-82,149-199; lvt/src/lvt_image_features_struct.cpp:68-120).
 * pose = camera->world (q wxyz, t).  matched_flags (n bytes) is read as the initial marks
 * and updated.  out_match_idx[i] = -2 not visible, -1 no match, else feature index.
 * out_d1/out_d2 may be NULL.  retried (may be NULL) = 1 if the second pass ran. */
LVT_API int lvtk_match_projected(lvtk_ctx *ctx, const double *pts_xyz, const uint8_t *pts_desc, int m,
                                 const double q_wxyz[4], const double t[3], const lvtk_keypoint *kps,
                                 const uint8_t *desc, int n, uint8_t *matched_flags, int retry_below,
                                 int *out_match_idx, float *out_d1, float *out_d2, int *out_count,
                                 int *retried);

/* Stereo row matching, left features in index order, skipping marked ones; marks both sides
 * (lvt/src/lvt_image_features_handler.cpp:302-323, lvt/src/lvt_image_features_struct.cpp:122-148).
 * out_query/out_train need room for n_left entries. */
LVT_API int lvtk_row_match(lvtk_ctx *ctx, const lvtk_keypoint *kps_left, const uint8_t *desc_left, int n_left,
                           uint8_t *matched_left, const lvtk_keypoint *kps_right, const uint8_t *desc_right,
                           int n_right, uint8_t *matched_right, int *out_query, int *out_train, int *n_matches);

/* Motion-only bundle adjustment: 2 passes x 5 LM iterations, Cauchy kernel, chi2 > 5.991
 * demotion (lvt/src/lvt_pnp_solver.cpp:60-128).  uv = matched keypoint locations (m x 2 float).
 * inlier_marks (m bytes, may be NULL). */
LVT_API int lvtk_solve_pose(lvtk_ctx *ctx, const double *pts_xyz, const float *uv, int m, const double q_in[4],
                            const double t_in[3], double q_out[4], double t_out[3], uint8_t *inlier_marks);

/* Linear-LS stereo triangulation with the visibility and reprojection gates
 * (lvt/src/lvt_local_map.cpp:258-329).  uv_left/uv_right: n x 2 float.  out_valid[i] = 1 if the
 * point passed every gate; out_xyz (n x 3 double) is written for every i. */
LVT_API int lvtk_triangulate(lvtk_ctx *ctx, const double q_wxyz[4], const double t[3], const float *uv_left,
                             const float *uv_right, int n, double *out_xyz, uint8_t *out_valid);

/* cv::initUndistortRectifyMap(K, D, R, P, (cols, rows), CV_32F, map_x, map_y)
 * (examples/euroc/euroc_example.cpp:106-107): for every rectified pixel the raw-image position it
 * is sampled from.  map_x / map_y: rows x cols float, tightly packed. */
LVT_API int lvtk_rectify_maps(lvtk_ctx *ctx, const lvt_rectify_c *r, int rows, int cols, float *map_x, float *map_y);

/* cv::remap(raw, out, map_x, map_y, cv::INTER_LINEAR) with the maps of lvtk_rectify_maps
 * (examples/euroc/euroc_example.cpp:142-143); the maps are never materialised.  out: rows x cols,
 * tightly packed. */
LVT_API int lvtk_rectify(lvtk_ctx *ctx, const uint8_t *raw, int rows, int cols, int stride, const lvt_rectify_c *r,
                         uint8_t *out);

#ifdef __cplusplus
}
#endif

#endif /* LVT_B200_KERNELS_H__ */
