/*
 * lvt_c.h -- public C ABI of the B200-native LVT front end.
 *
 * The first five entry points are the reference's C interface, unchanged in
 * name, argument order and meaning (reference: lvt/src/lvt_c.h:55-62,
 * behaviour: lvt/src/lvt_c.cpp:33-148).  A C caller or FFI binding written
 * against the reference's liblvt_c links against this library unmodified.
 *
 * Everything below the "additive extensions" banner is new and does not
 * alter the five reference entry points.
 *
 * Two libraries export this ABI:
 *   - lvt_b200/lib/liblvt_b200.so : the product (CUDA, sm_100a).  No CPU path.
 *   - oracle/_build/liblvt_oracle.so : the CPU oracle (test infrastructure).
 */
#ifndef LVT_B200_C_INTERFACE_H__
#define LVT_B200_C_INTERFACE_H__

#if defined(LVT_EXPORT_FUNCTIONS) && defined(__GNUC__)
#define LVT_API __attribute__((visibility("default")))
#else
#define LVT_API
#endif

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *lvt_handle;

/* ---- reference ABI (lvt/src/lvt_c.h:55-62) -------------------------------------------- */

/* sensor_type: 1 = stereo, 2 = RGB-D.  NULL on unreadable YAML / bad sensor / any failure
 * (lvt/src/lvt_c.cpp:33-48).  The YAML must carry fx,fy,cx,cy,baseline,img_width,img_height. */
LVT_API lvt_handle lvt_create(const char *config_file_name, int sensor_type);
/* lvt/src/lvt_c.cpp:50-61 */
LVT_API void lvt_destroy(lvt_handle vo_system);
/* 8-bit single-channel tightly packed images, borrowed for the call.  R = camera->world
 * rotation (row-major), t = camera position in the frame-0 world.  Outputs are left
 * untouched on an internal failure (lvt/src/lvt_c.cpp:63-88). */
LVT_API void lvt_track(lvt_handle vo_system, unsigned char *left_img, unsigned char *right_img,
                       int n_rows, int n_cols, double R[3][3], double t[3]);
/* lvt/src/lvt_c.cpp:90-134 */
LVT_API void lvt_track_with_external_corners(lvt_handle vo_system, unsigned char *left_img,
                                             unsigned char *right_img, int n_rows, int n_cols,
                                             double corners_left[][2], int n_corners_left,
                                             double corners_right[][2], int n_corners_right,
                                             double R[3][3], double t[3]);
/* 1 = not initialised, 2 = tracking, 3 = lost, -1 on failure (lvt/src/lvt_c.cpp:136-148) */
LVT_API int lvt_get_status(lvt_handle vo_system);

/* ---- additive extensions ------------------------------------------------------------ */

/* Plain-C mirror of struct lvt_parameters (lvt/src/lvt_parameters.h:29-64), same field
 * meaning and defaults (lvt/src/lvt_parameters.cpp:29-52).  bools are ints. */
typedef struct lvt_params_c
{
    float fx, fy, cx, cy;
    float baseline;
    int img_width, img_height;
    float k1, k2, p1, p2, k3;
    float near_plane_distance, far_plane_distance;
    float triangulation_ratio_test_threshold;
    float tracking_ratio_test_threshold;
    float descriptor_matching_threshold;
    int min_num_matches_for_tracking;
    int tracking_radius;
    int detection_cell_size;
    int max_keypoints_per_cell;
    int agast_threshold;
    int untracked_threshold;
    int staged_threshold;
    int enable_logging;
    int enable_visualization;
    int triangulation_policy; /* 1 decreasing matches, 2 always, 3 map size */
    float viewer_camera_size;
    int viewer_point_size;
} lvt_params_c;

/* fill with the reference defaults (lvt/src/lvt_parameters.cpp:29-52) */
LVT_API void lvt_params_default(lvt_params_c *p);
/* flat OpenCV-FileStorage YAML reader (lvt/src/lvt_parameters.cpp:54-93): missing keys read
 * as 0, exactly as cv::FileNode does.  Returns 1 on success, 0 if the file cannot be opened. */
LVT_API int lvt_params_from_file(lvt_params_c *p, const char *config_file_name);
/* lvt_system::create (lvt/src/lvt_system.cpp:70-127) from a struct, so that calibration can
 * arrive by broadcast rather than by file. */
LVT_API lvt_handle lvt_create_from_params(const lvt_params_c *p, int sensor_type);
/* lvt_system::reset (lvt/src/lvt_system.cpp:44-68) */
LVT_API void lvt_reset(lvt_handle vo_system);
/* RGB-D frame: gray u8 + depth in metres as float32 (lvt/src/lvt_system.cpp:179-183); the
 * reference's C ABI cannot carry a float image (lvt/src/lvt_c.cpp:69-70). */
LVT_API void lvt_track_rgbd(lvt_handle vo_system, const unsigned char *gray_img, const float *depth_m,
                            int n_rows, int n_cols, double R[3][3], double t[3]);

/* Per-frame bookkeeping numbers: the series the reference's value recorder registers
 * (lvt/src/lvt_system.cpp:336-350) plus what a parity test needs. */
typedef struct lvt_frame_info
{
    int frame_number;      /* as lvt_system::m_frame_number after the call */
    int state;             /* 1/2/3 */
    int n_features_left;   /* "image keypoints" */
    int n_features_right;
    int map_points_before; /* "map points count" at perform_tracking entry */
    int staged_before;     /* "staged points count" */
    int tracked;           /* "tracked map points" (matches fed to the pose solve) */
    int inliers;           /* "inlier count" */
    int map_points_after;
    int staged_after;
    int triangulated;      /* 1 if update_with_new_triangulation ran this frame */
    int new_points;        /* points it produced */
    int retried_matching;  /* 1 if the <50 matches, radius x2 pass ran */
} lvt_frame_info;

LVT_API int lvt_get_frame_info(lvt_handle vo_system, lvt_frame_info *out);
/* The reference's tracking calls are void and swallow every failure (lvt/src/lvt_c.cpp:63-134:
 * try { ... } catch (...) {}), leaving R / t untouched.  This returns the status of the last
 * lvt_track* / lvt_track_pool / lvt_track_batch call on the handle: 0 = ok, <0 = LVTK_ERR_*
 * (include/lvt_kernels.h): -1 wrong image size or an entry point that does not match the handle's
 * sensor type, -2 CUDA error, -3 a fixed capacity was exceeded (see "capacities" below). */
LVT_API int lvt_get_last_status(lvt_handle vo_system);
/* quaternion (w,x,y,z) + position of the last returned pose */
LVT_API int lvt_get_last_pose(lvt_handle vo_system, double q_wxyz[4], double t[3]);

/* Debug read-back of the state after the last track call (parity tests).
 * which: 0 = left/gray features, 1 = right features.  kps_xy: n x 2 float, desc: n x 32 bytes.
 * Returns the count (and copies min(count,cap) entries), <0 on error. */
LVT_API int lvt_debug_get_features(lvt_handle vo_system, int which, float *kps_xy, unsigned char *desc, int cap);
/* which: 0 = map points, 1 = staged points.  xyz: n x 3 double, counters/ages/match_idx: n int. */
LVT_API int lvt_debug_get_points(lvt_handle vo_system, int which, double *xyz, unsigned char *desc,
                                 int *counters, int *ages, int *match_idx, int cap);

/* ---- capacities ----------------------------------------------------------------------------
 * The reference keeps keypoints and map points in std::vectors that grow without bound
 * (lvt/src/lvt_local_map.cpp:331-353).  Here:
 *  - the local map and the staged-point store GROW: a frame that could overflow them is refused on
 *    the device before it changes anything, the stores are doubled (contents kept) and the frame
 *    runs again -- invisible to the caller apart from the time it takes;
 *  - keypoints per image are bounded by construction (ANMS keeps ~max_keypoints_per_cell + 1 per
 *    detection tile): the feature capacity is 2 x tiles x (max_keypoints_per_cell + 1), at least
 *    4096 and at most 24 576 (two owner arrays of the greedy matcher must fit in 227 KB of shared
 *    memory).  An image that yields more features than that fails the call with status -3
 *    (lvt_get_last_status), outputs untouched.
 * lvt_debug_point_capacity: current capacity of the map / staged stores (0 for the oracle). */
LVT_API int lvt_debug_point_capacity(lvt_handle vo_system);

/* Replace the 256 BRIEF test pairs, process-wide.  pairs[i] = {dy1, dx1, dy2, dx2},
 * |offset| <= 24, in the order of opencv_contrib's generated_32.i.  NULL restores the built-in
 * table.  Handles created afterwards use the new table; live handles (on any device, driven from
 * any thread) switch to it at their next tracking call -- do not call this while a tracking call
 * is in flight on another thread.  Returns 0 on success. */
LVT_API int lvt_set_brief_pairs(const signed char pairs[256][4]);

/* ---- in-pipeline stereo rectification (new) ---------------------------
 * The EuRoC driver rectifies every raw frame on the CPU before tracking it:
 * cv::initUndistortRectifyMap(K, D, R, P(0:3,0:3), size, CV_32F, M1, M2) once per camera
 * (examples/euroc/euroc_example.cpp:96-107) and cv::remap(raw, rect, M1, M2, cv::INTER_LINEAR) per
 * frame and camera (:142-143).  After lvt_set_rectification the images handed to lvt_track /
 * lvt_pool_upload are the RAW camera images: they are rectified on the way in (same arithmetic:
 * float maps from fp64 pinhole + radial-tangential math, 1/32-pixel fixed-point bilinear taps with
 * 15-bit weights, constant 0 outside the image) and tracking sees the rectified pair.
 * K, R, P: row-major 3x3 (P = the left 3x3 block of the projection matrix); D = k1 k2 p1 p2 k3.
 * Passing NULL for both cameras switches rectification off.  Returns 0 on success. */
typedef struct lvt_rectify_c
{
    double K[9];
    double D[5];
    double R[9];
    double P[9];
} lvt_rectify_c;
LVT_API int lvt_set_rectification(lvt_handle vo_system, const lvt_rectify_c *left, const lvt_rectify_c *right);

/* ---- resident-frame streaming (new; throughput path) -------------------------------------
 * The reference processes one frame per call and blocks (lvt/src/lvt_c.cpp:63-88).  When the
 * frames are already in device memory the same per-frame pipeline can run back to back with the
 * feature extraction of frame t+1 overlapped with the tracking of frame t (compute_features is
 * state-free, lvt/src/lvt_image_features_handler.cpp:196-209), and with the map maintenance of frame
 * t (staged points, triangulation: it only appends to the map) overlapped with the projection
 * matching of frame t+1 over the points that were there before.  Results are
 * identical to calling lvt_track (lvt_track_rgbd) on the same frames in the same order. */
LVT_API int lvt_pool_reserve(lvt_handle vo_system, int n_frames);
/* copy one stereo pair (tightly packed u8) into pool slot `frame` */
LVT_API int lvt_pool_upload(lvt_handle vo_system, int frame, const unsigned char *left, const unsigned char *right);
/* track pool frames [first_frame, first_frame + n_frames) in order; poses: n x 12 doubles
 * (R row-major, then t), infos: n entries; either may be NULL.  Returns 0 on success. */
LVT_API int lvt_track_pool(lvt_handle vo_system, int first_frame, int n_frames, double *poses, lvt_frame_info *infos);

/* RGB-D frame into pool slot `frame` (gray u8 + depth in metres, float32) */
LVT_API int lvt_pool_upload_rgbd(lvt_handle vo_system, int frame, const unsigned char *gray, const float *depth_m);

/* ---- batches of frames in HOST memory (new; the look-ahead path of SURVEY.md section 8b / 8f-2) -------
 * n_frames consecutive frames the caller holds in host memory, as lvt_track takes them (8-bit, tightly
 * packed; left[i] / right[i] point at frame i).  Equivalent to n_frames consecutive lvt_track calls on the
 * same frames -- same poses, same map, bit for bit -- but the frames go through the device in groups:
 * staging / H2D of later frames and their feature extraction (one launch per group of frames: the
 * extraction is state-free, lvt/src/lvt_image_features_handler.cpp:196-209) run behind the tracking of
 * earlier ones.  Page-locked buffers (lvt_alloc_pinned, cudaHostRegister) are read by the DMA engine
 * directly; pageable ones are staged.  Blocks until every pose is known.  poses: n x 12 doubles (R
 * row-major, then t), infos: n entries; either may be NULL.  Returns 0 on success (see lvt_get_last_status). */
LVT_API int lvt_track_batch(lvt_handle vo_system, int n_frames, const unsigned char *const *left,
                            const unsigned char *const *right, int n_rows, int n_cols, double *poses, lvt_frame_info *infos);
LVT_API int lvt_track_batch_rgbd(lvt_handle vo_system, int n_frames, const unsigned char *const *gray,
                                 const float *const *depth_m, int n_rows, int n_cols, double *poses, lvt_frame_info *infos);
/* page-locked host memory for frames handed to lvt_track* (optional: any host memory is accepted) */
LVT_API void *lvt_alloc_pinned(size_t bytes);
LVT_API void lvt_free_pinned(void *p);

/* device time (CUDA events, ms) of the last lvt_track_pool / lvt_track_batch call, and kernels launched so far */
LVT_API double lvt_last_batch_ms(lvt_handle vo_system);
LVT_API long lvt_launch_count(void);
/* per-kernel device time, measured with CUDA event pairs on the launching stream */
LVT_API void lvt_set_profiling(int on);
LVT_API int lvt_get_kernel_times(double *ms, long *counts, int cap); /* returns the number of kernels */
LVT_API void lvt_reset_kernel_times(void);
LVT_API const char *lvt_kernel_name(int id);
/* last error text of the calling thread (CUDA library only; "" for the oracle) */
LVT_API const char *lvtk_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* LVT_B200_C_INTERFACE_H__ */
