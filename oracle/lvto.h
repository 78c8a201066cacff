/*
 * CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A dependency-free C++17 restatement of the reference's per-frame track() path
 * (SAR-Research-Lab/lvt @ 77940fa).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this directory.
 * Nothing under lvt_b200/ includes, links or calls it.
 *
 * PARITY STATUS: "parity unpinned" against the reference's own tests -- the reference has
 * none and cannot be compiled here (OpenCV C++, opencv_contrib, Eigen and
 * g2o are absent).  What IS pinned: AGAST (segment test + score + NMS) and masked Hamming
 * top-2 are checked bit-exactly against the genuine OpenCV code through Python cv2 4.13
 * (tests/test_oracle_cv2.py and the golden vectors under tests/golden/ made by
 * tools/make_golden.py).  BRIEF bit positions use a stand-in test-pair table
 * (tools/gen_brief_pairs.py); the g2o solve follows the published algorithm of g2o tag
 * 20170730_git (types_sba, OptimizationAlgorithmLevenberg, RobustKernelCauchy,
 * LinearSolverPCG); Eigen's quaternion / slerp formulas are restated.
 */
#ifndef LVT_ORACLE_H__
#define LVT_ORACLE_H__

#include "../include/lvt_kernels.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace lvto
{

/* ---- fp64 pose math (reference: lvt/src/lvt_pose.h:34-79, Eigen::Quaterniond) ---------- */
struct Vec3
{
    double x = 0, y = 0, z = 0;
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(const Vec3 &a, double s) { return {a.x * s, a.y * s, a.z * s}; }

struct Mat3
{
    double m[3][3];
};
inline Vec3 mul(const Mat3 &A, const Vec3 &v)
{
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline Mat3 transpose(const Mat3 &A)
{
    Mat3 T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            T.m[i][j] = A.m[j][i];
    return T;
}

struct Quat
{
    double w = 1, x = 0, y = 0, z = 0;
};
/* Eigen quaternion product */
inline Quat qmul(const Quat &a, const Quat &b)
{
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline double qdot(const Quat &a, const Quat &b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
inline Quat qnormalized(const Quat &q)
{
    /* Eigen::QuaternionBase::normalize -> coeffs() / coeffs().norm() (divide, not multiply by inverse) */
    double n = std::sqrt(qdot(q, q));
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}
/* Eigen::QuaternionBase::inverse: conjugate / squaredNorm */
inline Quat qinverse(const Quat &q)
{
    double n2 = qdot(q, q);
    if (n2 > 0)
        return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
    return {0, 0, 0, 0};
}
/* Eigen::QuaternionBase::toRotationMatrix */
inline Mat3 qmat(const Quat &q)
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    Mat3 R;
    R.m[0][0] = 1 - (tyy + tzz);
    R.m[0][1] = txy - twz;
    R.m[0][2] = txz + twy;
    R.m[1][0] = txy + twz;
    R.m[1][1] = 1 - (txx + tzz);
    R.m[1][2] = tyz - twx;
    R.m[2][0] = txz - twy;
    R.m[2][1] = tyz + twx;
    R.m[2][2] = 1 - (txx + tyy);
    return R;
}
/* Eigen::QuaternionBase::slerp */
Quat qslerp(const Quat &a, double t, const Quat &b);

/* camera -> world pose (lvt/src/lvt_pose.h:51-79) */
struct Pose
{
    Quat q;
    Vec3 p;
};
/* [R^T | -R^T p]  (lvt/src/lvt_pose.cpp:36-43) */
struct Mat34
{
    double m[3][4];
};
Mat34 world_to_camera(const Pose &pose);
/* lvt/src/lvt_pose.cpp:28-34 */
Pose right_camera_pose(const Pose &left, double baseline);

/* ---- features ---------------------------------------------------------------------- */
struct Image
{
    const uint8_t *data;
    int rows, cols, stride;
    uint8_t at(int y, int x) const { return data[(size_t)y * stride + x]; }
};

typedef lvtk_keypoint Keypoint;
struct Desc
{
    uint8_t b[32];
};

/* cv::AgastFeatureDetector (OAST_9_16) on one image == one tile */
void agast_detect(const Image &img, int threshold, bool nonmax, std::vector<Keypoint> &out);
/* _adaptive_non_maximal_suppresion (lvt/src/lvt_image_features_handler.cpp:34-83) */
void anms(std::vector<Keypoint> &kps, int num_to_keep, float tx, float ty);
/* tile rects (lvt/src/lvt_image_features_handler.cpp:95-114) */
struct Rect
{
    int x, y, w, h;
};
std::vector<Rect> tile_rects(int img_w, int img_h, int cell);
/* perform_detect_corners + retry (lvt/src/lvt_image_features_handler.cpp:131-154,161-169) */
void detect_corners(const Image &img, const lvt_params_c &p, std::vector<Keypoint> &out);
/* BriefDescriptorExtractor::compute (32 bytes) */
void brief_compute(const Image &img, std::vector<Keypoint> &kps, std::vector<Desc> &desc);
void brief_set_pairs(const signed char pairs[256][4]);

/* lvt_image_features_struct (lvt/src/lvt_image_features_struct.{h,cpp}) */
struct FeatureSet
{
    std::vector<Keypoint> kps;
    std::vector<Desc> desc;
    std::vector<uint8_t> matched;
    std::vector<float> depths;
    int rows = 0, cols = 0;
    int cell_size = 25, cells_x = -1, cells_y = -1, cell_search_radius = 0;
    int tracking_radius = 0, vertical_search_radius = 2;
    float triangulation_ratio_th = 0.6f, tracking_ratio_th = 0.8f, desc_dist_th = 25.0f;
    std::vector<std::vector<int>> grid; /* cells_y * cells_x */

    void init(int rows, int cols, std::vector<Keypoint> &kps, std::vector<Desc> &desc, const lvt_params_c &p,
              const std::vector<float> *kps_depth = nullptr);
    int find_match_index(double px, double py, const Desc &d, float *d1, float *d2) const;
    int row_match(float ptx, float pty, const Desc &d) const;
    void reset_matched() { matched.assign(kps.size(), 0); }
    int size() const { return (int)kps.size(); }
};

/* masked Hamming top-2 == BFMatcher(NORM_HAMMING).knnMatch(k=2, mask) over a candidate list */
struct Top2
{
    int n = 0;          /* min(2, #candidates) */
    int idx[2] = {-1, -1};
    int dist[2] = {0, 0};
};
int hamming256(const uint8_t *a, const uint8_t *b);

/* ---- local map (lvt/src/lvt_local_map.{h,cpp}) ------------------------------------------ */
struct MapPoint
{
    Desc desc;
    Vec3 pos;
    int counter = 0;
    int age = 0;
    int match_idx = 0;
};
struct ImageBounds
{
    float min_x, max_x, min_y, max_y;
};
ImageBounds compute_bounds(const lvt_params_c &p);
void undistort_keypoints(const lvt_params_c &p, std::vector<Keypoint> &kps);
bool is_point_visible(const Vec3 &pt, const Mat34 &w2c, const lvt_params_c &p, const ImageBounds &b, double *u,
                      double *v);
/* lvt/src/lvt_local_map.cpp:280-293: 4x3 least squares A[:, :3] x = -A[:, 3] */
Vec3 solve_ls_4x3(const double A[4][4]);

struct LocalMap
{
    lvt_params_c params;
    ImageBounds bounds;
    std::vector<MapPoint> map_points, staged_points;
    bool retried = false; /* last find_matches ran the radius x2 pass */
    int last_new_points = 0;

    int find_matches(const Pose &cam_pose, FeatureSet *left, std::vector<Vec3> *out_points,
                     std::vector<int> *out_matches_left);
    void triangulate(const Pose &cam_pose, FeatureSet *left, FeatureSet *right, std::vector<MapPoint> *out);
    void triangulate_rgbd(const Pose &cam_pose, FeatureSet *img, std::vector<MapPoint> *out);
    void update_with_new_triangulation(const Pose &cam_pose, FeatureSet *left, FeatureSet *right, bool dont_stage);
    void update_staged_map_points(const Pose &cam_pose, FeatureSet *left);
    void clean_untracked_points(FeatureSet *left);
};
void row_match_all(FeatureSet *left, FeatureSet *right, std::vector<int> *query, std::vector<int> *train);

/* ---- pose solver (lvt/src/lvt_pnp_solver.cpp) ------------------------------------------- */
Pose solve_pose(const lvt_params_c &p, const Pose &init, const std::vector<Vec3> &pts, const std::vector<float> &uv,
                std::vector<uint8_t> *inlier_marks);

/* ---- motion model (lvt/src/lvt_motion_model.cpp) ---------------------------------------- */
struct MotionModel
{
    Quat last_q, angular_velocity;
    Vec3 last_position, linear_velocity;
    void reset();
    Pose predict_next_pose(const Pose &current);
};

/* ---- params --------------------------------------------------------------------------- */
/* ---- stereo rectification (rectify.cpp; examples/euroc/euroc_example.cpp:106-107,142-143) */
bool rectify_inverse(const lvt_rectify_c &r, double ir[9]);
void rectify_maps(const lvt_rectify_c &r, int rows, int cols, float *map_x, float *map_y);
void rectify_image(const lvt_rectify_c &r, const uint8_t *raw, int rows, int cols, int stride, uint8_t *out);

void params_default(lvt_params_c *p);
int params_from_file(lvt_params_c *p, const char *file);

} // namespace lvto

#endif
