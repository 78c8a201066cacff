/*
 * CPU ORACLE (test infrastructure) -- local map: projection matching, staging, culling,
 * triangulation.  Restates lvt/src/lvt_local_map.cpp:60-123 (bounds), :62-82
 * (is_point_visible), :136-229 (find_matches), :231-256 (triangulate_rgbd), :258-329
 * (triangulate), :331-353, :355-391, :393-413.
 *
 * Third-party pieces restated from their published behaviour (not in /root/reference):
 * Eigen::JacobiSVD(...).solve at :292 -> a one-sided Jacobi SVD least-squares solve (same
 * minimum-norm least-squares solution, agreement to rounding); cv::undistortPoints at :116 ->
 * OpenCV 3.x's 5-iteration inverse of the radial-tangential model.
 */
#include "lvto.h"
#include <algorithm>

namespace lvto
{

/* cv::undistortPoints(src, dst, K, dist, noArray(), K) for one point, OpenCV 3.x:
 * 5 fixed-point iterations in double, result stored as float. */
static void undistort_point(const lvt_params_c &p, float u, float v, float *ou, float *ov)
{
    const double fx = p.fx, fy = p.fy, cx = p.cx, cy = p.cy;
    const double k1 = p.k1, k2 = p.k2, p1 = p.p1, p2 = p.p2, k3 = p.k3;
    double x = ((double)u - cx) / fx, y = ((double)v - cy) / fy;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++)
    {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ou = (float)(x * fx + cx);
    *ov = (float)(y * fy + cy);
}

/* lvt/src/lvt_local_map.cpp:84-123 (per instance here; file-static in the reference, :60) */
ImageBounds compute_bounds(const lvt_params_c &p)
{
    ImageBounds b;
    if (std::fabs(p.k1) < 1e-5)
    {
        b.min_x = 0.0f;
        b.max_x = (float)p.img_width;
        b.min_y = 0.0f;
        b.max_y = (float)p.img_height;
    }
    else
    {
        float x[4], y[4];
        undistort_point(p, 0.0f, 0.0f, &x[0], &y[0]);
        undistort_point(p, (float)p.img_width, 0.0f, &x[1], &y[1]);
        undistort_point(p, 0.0f, (float)p.img_height, &x[2], &y[2]);
        undistort_point(p, (float)p.img_width, (float)p.img_height, &x[3], &y[3]);
        b.min_x = std::min(x[0], x[2]);
        b.max_x = std::max(x[1], x[3]);
        b.min_y = std::min(y[0], y[1]);
        b.max_y = std::max(y[2], y[3]);
    }
    return b;
}

void undistort_keypoints(const lvt_params_c &p, std::vector<Keypoint> &kps)
{
    for (Keypoint &k : kps)
        undistort_point(p, k.x, k.y, &k.x, &k.y);
}

/* lvt/src/lvt_local_map.cpp:62-82 */
bool is_point_visible(const Vec3 &pt, const Mat34 &w2c, const lvt_params_c &p, const ImageBounds &b, double *u,
                      double *v)
{
    const double xc = w2c.m[0][0] * pt.x + w2c.m[0][1] * pt.y + w2c.m[0][2] * pt.z + w2c.m[0][3];
    const double yc = w2c.m[1][0] * pt.x + w2c.m[1][1] * pt.y + w2c.m[1][2] * pt.z + w2c.m[1][3];
    const double zc = w2c.m[2][0] * pt.x + w2c.m[2][1] * pt.y + w2c.m[2][2] * pt.z + w2c.m[2][3];
    if (zc < p.near_plane_distance || zc > p.far_plane_distance)
        return false;
    const double inv_z = 1.0 / zc;
    const double uu = p.fx * xc * inv_z + p.cx;
    const double vv = p.fy * yc * inv_z + p.cy;
    if (uu < b.min_x || uu > b.max_x || vv < b.min_y || vv > b.max_y)
        return false;
    *u = uu;
    *v = vv;
    return true;
}

/* lvt/src/lvt_local_map.cpp:136-229 */
int LocalMap::find_matches(const Pose &cam_pose, FeatureSet *left, std::vector<Vec3> *out_points,
                           std::vector<int> *out_matches_left)
{
    const Mat34 cml = world_to_camera(cam_pose);
    int matches_count = 0;
    const int M = (int)map_points.size();
    std::vector<int> matches(M, -2);
    std::vector<double> pu(M), pv(M);
    retried = false;

    for (int i = 0; i < M; i++)
    {
        double u, v;
        if (!is_point_visible(map_points[i].pos, cml, params, bounds, &u, &v))
        {
            map_points[i].counter += 1;
            matches[i] = -2;
            continue;
        }
        pu[i] = u;
        pv[i] = v;
        float d1, d2;
        const int idx = left->find_match_index(u, v, map_points[i].desc, &d1, &d2);
        matches[i] = idx;
        if (idx != -1)
        {
            matches_count++;
            left->matched[idx] = 1;
        }
    }

    if (matches_count < 50 /* LVT_N_MATCHES_TH, lvt/src/lvt_definitions.h:34 */)
    {
        retried = true;
        matches_count = 0;
        left->reset_matched();
        const int original_radius = left->tracking_radius;
        left->tracking_radius = 2 * original_radius; /* cell_search_radius is NOT recomputed */
        for (int i = 0; i < M; i++)
        {
            if (matches[i] == -2)
                continue;
            float d1, d2;
            const int idx = left->find_match_index(pu[i], pv[i], map_points[i].desc, &d1, &d2);
            matches[i] = idx;
            if (idx != -1)
            {
                matches_count++;
                left->matched[idx] = 1;
            }
        }
        left->tracking_radius = original_radius;
    }

    for (int i = 0; i < M; i++)
    {
        map_points[i].match_idx = matches[i];
        if (matches[i] == -2)
            continue;
        if (matches[i] == -1)
        {
            map_points[i].counter += 1;
            continue;
        }
        map_points[i].age += 1;
        out_points->push_back(map_points[i].pos);
        out_matches_left->push_back(matches[i]);
    }
    return matches_count;
}

/* lvt/src/lvt_local_map.cpp:231-256 -- note the fp32 back-projection */
void LocalMap::triangulate_rgbd(const Pose &cam_pose, FeatureSet *img, std::vector<MapPoint> *out)
{
    const float inv_fx = 1.0f / params.fx;
    const float inv_fy = 1.0f / params.fy;
    const Mat3 R = qmat(cam_pose.q);
    for (int i = 0, n = img->size(); i < n; i++)
    {
        const float u = img->kps[i].x, v = img->kps[i].y;
        const float z = img->depths[i];
        const float x = (u - params.cx) * z * inv_fx;
        const float y = (v - params.cy) * z * inv_fy;
        const Vec3 pc{(double)x, (double)y, (double)z};
        MapPoint mp;
        mp.pos = mul(R, pc) + cam_pose.p;
        mp.desc = img->desc[i];
        mp.counter = 0;
        mp.age = 0;
        out->push_back(mp);
    }
}

/* minimum-norm least-squares solution of A[:, 0:3] x = -A[:, 3] by one-sided Jacobi SVD */
Vec3 solve_ls_4x3(const double Ain[4][4])
{
    double a[4][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, b[4];
    for (int i = 0; i < 4; i++)
    {
        for (int j = 0; j < 3; j++)
            a[i][j] = Ain[i][j];
        b[i] = -Ain[i][3];
    }
    for (int sweep = 0; sweep < 30; sweep++)
    {
        bool rotated = false;
        for (int p = 0; p < 2; p++)
        {
            for (int q = p + 1; q < 3; q++)
            {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 4; i++)
                {
                    alpha += a[i][p] * a[i][p];
                    beta += a[i][q] * a[i][q];
                    gamma += a[i][p] * a[i][q];
                }
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-15 * std::sqrt(alpha * beta))
                    continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 4; i++)
                {
                    const double ap = a[i][p], aq = a[i][q];
                    a[i][p] = c * ap - s * aq;
                    a[i][q] = s * ap + c * aq;
                }
                for (int i = 0; i < 3; i++)
                {
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        }
        if (!rotated)
            break;
    }
    double s2[3], ab[3], s2max = 0;
    for (int j = 0; j < 3; j++)
    {
        s2[j] = 0;
        ab[j] = 0;
        for (int i = 0; i < 4; i++)
        {
            s2[j] += a[i][j] * a[i][j];
            ab[j] += a[i][j] * b[i];
        }
        s2max = std::max(s2max, s2[j]);
    }
    double x[3] = {0, 0, 0};
    /* Eigen's default rank threshold: sigma <= eps * max(rows, cols) * sigma_max is dropped */
    const double thr = 4.0 * 2.220446049250313e-16;
    for (int j = 0; j < 3; j++)
    {
        if (s2[j] <= thr * thr * s2max)
            continue;
        const double w = ab[j] / s2[j];
        for (int i = 0; i < 3; i++)
            x[i] += V[i][j] * w;
    }
    return {x[0], x[1], x[2]};
}

/* lvt/src/lvt_local_map.cpp:258-329 */
void LocalMap::triangulate(const Pose &cam_pose, FeatureSet *left, FeatureSet *right, std::vector<MapPoint> *out)
{
    std::vector<int> query, train;
    row_match_all(left, right, &query, &train);
    if (query.empty())
        return;

    const Pose cam_pose_right = right_camera_pose(cam_pose, params.baseline);
    const Mat34 cml = world_to_camera(cam_pose);
    const Mat34 cmr = world_to_camera(cam_pose_right);
    const double cx = params.cx, cy = params.cy;
    const double inv_fx = 1.0 / params.fx, inv_fy = 1.0 / params.fy;

    for (size_t i = 0; i < query.size(); i++)
    {
        const Keypoint &u1 = left->kps[query[i]];
        const Keypoint &u2 = right->kps[train[i]];
        const double u1_x = (u1.x - cx) * inv_fx;
        const double u1_y = (u1.y - cy) * inv_fy;
        const double u2_x = (u2.x - cx) * inv_fx;
        const double u2_y = (u2.y - cy) * inv_fy;

        double A[4][4];
        for (int c = 0; c < 4; c++)
        {
            A[0][c] = u1_x * cml.m[2][c] - cml.m[0][c];
            A[1][c] = u1_y * cml.m[2][c] - cml.m[1][c];
            A[2][c] = u2_x * cmr.m[2][c] - cmr.m[0][c];
            A[3][c] = u2_y * cmr.m[2][c] - cmr.m[1][c];
        }
        const Vec3 world_pt = solve_ls_4x3(A);

        double ul, vl, ur, vr;
        if (!is_point_visible(world_pt, cml, params, bounds, &ul, &vl) ||
            !is_point_visible(world_pt, cmr, params, bounds, &ur, &vr))
            continue;
        {
            const double ex = ul - u1.x, ey = vl - u1.y;
            if ((ex * ex + ey * ey) > 5.991 /* LVT_REPROJECTION_TH2 */)
                continue;
        }
        {
            const double ex = ur - u2.x, ey = vr - u2.y;
            if ((ex * ex + ey * ey) > 5.991)
                continue;
        }
        MapPoint mp;
        mp.pos = world_pt;
        mp.desc = left->desc[query[i]];
        mp.counter = 0;
        mp.age = 0;
        out->push_back(mp);
    }
}

/* lvt/src/lvt_local_map.cpp:331-353 */
void LocalMap::update_with_new_triangulation(const Pose &cam_pose, FeatureSet *left, FeatureSet *right,
                                             bool dont_stage)
{
    std::vector<MapPoint> fresh;
    if (!left->depths.empty())
        triangulate_rgbd(cam_pose, left, &fresh);
    else
        triangulate(cam_pose, left, right, &fresh);
    last_new_points = (int)fresh.size();
    if (dont_stage || params.staged_threshold == 0 || (int)map_points.size() < 250 /* LVT_N_MAP_POINTS */)
        map_points.insert(map_points.end(), fresh.begin(), fresh.end());
    else
        staged_points.insert(staged_points.end(), fresh.begin(), fresh.end());
}

/* lvt/src/lvt_local_map.cpp:355-391 */
void LocalMap::update_staged_map_points(const Pose &cam_pose, FeatureSet *left)
{
    const Mat34 cml = world_to_camera(cam_pose);
    std::vector<MapPoint> remain;
    for (size_t i = 0; i < staged_points.size(); i++)
    {
        MapPoint &mp = staged_points[i];
        double u, v;
        float d1, d2;
        int idx = -1;
        if (!is_point_visible(mp.pos, cml, params, bounds, &u, &v) ||
            (idx = left->find_match_index(u, v, mp.desc, &d1, &d2)) == -1)
            continue; /* erased */
        left->matched[idx] = 1;
        mp.counter += 1;
        if (mp.counter == params.staged_threshold || (int)map_points.size() < 250)
            map_points.push_back(mp); /* upgraded (and erased from the staged list) */
        else
            remain.push_back(mp);
    }
    staged_points.swap(remain);
}

/* lvt/src/lvt_local_map.cpp:393-413 */
void LocalMap::clean_untracked_points(FeatureSet *left)
{
    const int th = params.untracked_threshold;
    std::vector<MapPoint> cleaned;
    cleaned.reserve(map_points.size());
    for (const MapPoint &mp : map_points)
    {
        if (mp.counter >= th)
        {
            if (mp.match_idx >= 0)
                left->matched[mp.match_idx] = 0;
        }
        else
            cleaned.push_back(mp);
    }
    cleaned.swap(map_points);
}

} // namespace lvto
