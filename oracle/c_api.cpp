/*
 * CPU ORACLE (test infrastructure) -- frame pipeline and the exported C ABI.
 *
 * System restates lvt/src/lvt_system.cpp:34-68 (ctor/reset), :70-127 (create), :157-207
 * (track), :209-250 (track_with_external_corners), :252-306 (perform_tracking), :308-334
 * (triangulation policies); compute_features* restate
 * lvt/src/lvt_image_features_handler.cpp:156-225 (stereo, 2 threads) and :227-300 (RGB-D).
 * The extern "C" block restates lvt/src/lvt_c.cpp:33-148 and exports the seam ABI of
 * include/lvt_kernels.h with the same symbols as the CUDA library.
 */
#define LVT_EXPORT_FUNCTIONS
#include "lvto.h"
#include <algorithm>
#include <climits>
#include <deque>
#include <map>
#include <thread>

namespace lvto
{


struct System
{
    lvt_params_c params;
    int sensor = 1;
    LocalMap map;
    MotionModel motion;
    Pose last_pose;
    int frame_number = 0;
    int state = 1;
    std::deque<int> last_matches;
    lvt_frame_info info;
    FeatureSet left, right; /* kept after the call for the debug getters */
    bool rectify = false;   /* lvt_set_rectification: the images of lvt_track are raw */
    lvt_rectify_c rect[2];
    std::vector<uint8_t> rect_buf[2];

    explicit System(const lvt_params_c &p, int sensor_type) : params(p), sensor(sensor_type)
    {
        map.params = p;
        map.bounds = compute_bounds(p);
        reset();
    }

    /* lvt/src/lvt_system.cpp:44-68 */
    void reset()
    {
        map.map_points.clear();
        map.staged_points.clear();
        motion.reset();
        last_pose = Pose{};
        frame_number = 0;
        last_matches = std::deque<int>(3, INT_MAX);
        state = 1;
        std::memset(&info, 0, sizeof(info));
        info.state = 1;
    }

    /* perform_compute_features, lvt/src/lvt_image_features_handler.cpp:156-176 */
    void compute_one(const Image &img, FeatureSet *out)
    {
        std::vector<Keypoint> kps;
        detect_corners(img, params, kps);
        std::vector<Desc> desc;
        brief_compute(img, kps, desc);
        out->init(img.rows, img.cols, kps, desc, params);
    }
    /* perform_compute_descriptors_only, :178-194 */
    void describe_one(const Image &img, const double (*corners)[2], int n, FeatureSet *out)
    {
        std::vector<Keypoint> kps(n);
        for (int i = 0; i < n; i++)
            kps[i] = Keypoint{(float)corners[i][0], (float)corners[i][1], 0.0f};
        std::vector<Desc> desc;
        brief_compute(img, kps, desc);
        out->init(img.rows, img.cols, kps, desc, params);
    }
    /* compute_features_rgbd, :227-300 */
    void compute_rgbd(const Image &gray, const float *depth, FeatureSet *out)
    {
        std::vector<Keypoint> kps;
        detect_corners(gray, params, kps);
        std::vector<Desc> desc;
        brief_compute(gray, kps, desc);
        std::vector<float> depths;
        std::vector<Keypoint> fk;
        std::vector<Desc> fd;
        for (size_t i = 0; i < kps.size(); i++)
        {
            const float d = depth[(size_t)(int)kps[i].y * gray.cols + (int)kps[i].x];
            if (d >= params.near_plane_distance && d <= params.far_plane_distance)
            {
                depths.push_back(d);
                fk.push_back(kps[i]);
                fd.push_back(desc[i]);
            }
        }
        if (std::fabs(params.k1) > 1e-5)
            undistort_keypoints(params, fk);
        out->init(gray.rows, gray.cols, fk, fd, params, &depths);
    }

    bool need_new_triangulation()
    {
        if (params.triangulation_policy == 2)
            return true;
        if (params.triangulation_policy == 3)
            return (int)map.map_points.size() < 1000;
        /* triangulation_policy_decreasing_matches, lvt/src/lvt_system.cpp:313-324 */
        const float ratio = 0.99;
        for (int i = 3 - 1; i > 0; --i)
            if (float(last_matches[i]) > ratio * float(last_matches[i - 1]))
                return false;
        return true;
    }

    /* lvt/src/lvt_system.cpp:252-306 */
    Pose perform_tracking(const Pose &estimated, bool *is_tracking)
    {
        info.map_points_before = (int)map.map_points.size();
        info.staged_before = (int)map.staged_points.size();
        std::vector<Vec3> pts;
        std::vector<int> matches_left;
        map.find_matches(estimated, &left, &pts, &matches_left);
        info.retried_matching = map.retried;
        const int matches_count = (int)pts.size();
        info.tracked = matches_count;
        if (matches_count < params.min_num_matches_for_tracking)
        {
            *is_tracking = false;
            return last_pose;
        }
        last_matches.push_back(matches_count);
        last_matches.pop_front();

        std::vector<float> uv(2 * pts.size());
        for (size_t i = 0; i < pts.size(); i++)
        {
            uv[2 * i] = left.kps[matches_left[i]].x;
            uv[2 * i + 1] = left.kps[matches_left[i]].y;
        }
        std::vector<uint8_t> marks;
        const Pose optimized = solve_pose(params, estimated, pts, uv, &marks);
        info.inliers = 0;
        for (uint8_t m : marks)
            info.inliers += m;

        map.clean_untracked_points(&left);
        if (params.staged_threshold > 0)
            map.update_staged_map_points(optimized, &left);
        if (need_new_triangulation())
        {
            map.update_with_new_triangulation(optimized, &left, &right, false);
            info.triangulated = 1;
            info.new_points = map.last_new_points;
        }
        *is_tracking = true;
        return optimized;
    }

    Pose finish_frame()
    {
        info.n_features_left = left.size();
        info.n_features_right = right.size();
        if (state == 1)
        {
            const Pose identity;
            map.update_with_new_triangulation(identity, &left, &right, true);
            state = 2;
            last_matches[0] = (int)map.map_points.size();
            info.triangulated = 1;
            info.new_points = map.last_new_points;
            return identity;
        }
        bool is_tracking = false;
        const Pose predicted = motion.predict_next_pose(last_pose);
        const Pose computed = perform_tracking(predicted, &is_tracking);
        if (!is_tracking)
        {
            state = 3;
            return last_pose;
        }
        last_pose = computed;
        return computed;
    }

    void end_info()
    {
        info.frame_number = frame_number;
        info.state = state;
        info.map_points_after = (int)map.map_points.size();
        info.staged_after = (int)map.staged_points.size();
    }

    /* lvt/src/lvt_system.cpp:157-207 */
    Pose track(const Image &img1, const Image &img2, const float *depth)
    {
        std::memset(&info, 0, sizeof(info));
        frame_number++;
        if (state == 3)
        {
            end_info();
            return last_pose;
        }
        left = FeatureSet();
        right = FeatureSet();
        if (sensor == 1)
        {
            std::thread th([&]() { compute_one(img2, &right); });
            compute_one(img1, &left);
            th.join();
        }
        else
        {
            compute_rgbd(img1, depth, &left);
        }
        const Pose r = finish_frame();
        end_info();
        return r;
    }

    /* lvt/src/lvt_system.cpp:209-250 */
    Pose track_external(const Image &img1, const Image &img2, const double (*cl)[2], int nl, const double (*cr)[2],
                        int nr)
    {
        std::memset(&info, 0, sizeof(info));
        frame_number++;
        if (state == 3)
        {
            end_info();
            return last_pose;
        }
        left = FeatureSet();
        right = FeatureSet();
        std::thread th([&]() { describe_one(img2, cr, nr, &right); });
        describe_one(img1, cl, nl, &left);
        th.join();
        const Pose r = finish_frame();
        end_info();
        return r;
    }
};

} // namespace lvto

using namespace lvto;

struct lvtk_ctx
{
    lvt_params_c params;
    ImageBounds bounds;
};

namespace
{
struct HostPool
{
    std::vector<std::vector<uint8_t>> left, right;
    std::vector<std::vector<float>> depth;
};
std::map<void *, HostPool> g_pools;
} // namespace

static void write_pose(const Pose &pose, double R[3][3], double t[3])
{
    const Mat3 m = qmat(pose.q);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i][j] = m.m[i][j];
    t[0] = pose.p.x;
    t[1] = pose.p.y;
    t[2] = pose.p.z;
}

extern "C"
{

/* ---- reference ABI (lvt/src/lvt_c.cpp:33-148) -------------------------------------------- */
LVT_API lvt_handle lvt_create(const char *config_file_name, int sensor_type)
{
    System *vo = nullptr;
    try
    {
        lvt_params_c p;
        if (params_from_file(&p, config_file_name) && (sensor_type == 1 || sensor_type == 2))
            vo = new System(p, sensor_type);
    }
    catch (...)
    {
    }
    return static_cast<lvt_handle>(vo);
}

LVT_API lvt_handle lvt_create_from_params(const lvt_params_c *p, int sensor_type)
{
    System *vo = nullptr;
    try
    {
        if (p && (sensor_type == 1 || sensor_type == 2))
            vo = new System(*p, sensor_type);
    }
    catch (...)
    {
    }
    return static_cast<lvt_handle>(vo);
}

LVT_API void lvt_destroy(lvt_handle h)
{
    try
    {
        g_pools.erase(h);
        delete static_cast<System *>(h);
    }
    catch (...)
    {
    }
}

LVT_API void lvt_reset(lvt_handle h)
{
    try
    {
        static_cast<System *>(h)->reset();
    }
    catch (...)
    {
    }
}

LVT_API void lvt_track(lvt_handle h, unsigned char *left, unsigned char *right, int n_rows, int n_cols, double R[3][3],
                       double t[3])
{
    try
    {
        System *vo = static_cast<System *>(h);
        Image l{left, n_rows, n_cols, n_cols}, r{right, n_rows, n_cols, n_cols};
        if (vo->rectify)
        {
            /* examples/euroc/euroc_example.cpp:142-143 */
            const unsigned char *raw[2] = {left, right};
            for (int k = 0; k < 2; k++)
            {
                vo->rect_buf[k].resize((size_t)n_rows * n_cols);
                rectify_image(vo->rect[k], raw[k], n_rows, n_cols, n_cols, vo->rect_buf[k].data());
            }
            l.data = vo->rect_buf[0].data();
            r.data = vo->rect_buf[1].data();
        }
        write_pose(vo->track(l, r, nullptr), R, t);
    }
    catch (...)
    {
    }
}

LVT_API void lvt_track_rgbd(lvt_handle h, const unsigned char *gray, const float *depth_m, int n_rows, int n_cols,
                            double R[3][3], double t[3])
{
    try
    {
        System *vo = static_cast<System *>(h);
        const Image g{gray, n_rows, n_cols, n_cols};
        write_pose(vo->track(g, g, depth_m), R, t);
    }
    catch (...)
    {
    }
}

LVT_API void lvt_track_with_external_corners(lvt_handle h, unsigned char *left, unsigned char *right, int n_rows,
                                             int n_cols, double corners_left[][2], int n_left,
                                             double corners_right[][2], int n_right, double R[3][3], double t[3])
{
    try
    {
        System *vo = static_cast<System *>(h);
        const Image l{left, n_rows, n_cols, n_cols}, r{right, n_rows, n_cols, n_cols};
        write_pose(vo->track_external(l, r, corners_left, n_left, corners_right, n_right), R, t);
    }
    catch (...)
    {
    }
}

LVT_API int lvt_get_status(lvt_handle h)
{
    try
    {
        return static_cast<System *>(h)->state;
    }
    catch (...)
    {
    }
    return -1;
}

/* the oracle has no failure mode of its own: a call that went wrong threw and left the outputs untouched */
LVT_API int lvt_get_last_status(lvt_handle h) { return h ? 0 : -1; }

/* ---- extensions ------------------------------------------------------------------------ */
LVT_API void lvt_params_default(lvt_params_c *p) { params_default(p); }
LVT_API int lvt_params_from_file(lvt_params_c *p, const char *f) { return params_from_file(p, f); }

LVT_API int lvt_get_frame_info(lvt_handle h, lvt_frame_info *out)
{
    if (!h || !out)
        return -1;
    *out = static_cast<System *>(h)->info;
    return 0;
}

LVT_API int lvt_get_last_pose(lvt_handle h, double q[4], double t[3])
{
    if (!h)
        return -1;
    const Pose &p = static_cast<System *>(h)->last_pose;
    q[0] = p.q.w;
    q[1] = p.q.x;
    q[2] = p.q.y;
    q[3] = p.q.z;
    t[0] = p.p.x;
    t[1] = p.p.y;
    t[2] = p.p.z;
    return 0;
}

LVT_API int lvt_debug_get_features(lvt_handle h, int which, float *kps_xy, unsigned char *desc, int cap)
{
    if (!h)
        return -1;
    const FeatureSet &f = which ? static_cast<System *>(h)->right : static_cast<System *>(h)->left;
    const int n = f.size();
    for (int i = 0; i < std::min(n, cap); i++)
    {
        if (kps_xy)
        {
            kps_xy[2 * i] = f.kps[i].x;
            kps_xy[2 * i + 1] = f.kps[i].y;
        }
        if (desc)
            std::memcpy(desc + 32 * (size_t)i, f.desc[i].b, 32);
    }
    return n;
}

LVT_API int lvt_debug_get_points(lvt_handle h, int which, double *xyz, unsigned char *desc, int *counters, int *ages,
                                 int *match_idx, int cap)
{
    if (!h)
        return -1;
    const std::vector<MapPoint> &v =
        which ? static_cast<System *>(h)->map.staged_points : static_cast<System *>(h)->map.map_points;
    const int n = (int)v.size();
    for (int i = 0; i < std::min(n, cap); i++)
    {
        if (xyz)
        {
            xyz[3 * i] = v[i].pos.x;
            xyz[3 * i + 1] = v[i].pos.y;
            xyz[3 * i + 2] = v[i].pos.z;
        }
        if (desc)
            std::memcpy(desc + 32 * (size_t)i, v[i].desc.b, 32);
        if (counters)
            counters[i] = v[i].counter;
        if (ages)
            ages[i] = v[i].age;
        if (match_idx)
            match_idx[i] = v[i].match_idx;
    }
    return n;
}

LVT_API int lvt_debug_point_capacity(lvt_handle) { return 0; } /* std::vector: unbounded */

LVT_API int lvt_set_brief_pairs(const signed char pairs[256][4])
{
    if (pairs)
        for (int i = 0; i < 256; i++)
            for (int j = 0; j < 4; j++)
                if (pairs[i][j] < -24 || pairs[i][j] > 24)
                    return -1;
    brief_set_pairs(pairs);
    return 0;
}

LVT_API int lvt_set_rectification(lvt_handle h, const lvt_rectify_c *left, const lvt_rectify_c *right)
{
    System *vo = static_cast<System *>(h);
    if (!vo || (left == nullptr) != (right == nullptr))
        return -1;
    vo->rectify = left != nullptr;
    if (left)
    {
        double ir[9];
        if (!rectify_inverse(*left, ir) || !rectify_inverse(*right, ir))
        {
            vo->rectify = false;
            return -1;
        }
        vo->rect[0] = *left;
        vo->rect[1] = *right;
    }
    return 0;
}

/* resident-frame streaming: the oracle keeps the "pool" in host memory and runs lvt_track per frame */
LVT_API int lvt_pool_reserve(lvt_handle h, int n_frames)
{
    if (!h || n_frames <= 0)
        return -1;
    HostPool &p = g_pools[h];
    p.left.assign(n_frames, {});
    p.right.assign(n_frames, {});
    p.depth.assign(n_frames, {});
    return 0;
}
LVT_API int lvt_pool_upload_rgbd(lvt_handle h, int frame, const unsigned char *gray, const float *depth_m)
{
    auto it = g_pools.find(h);
    if (it == g_pools.end() || frame < 0 || frame >= (int)it->second.left.size())
        return -1;
    const lvt_params_c &prm = static_cast<System *>(h)->params;
    const size_t n = (size_t)prm.img_width * prm.img_height;
    it->second.left[frame].assign(gray, gray + n);
    it->second.depth[frame].assign(depth_m, depth_m + n);
    return 0;
}
/* host batches: the oracle simply runs the blocking call frame by frame */
LVT_API int lvt_track_batch(lvt_handle h, int n_frames, const unsigned char *const *left, const unsigned char *const *right,
                            int n_rows, int n_cols, double *poses, lvt_frame_info *infos)
{
    System *vo = static_cast<System *>(h);
    if (!vo || n_frames <= 0 || !left || !right || vo->sensor != 1)
        return -1;
    for (int i = 0; i < n_frames; i++)
    {
        double R[3][3], t[3];
        lvt_track(h, const_cast<unsigned char *>(left[i]), const_cast<unsigned char *>(right[i]), n_rows, n_cols, R, t);
        if (poses)
        {
            std::memcpy(poses + 12 * (size_t)i, R, sizeof(R));
            std::memcpy(poses + 12 * (size_t)i + 9, t, sizeof(t));
        }
        if (infos)
            infos[i] = vo->info;
    }
    return 0;
}
LVT_API int lvt_track_batch_rgbd(lvt_handle h, int n_frames, const unsigned char *const *gray, const float *const *depth_m,
                                 int n_rows, int n_cols, double *poses, lvt_frame_info *infos)
{
    System *vo = static_cast<System *>(h);
    if (!vo || n_frames <= 0 || !gray || !depth_m || vo->sensor != 2)
        return -1;
    for (int i = 0; i < n_frames; i++)
    {
        double R[3][3], t[3];
        lvt_track_rgbd(h, gray[i], depth_m[i], n_rows, n_cols, R, t);
        if (poses)
        {
            std::memcpy(poses + 12 * (size_t)i, R, sizeof(R));
            std::memcpy(poses + 12 * (size_t)i + 9, t, sizeof(t));
        }
        if (infos)
            infos[i] = vo->info;
    }
    return 0;
}
LVT_API void *lvt_alloc_pinned(size_t bytes) { return std::malloc(bytes ? bytes : 1); } /* no device: plain host memory */
LVT_API void lvt_free_pinned(void *p) { std::free(p); }
LVT_API int lvt_pool_upload(lvt_handle h, int frame, const unsigned char *left, const unsigned char *right)
{
    auto it = g_pools.find(h);
    if (it == g_pools.end() || frame < 0 || frame >= (int)it->second.left.size())
        return -1;
    const lvt_params_c &prm = static_cast<System *>(h)->params;
    const size_t n = (size_t)prm.img_width * prm.img_height;
    it->second.left[frame].assign(left, left + n);
    it->second.right[frame].assign(right, right + n);
    return 0;
}
LVT_API int lvt_track_pool(lvt_handle h, int first, int n, double *poses, lvt_frame_info *infos)
{
    auto it = g_pools.find(h);
    if (it == g_pools.end() || first < 0 || n <= 0 || first + n > (int)it->second.left.size())
        return -1;
    System *vo = static_cast<System *>(h);
    for (int i = 0; i < n; i++)
    {
        double R[3][3], t[3];
        if (vo->sensor == 2)
            lvt_track_rgbd(h, it->second.left[first + i].data(), it->second.depth[first + i].data(), vo->params.img_height,
                           vo->params.img_width, R, t);
        else
            lvt_track(h, it->second.left[first + i].data(), it->second.right[first + i].data(), vo->params.img_height,
                      vo->params.img_width, R, t); /* rectifies first when lvt_set_rectification is on */
        if (poses)
        {
            std::memcpy(poses + 12 * (size_t)i, R, sizeof(R));
            std::memcpy(poses + 12 * (size_t)i + 9, t, sizeof(t));
        }
        if (infos)
            infos[i] = vo->info;
    }
    return 0;
}
LVT_API double lvt_last_batch_ms(lvt_handle) { return 0.0; }
LVT_API long lvt_launch_count(void) { return 0; }
LVT_API void lvt_set_profiling(int) {}
LVT_API int lvt_get_kernel_times(double *, long *, int) { return 0; }
LVT_API void lvt_reset_kernel_times(void) {}
LVT_API const char *lvt_kernel_name(int) { return ""; }
LVT_API const char *lvtk_last_error(void) { return ""; }

/* ---- seam ABI (include/lvt_kernels.h) ---------------------------------------------------- */
LVT_API lvtk_ctx *lvtk_ctx_create(const lvt_params_c *p, int device)
{
    (void)device;
    if (!p)
        return nullptr;
    lvtk_ctx *c = new lvtk_ctx;
    c->params = *p;
    c->bounds = compute_bounds(*p);
    return c;
}
LVT_API void lvtk_ctx_destroy(lvtk_ctx *c) { delete c; }
LVT_API int lvtk_is_gpu(void) { return 0; }

static int emit(const std::vector<Keypoint> &k, lvtk_keypoint *out, int cap, int *n_out)
{
    *n_out = (int)k.size();
    if ((int)k.size() > cap)
        return LVTK_ERR_CAPACITY;
    if (!k.empty())
        std::memcpy(out, k.data(), k.size() * sizeof(Keypoint));
    return LVTK_OK;
}

LVT_API int lvtk_agast(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, int threshold, int nonmax,
                       lvtk_keypoint *out, int cap, int *n_out)
{
    if (!ctx || !img || !n_out || rows <= 0 || cols <= 0 || stride < cols)
        return LVTK_ERR_ARG;
    std::vector<Keypoint> k;
    agast_detect(Image{img, rows, cols, stride}, threshold, nonmax != 0, k);
    return emit(k, out, cap, n_out);
}

LVT_API int lvtk_detect(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, lvtk_keypoint *out, int cap,
                        int *n_out)
{
    if (!ctx || !img || !n_out || rows != ctx->params.img_height || cols != ctx->params.img_width || stride < cols)
        return LVTK_ERR_ARG;
    std::vector<Keypoint> k;
    detect_corners(Image{img, rows, cols, stride}, ctx->params, k);
    return emit(k, out, cap, n_out);
}

LVT_API int lvtk_brief(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, const lvtk_keypoint *in,
                       int n_in, lvtk_keypoint *out_kps, uint8_t *out_desc, int *n_out)
{
    if (!ctx || !img || !n_out || n_in < 0 || stride < cols)
        return LVTK_ERR_ARG;
    std::vector<Keypoint> k(in, in + n_in);
    std::vector<Desc> d;
    brief_compute(Image{img, rows, cols, stride}, k, d);
    *n_out = (int)k.size();
    if (!k.empty())
    {
        std::memcpy(out_kps, k.data(), k.size() * sizeof(Keypoint));
        std::memcpy(out_desc, d.data(), d.size() * 32);
    }
    return LVTK_OK;
}

LVT_API int lvtk_extract(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, lvtk_keypoint *out_kps,
                         uint8_t *out_desc, int cap, int *n_out)
{
    if (!ctx || !img || !n_out || rows != ctx->params.img_height || cols != ctx->params.img_width || stride < cols)
        return LVTK_ERR_ARG;
    std::vector<Keypoint> k;
    const Image im{img, rows, cols, stride};
    detect_corners(im, ctx->params, k);
    std::vector<Desc> d;
    brief_compute(im, k, d);
    *n_out = (int)k.size();
    if ((int)k.size() > cap)
        return LVTK_ERR_CAPACITY;
    if (!k.empty())
    {
        std::memcpy(out_kps, k.data(), k.size() * sizeof(Keypoint));
        std::memcpy(out_desc, d.data(), d.size() * 32);
    }
    return LVTK_OK;
}

static void make_set(const lvtk_ctx *ctx, const lvtk_keypoint *kps, const uint8_t *desc, int n, const uint8_t *flags,
                     FeatureSet *fs)
{
    std::vector<Keypoint> k(kps, kps + n);
    std::vector<Desc> d(n);
    if (n)
        std::memcpy(d.data(), desc, (size_t)n * 32);
    fs->init(ctx->params.img_height, ctx->params.img_width, k, d, ctx->params);
    if (flags)
        for (int i = 0; i < n; i++)
            fs->matched[i] = flags[i] != 0;
}

LVT_API int lvtk_match_projected(lvtk_ctx *ctx, const double *pts_xyz, const uint8_t *pts_desc, int m,
                                 const double q[4], const double t[3], const lvtk_keypoint *kps, const uint8_t *desc,
                                 int n, uint8_t *matched_flags, int retry_below, int *out_match_idx, float *out_d1,
                                 float *out_d2, int *out_count, int *retried)
{
    if (!ctx || m < 0 || n < 0 || !out_match_idx)
        return LVTK_ERR_ARG;
    FeatureSet fs;
    make_set(ctx, kps, desc, n, matched_flags, &fs);
    Pose pose;
    pose.q = Quat{q[0], q[1], q[2], q[3]};
    pose.p = Vec3{t[0], t[1], t[2]};
    const Mat34 cml = world_to_camera(pose);
    std::vector<double> pu(m), pv(m);
    int count = 0;
    auto run = [&](bool second) {
        count = 0;
        for (int i = 0; i < m; i++)
        {
            if (!second)
            {
                const Vec3 p{pts_xyz[3 * i], pts_xyz[3 * i + 1], pts_xyz[3 * i + 2]};
                if (!is_point_visible(p, cml, ctx->params, ctx->bounds, &pu[i], &pv[i]))
                {
                    out_match_idx[i] = -2;
                    continue;
                }
            }
            else if (out_match_idx[i] == -2)
                continue;
            Desc d;
            std::memcpy(d.b, pts_desc + 32 * (size_t)i, 32);
            float d1 = 0, d2 = 0;
            const int idx = fs.find_match_index(pu[i], pv[i], d, &d1, &d2);
            out_match_idx[i] = idx;
            if (out_d1)
                out_d1[i] = idx >= 0 ? d1 : 0.f;
            if (out_d2)
                out_d2[i] = idx >= 0 ? d2 : 0.f;
            if (idx != -1)
            {
                count++;
                fs.matched[idx] = 1;
            }
        }
    };
    run(false);
    int did_retry = 0;
    if (count < retry_below)
    {
        did_retry = 1;
        fs.reset_matched();
        fs.tracking_radius *= 2;
        run(true);
    }
    if (matched_flags)
        for (int i = 0; i < n; i++)
            matched_flags[i] = fs.matched[i];
    if (out_count)
        *out_count = count;
    if (retried)
        *retried = did_retry;
    return LVTK_OK;
}

LVT_API int lvtk_row_match(lvtk_ctx *ctx, const lvtk_keypoint *kl, const uint8_t *dl, int nl, uint8_t *ml,
                           const lvtk_keypoint *kr, const uint8_t *dr, int nr, uint8_t *mr, int *out_query,
                           int *out_train, int *n_matches)
{
    if (!ctx || nl < 0 || nr < 0 || !n_matches)
        return LVTK_ERR_ARG;
    FeatureSet L, R;
    make_set(ctx, kl, dl, nl, ml, &L);
    make_set(ctx, kr, dr, nr, mr, &R);
    std::vector<int> q, t;
    row_match_all(&L, &R, &q, &t);
    *n_matches = (int)q.size();
    for (size_t i = 0; i < q.size(); i++)
    {
        out_query[i] = q[i];
        out_train[i] = t[i];
    }
    if (ml)
        for (int i = 0; i < nl; i++)
            ml[i] = L.matched[i];
    if (mr)
        for (int i = 0; i < nr; i++)
            mr[i] = R.matched[i];
    return LVTK_OK;
}

LVT_API int lvtk_solve_pose(lvtk_ctx *ctx, const double *pts_xyz, const float *uv, int m, const double q_in[4],
                            const double t_in[3], double q_out[4], double t_out[3], uint8_t *inlier_marks)
{
    if (!ctx || m < 0)
        return LVTK_ERR_ARG;
    std::vector<Vec3> pts(m);
    for (int i = 0; i < m; i++)
        pts[i] = Vec3{pts_xyz[3 * i], pts_xyz[3 * i + 1], pts_xyz[3 * i + 2]};
    std::vector<float> vuv(uv, uv + 2 * (size_t)m);
    Pose init;
    init.q = Quat{q_in[0], q_in[1], q_in[2], q_in[3]};
    init.p = Vec3{t_in[0], t_in[1], t_in[2]};
    std::vector<uint8_t> marks;
    const Pose out = solve_pose(ctx->params, init, pts, vuv, &marks);
    q_out[0] = out.q.w;
    q_out[1] = out.q.x;
    q_out[2] = out.q.y;
    q_out[3] = out.q.z;
    t_out[0] = out.p.x;
    t_out[1] = out.p.y;
    t_out[2] = out.p.z;
    if (inlier_marks && m)
        std::memcpy(inlier_marks, marks.data(), m);
    return LVTK_OK;
}

LVT_API int lvtk_rectify_maps(lvtk_ctx *ctx, const lvt_rectify_c *r, int rows, int cols, float *map_x, float *map_y)
{
    if (!ctx || !r || !map_x || !map_y || rows <= 0 || cols <= 0)
        return LVTK_ERR_ARG;
    rectify_maps(*r, rows, cols, map_x, map_y);
    return LVTK_OK;
}

LVT_API int lvtk_rectify(lvtk_ctx *ctx, const uint8_t *raw, int rows, int cols, int stride, const lvt_rectify_c *r,
                         uint8_t *out)
{
    if (!ctx || !raw || !r || !out || rows <= 0 || cols <= 0 || stride < cols)
        return LVTK_ERR_ARG;
    rectify_image(*r, raw, rows, cols, stride, out);
    return LVTK_OK;
}

LVT_API int lvtk_triangulate(lvtk_ctx *ctx, const double q[4], const double t[3], const float *uv_left,
                             const float *uv_right, int n, double *out_xyz, uint8_t *out_valid)
{
    if (!ctx || n < 0)
        return LVTK_ERR_ARG;
    Pose pose;
    pose.q = Quat{q[0], q[1], q[2], q[3]};
    pose.p = Vec3{t[0], t[1], t[2]};
    const lvt_params_c &prm = ctx->params;
    const Pose pr = right_camera_pose(pose, prm.baseline);
    const Mat34 cml = world_to_camera(pose), cmr = world_to_camera(pr);
    const double cx = prm.cx, cy = prm.cy, inv_fx = 1.0 / prm.fx, inv_fy = 1.0 / prm.fy;
    for (int i = 0; i < n; i++)
    {
        const float u1x = uv_left[2 * i], u1y = uv_left[2 * i + 1], u2x = uv_right[2 * i], u2y = uv_right[2 * i + 1];
        const double a = (u1x - cx) * inv_fx, b = (u1y - cy) * inv_fy, c = (u2x - cx) * inv_fx,
                     d = (u2y - cy) * inv_fy;
        double A[4][4];
        for (int k = 0; k < 4; k++)
        {
            A[0][k] = a * cml.m[2][k] - cml.m[0][k];
            A[1][k] = b * cml.m[2][k] - cml.m[1][k];
            A[2][k] = c * cmr.m[2][k] - cmr.m[0][k];
            A[3][k] = d * cmr.m[2][k] - cmr.m[1][k];
        }
        const Vec3 w = solve_ls_4x3(A);
        out_xyz[3 * i] = w.x;
        out_xyz[3 * i + 1] = w.y;
        out_xyz[3 * i + 2] = w.z;
        double ul, vl, ur, vr;
        bool ok = is_point_visible(w, cml, prm, ctx->bounds, &ul, &vl) &&
                  is_point_visible(w, cmr, prm, ctx->bounds, &ur, &vr);
        if (ok)
        {
            const double ex = ul - u1x, ey = vl - u1y;
            ok = !((ex * ex + ey * ey) > 5.991);
        }
        if (ok)
        {
            const double ex = ur - u2x, ey = vr - u2y;
            ok = !((ex * ex + ey * ey) > 5.991);
        }
        out_valid[i] = ok ? 1 : 0;
    }
    return LVTK_OK;
}

} /* extern "C" */
