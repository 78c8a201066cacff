/*
 * lvt_system.hpp -- header-only C++ mirror of the reference's lvt_system (lvt/src/lvt_system.h:44-70)
 * over the C ABI of lvt_c.h.  Same method names, argument meaning and state values; images are passed
 * as views (pointer, rows, cols) instead of cv::Mat because this library has no OpenCV dependency, and
 * poses come back as lvt_pose_view {R row-major camera->world, t} instead of the Eigen-based lvt_pose.
 * Link with liblvt_b200.so (or, in tests, the CPU oracle: it exports the same symbols).
 */
#ifndef LVT_B200_SYSTEM_HPP__
#define LVT_B200_SYSTEM_HPP__

#include "lvt_c.h"
#include <cstddef>
#include <vector>

struct lvt_image_view
{
    const unsigned char *data; // 8-bit, single channel, tightly packed (lvt/src/lvt_c.cpp:69-70)
    int rows, cols;
};

struct lvt_depth_view
{
    const float *data; // metres, CV_32F in the reference (lvt/src/lvt_system.h:62)
    int rows, cols;
};

struct lvt_pose_view
{
    double R[3][3]; // camera -> world
    double t[3];    // camera position in the world of frame 0
};

struct lvt_point2f
{
    float x, y;
};

class lvt_system
{
  public:
    enum eState // lvt/src/lvt_system.h:44-50
    {
        eState_NOT_INITIALIZED = 1,
        eState_TRACKING,
        eState_LOST
    };
    enum eSensor // lvt/src/lvt_system.h:51-55
    {
        eSensor_STEREO = 1,
        eSensor_RGBD
    };

    // lvt_system::create (lvt/src/lvt_system.cpp:70-127): nullptr when the parameters are rejected or no GPU is there
    static lvt_system *create(const lvt_params_c &params, eSensor sensor_type)
    {
        lvt_handle h = lvt_create_from_params(&params, (int)sensor_type);
        if (!h)
            return nullptr;
        lvt_system *s = new lvt_system();
        s->m_handle = h;
        s->m_sensor = sensor_type;
        return s;
    }
    static void destroy(lvt_system *s)
    {
        if (!s)
            return;
        lvt_destroy(s->m_handle);
        delete s;
    }
    void reset() { lvt_reset(m_handle); }

    // stereo: two rectified grayscale images (raw ones after set_rectification)
    lvt_pose_view track(const lvt_image_view &img1, const lvt_image_view &img2)
    {
        lvt_pose_view p = identity();
        lvt_track(m_handle, const_cast<unsigned char *>(img1.data), const_cast<unsigned char *>(img2.data), img1.rows, img1.cols,
                  p.R, p.t);
        return p;
    }
    // RGB-D: grayscale image + depth image in metres
    lvt_pose_view track(const lvt_image_view &gray, const lvt_depth_view &depth)
    {
        lvt_pose_view p = identity();
        lvt_track_rgbd(m_handle, gray.data, depth.data, gray.rows, gray.cols, p.R, p.t);
        return p;
    }
    lvt_pose_view track_with_external_corners(const lvt_image_view &left, const lvt_image_view &right,
                                              const std::vector<lvt_point2f> &corners_left,
                                              const std::vector<lvt_point2f> &corners_right)
    {
        std::vector<double> cl(2 * corners_left.size() + 2), cr(2 * corners_right.size() + 2);
        for (size_t i = 0; i < corners_left.size(); i++)
            cl[2 * i] = corners_left[i].x, cl[2 * i + 1] = corners_left[i].y;
        for (size_t i = 0; i < corners_right.size(); i++)
            cr[2 * i] = corners_right[i].x, cr[2 * i + 1] = corners_right[i].y;
        lvt_pose_view p = identity();
        lvt_track_with_external_corners(m_handle, const_cast<unsigned char *>(left.data), const_cast<unsigned char *>(right.data),
                                        left.rows, left.cols, reinterpret_cast<double(*)[2]>(cl.data()), (int)corners_left.size(),
                                        reinterpret_cast<double(*)[2]>(cr.data()), (int)corners_right.size(), p.R, p.t);
        return p;
    }
    // a run of consecutive stereo frames the caller already holds (a recorded sequence, a look-ahead buffer):
    // the same poses as calling track() on each pair in turn, with the upload and feature extraction of later
    // frames running behind the tracking of earlier ones (lvt_track_batch).  Empty on failure (last_status()).
    std::vector<lvt_pose_view> track_batch(const std::vector<lvt_image_view> &left, const std::vector<lvt_image_view> &right)
    {
        std::vector<lvt_pose_view> out;
        const size_t n = left.size();
        if (n == 0 || right.size() != n)
            return out;
        std::vector<const unsigned char *> l(n), r(n);
        for (size_t i = 0; i < n; i++)
            l[i] = left[i].data, r[i] = right[i].data;
        std::vector<double> poses(12 * n);
        if (lvt_track_batch(m_handle, (int)n, l.data(), r.data(), left[0].rows, left[0].cols, poses.data(), nullptr) != 0)
            return out;
        out.resize(n);
        for (size_t i = 0; i < n; i++)
        {
            for (int k = 0; k < 9; k++)
                out[i].R[k / 3][k % 3] = poses[12 * i + k];
            for (int k = 0; k < 3; k++)
                out[i].t[k] = poses[12 * i + 9 + k];
        }
        return out;
    }
    // 0 = the last tracking call succeeded; < 0: LVTK_ERR_* (the reference swallows failures, lvt_c.cpp:63-134)
    int last_status() const { return lvt_get_last_status(m_handle); }

    // examples/euroc/euroc_example.cpp:96-107,142-143 moved behind track()
    bool set_rectification(const lvt_rectify_c *left, const lvt_rectify_c *right)
    {
        return lvt_set_rectification(m_handle, left, right) == 0;
    }

    eSensor get_sensor_type() const { return m_sensor; }
    eState get_state() const { return (eState)lvt_get_status(m_handle); }
    bool should_quit() const { return false; } // the reference's viewer asks to quit; there is no viewer here
    lvt_frame_info frame_info() const
    {
        lvt_frame_info fi = {};
        lvt_get_frame_info(m_handle, &fi);
        return fi;
    }

    lvt_system(const lvt_system &) = delete;
    lvt_system &operator=(const lvt_system &) = delete;

  private:
    lvt_system() = default;
    ~lvt_system() = default;
    static lvt_pose_view identity()
    {
        lvt_pose_view p = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, {0, 0, 0}};
        return p;
    }
    lvt_handle m_handle = nullptr;
    eSensor m_sensor = eSensor_STEREO;
};

#endif /* LVT_B200_SYSTEM_HPP__ */
