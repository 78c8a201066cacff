/*
 * CPU ORACLE (test infrastructure) -- pose helpers and the constant-velocity motion model.
 * Restates lvt/src/lvt_pose.cpp:28-43 and lvt/src/lvt_motion_model.cpp:34-65; the Eigen
 * quaternion operations those call (slerp, inverse, product, normalize, toRotationMatrix)
 * are restated from Eigen 3's Geometry/Quaternion.h (Eigen is not in this image).
 */
#include "lvto.h"

namespace lvto
{

Quat qslerp(const Quat &a, double t, const Quat &b)
{
    const double one = 1.0 - 2.220446049250313e-16;
    const double d = qdot(a, b);
    const double abs_d = std::fabs(d);
    double scale0, scale1;
    if (abs_d >= one)
    {
        scale0 = 1.0 - t;
        scale1 = t;
    }
    else
    {
        const double theta = std::acos(abs_d);
        const double sin_theta = std::sin(theta);
        scale0 = std::sin((1.0 - t) * theta) / sin_theta;
        scale1 = std::sin((t * theta)) / sin_theta;
    }
    if (d < 0)
        scale1 = -scale1;
    return {scale0 * a.w + scale1 * b.w, scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y,
            scale0 * a.z + scale1 * b.z};
}

Mat34 world_to_camera(const Pose &pose)
{
    const Mat3 Rt = transpose(qmat(pose.q));
    Mat34 W;
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++)
            W.m[i][j] = Rt.m[i][j];
        W.m[i][3] = (-Rt.m[i][0]) * pose.p.x + (-Rt.m[i][1]) * pose.p.y + (-Rt.m[i][2]) * pose.p.z;
    }
    return W;
}

Pose right_camera_pose(const Pose &left, double baseline)
{
    const Mat3 R = qmat(left.q);
    Pose r;
    r.q = left.q;
    r.p = mul(R, Vec3{baseline, 0.0, 0.0}) + left.p;
    return r;
}

void MotionModel::reset()
{
    last_q = Quat{};
    angular_velocity = Quat{};
    last_position = Vec3{};
    linear_velocity = Vec3{};
}

Pose MotionModel::predict_next_pose(const Pose &current)
{
    Vec3 new_lin = current.p - last_position;
    new_lin = (new_lin + linear_velocity) * 0.5;

    const Quat current_q = current.q;
    const Quat diff = qmul(current_q, qinverse(last_q));
    Quat new_ang = qnormalized(qslerp(diff, 0.5, angular_velocity));

    last_q = current_q;
    angular_velocity = new_ang;
    last_position = current.p;
    linear_velocity = new_lin;

    Pose out;
    out.p = last_position + linear_velocity;
    out.q = qnormalized(qmul(current_q, new_ang));
    return out;
}

} // namespace lvto
