/*
 * CPU ORACLE -- TEST INFRASTRUCTURE (see lvto.h).
 *
 * Stereo rectification as the EuRoC driver does it before every track() call:
 *   cv::initUndistortRectifyMap(K, D, R, P(0:3,0:3), size, CV_32F, M1, M2)   examples/euroc/euroc_example.cpp:106-107
 *   cv::remap(raw, rect, M1, M2, cv::INTER_LINEAR)                           examples/euroc/euroc_example.cpp:142-143
 * OpenCV (imgproc/src/undistort.dispatch.cpp, imgwarp.cpp) is not part of /root/reference; the
 * published algorithm is restated here and PINNED bit-exactly against the genuine code through
 * Python cv2 4.13: tests/golden/rectify_cv2.npz (tools/make_golden.py) and, where cv2 is importable,
 * live in tests/test_oracle_cv2.py.
 */
#include "lvto.h"
#include <climits>

namespace lvto
{

/* (P R)^-1 as cv::invert does it for a 3x3 double matrix: cofactors times 1/det */
bool rectify_inverse(const lvt_rectify_c &r, double ir[9])
{
    double m[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
        {
            double s = 0;
            for (int k = 0; k < 3; k++)
                s += r.P[3 * i + k] * r.R[3 * k + j];
            m[i][j] = s;
        }
    double d = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
               m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    if (d == 0.0 || !std::isfinite(d))
        return false;
    d = 1.0 / d;
    ir[0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * d;
    ir[1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * d;
    ir[2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * d;
    ir[3] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * d;
    ir[4] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * d;
    ir[5] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * d;
    ir[6] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * d;
    ir[7] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * d;
    ir[8] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * d;
    return true;
}

/* initUndistortRectifyMap, one pixel (u, v) of the rectified image -> position in the raw image */
static inline void rectify_source(const lvt_rectify_c &r, const double ir[9], int u, int v, float *mx, float *my)
{
    const double fx = r.K[0], fy = r.K[4], cx = r.K[2], cy = r.K[5];
    const double k1 = r.D[0], k2 = r.D[1], p1 = r.D[2], p2 = r.D[3], k3 = r.D[4];
    const double _x = ((double)v * ir[1] + ir[2]) + (double)u * ir[0];
    const double _y = ((double)v * ir[4] + ir[5]) + (double)u * ir[3];
    const double _w = ((double)v * ir[7] + ir[8]) + (double)u * ir[6];
    const double w = 1.0 / _w, x = _x * w, y = _y * w;
    const double x2 = x * x, y2 = y * y;
    const double r2 = x2 + y2, _2xy = 2 * x * y;
    const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2; /* k4..k6 = 0: the denominator is 1 */
    const double xd = (x * kr + p1 * _2xy) + p2 * (r2 + 2 * x2);
    const double yd = (y * kr + p1 * (r2 + 2 * y2)) + p2 * _2xy;
    *mx = (float)(xd * fx + cx);
    *my = (float)(yd * fy + cy);
}

void rectify_maps(const lvt_rectify_c &r, int rows, int cols, float *map_x, float *map_y)
{
    double ir[9];
    if (!rectify_inverse(r, ir))
    {
        std::fill(map_x, map_x + (size_t)rows * cols, -1.f);
        std::fill(map_y, map_y + (size_t)rows * cols, -1.f);
        return;
    }
    for (int v = 0; v < rows; v++)
        for (int u = 0; u < cols; u++)
            rectify_source(r, ir, u, v, &map_x[(size_t)v * cols + u], &map_y[(size_t)v * cols + u]);
}

/* cvRound(float) as SSE cvtss2si: nearest-even, INT_MIN for NaN and out-of-range values */
static inline int cv_round(float v)
{
    if (!(v >= -2147483648.f && v < 2147483648.f))
        return INT_MIN;
    return (int)std::nearbyintf(v);
}
static inline int saturate_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

/* remap, INTER_LINEAR, BORDER_CONSTANT(0), 8-bit: the float map is quantised to 1/32 pixel
 * (INTER_BITS = 5), the four taps are weighted with (32-fx)(32-fy)*32 ... (exactly the 15-bit table
 * of initInterTab2D, whose entries sum to 1 << 15 without correction for the bilinear kernel) and the
 * sum is rounded by (s + (1 << 14)) >> 15 */
static inline uint8_t remap_pixel(const uint8_t *raw, int rows, int cols, int stride, float mx, float my)
{
    const int sx = cv_round(mx * 32.f), sy = cv_round(my * 32.f);
    const int ix = saturate_short(sx >> 5), iy = saturate_short(sy >> 5);
    const int fx = sx & 31, fy = sy & 31;
    auto tap = [&](int yy, int xx) -> int {
        return (yy >= 0 && yy < rows && xx >= 0 && xx < cols) ? raw[(size_t)yy * stride + xx] : 0;
    };
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const int s = w00 * tap(iy, ix) + w01 * tap(iy, ix + 1) + w10 * tap(iy + 1, ix) + w11 * tap(iy + 1, ix + 1);
    return (uint8_t)((s + (1 << 14)) >> 15);
}

void rectify_image(const lvt_rectify_c &r, const uint8_t *raw, int rows, int cols, int stride, uint8_t *out)
{
    double ir[9];
    if (!rectify_inverse(r, ir))
    {
        std::fill(out, out + (size_t)rows * cols, (uint8_t)0);
        return;
    }
    for (int v = 0; v < rows; v++)
        for (int u = 0; u < cols; u++)
        {
            float mx, my;
            rectify_source(r, ir, u, v, &mx, &my);
            out[(size_t)v * cols + u] = remap_pixel(raw, rows, cols, stride, mx, my);
        }
}

} // namespace lvto
