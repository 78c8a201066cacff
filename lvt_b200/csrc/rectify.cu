// Stereo rectification in front of the detector: the EuRoC driver's
//   cv::initUndistortRectifyMap(K, D, R, P(0:3,0:3), size, CV_32F, M1, M2)   examples/euroc/euroc_example.cpp:106-107
//   cv::remap(raw, rect, M1, M2, cv::INTER_LINEAR)                           examples/euroc/euroc_example.cpp:142-143
// as one kernel per image.  The maps are never stored: every thread recomputes the source position
// of its pixels in fp64 (about 40 operations, free next to the memory traffic), quantises it to
// 1/32 pixel like cv::remap's fixed-point path and blends the four taps with the 15-bit weights
// (32-fx)(32-fy)*32 ...; taps outside the raw image count as 0 (BORDER_CONSTANT).  HBM traffic is
// the raw image in, the rectified image out: 2 W H bytes per image (the taps of neighbouring
// pixels overlap and are served by L1/L2).  Bit-exact with the CPU oracle, which is pinned against
// cv2 (tests/golden/rectify_cv2.npz).
#include "extract.cuh"
#include <climits>
#include <cmath>

namespace lvtb
{

// (P R)^-1 as cv::invert does it for a 3x3 double matrix (cofactors times 1/det); host side, once
int make_rectify_dev(const lvt_rectify_c &r, RectifyDev *out)
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
        return LVTK_ERR_ARG;
    d = 1.0 / d;
    double *ir = out->ir;
    ir[0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * d;
    ir[1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * d;
    ir[2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * d;
    ir[3] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * d;
    ir[4] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * d;
    ir[5] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * d;
    ir[6] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * d;
    ir[7] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * d;
    ir[8] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * d;
    out->fx = r.K[0];
    out->fy = r.K[4];
    out->cx = r.K[2];
    out->cy = r.K[5];
    out->k1 = r.D[0];
    out->k2 = r.D[1];
    out->p1 = r.D[2];
    out->p2 = r.D[3];
    out->k3 = r.D[4];
    return LVTK_OK;
}

// initUndistortRectifyMap, one pixel (u, v) of the rectified image -> position in the raw image.
// Same operation order as oracle/rectify.cpp; the file is compiled with -fmad=false.
__device__ __forceinline__ void rectify_source(const RectifyDev &r, int u, int v, float *mx, float *my)
{
    const double _x = ((double)v * r.ir[1] + r.ir[2]) + (double)u * r.ir[0];
    const double _y = ((double)v * r.ir[4] + r.ir[5]) + (double)u * r.ir[3];
    const double _w = ((double)v * r.ir[7] + r.ir[8]) + (double)u * r.ir[6];
    const double w = 1.0 / _w, x = _x * w, y = _y * w;
    const double x2 = x * x, y2 = y * y;
    const double r2 = x2 + y2, _2xy = 2 * x * y;
    const double kr = 1 + ((r.k3 * r2 + r.k2) * r2 + r.k1) * r2;
    const double xd = (x * kr + r.p1 * _2xy) + r.p2 * (r2 + 2 * x2);
    const double yd = (y * kr + r.p1 * (r2 + 2 * y2)) + r.p2 * _2xy;
    *mx = (float)(xd * r.fx + r.cx);
    *my = (float)(yd * r.fy + r.cy);
}

// cvRound(float) as SSE cvtss2si: nearest-even, INT_MIN for NaN and out-of-range values
__device__ __forceinline__ int cv_round(float v)
{
    return (v >= -2147483648.f && v < 2147483648.f) ? __float2int_rn(v) : INT_MIN;
}

__device__ __forceinline__ uint32_t remap_pixel(const uint8_t *__restrict__ raw, int rows, int cols, int pitch, float mx,
                                                float my)
{
    const int sx = cv_round(__fmul_rn(mx, 32.f)), sy = cv_round(__fmul_rn(my, 32.f));
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
    const int fx = sx & 31, fy = sy & 31;
    const bool x0 = ix >= 0 && ix < cols, x1 = ix + 1 >= 0 && ix + 1 < cols;
    const bool y0 = iy >= 0 && iy < rows, y1 = iy + 1 >= 0 && iy + 1 < rows;
    const uint8_t *p = raw + (ptrdiff_t)iy * pitch + ix;
    const int t00 = (y0 && x0) ? __ldg(p) : 0, t01 = (y0 && x1) ? __ldg(p + 1) : 0;
    const int t10 = (y1 && x0) ? __ldg(p + pitch) : 0, t11 = (y1 && x1) ? __ldg(p + pitch + 1) : 0;
    const int s = ((32 - fx) * (32 - fy) * t00 + fx * (32 - fy) * t01 + (32 - fx) * fy * t10 + fx * fy * t11) * 32;
    return (uint32_t)((s + (1 << 14)) >> 15);
}

struct RectifyArgs
{
    const uint8_t *raw[2]; // up to two raw images, pitched like the pool
    uint8_t *dst[2];       // their pool slots
    int rows, cols, pitch;
    RectifyDev cam[2];     // camera of image i
};

constexpr int kRectThreads = 256;
constexpr int kRectRowsPerCta = 8; // 32 threads x 4 pixels wide, 8 rows

__global__ void __launch_bounds__(kRectThreads) rectify_kernel(RectifyArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    const int img = blockIdx.z;
    const RectifyDev &r = a.cam[img];
    const uint8_t *raw = a.raw[img];
    uint8_t *dst = a.dst[img];
    const int u0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int v = blockIdx.y * kRectRowsPerCta + (threadIdx.x >> 5);
    if (v >= a.rows || u0 >= a.cols)
        return;
    uint32_t px = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        uint32_t q = 0;
        if (u0 + k < a.cols)
        {
            float mx, my;
            rectify_source(r, u0 + k, v, &mx, &my);
            q = remap_pixel(raw, a.rows, a.cols, a.pitch, mx, my);
        }
        px |= q << (8 * k);
    }
    // pitch is a multiple of 128 and u0 of 4: one aligned 4-byte store (the padding columns get zeros)
    *reinterpret_cast<uint32_t *>(dst + (size_t)v * a.pitch + u0) = px;
}

struct RectifyMapArgs
{
    RectifyDev r;
    int rows, cols;
    float *map_x, *map_y;
};

__global__ void rectify_map_kernel(RectifyMapArgs a)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= a.cols)
        return;
    float mx, my;
    rectify_source(a.r, u, v, &mx, &my);
    a.map_x[(size_t)v * a.cols + u] = mx;
    a.map_y[(size_t)v * a.cols + u] = my;
}

int launch_rectify(const uint8_t *const raw[2], uint8_t *const dst[2], int n_images, const RectifyDev *const cam[2],
                   const ImagePool &pool, cudaStream_t stream)
{
    if (n_images < 1 || n_images > 2)
        return LVTK_ERR_ARG;
    RectifyArgs a{{raw[0], raw[n_images - 1]}, {dst[0], dst[n_images - 1]}, pool.rows, pool.cols, pool.pitch,
                  {*cam[0], *cam[n_images - 1]}};
    const dim3 grid((pool.cols + 127) / 128, (pool.rows + kRectRowsPerCta - 1) / kRectRowsPerCta, n_images);
    LVT_TIMED(stream, K_RECTIFY, launch_chained(rectify_kernel, grid, dim3(kRectThreads), 0, stream, a));
    LVT_LAUNCH_CHECK(stream, "rectify_kernel");
    return LVTK_OK;
}

// ---- tightly packed image (as the caller holds it, lvt/src/lvt_c.cpp:69-70) -> pitched pool slot ----------
// A page-locked caller buffer is fetched with ONE contiguous DMA (15 us for 1242x375; the pitched 2-D copy of
// the same image takes 55 us: measured, tools/probe/h2d_probe.cu) and re-pitched here: every thread writes one
// aligned 16-byte piece of a destination row from (unaligned) source bytes.
struct RepitchArgs
{
    const uint8_t *src;
    uint8_t *dst;
    int rows, cols, pitch;
};

__global__ void __launch_bounds__(128) repitch_kernel(RepitchArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    const int y = blockIdx.y, x0 = 16 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x0 >= a.cols)
        return;
    const uint8_t *s = a.src + (size_t)y * a.cols + x0;
    uint32_t w[4] = {0, 0, 0, 0};
    const int n = min(16, a.cols - x0);
    // the source row starts at an arbitrary byte: 2-byte loads when the row offset is even, bytes otherwise
    if ((((size_t)s) & 1) == 0 && n == 16)
    {
        const uint16_t *s2 = reinterpret_cast<const uint16_t *>(s);
#pragma unroll
        for (int k = 0; k < 4; k++)
            w[k] = (uint32_t)s2[2 * k] | ((uint32_t)s2[2 * k + 1] << 16);
    }
    else
        for (int k = 0; k < n; k++)
            w[k >> 2] |= (uint32_t)s[k] << (8 * (k & 3));
    *reinterpret_cast<uint4 *>(a.dst + (size_t)y * a.pitch + x0) = make_uint4(w[0], w[1], w[2], w[3]);
}

int launch_repitch(const uint8_t *d_packed, uint8_t *d_dst, int rows, int cols, int pitch, cudaStream_t stream)
{
    RepitchArgs a{d_packed, d_dst, rows, cols, pitch};
    const dim3 grid(((cols + 15) / 16 + 127) / 128, rows);
    count_launch();
    if (launch_chained(repitch_kernel, grid, dim3(128), 0, stream, a) != cudaSuccess)
        return LVTK_ERR_CUDA;
    LVT_LAUNCH_CHECK(stream, "repitch_kernel");
    return LVTK_OK;
}

int launch_rectify_maps(const RectifyDev &r, int rows, int cols, float *d_map_x, float *d_map_y, cudaStream_t stream)
{
    RectifyMapArgs a{r, rows, cols, d_map_x, d_map_y};
    rectify_map_kernel<<<dim3((cols + 127) / 128, rows), 128, 0, stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "rectify_map_kernel");
    return LVTK_OK;
}

} // namespace lvtb
