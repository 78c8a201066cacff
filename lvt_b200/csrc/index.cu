// Feature container build-up: the 25-px hash grid of lvt_image_features_struct::init
// (lvt/src/lvt_image_features_struct.cpp:35-66, .h:82-85) as a CSR, a row CSR for the stereo band
// search (:124-137), cleared match marks (:62), and the RGB-D depth gate
// (lvt/src/lvt_image_features_handler.cpp:249-294).
#include "index.cuh"

namespace lvtb
{

struct IndexArgs
{
    const FeatDev *feats;
    CamParams cam;
};

__global__ void __launch_bounds__(1024) index_kernel(IndexArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    extern __shared__ int s_dyn[];
    __shared__ int s_scan[34];
    block_build_index(a.feats[blockIdx.x], a.cam, s_dyn, s_scan);
}

int launch_index(const FeatDev *d_feats, int n_images, const CamParams &cam, cudaStream_t stream)
{
    const int bytes = index_smem_ints(cam) * (int)sizeof(int);
    // once per device: the largest dynamic size the device allows (the size of a launch depends on the image)
    static DeviceOnce once;
    if (int rc = once.run([](int dev) {
            int optin = 0;
            LVT_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            cudaFuncAttributes fa;
            LVT_CUDA_TRY(cudaFuncGetAttributes(&fa, index_kernel));
            LVT_CUDA_TRY(cudaFuncSetAttribute(index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              optin - (int)fa.sharedSizeBytes));
            return (int)LVTK_OK;
        }))
        return rc;
    IndexArgs ia{d_feats, cam};
    LVT_TIMED(stream, K_INDEX, launch_chained(index_kernel, dim3(n_images), dim3(1024), bytes, stream, ia));
    LVT_LAUNCH_CHECK(stream, "index_kernel");
    return LVTK_OK;
}

// ---- RGB-D: keep corners with a valid depth, attach it, undistort the survivors ----------------
struct DepthArgs
{
    FeatDev f;
    const float *depth; // rows x cols metres
    int cols;
    float near_plane, far_plane;
    int undistort;
    float fx, fy, cx, cy, k1, k2, p1, p2, k3;
};

// cv::undistortPoints(src, dst, K, dist, noArray(), K): 5 fixed-point iterations in fp64
__device__ float2 undistort_point(const DepthArgs &a, float2 p)
{
    const double fx = a.fx, fy = a.fy, cx = a.cx, cy = a.cy;
    double x = ((double)p.x - cx) / fx, y = ((double)p.y - cy) / fy;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++)
    {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + (((double)a.k3 * r2 + (double)a.k2) * r2 + (double)a.k1) * r2);
        const double dX = 2 * (double)a.p1 * x * y + (double)a.p2 * (r2 + 2 * x * x);
        const double dY = (double)a.p1 * (r2 + 2 * y * y) + 2 * (double)a.p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    return make_float2((float)(x * fx + cx), (float)(y * fy + cy));
}

__global__ void __launch_bounds__(1024) depth_gate_kernel(DepthArgs a)
{
    __shared__ int s_scan[34];
    const FeatDev &f = a.f;
    const int n = *f.n;
    int running = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x)
    {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        float2 p = make_float2(0, 0);
        float r = 0, d = 0;
        uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
        if (i < n)
        {
            p = f.xy[i];
            d = a.depth[(size_t)(int)p.y * a.cols + (int)p.x]; // Mat::at<float>(kp.pt.y, kp.pt.x): truncation
            keep = d >= a.near_plane && d <= a.far_plane;
            if (keep)
            {
                r = f.resp[i];
                d0 = *reinterpret_cast<const uint4 *>(f.desc + 8 * (size_t)i);
                d1 = *reinterpret_cast<const uint4 *>(f.desc + 8 * (size_t)i + 4);
            }
        }
        int total;
        const int pos = block_exclusive_scan(keep, s_scan, &total);
        if (keep)
        {
            const int o = running + pos;
            f.xy[o] = a.undistort ? undistort_point(a, p) : p;
            f.resp[o] = r;
            f.depth[o] = d;
            *reinterpret_cast<uint4 *>(f.desc + 8 * (size_t)o) = d0;
            *reinterpret_cast<uint4 *>(f.desc + 8 * (size_t)o + 4) = d1;
        }
        running += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *f.n = running;
}

int launch_depth_gate(const FeatDev &f, const float *d_depth, const lvt_params_c &p, cudaStream_t stream)
{
    DepthArgs a{f, d_depth, p.img_width, p.near_plane_distance, p.far_plane_distance, fabsf(p.k1) > 1e-5f ? 1 : 0,
                p.fx, p.fy, p.cx, p.cy, p.k1, p.k2, p.p1, p.p2, p.k3};
    depth_gate_kernel<<<1, 1024, 0, stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "depth_gate_kernel");
    return LVTK_OK;
}

} // namespace lvtb
