// Common device/host helpers for the sm_100a LVT front end.
#pragma once

#include "../../include/lvt_kernels.h"
#include <atomic>
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>

namespace lvtb
{

// ---------------------------------------------------------------------------------------------
// error handling: every CUDA call is checked; the C ABI turns failures into LVTK_ERR_CUDA.
// ---------------------------------------------------------------------------------------------
void set_last_error(const char *file, int line, const char *what);
const char *last_error();

#define LVT_CUDA_TRY(expr)                                                                                            \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess)                                                                                        \
        {                                                                                                             \
            lvtb::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));                                         \
            return LVTK_ERR_CUDA;                                                                                     \
        }                                                                                                             \
    } while (0)

// per-kernel timing (CUDA event pairs on the launching stream), enabled by lvt_set_profiling(1)
enum KernelId
{
    K_SCORE,
    K_NMS,
    K_TILE,
    K_GATHER,
    K_BRIEF,
    K_INDEX,
    K_TRACK_A,
    K_MAPCAND,
    K_ROWCAND,
    K_POSE,
    K_STAGEDCAND,
    K_TRACK_B,
    K_RECTIFY,
    K_TRACK_A_EARLY, // track_a_kernel, early part of the map pass (batched engine)
    K_MAPCAND_EARLY, // mapcand_kernel for the coming frame, behind the pose
    K_TAILCAND,      // mapcand_kernel over the points track_b appended
    K_COUNT
};
// Function attributes (opt-in shared memory) are per device and per process: every launcher sets
// them once per device through one of these, from whichever thread gets there first.  Distinct
// handles may be driven from distinct threads (lvt/src/lvt_c.cpp:33-148 has no shared state).
struct DeviceOnce
{
    static constexpr int kMaxDevices = 64;
    std::atomic<int> done[kMaxDevices];
    std::mutex mu;
    DeviceOnce()
    {
        for (auto &d : done)
            d.store(0);
    }
    // f(device) -> LVTK_* ; runs once per device, other threads wait for it
    template <class F>
    int run(F f)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
            return LVTK_ERR_CUDA;
        if (done[dev].load(std::memory_order_acquire))
            return LVTK_OK;
        std::lock_guard<std::mutex> lk(mu);
        if (done[dev].load(std::memory_order_relaxed))
            return LVTK_OK;
        if (int rc = f(dev))
            return rc;
        done[dev].store(1, std::memory_order_release);
        return LVTK_OK;
    }
};

bool prof_enabled();
void count_launch();
long launch_count();
void prof_begin(cudaStream_t s, int id);
void prof_end(cudaStream_t s, int id);
// Programmatic dependent launch: every kernel of the per-frame chain is launched with the
// programmatic-stream-serialization attribute and starts with LVT_GRID_DEP_SYNC(), so its CTAs are
// scheduled (and its launch latency is paid) while the previous kernel of the stream is still
// running; griddepcontrol.wait returns once that kernel has completed and its writes are visible.
template <class... KArgs>
inline cudaError_t launch_chained_impl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                       int cluster, KArgs... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
    if (cluster > 1)
    {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = (unsigned)cluster;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
template <class... KArgs, class... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args &&...args)
{
    return launch_chained_impl<KArgs...>(kernel, grid, block, smem, stream, 1, static_cast<KArgs>(args)...);
}
// the same as a launch of thread-block clusters of `cluster` CTAs along x
template <class... KArgs, class... Args>
inline cudaError_t launch_chained_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                          int cluster, Args &&...args)
{
    return launch_chained_impl<KArgs...>(kernel, grid, block, smem, stream, cluster, static_cast<KArgs>(args)...);
}
#ifdef LVT_NO_EARLY_TRIGGER
#define LVT_GRID_DEP_SYNC() cudaGridDependencySynchronize()
#else
#define LVT_GRID_DEP_SYNC()                                                                                           \
    do                                                                                                                \
    {                                                                                                                 \
        cudaGridDependencySynchronize();                                                                              \
        cudaTriggerProgrammaticLaunchCompletion();                                                                    \
    } while (0)
#endif

#define LVT_TIMED(stream, id, launch)                                                                                 \
    do                                                                                                                \
    {                                                                                                                 \
        const bool _p = lvtb::prof_enabled();                                                                         \
        lvtb::count_launch();                                                                                         \
        if (_p)                                                                                                       \
            lvtb::prof_begin(stream, id);                                                                             \
        launch;                                                                                                       \
        if (_p)                                                                                                       \
            lvtb::prof_end(stream, id);                                                                               \
    } while (0)

// LVT_B200_SYNC=1: synchronise and check after every kernel launch (debug aid)
bool debug_sync_enabled();
#define LVT_LAUNCH_CHECK(stream, name)                                                                                \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t _e = cudaGetLastError();                                                                          \
        if (_e == cudaSuccess && lvtb::debug_sync_enabled())                                                          \
            _e = cudaStreamSynchronize(stream);                                                                       \
        if (_e != cudaSuccess)                                                                                        \
        {                                                                                                             \
            lvtb::set_last_error(name, __LINE__, cudaGetErrorString(_e));                                             \
            return LVTK_ERR_CUDA;                                                                                     \
        }                                                                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------------
// reference constants (lvt/src/lvt_definitions.h:29-34)
// ---------------------------------------------------------------------------------------------
constexpr double kReprojectionTh2 = 5.991; // LVT_REPROJECTION_TH2
constexpr int kNMapPoints = 250;           // LVT_N_MAP_POINTS
constexpr int kRowSearchRadius = 2;        // LVT_ROW_MATCHING_VERTICAL_SEARCH_RADIUS
constexpr int kHashCell = 25;              // LVT_HASHING_CELL_SIZE
constexpr int kCornersLowTh = 200;         // LVT_CORNERS_LOW_TH
constexpr int kNMatchesTh = 50;            // LVT_N_MATCHES_TH
constexpr int kBriefBorder = 28;           // PATCH_SIZE/2 + KERNEL_SIZE/2 (opencv_contrib brief.cpp)

// ---------------------------------------------------------------------------------------------
// fp64 pose math on the device (Eigen formulas, see lvt/src/lvt_pose.h:34-79)
// ---------------------------------------------------------------------------------------------
struct Quat
{
    double w, x, y, z;
};
struct PoseD
{
    Quat q;
    double t[3];
};

__host__ __device__ inline Quat quat_mul(const Quat &a, const Quat &b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__host__ __device__ inline double quat_dot(const Quat &a, const Quat &b)
{
    return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
}
__host__ __device__ inline Quat quat_normalized(const Quat &q)
{
    const double n = sqrt(quat_dot(q, q));
    Quat r;
    r.w = q.w / n;
    r.x = q.x / n;
    r.y = q.y / n;
    r.z = q.z / n;
    return r;
}
__host__ __device__ inline Quat quat_inverse(const Quat &q)
{
    const double n2 = quat_dot(q, q);
    Quat r;
    if (n2 > 0)
    {
        r.w = q.w / n2;
        r.x = -q.x / n2;
        r.y = -q.y / n2;
        r.z = -q.z / n2;
    }
    else
    {
        r.w = r.x = r.y = r.z = 0;
    }
    return r;
}
// rotation matrix, row-major R[3*i+j]
__host__ __device__ inline void quat_to_mat(const Quat &q, double R[9])
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1 - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1 - (txx + tyy);
}
// world -> camera 3x4, row-major W[4*i+j]: [R^T | -R^T t]  (lvt/src/lvt_pose.cpp:36-43)
__host__ __device__ inline void world_to_camera(const PoseD &p, double W[12])
{
    double R[9];
    quat_to_mat(p.q, R);
    for (int i = 0; i < 3; i++)
    {
        const double a = R[0 + i], b = R[3 + i], c = R[6 + i]; // row i of R^T
        W[4 * i + 0] = a;
        W[4 * i + 1] = b;
        W[4 * i + 2] = c;
        W[4 * i + 3] = (-a) * p.t[0] + (-b) * p.t[1] + (-c) * p.t[2];
    }
}

// ---------------------------------------------------------------------------------------------
// camera / matching parameters handed to kernels by value
// ---------------------------------------------------------------------------------------------
struct CamParams
{
    float fx, fy, cx, cy, baseline;
    float near_plane, far_plane;
    float min_x, max_x, min_y, max_y; // image bounds (lvt/src/lvt_local_map.cpp:84-123), per instance
    int img_w, img_h;
    int cells_x, cells_y;     // 25-px hash grid
    int cell_search_radius;   // lvt/src/lvt_image_features_struct.cpp:53
    int tracking_radius;
    float tracking_ratio_th, triangulation_ratio_th, desc_dist_th;
};

// is_point_visible (lvt/src/lvt_local_map.cpp:62-82)
__device__ inline bool point_visible(const double W[12], const CamParams &c, double x, double y, double z, double *u,
                                     double *v)
{
    const double xc = W[0] * x + W[1] * y + W[2] * z + W[3];
    const double yc = W[4] * x + W[5] * y + W[6] * z + W[7];
    const double zc = W[8] * x + W[9] * y + W[10] * z + W[11];
    if (zc < (double)c.near_plane || zc > (double)c.far_plane)
        return false;
    const double inv_z = 1.0 / zc;
    const double uu = (double)c.fx * xc * inv_z + (double)c.cx;
    const double vv = (double)c.fy * yc * inv_z + (double)c.cy;
    if (uu < (double)c.min_x || uu > (double)c.max_x || vv < (double)c.min_y || vv > (double)c.max_y)
        return false;
    *u = uu;
    *v = vv;
    return true;
}

// ---------------------------------------------------------------------------------------------
// warp / block primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exclusive prefix sum of one int per thread over the block; returns the exclusive value and the
// block total through *total.  scratch: >= 33 ints of shared memory.  All threads must call.
__device__ inline int block_exclusive_scan(int v, int *scratch, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += n;
    }
    __syncthreads(); // scratch may still be read from a previous call
    if (lane == 31)
        scratch[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        int w = lane < nwarps ? scratch[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o)
                wi += n;
        }
        scratch[lane] = wi - w; // exclusive per-warp offset
        if (lane == 31)
            scratch[32] = wi;
    }
    __syncthreads();
    *total = scratch[32];
    return scratch[warp] + incl - v;
}

// Four exclusive prefix sums at once: v = four 16-bit counts packed into 64 bits (field j in bits
// 16j..16j+15; every field's block total must stay below 65536).  Returns the packed exclusive values,
// *total the packed block totals.  scratch: >= 33 uint64 of shared memory.  All threads must call.
// Used for 4 x blockDim items handled as item(j, t) = base + j * blockDim + t (coalesced accesses), whose
// order is j-major: position = sum of the totals of the fields before j + the field's exclusive value.
__device__ inline unsigned long long block_exclusive_scan4(unsigned long long v, unsigned long long *scratch,
                                                           unsigned long long *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += n;
    }
    __syncthreads(); // scratch may still be read from a previous call
    if (lane == 31)
        scratch[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        const unsigned long long w = lane < nwarps ? scratch[lane] : 0ull;
        unsigned long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o)
                wi += n;
        }
        scratch[lane] = wi - w; // exclusive per-warp offset
        if (lane == 31)
            scratch[32] = wi;
    }
    __syncthreads();
    *total = scratch[32];
    return scratch[warp] + incl - v;
}
// offsets of the four fields of a packed total (j-major order) and the position of item (j, t)
__device__ __forceinline__ int scan4_position(unsigned long long excl, unsigned long long total, int j)
{
    int before = 0;
#pragma unroll
    for (int jj = 0; jj < 4; jj++)
        if (jj < j)
            before += (int)((total >> (16 * jj)) & 0xFFFFull);
    return before + (int)((excl >> (16 * j)) & 0xFFFFull);
}
__device__ __forceinline__ int scan4_sum(unsigned long long total)
{
    return (int)(total & 0xFFFFull) + (int)((total >> 16) & 0xFFFFull) + (int)((total >> 32) & 0xFFFFull) +
           (int)((total >> 48) & 0xFFFFull);
}

// asynchronous global -> shared copies of 4 / 8 / 16 bytes (LDGSTS): no register in between, so a thread
// can have many in flight; cp_async_wait_all() makes the calling thread's copies visible to itself (a
// barrier publishes them to the CTA)
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src)
{
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async copies 4, 8 or 16 bytes");
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) + mbarrier, hand-written PTX
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// bounded wait: a TMA that never lands traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24))
            __trap();
}
// 3-D tiled load: tensor (x, y, image) -> dense box in shared memory
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
                 "l"((uint64_t)map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

} // namespace lvtb
