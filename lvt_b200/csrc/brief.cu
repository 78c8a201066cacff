// BRIEF-32 (256 binary tests on 9x9 box sums) on sm_100a.
//
// Replaces cv::xfeatures2d::BriefDescriptorExtractor::compute as called at
// lvt/src/lvt_image_features_handler.cpp:172 (also :190, :247): 32 bytes, no orientation.
//
// One warp per keypoint.  TMA stages the keypoint's 57x57 support (an 80x57 box starting at the
// 16-byte aligned column left of it) into shared memory; the warp turns it into a 58x58 integral image held in u16
// -- arithmetic modulo 2^16 is exact because every 9x9 box sum is <= 81*255 = 20655 -- and
// each lane then evaluates 8 tests.  The 32 outcomes of a group of tests are packed with one
// __ballot_sync per 32-bit descriptor word; the lane -> test mapping is chosen so that the
// ballot word already has OpenCV's bit order (test 8b+j -> byte b, bit 7-j).
// The next keypoint's patch is in flight while the current one is being integrated.
#include "index.cuh"

namespace lvtb
{

constexpr int kIntegStride = 58;                                       // u16 elements; 29 words -> conflict-free rows
constexpr int kPatchBytes = kPatchW * kPatchH;                         // 3648
constexpr int kPatchSlot = (kPatchBytes + 127) / 128 * 128;            // 3712
constexpr int kIntegBytes = 58 * kIntegStride * 2;                     // 6728
constexpr int kIntegSlot = (kIntegBytes + 127) / 128 * 128;            // 6784
constexpr int kBriefWarps = 4;
constexpr int kBriefSmem = kBriefWarps * (kPatchSlot + kIntegSlot) + 128;

// The test-pair table as the kernel wants it -- per (word w, lane l): offsets of the two boxes'
// top-left integral corner, lo = first box.  Every context owns a copy in device memory (no
// process-wide device symbol: contexts on any number of devices and threads stay independent).

static const signed char kDefaultPairs[256][4] = {
#include "brief_pairs.inc"
};

void make_brief_offsets(const signed char pairs_in[256][4], uint32_t h[8 * 32])
{
    const signed char(*pairs)[4] = pairs_in ? pairs_in : kDefaultPairs;
    for (int w = 0; w < 8; w++)
    {
        for (int l = 0; l < 32; l++)
        {
            const int t = 32 * w + (l & ~7) + (7 - (l & 7)); // ballot bit l <-> test t
            const int o1 = (pairs[t][0] + 24) * kIntegStride + (pairs[t][1] + 24);
            const int o2 = (pairs[t][2] + 24) * kIntegStride + (pairs[t][3] + 24);
            h[w * 32 + l] = (uint32_t)o1 | ((uint32_t)o2 << 16);
        }
    }
}

struct BriefArgs
{
    const int *slots;
    const FeatDev *feats;
    int rows, cols;
    int with_index; // 1: the last CTA of every image builds the feature index (index.cuh) instead of describing
    CamParams cam;
    const uint32_t *offsets; // [8 * 32], make_brief_offsets
};

__global__ void __launch_bounds__(kBriefWarps * 32) brief_kernel(const __grid_constant__ CUtensorMap tmap, BriefArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.with_index && blockIdx.x == gridDim.x - 1)
    {
        // the hash grid / row CSR need the keypoint positions only: built next to the descriptors
        // instead of in a kernel of its own behind them
        __shared__ int s_scan[34];
        block_build_index(a.feats[blockIdx.y], a.cam, reinterpret_cast<int *>(smem), s_scan);
        return;
    }
    uint8_t *patch = smem + warp * (kPatchSlot + kIntegSlot);
    uint16_t *integ = reinterpret_cast<uint16_t *>(patch + kPatchSlot);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kBriefWarps * (kPatchSlot + kIntegSlot)) + warp;

    const int b = blockIdx.y;
    const FeatDev f = a.feats[b];
    const int n = *f.n;
    const int slot = a.slots[b];
    const int stride = (gridDim.x - a.with_index) * kBriefWarps;
    int kp = blockIdx.x * kBriefWarps + warp;

    uint32_t offs[8];
#pragma unroll
    for (int w = 0; w < 8; w++)
        offs[w] = a.offsets[w * 32 + lane];

    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    // row 0 and column 0 of the integral image are zero and never written again
    for (int i = lane; i < 58; i += 32)
    {
        integ[i] = 0;
        integ[i * kIntegStride] = 0;
    }
    __syncwarp();

    auto centre = [&](int k, int *X, int *Y) {
        const float2 p = f.xy[k];
        // (int)(pt + 0.5) with a double add, clamped to the last valid centre (see oracle/brief.cpp)
        *X = min((int)((double)p.x + 0.5), a.cols - kBriefBorder - 1);
        *Y = min((int)((double)p.y + 0.5), a.rows - kBriefBorder - 1);
    };
    // the box starts at the 16-byte aligned column at or left of X-28 (TMA alignment rule)
    auto issue = [&](int k) {
        int X, Y;
        centre(k, &X, &Y);
        mbar_arrive_expect_tx(bar, kPatchBytes);
        tma_load_3d(patch, &tmap, (X - kBriefBorder) & ~15, Y - kBriefBorder, slot, bar);
    };

    uint32_t phase = 0;
    if (kp < n && lane == 0)
        issue(kp);
    for (; kp < n; kp += stride)
    {
        int Xc, Yc;
        centre(kp, &Xc, &Yc);
        const int xoff = (Xc - kBriefBorder) & 15; // column of the support inside the box
        mbar_wait(bar, phase);
        phase ^= 1;
        // pass 1: running column sums, lane = column (bytes of a row are consecutive: no conflicts)
#pragma unroll
        for (int cc = 0; cc < 2; cc++)
        {
            const int c = lane + 32 * cc;
            if (c < 57)
            {
                uint32_t acc = 0;
#pragma unroll 19
                for (int r = 0; r < 57; r++)
                {
                    acc += patch[r * kPatchW + xoff + c];
                    integ[(r + 1) * kIntegStride + c + 1] = (uint16_t)acc;
                }
            }
        }
        __syncwarp();
        // the patch buffer is free: put the next keypoint's support in flight
        if (lane == 0 && kp + stride < n)
            issue(kp + stride);
        // pass 2: running row sums, lane = row (stride 29 words: no conflicts)
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
        {
            const int r = 1 + lane + 32 * rr;
            if (r <= 57)
            {
                uint16_t *row = integ + r * kIntegStride;
                uint32_t acc = 0;
#pragma unroll 19
                for (int c = 1; c <= 57; c++)
                {
                    acc += row[c];
                    row[c] = (uint16_t)acc;
                }
            }
        }
        __syncwarp();
        // 8 ballots = 8 descriptor words
        uint32_t mine = 0;
#pragma unroll
        for (int w = 0; w < 8; w++)
        {
            const uint32_t o1 = offs[w] & 0xFFFFu, o2 = offs[w] >> 16;
            const uint16_t s1 = (uint16_t)(integ[o1 + 9 * kIntegStride + 9] - integ[o1 + 9] - integ[o1 + 9 * kIntegStride] + integ[o1]);
            const uint16_t s2 = (uint16_t)(integ[o2 + 9 * kIntegStride + 9] - integ[o2 + 9] - integ[o2 + 9 * kIntegStride] + integ[o2]);
            const uint32_t word = __ballot_sync(0xffffffffu, s1 < s2);
            if (lane == w)
                mine = word;
        }
        if (lane < 8)
            f.desc[(size_t)kp * 8 + lane] = mine;
        __syncwarp(); // integ is rewritten by the next keypoint
    }
}

// KeyPointsFilter::runByImageBorder for arbitrary (external) keypoints: keep iff the
// round-half-even integer position lies in [28, W-28) x [28, H-28); order preserved.
struct FilterArgs
{
    const float2 *src_xy;
    const float *src_resp; // may be null
    const int *src_n;
    int src_stride; // elements between images in src arrays
    const FeatDev *feats;
    int *error;
    int rows, cols;
};

__global__ void __launch_bounds__(1024) border_filter_kernel(FilterArgs a)
{
    __shared__ int s_scan[34];
    const int b = blockIdx.x;
    const FeatDev f = a.feats[b];
    const int n_in = a.src_n[b];
    const float2 *xy = a.src_xy + (size_t)b * a.src_stride;
    const bool any = a.rows > 2 * kBriefBorder && a.cols > 2 * kBriefBorder;
    int running = 0;
    for (int i0 = 0; i0 < n_in; i0 += blockDim.x)
    {
        const int i = i0 + threadIdx.x;
        int keep = 0;
        float2 p = make_float2(0, 0);
        if (i < n_in)
        {
            p = xy[i];
            const int rx = __float2int_rn(p.x), ry = __float2int_rn(p.y);
            keep = any && rx >= kBriefBorder && rx < a.cols - kBriefBorder && ry >= kBriefBorder &&
                   ry < a.rows - kBriefBorder;
        }
        int total;
        const int pos = block_exclusive_scan(keep, s_scan, &total);
        if (keep && running + pos < f.cap)
        {
            f.xy[running + pos] = p;
            f.resp[running + pos] = a.src_resp ? a.src_resp[(size_t)b * a.src_stride + i] : 0.0f;
        }
        running += total;
    }
    if (threadIdx.x == 0)
    {
        if (running > f.cap)
        {
            *a.error = LVTK_ERR_CAPACITY;
            running = f.cap;
        }
        *f.n = running;
    }
}

int launch_border_filter(const float2 *src_xy, const float *src_resp, const int *src_n, int src_stride,
                         const FeatDev *d_feats, int n_images, int rows, int cols, int *error, cudaStream_t stream)
{
    FilterArgs fa{src_xy, src_resp, src_n, src_stride, d_feats, error, rows, cols};
    border_filter_kernel<<<n_images, 1024, 0, stream>>>(fa);
    LVT_LAUNCH_CHECK(stream, "border_filter_kernel");
    return LVTK_OK;
}

bool brief_can_index(const CamParams &cam) { return index_smem_ints(cam) * (int)sizeof(int) <= kBriefSmem; }

int launch_brief(const ImagePool &pool, const int *d_slots, int n_images, const FeatDev *d_feats, const uint32_t *d_offsets,
                 cudaStream_t stream, const CamParams *index_cam)
{
    static DeviceOnce once;
    if (int rc = once.run([](int) {
            LVT_CUDA_TRY(cudaFuncSetAttribute(brief_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBriefSmem));
            return (int)LVTK_OK;
        }))
        return rc;
    BriefArgs ba{d_slots, d_feats, pool.rows, pool.cols, 0, CamParams{}, d_offsets};
    if (index_cam)
    {
        if (index_smem_ints(*index_cam) * (int)sizeof(int) > kBriefSmem)
            return LVTK_ERR_ARG; // the caller launches index_kernel instead (brief_can_index)
        ba.with_index = 1;
        ba.cam = *index_cam;
    }
    // persistent-style: 148 SMs x 4 CTAs of 4 warps per image; warps stride over the keypoints
    LVT_TIMED(stream, K_BRIEF, launch_chained(brief_kernel, dim3(148 * 2 + ba.with_index, n_images), dim3(kBriefWarps * 32), kBriefSmem, stream, pool.tmap_patch, ba));
    LVT_LAUNCH_CHECK(stream, "brief_kernel");
    return LVTK_OK;
}

} // namespace lvtb
