// Block-level (one CTA) greedy Hamming matchers, bit-exact with the reference's sequential,
// order-dependent loops.  Header-only device code shared by the seam kernels (match.cu) and the
// fused per-frame tracking kernel (track.cu).
//
// The reference matches map points one after another; a feature taken by an earlier point is
// masked for all later ones (lvt/src/lvt_local_map.cpp:149-170,
// lvt/src/lvt_image_features_struct.cpp:85-101), and stereo row matching does the same over left
// features (lvt/src/lvt_image_features_handler.cpp:305-322).  The exact sequential result is the
// unique fixed point of
//     choice[i] = ratio_test( best2( candidates(i) \ { f : exists j < i, choice[j] == f } ) )
// (induction on i), so it is computed by parallel rounds: every query re-selects against the
// owners (= lowest query index choosing each feature) of the previous round until a round
// changes nothing.  Round r fixes at least queries 0..r; in practice 2-4 rounds suffice.
// Inside a round one warp serves one query: lanes stride over the candidates, each computes a
// 256-bit Hamming distance with 8 __popc, and two redux.sync.min give the best two
// (distance << 20 | index) keys == knnMatch's (distance, index) order.
#pragma once
#include "extract.cuh"

namespace lvtb
{

constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kFree = 0x7FFFFFFF; // owner of a feature nobody has chosen
constexpr int kTaken = -1;        // owner of a feature that was marked before the pass started

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint32_t *b)
{
    const uint4 b0 = *reinterpret_cast<const uint4 *>(b), b1 = *reinterpret_cast<const uint4 *>(b + 4);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ void top2_insert(uint32_t key, uint32_t &k1, uint32_t &k2)
{
    if (key < k1)
    {
        k2 = k1;
        k1 = key;
    }
    else if (key < k2)
        k2 = key;
}

// per-lane (k1 <= k2) -> warp-wide best two keys (keys are unique: they embed the index)
__device__ __forceinline__ void warp_top2(uint32_t k1, uint32_t k2, uint32_t &b1, uint32_t &b2)
{
    b1 = __reduce_min_sync(0xffffffffu, k1);
    b2 = __reduce_min_sync(0xffffffffu, k1 == b1 ? k2 : k1);
}

// the acceptance rule of struct.cpp:105-116 and :141-144; -1 = no match
__device__ __forceinline__ int accept_match(uint32_t b1, uint32_t b2, float ratio_th, float dist_th)
{
    if (b1 == kNoKey)
        return -1;
    const float d0 = (float)(b1 >> 20);
    if (b2 != kNoKey)
        return (__fdiv_rn(d0, (float)(b2 >> 20)) < ratio_th) ? (int)(b1 & 0xFFFFFu) : -1;
    return (d0 <= dist_th) ? (int)(b1 & 0xFFFFFu) : -1;
}

struct MatchScratch
{
    float2 *proj; // [cap] projected pixel (double -> float, struct.cpp:70)
    uint8_t *vis; // [cap] is_point_visible
    int *choice;  // [cap]
};

// is_point_visible for every point (lvt_local_map.cpp:62-82, :149-157)
__device__ inline void block_project(const double *xyz, int m, const double *W /* smem, 12 */, const CamParams &cam,
                                     const MatchScratch &ms)
{
    for (int i = threadIdx.x; i < m; i += blockDim.x)
    {
        double u, v;
        const bool ok = point_visible(W, cam, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &u, &v);
        ms.vis[i] = ok;
        ms.proj[i] = ok ? make_float2((float)u, (float)v) : make_float2(0.f, 0.f);
    }
    __syncthreads();
}

// One pass of find_match_index over all visible points, in order, greedy (see file header).
// use_marks: start from f.matched (true) or from a cleared mark vector (false).
// On return: ms.choice[i] = feature or -1 for visible points; owner_a[f] != kFree <=> f is marked.
// owner_a / owner_b: int[>= n] each (shared memory).  Returns the number of matches.
__device__ inline int block_match_projected(const uint32_t *pdesc, const MatchScratch &ms, int m, const FeatDev &f,
                                            int n, const CamParams &cam, float r2, bool use_marks, int *owner_a,
                                            int *owner_b, int *s_flag, float *out_d1, float *out_d2)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int *cur = owner_a, *nxt = owner_b;
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        cur[j] = (use_marks && f.matched[j]) ? kTaken : kFree;
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        ms.choice[i] = -1;
    __syncthreads();

    const float cell = (float)kHashCell;
    int count = 0;
    for (int round = 0;; round++)
    {
        for (int j = threadIdx.x; j < n; j += blockDim.x)
            nxt[j] = cur[j] == kTaken ? kTaken : kFree;
        if (threadIdx.x == 0)
            s_flag[0] = 0, s_flag[1] = 0;
        __syncthreads();

        int my_count = 0;
        for (int i = warp; i < m; i += nwarps)
        {
            if (!ms.vis[i])
                continue;
            const float2 p = ms.proj[i];
            const int hy = (int)floorf(__fdiv_rn(p.y, cell)), hx = (int)floorf(__fdiv_rn(p.x, cell));
            const int sy = max(hy - cam.cell_search_radius, 0), ey = min(hy + cam.cell_search_radius + 1, cam.cells_y);
            const int sx = max(hx - cam.cell_search_radius, 0), ex = min(hx + cam.cell_search_radius + 1, cam.cells_x);
            const uint4 q0 = *reinterpret_cast<const uint4 *>(pdesc + 8 * (size_t)i);
            const uint4 q1 = *reinterpret_cast<const uint4 *>(pdesc + 8 * (size_t)i + 4);
            uint32_t k1 = kNoKey, k2 = kNoKey;
            if (sx < ex)
            {
                for (int cy = sy; cy < ey; cy++)
                {
                    // cells of one grid row are contiguous in the CSR
                    const int s = f.cell_start[cy * cam.cells_x + sx], e = f.cell_start[cy * cam.cells_x + ex];
                    for (int pos = s + lane; pos < e; pos += 32)
                    {
                        const int j = f.cell_items[pos];
                        if (cur[j] < i)
                            continue; // marked, or taken by an earlier point
                        const float2 k = f.xy[j];
                        const float dx = __fsub_rn(k.x, p.x), dy = __fsub_rn(k.y, p.y);
                        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2)
                            top2_insert(((uint32_t)hamming256(q0, q1, f.desc + 8 * (size_t)j) << 20) | (uint32_t)j, k1, k2);
                    }
                }
            }
            uint32_t b1, b2;
            warp_top2(k1, k2, b1, b2);
            const int c = accept_match(b1, b2, cam.tracking_ratio_th, cam.desc_dist_th);
            if (lane == 0)
            {
                if (c != ms.choice[i])
                {
                    ms.choice[i] = c;
                    s_flag[0] = 1;
                }
                if (c >= 0)
                {
                    atomicMin(&nxt[c], i);
                    my_count++;
                    if (out_d1)
                    {
                        out_d1[i] = (float)(b1 >> 20);
                        out_d2[i] = b2 != kNoKey ? (float)(b2 >> 20) : -1.0f;
                    }
                }
            }
        }
        if (lane == 0 && my_count)
            atomicAdd(&s_flag[1], my_count);
        __syncthreads();
        const int changed = s_flag[0];
        count = s_flag[1];
        int *t = cur;
        cur = nxt;
        nxt = t;
        __syncthreads();
        if (!changed)
            break;
    }
    if (cur != owner_a)
    {
        for (int j = threadIdx.x; j < n; j += blockDim.x)
            owner_a[j] = cur[j];
        __syncthreads();
    }
    return count;
}

// Stereo row matching pass (handler.cpp:302-323 + struct.cpp:122-148).  Queries = left features
// in index order that are not marked; candidates = unmarked right features whose y lies in
// [max(0,(int)y-2), min(rows,(int)y+2)] -- no x constraint.  choice: int[>= nl] (global).
// On return the marks of both sides are updated and the pairs are written in left-index order.
__device__ inline int block_row_match(const FeatDev &fl, int nl, const FeatDev &fr, int nr, const CamParams &cam,
                                      int *choice, int *owner_a, int *owner_b, int *s_flag, int *s_scan,
                                      int *out_query, int *out_train)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int *cur = owner_a, *nxt = owner_b;
    for (int j = threadIdx.x; j < nr; j += blockDim.x)
        cur[j] = fr.matched[j] ? kTaken : kFree;
    for (int i = threadIdx.x; i < nl; i += blockDim.x)
        choice[i] = -1;
    __syncthreads();

    for (int round = 0;; round++)
    {
        for (int j = threadIdx.x; j < nr; j += blockDim.x)
            nxt[j] = cur[j] == kTaken ? kTaken : kFree;
        if (threadIdx.x == 0)
            s_flag[0] = 0;
        __syncthreads();
        for (int i = warp; i < nl; i += nwarps)
        {
            if (fl.matched[i])
                continue; // tracked from the map this frame (handler.cpp:307-310)
            const float2 p = fl.xy[i];
            const int start_y = max((int)p.y - kRowSearchRadius, 0);
            const int end_y = min((int)p.y + kRowSearchRadius, cam.img_h);
            const uint4 q0 = *reinterpret_cast<const uint4 *>(fl.desc + 8 * (size_t)i);
            const uint4 q1 = *reinterpret_cast<const uint4 *>(fl.desc + 8 * (size_t)i + 4);
            uint32_t k1 = kNoKey, k2 = kNoKey;
            if (start_y <= end_y)
            {
                // bins floor(y) in [start_y, end_y] are contiguous in the row CSR
                const int s = fr.row_start[start_y], e = fr.row_start[end_y + 1];
                for (int pos = s + lane; pos < e; pos += 32)
                {
                    const int j = fr.row_items[pos];
                    if (cur[j] < i)
                        continue;
                    const float yj = fr.xy[j].y;
                    if (yj >= (float)start_y && yj <= (float)end_y)
                        top2_insert(((uint32_t)hamming256(q0, q1, fr.desc + 8 * (size_t)j) << 20) | (uint32_t)j, k1, k2);
                }
            }
            uint32_t b1, b2;
            warp_top2(k1, k2, b1, b2);
            const int c = accept_match(b1, b2, cam.triangulation_ratio_th, cam.desc_dist_th);
            if (lane == 0)
            {
                if (c != choice[i])
                {
                    choice[i] = c;
                    s_flag[0] = 1;
                }
                if (c >= 0)
                    atomicMin(&nxt[c], i);
            }
        }
        __syncthreads();
        const int changed = s_flag[0];
        int *t = cur;
        cur = nxt;
        nxt = t;
        __syncthreads();
        if (!changed)
            break;
    }
    for (int j = threadIdx.x; j < nr; j += blockDim.x)
        if (cur[j] != kFree)
            fr.matched[j] = 1;
    // pairs in left-index order + left marks (handler.cpp:313-321)
    int running = 0;
    for (int i0 = 0; i0 < nl; i0 += blockDim.x)
    {
        const int i = i0 + threadIdx.x;
        const int c = i < nl ? choice[i] : -1;
        int total;
        const int pos = block_exclusive_scan(c >= 0, s_scan, &total);
        if (c >= 0)
        {
            out_query[running + pos] = i;
            out_train[running + pos] = c;
            fl.matched[i] = 1;
        }
        running += total;
    }
    __syncthreads();
    return running;
}

} // namespace lvtb
