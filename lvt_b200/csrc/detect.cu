// Corner detection on sm_100a: AGAST(OAST_9_16) score -> AGAST NMS -> per-tile ANMS -> gather.
//
// Replaces, with identical results, the calls at lvt/src/lvt_image_features_handler.cpp:139
// (cv::AgastFeatureDetector::detect per tile), :142 (_adaptive_non_maximal_suppresion, :34-83),
// :146-152 (tile offset + concatenation), :161-169 (low-corner retry) and the border filter
// that BriefDescriptorExtractor::compute applies at :172.
//
//   K1 score_kernel  : TMA stages a 70x38 halo tile (96x38 box, 16-B aligned start) into shared memory; every pixel's
//                      exact AGAST score (max threshold for which it is a 9-of-16 segment-test
//                      corner) is computed branch-free with packed 2 x s16 min/max; the dense
//                      u8 score map goes to HBM (stays in L2).
//   K2 nms_kernel    : one survivor per 4-connected component of corner pixels, OpenCV's
//                      union-find tie-breaking replayed per component; survivors are appended to
//                      their detection tile's list.
//   K2b nms_fallback : the sequential algorithm for a tile with a component too large for K2.
//   K3 tile_kernel   : per tile: raster sort, then (n > k) ANMS = all-pairs suppression radius,
//                      radix-select of the (k+1)-th largest, emission in std::sort's order.
//   K4 gather_kernel : concatenates the tiles, applies BRIEF's 28-px border filter, decides the
//                      <200-corner retry.
#include "extract.cuh"
#include "introsort.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace lvtb
{

// ---------------------------------------------------------------------------------------------
// K1: score
// ---------------------------------------------------------------------------------------------
struct ScoreArgs
{
    const int *slots; // pool slot of image b
    uint8_t *score;   // [batch][rows][pitch]
    TileGrid grid;
    int pitch, rows, cols;
    int min_score; // scores below this are stored as 0
};

// ring offsets (dx, dy) of OAST_9_16 in OpenCV's order
__device__ __constant__ int8_t c_ring[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                               {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// packed 2 x s16 min / max: one VIMNMX each on sm_90+
__device__ __forceinline__ uint32_t min_s16x2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("min.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t max_s16x2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// exact AGAST score of the pixel at smem row pointer p (stride S): both polarities ride in one
// register as 2 x s16 biased by 256: lo = v - c + 256, hi = c - v + 256.
template <int S>
__device__ __forceinline__ int agast_score(const uint8_t *p)
{
    const uint32_t c = p[0];
    const uint32_t bias = ((256u + c) << 16) | (256u - c);
    uint32_t t[16];
#define LVT_RING(i, dx, dy) t[i] = (uint32_t)p[(dy)*S + (dx)] * 0xFFFF0001u + bias;
    LVT_RING(0, 0, 3)
    LVT_RING(1, 1, 3)
    LVT_RING(2, 2, 2)
    LVT_RING(3, 3, 1)
    LVT_RING(4, 3, 0)
    LVT_RING(5, 3, -1)
    LVT_RING(6, 2, -2)
    LVT_RING(7, 1, -3)
    LVT_RING(8, 0, -3)
    LVT_RING(9, -1, -3)
    LVT_RING(10, -2, -2)
    LVT_RING(11, -3, -1)
    LVT_RING(12, -3, 0)
    LVT_RING(13, -3, 1)
    LVT_RING(14, -2, 2)
    LVT_RING(15, -1, 3)
#undef LVT_RING
    uint32_t m2[16], m4[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        m2[i] = min_s16x2(t[i], t[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++)
        m4[i] = min_s16x2(m2[i], m2[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++)
        m2[i] = min_s16x2(m4[i], m4[(i + 4) & 15]); // min over 8
    uint32_t best = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        best = max_s16x2(best, min_s16x2(m2[i], t[(i + 8) & 15])); // min over 9, max over arcs
    const int lo = (int)(best & 0xFFFFu), hi = (int)(best >> 16);
    return max(lo, hi) - 257;
}

__global__ void __launch_bounds__(256) score_kernel(const __grid_constant__ CUtensorMap tmap, ScoreArgs a)
{
    __shared__ __align__(128) uint8_t tile[kScoreBoxH][kScoreBoxW];
    __shared__ __align__(8) uint64_t bar;

    const int x0 = blockIdx.x * kScoreTileW, y0 = blockIdx.y * kScoreTileH, b = blockIdx.z;
    if (threadIdx.x == 0)
    {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&bar, kScoreBoxH * kScoreBoxW);
        tma_load_3d(&tile[0][0], &tmap, x0 - kScoreBoxX, y0 - 3, a.slots[b], &bar); // out-of-bounds -> 0
    }
    mbar_wait(&bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *out = a.score + (size_t)b * a.rows * a.pitch;
#pragma unroll
    for (int rr = 0; rr < kScoreTileH / 8; rr++)
    {
        const int r = warp + 8 * rr, gy = y0 + r;
        if (gy >= a.rows)
            break;
        // tile-local row coordinate within its detection tile
        const int ty = gy / a.grid.cell, ly = gy - ty * a.grid.cell, th = a.grid.tile_h(ty);
        const bool row_ok = (ly >= 3) && (ly <= th - 4);
#pragma unroll
        for (int cc = 0; cc < 2; cc++)
        {
            const int lx0 = lane + 32 * cc, gx = x0 + lx0;
            if (gx >= a.cols)
                continue;
            const int tx = gx / a.grid.cell, lx = gx - tx * a.grid.cell, tw = a.grid.tile_w(tx);
            int s = agast_score<kScoreBoxW>(&tile[r + 3][lx0 + kScoreBoxX]);
            if (!(row_ok && lx >= 3 && lx <= tw - 4) || s < a.min_score)
                s = 0;
            out[(size_t)gy * a.pitch + gx] = (uint8_t)s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: non-maximum suppression
// ---------------------------------------------------------------------------------------------
struct NmsArgs
{
    const uint8_t *score;
    uint32_t *tile_list;
    int *tile_count, *tile_overflow, *error;
    uint32_t *cand;   // [batch][rows * pitch] local maxima that need their component examined
    int *cand_count;  // [batch]
    const int *retry; // nullptr on the first pass
    TileGrid grid;
    int pitch, rows, cols, tile_cap, n_tiles;
    int threshold;
    int nonmax;
};

constexpr int kCompCap = 96; // largest component replayed in registers/local memory

__device__ __forceinline__ void emit_survivor(const NmsArgs &a, int b, int x, int y, int s)
{
    const int tx = x / a.grid.cell, ty = y / a.grid.cell, t = ty * a.grid.nx + tx;
    const int raster = (y - ty * a.grid.cell) * a.grid.tile_w(tx) + (x - tx * a.grid.cell);
    const int pos = atomicAdd(&a.tile_count[b * a.n_tiles + t], 1);
    if (pos < a.tile_cap)
        a.tile_list[((size_t)b * a.n_tiles + t) * a.tile_cap + pos] = ((uint32_t)raster << 8) | (uint32_t)s;
    else
        *a.error = LVTK_ERR_CAPACITY;
}

// OpenCV's AGAST NMS restricted to one component: pts sorted in raster order, sc their scores.
// Returns the index of the surviving root.
__device__ int replay_component(const uint32_t *pts, const uint8_t *sc, int n)
{
    int8_t parent[kCompCap];
    for (int i = 0; i < n; i++)
        parent[i] = -1;
    for (int c = 0; c < n; c++)
    {
        const uint32_t p = pts[c];
        // corner directly above: (x, y-1) = p - (1 << 16)
        int above = -1;
        if ((p >> 16) > 0)
        {
            const uint32_t q = p - (1u << 16);
            for (int j = c - 1; j >= 0; j--)
            {
                if (pts[j] == q)
                {
                    above = j;
                    break;
                }
                if (pts[j] < q)
                    break;
            }
        }
        if (above >= 0)
        {
            int w = above;
            while (parent[w] != -1)
                w = parent[w];
            if (sc[c] < sc[w])
                parent[c] = (int8_t)w;
            else
                parent[w] = (int8_t)c;
        }
        // corner directly left: (x-1, y) is the previous point in raster order if present
        if (c != 0 && pts[c - 1] + 1 == p && (p & 0xFFFFu) != 0)
        {
            const int pa = parent[c];
            int t = c - 1;
            while (parent[t] != -1)
                t = parent[t];
            if (pa == -1)
            {
                if (t != c)
                {
                    if (sc[c] < sc[t])
                        parent[c] = (int8_t)t;
                    else
                        parent[t] = (int8_t)c;
                }
            }
            else if (t != pa)
            {
                if (sc[pa] < sc[t])
                {
                    parent[pa] = (int8_t)t;
                    parent[c] = (int8_t)t;
                }
                else
                {
                    parent[t] = (int8_t)pa;
                    parent[c] = (int8_t)pa;
                }
            }
        }
    }
    for (int i = 0; i < n; i++)
        if (parent[i] == -1)
            return i;
    return 0;
}

// first look at a corner pixel: beaten by a 4-neighbour -> dropped; isolated -> survivor; otherwise
// it is a local maximum whose component has to be examined (nms_resolve_kernel)
__device__ void nms_pixel(const NmsArgs &a, int b, const uint8_t *sm, int x, int y, int s)
{
    if (!a.nonmax)
    {
        emit_survivor(a, b, x, y, s);
        return;
    }
    auto score_at = [&](int xx, int yy) -> int {
        if (xx < 0 || yy < 0 || xx >= a.cols || yy >= a.rows)
            return 0;
        const int v = sm[(size_t)yy * a.pitch + xx];
        return v >= a.threshold ? v : 0;
    };
    const int up = score_at(x, y - 1), dn = score_at(x, y + 1), lf = score_at(x - 1, y), rt = score_at(x + 1, y);
    if (max(max(up, dn), max(lf, rt)) > s)
        return; // a neighbour in the same component beats it
    if ((up | dn | lf | rt) == 0)
    {
        emit_survivor(a, b, x, y, s); // isolated corner
        return;
    }
    const int pos = atomicAdd(&a.cand_count[b], 1);
    a.cand[(size_t)b * a.rows * a.pitch + pos] = ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x;
}

// a local maximum: flood its component, decide whether it is the component's survivor.
// score_at(x, y) = corner score at the current threshold, 0 outside the image / below it.
template <class ScoreAt>
__device__ void nms_resolve(const NmsArgs &a, int b, ScoreAt score_at, int x, int y, int s)
{
    // flood the component; give up as soon as anything larger shows up.  Visited test: linear
    // search while the component is small, then a 64x32-pixel bitmap around the start pixel
    // (pixels outside that window keep the linear search).
    uint32_t pts[kCompCap];
    uint8_t sc[kCompCap];
    uint32_t bm[64];
    bool use_bm = false;
    const int wx0 = x - 32, wy0 = y - 16;
    auto in_win = [&](int qx, int qy) { return (unsigned)(qx - wx0) < 64u && (unsigned)(qy - wy0) < 32u; };
    auto bm_set = [&](int qx, int qy) { bm[(qy - wy0) * 2 + ((qx - wx0) >> 5)] |= 1u << ((qx - wx0) & 31); };
    int n = 1, head = 0;
    bool tie = false;
    pts[0] = ((uint32_t)y << 16) | (uint32_t)x;
    sc[0] = (uint8_t)s;
    while (head < n)
    {
        const uint32_t p = pts[head++];
        const int px = (int)(p & 0xFFFFu), py = (int)(p >> 16);
        // the four neighbour loads are independent: issue them together
        int vv[4];
#pragma unroll
        for (int d = 0; d < 4; d++)
            vv[d] = score_at(px + (d == 0) - (d == 1), py + (d == 2) - (d == 3));
        if (max(max(vv[0], vv[1]), max(vv[2], vv[3])) > s)
            return;
#pragma unroll
        for (int d = 0; d < 4; d++)
        {
            const int qx = px + (d == 0) - (d == 1), qy = py + (d == 2) - (d == 3);
            const int v = vv[d];
            if (v == 0)
                continue;
            const uint32_t q = ((uint32_t)qy << 16) | (uint32_t)qx;
            bool seen = false;
            if (use_bm && in_win(qx, qy))
                seen = (bm[(qy - wy0) * 2 + ((qx - wx0) >> 5)] >> ((qx - wx0) & 31)) & 1u;
            else
                for (int j = 0; j < n; j++)
                    seen |= (pts[j] == q);
            if (seen)
                continue;
            if (n == kCompCap)
            {
                const int t = (y / a.grid.cell) * a.grid.nx + (x / a.grid.cell);
                a.tile_overflow[b * a.n_tiles + t] = 1; // K2b redoes this tile sequentially
                return;
            }
            tie |= (v == s);
            pts[n] = q;
            sc[n] = (uint8_t)v;
            n++;
            if (use_bm && in_win(qx, qy))
                bm_set(qx, qy);
            if (!use_bm && n == 16)
            {
                for (int k = 0; k < 64; k++)
                    bm[k] = 0;
                for (int j = 0; j < n; j++)
                {
                    const int jx = (int)(pts[j] & 0xFFFFu), jy = (int)(pts[j] >> 16);
                    if (in_win(jx, jy))
                        bm_set(jx, jy);
                }
                use_bm = true;
            }
        }
    }
    if (!tie)
    {
        emit_survivor(a, b, x, y, s); // unique maximum of its component
        return;
    }
    // several pixels share the maximum: OpenCV's merge order decides.  Insertion sort to
    // raster order, replay, and emit only if this pixel is the root.
    const uint32_t self = pts[0];
    for (int i = 1; i < n; i++)
    {
        const uint32_t p = pts[i];
        const uint8_t v = sc[i];
        int j = i - 1;
        while (j >= 0 && pts[j] > p)
        {
            pts[j + 1] = pts[j];
            sc[j + 1] = sc[j];
            j--;
        }
        pts[j + 1] = p;
        sc[j + 1] = v;
    }
    const int root = replay_component(pts, sc, n);
    if (pts[root] == self)
        emit_survivor(a, b, x, y, s);
}

__global__ void __launch_bounds__(128) nms_kernel(NmsArgs a)
{
    const int b = blockIdx.z;
    if (a.retry && !a.retry[b])
        return;
    const int y = blockIdx.y, xw = blockIdx.x * blockDim.x + threadIdx.x;
    if (xw * 4 >= a.cols)
        return;
    const uint8_t *sm = a.score + (size_t)b * a.rows * a.pitch;
    const uint32_t word = *reinterpret_cast<const uint32_t *>(sm + (size_t)y * a.pitch + xw * 4);
    if (word == 0)
        return;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const int s = (word >> (8 * k)) & 0xFF;
        if (s >= a.threshold && s != 0 && xw * 4 + k < a.cols)
            nms_pixel(a, b, sm, xw * 4 + k, y, s);
    }
}

// one thread per local maximum, densely packed (the marking kernel above is a coalesced sweep)
__global__ void __launch_bounds__(128) nms_resolve_kernel(NmsArgs a)
{
    const int b = blockIdx.y;
    if (a.retry && !a.retry[b])
        return;
    const uint8_t *sm = a.score + (size_t)b * a.rows * a.pitch;
    const int n = a.cand_count[b];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t c = a.cand[(size_t)b * a.rows * a.pitch + i];
        auto score_at = [&](int xx, int yy) -> int {
            if (xx < 0 || yy < 0 || xx >= a.cols || yy >= a.rows)
                return 0;
            const int v = sm[(size_t)yy * a.pitch + xx];
            return v >= a.threshold ? v : 0;
        };
        nms_resolve(a, b, score_at, (int)(c & 0xFFFu), (int)((c >> 12) & 0xFFFu), (int)(c >> 24));
    }
}

// Fused NMS: a CTA stages a 64x64 window of the score map (its 32x32 pixels + 16 px halo) in shared
// memory, sweeps its pixels (beaten / isolated / local maximum), then resolves its local maxima with
// the component flood reading shared memory (global memory only outside the window).
constexpr int kNmsTile = 32, kNmsHalo = 16, kNmsWin = kNmsTile + 2 * kNmsHalo;

__global__ void __launch_bounds__(256) nms_tile_kernel(NmsArgs a)
{
    __shared__ __align__(16) uint8_t win[kNmsWin * kNmsWin];
    __shared__ uint32_t s_cand[kNmsTile * kNmsTile];
    __shared__ int s_ncand;
    const int b = blockIdx.z;
    if (a.retry && !a.retry[b])
        return;
    const uint8_t *sm = a.score + (size_t)b * a.rows * a.pitch;
    const int x0 = blockIdx.x * kNmsTile - kNmsHalo, y0 = blockIdx.y * kNmsTile - kNmsHalo; // x0 is a multiple of 16
    if (threadIdx.x == 0)
        s_ncand = 0;
    for (int w = threadIdx.x; w < kNmsWin * kNmsWin / 4; w += blockDim.x)
    {
        const int row = w / (kNmsWin / 4), gx = x0 + 4 * (w % (kNmsWin / 4)), gy = y0 + row;
        uint32_t v = 0;
        if (gy >= 0 && gy < a.rows && gx >= 0 && gx + 3 < a.pitch)
            v = *reinterpret_cast<const uint32_t *>(sm + (size_t)gy * a.pitch + gx); // pitch padding is zero
        reinterpret_cast<uint32_t *>(win)[w] = v;
    }
    __syncthreads();
    auto score_at = [&](int xx, int yy) -> int {
        if (xx < 0 || yy < 0 || xx >= a.cols || yy >= a.rows)
            return 0;
        const int lx = xx - x0, ly = yy - y0;
        const int v = ((unsigned)lx < (unsigned)kNmsWin && (unsigned)ly < (unsigned)kNmsWin) ? win[ly * kNmsWin + lx]
                                                                                              : sm[(size_t)yy * a.pitch + xx];
        return v >= a.threshold ? v : 0;
    };
    // sweep: thread = one 4-pixel word of the 32x32 interior
    {
        const int row = threadIdx.x / (kNmsTile / 4), c4 = threadIdx.x % (kNmsTile / 4);
        const int y = y0 + kNmsHalo + row;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int x = x0 + kNmsHalo + 4 * c4 + k;
            if (x >= a.cols || y >= a.rows)
                continue;
            const int s = score_at(x, y);
            if (s == 0)
                continue;
            if (!a.nonmax)
            {
                emit_survivor(a, b, x, y, s);
                continue;
            }
            const int up = score_at(x, y - 1), dn = score_at(x, y + 1), lf = score_at(x - 1, y), rt = score_at(x + 1, y);
            if (max(max(up, dn), max(lf, rt)) > s)
                continue; // a neighbour in the same component beats it
            if ((up | dn | lf | rt) == 0)
                emit_survivor(a, b, x, y, s); // isolated corner
            else
                s_cand[atomicAdd(&s_ncand, 1)] = ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x;
        }
    }
    __syncthreads();
    const int n = s_ncand;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        const uint32_t c = s_cand[i];
        nms_resolve(a, b, score_at, (int)(c & 0xFFFu), (int)((c >> 12) & 0xFFFu), (int)(c >> 24));
    }
}

// K2b: sequential fallback, one thread per flagged tile (pathological inputs only)
__global__ void nms_fallback_kernel(NmsArgs a, int *parent_all)
{
    const int t = blockIdx.x, b = blockIdx.y;
    if (threadIdx.x != 0 || !a.tile_overflow[b * a.n_tiles + t])
        return;
    if (a.retry && !a.retry[b])
        return;
    const uint8_t *sm = a.score + (size_t)b * a.rows * a.pitch;
    int *parent = parent_all + (size_t)b * a.rows * a.pitch;
    const int tx = t % a.grid.nx, ty = t / a.grid.nx;
    const int x0 = tx * a.grid.cell, y0 = ty * a.grid.cell, tw = a.grid.tile_w(tx), th = a.grid.tile_h(ty);
    auto is_corner = [&](int x, int y) { const int v = sm[(size_t)y * a.pitch + x]; return v != 0 && v >= a.threshold; };
    auto idx = [&](int x, int y) { return y * a.pitch + x; };
    auto root = [&](int i) { while (parent[i] != -1) i = parent[i]; return i; };
    for (int y = y0 + 3; y <= y0 + th - 4; y++)
        for (int x = x0 + 3; x <= x0 + tw - 4; x++)
            if (is_corner(x, y))
                parent[idx(x, y)] = -1;
    for (int y = y0 + 3; y <= y0 + th - 4; y++)
    {
        for (int x = x0 + 3; x <= x0 + tw - 4; x++)
        {
            if (!is_corner(x, y))
                continue;
            const int c = idx(x, y);
            const int sc = sm[c];
            if (y - 1 >= y0 + 3 && is_corner(x, y - 1))
            {
                const int w = root(idx(x, y - 1));
                if (sc < sm[w])
                    parent[c] = w;
                else
                    parent[w] = c;
            }
            if (x - 1 >= x0 + 3 && is_corner(x - 1, y))
            {
                const int pa = parent[c];
                const int l = root(idx(x - 1, y));
                if (pa == -1)
                {
                    if (l != c)
                    {
                        if (sc < sm[l])
                            parent[c] = l;
                        else
                            parent[l] = c;
                    }
                }
                else if (l != pa)
                {
                    if (sm[pa] < sm[l])
                    {
                        parent[pa] = l;
                        parent[c] = l;
                    }
                    else
                    {
                        parent[l] = pa;
                        parent[c] = pa;
                    }
                }
            }
        }
    }
    // rewrite the tile's list from scratch
    int n = 0;
    uint32_t *list = a.tile_list + ((size_t)b * a.n_tiles + t) * a.tile_cap;
    for (int y = y0 + 3; y <= y0 + th - 4; y++)
        for (int x = x0 + 3; x <= x0 + tw - 4; x++)
            if (is_corner(x, y) && parent[idx(x, y)] == -1)
            {
                if (n < a.tile_cap)
                    list[n] = ((uint32_t)((y - y0) * tw + (x - x0)) << 8) | sm[idx(x, y)];
                n++;
            }
    if (n > a.tile_cap)
        *a.error = LVTK_ERR_CAPACITY;
    a.tile_count[b * a.n_tiles + t] = min(n, a.tile_cap);
}

// ---------------------------------------------------------------------------------------------
// K3: per-tile ordering + ANMS
// ---------------------------------------------------------------------------------------------
struct TileArgs
{
    uint32_t *tile_list, *tile_aux, *tile_out;
    const int *tile_count;
    int *tile_out_count;
    const int *retry;
    TileGrid grid;
    int tile_cap, n_tiles, max_per_cell;
    long long *dbg; // optional clock64() marks [tile][8]
};

constexpr int kTileSmemCap = 8192; // tiles with more survivors work out of global scratch
constexpr int kTileThreads = 1024;
constexpr int kTileRanges = kTileSmemCap / 8 + 2; // ranges alive in one level of the introsort schedule
constexpr int kTileSmemBytes = 3 * kTileSmemCap * (int)sizeof(uint32_t) + 2 * kTileRanges * (int)sizeof(isort::LevelRange) +
                               2 * kTileSmemCap * (int)sizeof(uint16_t);

// std::__unguarded_partition(first + 1, last, first) computed by a whole warp with the SAME result as
// the sequential two-pointer loop.  With pv = key(pivot), the forward pointer stops at elements with
// key <= pv ("L-stoppers"), the backward pointer at elements with key >= pv ("R-stoppers"); the loop
// swaps the i-th L-stopper from the left with the i-th R-stopper from the right while the former lies
// left of the latter, and nothing it swaps is ever looked at again.  So with posL / posR = the two
// position lists of the ORIGINAL range, K = #{i : posL[i] < posR[i]} swaps happen (the predicate is
// monotone) and the returned cut is posL[K] if that lies left of posR[K-1], else posR[K-1].
// Ranks come from ballots + running counts; posL / posR are shared scratch (u16, range length).
__device__ int warp_partition(uint32_t *a, int first, int last, uint16_t *posL, uint16_t *posR)
{
    const int lane = threadIdx.x & 31;
    const uint32_t pv = a[first] >> 24;
    const int lo = first + 1, m = last - lo;
    int nL = 0;
    for (int base = 0; base < m; base += 32)
    {
        const int p = base + lane;
        const bool isL = p < m && (a[lo + p] >> 24) <= pv;
        const uint32_t bl = __ballot_sync(0xffffffffu, isL);
        if (isL)
            posL[nL + __popc(bl & ((1u << lane) - 1u))] = (uint16_t)p;
        nL += __popc(bl);
    }
    int nR = 0;
    for (int base = 0; base < m; base += 32)
    {
        const int p = m - 1 - (base + lane); // from the right
        const bool isR = p >= 0 && (a[lo + p] >> 24) >= pv;
        const uint32_t br = __ballot_sync(0xffffffffu, isR);
        if (isR)
            posR[nR + __popc(br & ((1u << lane) - 1u))] = (uint16_t)p;
        nR += __popc(br);
    }
    __syncwarp();
    const int nmin = min(nL, nR);
    int K = 0;
    for (int base = 0; base < nmin; base += 32)
    {
        const int i = base + lane;
        const bool sw = i < nmin && posL[i] < posR[i];
        const uint32_t bs = __ballot_sync(0xffffffffu, sw);
        if (sw)
        {
            const int pl = lo + posL[i], pr = lo + posR[i];
            const uint32_t t = a[pl];
            a[pl] = a[pr];
            a[pr] = t;
        }
        K += __popc(bs);
        if (bs != 0xffffffffu)
            break; // monotone: no further swaps
    }
    int cut;
    const int prevR = K > 0 ? (int)posR[K - 1] : m; // m = one past the range
    if (K < nL && (int)posL[K] < prevR)
        cut = lo + posL[K];
    else
        cut = lo + prevR;
    __syncwarp();
    return cut;
}

constexpr int kCoopRange = 96; // ranges at least this long are partitioned by the whole warp

// the same schedule run by ONE warp so that the other warps of the CTA can compute the suppression
// radii at the same time: long ranges are partitioned cooperatively (warp_partition), short ones
// lane-per-range, __syncwarp between levels
__device__ void warp_introsort(uint32_t *a, int n, isort::LevelRange *q0, isort::LevelRange *q1, volatile int *s_cnt,
                               uint16_t *posL, uint16_t *posR)
{
    const int lane = threadIdx.x & 31;
    if (n <= 1)
        return;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1)
        lg++;
    if (lane == 0)
    {
        q0[0] = isort::LevelRange{0, n, 2 * lg};
        s_cnt[0] = 1;
        s_cnt[1] = 0;
    }
    __syncwarp();
    isort::LevelRange *cur = q0, *nxt = q1;
    int which = 0;
    while (true)
    {
        const int ncur = s_cnt[which];
        if (ncur == 0)
            break;
        // long ranges: one after another, all lanes together
        for (int i = 0; i < ncur; i++)
        {
            const isort::LevelRange r = cur[i];
            if (r.last - r.first < kCoopRange || r.depth == 0)
                continue;
            if (lane == 0)
            {
                uint32_t *f = a + r.first, *l = a + r.last;
                isort::move_median_to_first(f, f + 1, f + (l - f) / 2, l - 1);
            }
            __syncwarp();
            const int cut = warp_partition(a, r.first, r.last, posL, posR);
            if (lane == 0)
            {
                const int slot = atomicAdd((int *)&s_cnt[which ^ 1], 2);
                nxt[slot] = isort::LevelRange{r.first, cut, r.depth - 1};
                nxt[slot + 1] = isort::LevelRange{cut, r.last, r.depth - 1};
            }
        }
        __syncwarp();
        // short ranges (and heap-sort fallbacks): lane per range
        for (int i = lane; i < ncur; i += 32)
        {
            const isort::LevelRange r = cur[i];
            if (r.last - r.first >= kCoopRange && r.depth != 0)
                continue;
            if (r.last - r.first <= 16)
            {
                isort::insertion_sort(a + r.first, a + r.last);
                continue;
            }
            isort::LevelRange l, rr;
            if (isort::split_range(a, r, l, rr))
            {
                const int slot = atomicAdd((int *)&s_cnt[which ^ 1], 2);
                nxt[slot] = l;
                nxt[slot + 1] = rr;
            }
        }
        __syncwarp();
        if (lane == 0)
            s_cnt[which] = 0;
        which ^= 1;
        isort::LevelRange *t = cur;
        cur = nxt;
        nxt = t;
        __syncwarp();
    }
}

// stable sort of a leaf range (<= 16 elements) by rank: what the final insertion sort does to it
// (a stable sort's result is unique).  Lanes 0..len-1 hold one element each.
__device__ __forceinline__ void warp_leaf_sort(uint32_t *a, int first, int len)
{
    const int lane = threadIdx.x & 31;
    const uint32_t mine = lane < len ? a[first + lane] : 0u;
    const uint32_t key = mine >> 24;
    int rank = 0;
    for (int j = 0; j < len; j++)
    {
        const uint32_t other = __shfl_sync(0xffffffffu, key, j);
        rank += (other > key) || (other == key && j < lane);
    }
    __syncwarp();
    if (lane < len)
        a[first + rank] = mine;
}

// std::sort's permutation (introsort.cuh), all warps of the CTA: every warp takes ranges of the current
// recursion level -- partition by warp_partition (exact), leaves by rank sort, heap-sort fallback by
// one lane -- with a block barrier between levels.  posL / posR: u16 scratch indexed like the array
// (ranges of one level are disjoint, so each uses its own slice).
__device__ void block_introsort(uint32_t *a, int n, isort::LevelRange *q0, isort::LevelRange *q1, int *s_cnt,
                                uint16_t *posL, uint16_t *posR)
{
    if (n <= 1)
        return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1)
        lg++;
    if (threadIdx.x == 0)
    {
        q0[0] = isort::LevelRange{0, n, 2 * lg};
        s_cnt[0] = 1;
        s_cnt[1] = 0;
    }
    __syncthreads();
    isort::LevelRange *cur = q0, *nxt = q1;
    int which = 0;
    while (true)
    {
        const int ncur = s_cnt[which];
        if (ncur == 0)
            break;
        for (int i = warp; i < ncur; i += nwarps)
        {
            const isort::LevelRange r = cur[i];
            const int len = r.last - r.first;
            if (len <= 16)
            {
                warp_leaf_sort(a, r.first, len);
                continue;
            }
            if (r.depth == 0)
            {
                if (lane == 0)
                    isort::heap_sort(a + r.first, a + r.last);
                continue;
            }
            if (lane == 0)
            {
                uint32_t *f = a + r.first, *l = a + r.last;
                isort::move_median_to_first(f, f + 1, f + (l - f) / 2, l - 1);
            }
            __syncwarp();
            const int cut = warp_partition(a, r.first, r.last, posL + r.first, posR + r.first);
            if (lane == 0)
            {
                const int slot = atomicAdd(&s_cnt[which ^ 1], 2);
                nxt[slot] = isort::LevelRange{r.first, cut, r.depth - 1};
                nxt[slot + 1] = isort::LevelRange{cut, r.last, r.depth - 1};
            }
        }
        __syncthreads();
        if (threadIdx.x == 0)
            s_cnt[which] = 0;
        which ^= 1;
        isort::LevelRange *t = cur;
        cur = nxt;
        nxt = t;
        __syncthreads();
    }
}

__device__ void block_bitonic_sort(uint32_t *keys, int P)
{
    for (int k = 2; k <= P; k <<= 1)
    {
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int i = threadIdx.x; i < P; i += blockDim.x)
            {
                const int ixj = i ^ j;
                if (ixj > i)
                {
                    const uint32_t va = keys[i], vb = keys[ixj];
                    const bool asc = (i & k) == 0;
                    if ((va > vb) == asc)
                    {
                        keys[i] = vb;
                        keys[ixj] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kTileThreads) tile_kernel(TileArgs a)
{
    extern __shared__ uint32_t s_dyn[];
    uint32_t *s_keys = s_dyn, *s_rad = s_dyn + kTileSmemCap, *s_perm = s_dyn + 2 * kTileSmemCap;
    isort::LevelRange *s_q0 = reinterpret_cast<isort::LevelRange *>(s_dyn + 3 * kTileSmemCap), *s_q1 = s_q0 + kTileRanges;
    uint16_t *s_posL = reinterpret_cast<uint16_t *>(s_q1 + kTileRanges), *s_posR = s_posL + kTileSmemCap;
    __shared__ int s_cnt[2];
    __shared__ int s_hist[257];
    __shared__ int s_scan[34];
    __shared__ uint32_t s_prefix;
    __shared__ int s_k;

    const int t = blockIdx.x, b = blockIdx.y;
    if (a.retry && !a.retry[b])
        return;
#define LVT_TDBG(k)                                                                                                   \
    if (a.dbg && threadIdx.x == 0)                                                                                    \
    a.dbg[(b * a.n_tiles + t) * 8 + k] = clock64()
    LVT_TDBG(0);
    const int tx = t % a.grid.nx, ty = t / a.grid.nx;
    const int x0 = tx * a.grid.cell, y0 = ty * a.grid.cell, tw = a.grid.tile_w(tx);
    const size_t base = ((size_t)b * a.n_tiles + t) * a.tile_cap;
    const int n = min(a.tile_count[b * a.n_tiles + t], a.tile_cap);
    uint32_t *out = a.tile_out + base;
    if (n == 0)
    {
        if (threadIdx.x == 0)
            a.tile_out_count[b * a.n_tiles + t] = 0;
        return;
    }
    const bool small = n <= kTileSmemCap;
    uint32_t *keys = small ? s_keys : a.tile_list + base;
    uint32_t *rad = small ? s_rad : a.tile_aux + 2 * base;
    uint32_t *perm = small ? s_perm : a.tile_aux + 2 * base + a.tile_cap;

    // ---- raster order -------------------------------------------------------------------------
    int P = 1;
    while (P < n)
        P <<= 1;
    if (small)
        for (int i = threadIdx.x; i < P; i += blockDim.x)
            keys[i] = i < n ? a.tile_list[base + i] : 0xFFFFFFFFu;
    else
        for (int i = n + threadIdx.x; i < P; i += blockDim.x)
            keys[i] = 0xFFFFFFFFu; // tile_cap is a power of two >= n
    __syncthreads();
    block_bitonic_sort(keys, P);
    // re-encode in place: (raster << 8 | response) -> (ly << 20 | lx << 8 | response); same order
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        const uint32_t key = keys[i];
        const int raster = (int)(key >> 8), ly = raster / tw, lx = raster - ly * tw;
        keys[i] = ((uint32_t)ly << 20) | ((uint32_t)lx << 8) | (key & 0xFFu);
    }
    __syncthreads();
    const uint32_t origin = ((uint32_t)y0 << 20) | ((uint32_t)x0 << 8);
    auto pack_out = [&](uint32_t key) -> uint32_t { return key + origin; };

    if (n <= a.max_per_cell)
    {
        // lvt_image_features_handler.cpp:144-151: raster order, tile offset added
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            out[i] = pack_out(keys[i]);
        if (threadIdx.x == 0)
            a.tile_out_count[b * a.n_tiles + t] = n;
        return;
    }

    LVT_TDBG(1);
    // ---- ANMS (lvt_image_features_handler.cpp:34-83) ------------------------------------------
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        perm[i] = ((keys[i] & 0xFFu) << 24) | (uint32_t)i;
    __syncthreads();
    // :38-41, std::sort's exact permutation
    if (small)
        block_introsort(perm, n, s_q0, s_q1, s_cnt, s_posL, s_posR);
    else if (threadIdx.x == 0)
        isort::sort(perm, n); // more survivors than fit in shared memory (pathological): sequential replay
    __syncthreads();
    LVT_TDBG(2);
    // :52-64  radius^2 = min squared distance to any corner with response > 1.11f * own.  In the sorted
    // order those corners are a prefix: s_hist[v] = number of corners with response >= v.
    for (int i = threadIdx.x; i < 257; i += blockDim.x)
        s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        atomicAdd(&s_hist[keys[i] & 0xFFu], 1);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int acc = 0;
        for (int v = 256; v >= 0; v--)
        {
            acc += s_hist[v];
            s_hist[v] = acc;
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x)
    {
        const int i = (int)(perm[p] & 0xFFFFFFu);
        const uint32_t ki = keys[i];
        // response_j > 1.11f * response_i in fp32  <=>  response_j >= floor(thr) + 1 (integers)
        const float thr = __fmul_rn((float)(ki & 0xFFu), 1.11f);
        const int need = (int)floorf(thr) + 1;
        const int prefix_len = need > 255 ? 0 : s_hist[need];
        const int yi = (int)(ki >> 20), xi = (int)((ki >> 8) & 0xFFFu);
        uint32_t best = 0xFFFFFFFFu; // FLT_MAX
        for (int q = 0; q < prefix_len; q++)
        {
            const uint32_t kj = keys[perm[q] & 0xFFFFFFu];
            const int dx = xi - (int)((kj >> 8) & 0xFFFu), dy = yi - (int)(kj >> 20);
            best = min(best, (uint32_t)(dx * dx + dy * dy));
        }
        rad[i] = best;
    }
    LVT_TDBG(3);
    __syncthreads();
    LVT_TDBG(4);

    // :66-71  decision = radiiSorted[num_to_keep]  (descending) -> MSB-first radix select
    if (threadIdx.x == 0)
    {
        s_prefix = 0;
        s_k = a.max_per_cell;
    }
    for (int shift = 24; shift >= 0; shift -= 8)
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            s_hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if ((rad[i] & mask) == prefix)
                atomicAdd(&s_hist[(rad[i] >> shift) & 0xFF], 1);
        __syncthreads();
        if (threadIdx.x < 32)
        {
            // walk the 256 bins from the top: lane l owns bins 255-8l .. 248-8l
            const int lane = threadIdx.x;
            int mine = 0;
#pragma unroll
            for (int u = 0; u < 8; u++)
                mine += s_hist[255 - 8 * lane - u];
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += v;
            }
            const int k0 = s_k;
            const bool here = (incl - mine) <= k0 && k0 < incl; // the k-th largest falls into my bins
            const uint32_t who = __ballot_sync(0xffffffffu, here);
            if (who == 0)
            {
                if (lane == 0) // fewer than k+1 elements left (cannot happen for n > k): bin 0
                {
                    s_k = k0 - incl; // unused
                    s_prefix = prefix;
                }
            }
            else if (lane == __ffs(who) - 1)
            {
                int k = k0 - (incl - mine), d = 255 - 8 * lane;
                for (int u = 0; u < 8; u++, d--)
                {
                    if (k < s_hist[d])
                        break;
                    k -= s_hist[d];
                }
                s_k = k;
                s_prefix = prefix | ((uint32_t)d << shift);
            }
        }
        __syncthreads();
    }
    const uint32_t decision = s_prefix;
    LVT_TDBG(5);

    // :72-80  keep radius >= decision, in sorted order
    int running = 0;
    for (int p0 = 0; p0 < n; p0 += blockDim.x)
    {
        const int p = p0 + threadIdx.x;
        uint32_t key = 0;
        int keep = 0;
        if (p < n)
        {
            const int i = (int)(perm[p] & 0xFFFFFFu);
            key = keys[i];
            keep = rad[i] >= decision;
        }
        int total;
        const int pos = block_exclusive_scan(keep, s_scan, &total);
        if (keep)
            out[running + pos] = pack_out(key);
        running += total;
    }
    if (threadIdx.x == 0)
        a.tile_out_count[b * a.n_tiles + t] = running;
    LVT_TDBG(6);
    if (a.dbg && threadIdx.x == 0)
        a.dbg[(b * a.n_tiles + t) * 8 + 7] = n;
}

// ---------------------------------------------------------------------------------------------
// K4: gather tiles -> image keypoint list (+ BRIEF border filter, + retry decision)
// ---------------------------------------------------------------------------------------------
struct GatherArgs
{
    const uint32_t *tile_out;
    const int *tile_out_count;
    int *retry, *error;
    FeatDev *feats;
    int n_tiles, tile_cap, rows, cols;
    int border;      // 28 = BRIEF filter, 0 = none
    int pass;        // 0 first, 1 lowered threshold
    int retry_below; // LVT_CORNERS_LOW_TH, or 0 to disable the retry
};

__global__ void __launch_bounds__(1024) gather_kernel(GatherArgs a)
{
    __shared__ int s_pref[1025];
    __shared__ int s_scan[34];
    const int b = blockIdx.x;
    if (a.pass == 1 && !a.retry[b])
        return;
    const int nt = a.n_tiles; // <= 1024
    if (threadIdx.x == 0)
    {
        int acc = 0;
        for (int t = 0; t < nt; t++)
        {
            s_pref[t] = acc;
            acc += a.tile_out_count[b * nt + t];
        }
        s_pref[nt] = acc;
    }
    __syncthreads();
    const int total_in = s_pref[nt];
    if (a.pass == 0)
    {
        const int redo = total_in < a.retry_below;
        if (threadIdx.x == 0)
            a.retry[b] = redo;
        if (redo)
            return;
    }
    const FeatDev f = a.feats[b];
    int running = 0;
    for (int g0 = 0; g0 < total_in; g0 += blockDim.x)
    {
        const int g = g0 + threadIdx.x;
        int keep = 0;
        float x = 0, y = 0, r = 0;
        if (g < total_in)
        {
            int lo = 0, hi = nt - 1; // last tile with s_pref[t] <= g
            while (lo < hi)
            {
                const int mid = (lo + hi + 1) >> 1;
                if (s_pref[mid] <= g)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            const uint32_t e = a.tile_out[((size_t)b * nt + lo) * a.tile_cap + (g - s_pref[lo])];
            const int xi = (e >> 8) & 0xFFF, yi = e >> 20;
            x = (float)xi;
            y = (float)yi;
            r = (float)(e & 0xFF);
            keep = a.border == 0 || (a.rows > 2 * a.border && a.cols > 2 * a.border && xi >= a.border &&
                                     xi < a.cols - a.border && yi >= a.border && yi < a.rows - a.border);
        }
        int total;
        const int pos = block_exclusive_scan(keep, s_scan, &total);
        if (keep)
        {
            const int o = running + pos;
            if (o < f.cap)
            {
                f.xy[o] = make_float2(x, y);
                f.resp[o] = r;
            }
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

__global__ void clear_counts_kernel(int *tile_count, int *tile_overflow, int *cand_count, int n, const int *retry, int n_tiles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (retry && !retry[i / n_tiles])
        return;
    tile_count[i] = 0;
    tile_overflow[i] = 0;
    if (i % n_tiles == 0)
        cand_count[i / n_tiles] = 0;
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
int launch_detect(const ImagePool &pool, const DetectWorkspace &ws, const DetectParams &dp, const int *d_slots,
                  int n_images, FeatDev *d_feats, int border, int nonmax, cudaStream_t stream)
{
    if (n_images > ws.batch || dp.grid.count() != ws.n_tiles || ws.n_tiles > 1024)
        return LVTK_ERR_ARG;
    const int nt = ws.n_tiles;
    static bool smem_set = false;
    if (!smem_set)
    {
        LVT_CUDA_TRY(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
        smem_set = true;
    }
    const bool allow_retry = dp.threshold_low < dp.threshold;

    ScoreArgs sa{d_slots, ws.score, dp.grid, dp.pitch, dp.rows, dp.cols, allow_retry ? dp.threshold_low : dp.threshold};
    dim3 sgrid((dp.cols + kScoreTileW - 1) / kScoreTileW, (dp.rows + kScoreTileH - 1) / kScoreTileH, n_images);
    LVT_TIMED(stream, K_SCORE, (score_kernel<<<sgrid, 256, 0, stream>>>(pool.tmap_score, sa)));
    LVT_LAUNCH_CHECK(stream, "score_kernel");

    for (int pass = 0; pass < (allow_retry ? 2 : 1); pass++)
    {
        const int *retry = pass ? ws.retry : nullptr;
        const int th = pass ? dp.threshold_low : dp.threshold;
        LVT_TIMED(stream, K_CLEAR, (clear_counts_kernel<<<(n_images * nt + 255) / 256, 256, 0, stream>>>(ws.tile_count, ws.tile_overflow, ws.cand_count,
                                                                            n_images * nt, retry, nt)));
        NmsArgs na{ws.score, ws.tile_list, ws.tile_count, ws.tile_overflow, ws.error, reinterpret_cast<uint32_t *>(ws.parent),
                   ws.cand_count, retry, dp.grid,
                   dp.pitch, dp.rows,     dp.cols,       ws.tile_cap,      nt,       th,    nonmax};
        dim3 ngrid((dp.cols + kNmsTile - 1) / kNmsTile, (dp.rows + kNmsTile - 1) / kNmsTile, n_images);
        LVT_TIMED(stream, K_NMS, (nms_tile_kernel<<<ngrid, 256, 0, stream>>>(na)));
        LVT_LAUNCH_CHECK(stream, "nms_tile_kernel");
        if (nonmax)
        {
            LVT_TIMED(stream, K_NMS_FALLBACK, (nms_fallback_kernel<<<dim3(nt, n_images), 32, 0, stream>>>(na, ws.parent)));
            LVT_LAUNCH_CHECK(stream, "nms_fallback_kernel");
        }
        TileArgs ta{ws.tile_list, ws.tile_aux, ws.tile_out, ws.tile_count, ws.tile_out_count, retry,
                    dp.grid,      ws.tile_cap, nt,          dp.max_per_cell,
                    (pass == 0 && debug_sync_enabled()) ? reinterpret_cast<long long *>(ws.tile_aux) : nullptr}; // scratch unused for small tiles
        static bool dumped = false;
        LVT_TIMED(stream, K_TILE, (tile_kernel<<<dim3(nt, n_images), kTileThreads, kTileSmemBytes, stream>>>(ta)));
        LVT_LAUNCH_CHECK(stream, "tile_kernel");
        if (ta.dbg && !dumped && std::getenv("LVT_B200_TILEDBG"))
        {
            static int calls = 0;
            if (++calls == 8)
            {
                dumped = true;
                std::vector<long long> h((size_t)nt * n_images * 8);
                cudaMemcpy(h.data(), ta.dbg, h.size() * 8, cudaMemcpyDeviceToHost);
                for (int i = 0; i < nt * n_images; i++)
                    std::fprintf(stderr, "tile %2d n=%4lld: load+bitonic %6lld | sort %6lld (radii %6lld) | sync %6lld | select %6lld | compact %6lld cycles\n", i,
                                 h[i * 8 + 7], h[i * 8 + 1] - h[i * 8], h[i * 8 + 2] - h[i * 8 + 1], h[i * 8 + 3] - h[i * 8 + 1],
                                 h[i * 8 + 4] - h[i * 8 + 1], h[i * 8 + 5] - h[i * 8 + 4], h[i * 8 + 6] - h[i * 8 + 5]);
            }
        }
        GatherArgs ga{ws.tile_out, ws.tile_out_count, ws.retry, ws.error, d_feats, nt, ws.tile_cap,
                      dp.rows,     dp.cols,           border,   pass,     allow_retry ? kCornersLowTh : 0};
        LVT_TIMED(stream, K_GATHER, (gather_kernel<<<n_images, 1024, 0, stream>>>(ga)));
        LVT_LAUNCH_CHECK(stream, "gather_kernel");
    }
    LVT_CUDA_TRY(cudaGetLastError());
    return LVTK_OK;
}

} // namespace lvtb
