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
//   K2 nms_tile_kernel: one survivor per 4-connected component of corner pixels: components are
//                      labelled by max-propagation in a shared-memory window; OpenCV's union-find
//                      tie-breaking is replayed only where several pixels share the maximum.
//                      Survivors are appended to their detection tile's list.
//   K3 tile_kernel   : per tile: (sequential NMS redo if K2 flagged the tile,) raster sort, then
//                      (n > k) ANMS = suppression radius against the stronger prefix, radix-select
//                      of the (k+1)-th largest, emission in std::sort's order (introsort replay).
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
    int *tile_count, *tile_overflow; // [batch][n_tiles], zeroed here for the NMS that follows
    int *tiles_done;                 // [batch], zeroed here: tile CTAs of the image that have finished
    int n_tiles;
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
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    __shared__ __align__(128) uint8_t tile[kScoreBoxH][kScoreBoxW];
    __shared__ __align__(8) uint64_t bar;

    const int x0 = blockIdx.x * kScoreTileW, y0 = blockIdx.y * kScoreTileH, b = blockIdx.z;
    if (blockIdx.x == 0 && blockIdx.y == 0)
    {
        for (int i = threadIdx.x; i < a.n_tiles; i += blockDim.x)
            a.tile_count[b * a.n_tiles + i] = a.tile_overflow[b * a.n_tiles + i] = 0;
        if (threadIdx.x == 0)
            a.tiles_done[b] = 0;
    }
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
// tile_overflow[t] = number of "big" candidates queued for the tile (low bits) | kTileSequential.
// A big candidate is a local maximum whose component outgrew the per-thread / per-warp replay buffers
// (kCompCap pixels); tile_kernel floods it with the whole CTA (tile_big_candidates).  kTileSequential
// sends the whole tile through the one-thread fallback (more than kBigCandCap big candidates, more
// slow candidates than a CTA's lists hold, components beyond kBigMemberCap pixels).
constexpr int kTileSequential = 0x40000000;
constexpr int kBigCandCap = 64;

struct NmsArgs
{
    const uint8_t *score;
    uint32_t *tile_list;
    uint32_t *big_list; // [batch][n_tiles][kBigCandCap]  (s << 24 | y << 12 | x)
    int *tile_count, *tile_overflow, *error;
    const int *retry; // nullptr on the first pass
    TileGrid grid;
    int pitch, rows, cols, tile_cap, n_tiles;
    int threshold;
    int nonmax;
};


constexpr int kCompCap = 96; // largest component replayed in registers/local memory

__device__ __forceinline__ int tile_of(const NmsArgs &a, int x, int y) { return (y / a.grid.cell) * a.grid.nx + x / a.grid.cell; }

// the component of the local maximum (x, y, s) is too large for this kernel's buffers: tile_kernel settles it
__device__ void queue_big_candidate(const NmsArgs &a, int b, int x, int y, int s)
{
    const int t = b * a.n_tiles + tile_of(a, x, y);
    const int pos = atomicAdd(&a.tile_overflow[t], 1) & (kTileSequential - 1);
    if (pos < kBigCandCap)
        a.big_list[(size_t)t * kBigCandCap + pos] = ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x;
    else
        atomicOr(&a.tile_overflow[t], kTileSequential);
}

// Appends the CTA's survivors (s << 24 | y << 12 | x) to their detection tiles' lists.  Lanes of a
// warp that hit the same tile reserve their slots with one atomicAdd (a few hot counters would
// otherwise serialise tens of thousands of returning atomics per image).
__device__ void flush_survivors(const NmsArgs &a, int b, const uint32_t *surv, int n)
{
    for (int i0 = 0; i0 < n; i0 += blockDim.x)
    {
        const int i = i0 + threadIdx.x;
        const bool on = i < n;
        const unsigned active = __ballot_sync(0xFFFFFFFFu, on);
        if (!on)
            continue;
        const uint32_t c = surv[i];
        const int x = (int)(c & 0xFFFu), y = (int)((c >> 12) & 0xFFFu), s = (int)(c >> 24);
        const int tx = x / a.grid.cell, ty = y / a.grid.cell, t = ty * a.grid.nx + tx;
        const int raster = (y - ty * a.grid.cell) * a.grid.tile_w(tx) + (x - tx * a.grid.cell);
        const unsigned peers = __match_any_sync(active, t);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader)
            base = atomicAdd(&a.tile_count[b * a.n_tiles + t], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        const int pos = base + __popc(peers & ((1u << lane) - 1u));
        if (pos < a.tile_cap)
            a.tile_list[((size_t)b * a.n_tiles + t) * a.tile_cap + pos] = ((uint32_t)raster << 8) | (uint32_t)s;
        else
            *a.error = LVTK_ERR_CAPACITY;
    }
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

// Fused NMS.  A CTA stages a 64x64 window of the score map (its 32x32 pixels + 16 px halo) in shared
// memory and labels the 4-connected corner components inside it by max-propagation: every corner
// starts with (score << 12 | local index) and repeatedly takes the maximum over itself and its four
// neighbours until nothing changes, so each pixel ends up holding the score and position of its
// component's maximum.  Then, per pixel of the interior:
//   * score below the component maximum             -> suppressed
//   * the only pixel carrying the maximum           -> survivor          (the common case)
//   * one of several pixels carrying the maximum    -> OpenCV's merge order decides (warp_replay_component)
//   * component reaches the window's outer ring     -> labels may be incomplete: warp_nms_resolve
// The propagation is a few dozen shared-memory sweeps over the corner list; the sequential flood
// of warp_nms_resolve is left for components that leave the window.
constexpr int kNmsTile = 48, kNmsHalo = 8, kNmsWin = kNmsTile + 2 * kNmsHalo;
constexpr uint32_t kNmsOpen = 0xFFFFFFFFu; // label of a corner on the window's ring
constexpr int kNmsSlowCap = 256;           // per CTA: components with a shared maximum / open-label candidates;
                                           // beyond that the detection tile is redone sequentially


// One warp settles a component whose maximum is shared by several pixels.  The component is the set
// of window pixels labelled L (complete: its label is closed).  The lanes collect the members in
// raster order and link each to the member directly above; lane 0 then replays OpenCV's merge
// sequence (same decisions as replay_component) on shared-memory arrays.  Returns the local index of
// the surviving root, or -1 if the component has more than kCompCap pixels.
struct alignas(4) WarpComp
{
    uint16_t mem[kCompCap];
    uint8_t above[kCompCap], sc[kCompCap];
    int8_t par[kCompCap];
};
// the same per-warp shared memory while a warp floods a component that leaves the window (warp_nms_resolve: behind the
// warp's last warp_replay_component job, so the two uses never overlap)
struct WarpFlood
{
    uint32_t pts[kCompCap]; // y << 16 | x, image coordinates
    uint8_t sc[kCompCap];
};
static_assert(sizeof(WarpFlood) <= sizeof(WarpComp), "the flood's member list lives in the warp's WarpComp");

// A local maximum (x, y, s) whose component reaches the window's ring: one WARP floods the component through the
// score map, breadth first -- every step takes eight members of the current layer and their four neighbours each,
// one neighbour per lane, so the trips to the score map outside the window (L2) are paid per layer, not per pixel
// (the one-thread flood this replaces took up to 120 us for a single candidate on densely cornered images).  The
// outcome does not depend on the order of the walk: a stronger pixel anywhere in the component suppresses the
// candidate; more than kCompCap members send it to tile_kernel (which floods it with a whole CTA); a unique maximum
// survives; with several pixels at the maximum the members are sorted to raster order and lane 0 replays OpenCV's
// merge sequence (replay_component).  All lanes call; returns the verdict in every lane.
template <class ScoreAt>
__device__ bool warp_nms_resolve(const NmsArgs &a, int b, ScoreAt score_at, int x, int y, int s, WarpFlood &w)
{
    const int lane = threadIdx.x & 31;
    int n = 1, head = 0;
    bool tie = false;
    if (lane == 0)
    {
        w.pts[0] = ((uint32_t)y << 16) | (uint32_t)x;
        w.sc[0] = (uint8_t)s;
    }
    __syncwarp();
    while (head < n)
    {
        const int end = n; // the layer [head, end); members found on the way are appended behind it
        for (int base = head; base < end; base += 8)
        {
            const int j = base + (lane >> 2), d = lane & 3;
            bool cand = false;
            uint32_t q = 0;
            int v = 0;
            if (j < end)
            {
                const uint32_t p = w.pts[j];
                const int qx = (int)(p & 0xFFFFu) + (d == 0) - (d == 1), qy = (int)(p >> 16) + (d == 2) - (d == 3);
                v = score_at(qx, qy);
                q = ((uint32_t)qy << 16) | (uint32_t)qx;
                cand = v != 0;
            }
            if (__any_sync(0xFFFFFFFFu, cand && v > s))
                return false; // beaten inside its own component
            if (cand) // a member already?
                for (int k = 0; k < n; k++)
                    if (w.pts[k] == q)
                    {
                        cand = false;
                        break;
                    }
            // the same pixel reached from two members in this step: the lowest lane keeps it
            const unsigned active = __ballot_sync(0xFFFFFFFFu, cand);
            if (cand)
            {
                const unsigned peers = __match_any_sync(active, q);
                cand = lane == __ffs(peers) - 1;
            }
            const unsigned add = __ballot_sync(0xFFFFFFFFu, cand);
            const int cnt = __popc(add);
            if (n + cnt > kCompCap)
            {
                if (lane == 0)
                    queue_big_candidate(a, b, x, y, s); // tile_kernel floods it with the whole CTA
                return false;
            }
            if (cand)
            {
                const int pos = n + __popc(add & ((1u << lane) - 1u));
                w.pts[pos] = q;
                w.sc[pos] = (uint8_t)v;
            }
            tie |= __any_sync(0xFFFFFFFFu, cand && v == s);
            n += cnt;
            __syncwarp();
        }
        head = end;
    }
    if (!tie)
        return true; // unique maximum of its component
    // several pixels share the maximum: raster order (rank sort through registers: the keys are unique), then the replay
    const uint32_t self = ((uint32_t)y << 16) | (uint32_t)x;
    uint32_t mp[3];
    uint8_t ms[3];
    int rank[3];
#pragma unroll
    for (int u = 0; u < 3; u++)
    {
        const int i = lane + 32 * u;
        mp[u] = i < n ? w.pts[i] : 0xFFFFFFFFu;
        ms[u] = i < n ? w.sc[i] : (uint8_t)0;
        rank[u] = 0;
    }
    for (int k = 0; k < n; k++)
    {
        const uint32_t o = w.pts[k];
#pragma unroll
        for (int u = 0; u < 3; u++)
            rank[u] += o < mp[u];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 3; u++)
        if (lane + 32 * u < n)
        {
            w.pts[rank[u]] = mp[u];
            w.sc[rank[u]] = ms[u];
        }
    __syncwarp();
    int root = 0;
    if (lane == 0)
        root = replay_component(w.pts, w.sc, n);
    root = __shfl_sync(0xFFFFFFFFu, root, 0);
    return w.pts[root] == self;
}

__device__ int warp_replay_component(const uint32_t *lab, const uint8_t *win, int n_win, int row_w, uint32_t L, WarpComp &w)
{
    const int lane = threadIdx.x & 31;
    int n = 0;
    for (int base = 0; base < n_win; base += 32)
    {
        const bool m = lab[base + lane] == L;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, m);
        if (m)
        {
            const int pos = n + __popc(bal & ((1u << lane) - 1u));
            if (pos < kCompCap)
                w.mem[pos] = (uint16_t)(base + lane);
        }
        n += __popc(bal);
    }
    if (n > kCompCap)
        return -1;
    __syncwarp();
    for (int j = lane; j < n; j += 32)
    {
        const int lid = w.mem[j], up = lid - row_w;
        int ab = 255;
        if (lab[up] == L)
        {
            int lo = 0, hi = j - 1; // mem is ascending and contains up
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                if (w.mem[mid] < up)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            ab = lo;
        }
        w.above[j] = (uint8_t)ab;
        w.par[j] = -1;
        w.sc[j] = win[lid];
    }
    __syncwarp();
    int root = 0;
    if (lane == 0)
    {
        auto find = [&](int i) {
            while (w.par[i] != -1)
                i = w.par[i];
            return i;
        };
        for (int c = 0; c < n; c++)
        {
            if (w.above[c] != 255)
            {
                const int r = find(w.above[c]);
                if (w.sc[c] < w.sc[r])
                    w.par[c] = (int8_t)r;
                else
                    w.par[r] = (int8_t)c;
            }
            if (c != 0 && w.mem[c - 1] + 1 == w.mem[c])
            {
                const int pa = w.par[c], t = find(c - 1);
                if (pa == -1)
                {
                    if (t != c)
                    {
                        if (w.sc[c] < w.sc[t])
                            w.par[c] = (int8_t)t;
                        else
                            w.par[t] = (int8_t)c;
                    }
                }
                else if (t != pa)
                {
                    if (w.sc[pa] < w.sc[t])
                    {
                        w.par[pa] = (int8_t)t;
                        w.par[c] = (int8_t)t;
                    }
                    else
                    {
                        w.par[t] = (int8_t)pa;
                        w.par[c] = (int8_t)pa;
                    }
                }
            }
        }
        for (int i = 0; i < n; i++)
            if (w.par[i] == -1)
            {
                root = i;
                break;
            }
        root = w.mem[root];
    }
    return __shfl_sync(0xFFFFFFFFu, root, 0);
}

#ifdef LVT_NMS_STATS
// per-CTA trace of the last full launch: [0] start ns, [1] end ns, [2..6] cycles of stage / propagate / classify /
// resolve / flush, [7] slow candidates, [8] corners in the window, [9] survivors, [10] sweeps, [11] longest resolve
__device__ long long g_nms_trace[4096][12];
__device__ __forceinline__ long long nms_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define NMS_TR(k, v)                                                                                                  \
    if (threadIdx.x == 0)                                                                                             \
    g_nms_trace[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x][k] = (long long)(v)
#else
#define NMS_TR(k, v)
#endif

__global__ void __launch_bounds__(256) nms_tile_kernel(NmsArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
#ifdef LVT_NMS_STATS
    __shared__ unsigned long long s_longest;
    if (threadIdx.x == 0)
        s_longest = 0;
    long long ck = clock64(), cn;
    NMS_TR(0, nms_ns());
#define NMS_PHASE(k)                                                                                                  \
    cn = clock64();                                                                                                   \
    NMS_TR(k, cn - ck);                                                                                               \
    ck = cn
#else
#define NMS_PHASE(k)
#endif
    __shared__ __align__(16) uint8_t win[kNmsWin * kNmsWin];
    __shared__ uint32_t lab[kNmsWin * kNmsWin];
    __shared__ uint16_t s_list[kNmsWin * kNmsWin];
    __shared__ __align__(16) uint8_t s_tied[kNmsWin * kNmsWin];
    __shared__ uint32_t s_cand[kNmsSlowCap], s_surv[kNmsTile * kNmsTile];
    __shared__ uint16_t s_job[kNmsSlowCap];
    __shared__ WarpComp s_comp[8];
    __shared__ int s_nlist, s_ncand, s_nsurv, s_njob;
    const int b = blockIdx.z;
    if (a.retry && !a.retry[b])
        return;
    const uint8_t *sm = a.score + (size_t)b * a.rows * a.pitch;
    const int x0 = blockIdx.x * kNmsTile - kNmsHalo, y0 = blockIdx.y * kNmsTile - kNmsHalo; // x0 is a multiple of 16
    if (threadIdx.x == 0)
        s_nlist = s_ncand = s_nsurv = s_njob = 0;
    // stage the window, thresholded: 0 = not a corner
    const uint32_t th4 = (uint32_t)a.threshold * 0x01010101u;
    for (int w = threadIdx.x; w < kNmsWin * kNmsWin / 4; w += blockDim.x)
    {
        const int row = w / (kNmsWin / 4), gx = x0 + 4 * (w % (kNmsWin / 4)), gy = y0 + row;
        uint32_t v = 0;
        if (gy >= 0 && gy < a.rows && gx >= 0 && gx + 3 < a.pitch)
            v = *reinterpret_cast<const uint32_t *>(sm + (size_t)gy * a.pitch + gx); // pitch padding is zero
        if (v)
        {
            const uint32_t ge = __vcmpgeu4(v, th4); // 0xFF per byte >= threshold
            v &= ge;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (gx + k >= a.cols)
                    v &= ~(0xFFu << (8 * k));
        }
        reinterpret_cast<uint32_t *>(win)[w] = v;
        reinterpret_cast<uint32_t *>(s_tied)[w] = 0;
    }
    __syncthreads();
    auto score_at = [&](int xx, int yy) -> int {
        if (xx < 0 || yy < 0 || xx >= a.cols || yy >= a.rows)
            return 0;
        const int lx = xx - x0, ly = yy - y0;
        if ((unsigned)lx < (unsigned)kNmsWin && (unsigned)ly < (unsigned)kNmsWin)
            return win[ly * kNmsWin + lx];
        const int v = sm[(size_t)yy * a.pitch + xx];
        return v >= a.threshold ? v : 0;
    };
    if (!a.nonmax)
    {
        for (int i = threadIdx.x; i < kNmsTile * kNmsTile; i += blockDim.x)
        {
            const int lx = kNmsHalo + i % kNmsTile, ly = kNmsHalo + i / kNmsTile;
            const int s = win[ly * kNmsWin + lx];
            if (s && y0 + ly < a.rows)
                s_surv[atomicAdd(&s_nsurv, 1)] = ((uint32_t)s << 24) | ((uint32_t)(y0 + ly) << 12) | (uint32_t)(x0 + lx);
        }
        __syncthreads();
        flush_survivors(a, b, s_surv, s_nsurv);
        return;
    }
    // initial labels + list of the corners that take part in the propagation
    for (int i = threadIdx.x; i < kNmsWin * kNmsWin; i += blockDim.x)
    {
        const int v = win[i];
        const int lx = i % kNmsWin, ly = i / kNmsWin;
        const bool ring = lx == 0 || ly == 0 || lx == kNmsWin - 1 || ly == kNmsWin - 1;
        uint32_t l = 0;
        if (v)
        {
            l = ring ? kNmsOpen : (((uint32_t)v << 12) | (uint32_t)i);
            if (!ring)
                s_list[atomicAdd(&s_nlist, 1)] = (uint16_t)i;
        }
        lab[i] = l;
    }
    __syncthreads();
    const int nl = s_nlist;
    NMS_PHASE(2);
    NMS_TR(8, nl);
    NMS_TR(1, nms_ns());
    if (nl == 0)
        return;
    int sweeps = 0;
    // max-propagation to the fixed point (in place: the update is monotone, any order converges)
    for (;;)
    {
        int changed = 0;
        for (int k = threadIdx.x; k < nl; k += blockDim.x)
        {
            const int i = s_list[k];
            const uint32_t cur = lab[i];
            const uint32_t m = max(max(max(lab[i - 1], lab[i + 1]), max(lab[i - kNmsWin], lab[i + kNmsWin])), cur);
            if (m != cur)
            {
                lab[i] = m;
                changed = 1;
            }
        }
        sweeps++;
        if (!__syncthreads_or(changed))
            break;
    }
    NMS_PHASE(3);
    NMS_TR(10, sweeps);
    // a pixel that carries its component's maximum but is not the label holder: the maximum is shared
    for (int k = threadIdx.x; k < nl; k += blockDim.x)
    {
        const int i = s_list[k];
        const uint32_t l = lab[i];
        if (l != kNmsOpen && (l >> 12) == win[i] && (int)(l & 0xFFFu) != i)
            s_tied[l & 0xFFFu] = 1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nl; k += blockDim.x)
    {
        const int i = s_list[k];
        const int lx = i % kNmsWin, ly = i / kNmsWin;
        if (lx < kNmsHalo || lx >= kNmsHalo + kNmsTile || ly < kNmsHalo || ly >= kNmsHalo + kNmsTile)
            continue;
        const int s = win[i], x = x0 + lx, y = y0 + ly;
        const uint32_t l = lab[i];
        if (l != kNmsOpen)
        {
            if ((int)(l >> 12) != s)
                continue; // not the maximum of its component
            if ((int)(l & 0xFFFu) == i && !s_tied[i])
                s_surv[atomicAdd(&s_nsurv, 1)] = ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x; // unique maximum
            else
                s_tied[l & 0xFFFu] = 2; // shared maximum with a pixel in this CTA's interior: settle the component
            continue;
        }
        if (max(max(win[i - 1], win[i + 1]), max(win[i - kNmsWin], win[i + kNmsWin])) > s)
            continue; // beaten by a neighbour
        const int pos = atomicAdd(&s_ncand, 1);
        if (pos < kNmsSlowCap)
            s_cand[pos] = ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x;
        else
            atomicOr(&a.tile_overflow[b * a.n_tiles + tile_of(a, x, y)], kTileSequential);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nl; k += blockDim.x)
    {
        const int i = s_list[k];
        if (s_tied[i] == 2 && (int)(lab[i] & 0xFFFu) == i)
        {
            const int pos = atomicAdd(&s_njob, 1);
            if (pos < kNmsSlowCap)
                s_job[pos] = (uint16_t)i;
            else
            {
                // the component lies in one detection tile (tile borders carry no corners)
                const int gx = x0 + i % kNmsWin, gy = y0 + i / kNmsWin;
                atomicOr(&a.tile_overflow[b * a.n_tiles + tile_of(a, gx, gy)], kTileSequential);
            }
        }
    }
    __syncthreads();
    NMS_PHASE(4);
    {
        const int njob = min(s_njob, kNmsSlowCap);
        const int warp = threadIdx.x >> 5;
        for (int j = warp; j < njob; j += 8)
        {
            const int h = s_job[j];
            const uint32_t L = lab[h];
            const int r = warp_replay_component(lab, win, kNmsWin * kNmsWin, kNmsWin, L, s_comp[warp]);
            if (r < 0)
            {
                // more than kCompCap pixels: every interior pixel that carries the maximum becomes a big
                // candidate; the one that turns out to be the root is emitted by tile_kernel
                for (int i = threadIdx.x & 31; i < kNmsWin * kNmsWin; i += 32)
                {
                    const int lx = i % kNmsWin, ly = i / kNmsWin;
                    if (lab[i] == L && win[i] == (L >> 12) && lx >= kNmsHalo && lx < kNmsHalo + kNmsTile && ly >= kNmsHalo &&
                        ly < kNmsHalo + kNmsTile)
                        queue_big_candidate(a, b, x0 + lx, y0 + ly, (int)win[i]);
                }
            }
            else if ((threadIdx.x & 31) == 0)
            {
                const int q = r, lx = q % kNmsWin, ly = q / kNmsWin;
                if (lx >= kNmsHalo && lx < kNmsHalo + kNmsTile && ly >= kNmsHalo && ly < kNmsHalo + kNmsTile)
                    s_surv[atomicAdd(&s_nsurv, 1)] = ((uint32_t)win[r] << 24) | ((uint32_t)(y0 + ly) << 12) | (uint32_t)(x0 + lx);
            }
            __syncwarp();
        }
    }
    const int n = min(s_ncand, kNmsSlowCap);
    for (int i = threadIdx.x >> 5; i < n; i += 8) // a warp per candidate
    {
        const uint32_t c = s_cand[i];
#ifdef LVT_NMS_STATS
        const long long r0 = clock64();
#endif
        const bool keep = warp_nms_resolve(a, b, score_at, (int)(c & 0xFFFu), (int)((c >> 12) & 0xFFFu), (int)(c >> 24),
                                           *reinterpret_cast<WarpFlood *>(&s_comp[threadIdx.x >> 5]));
        if (keep && (threadIdx.x & 31) == 0)
            s_surv[atomicAdd(&s_nsurv, 1)] = c;
#ifdef LVT_NMS_STATS
        if ((threadIdx.x & 31) == 0)
            atomicMax(&s_longest, (unsigned long long)(clock64() - r0));
#endif
        __syncwarp();
    }
    __syncthreads();
    NMS_PHASE(5);
    flush_survivors(a, b, s_surv, s_nsurv);
#ifdef LVT_NMS_STATS
    __syncthreads();
    NMS_PHASE(6);
    NMS_TR(7, n);
    NMS_TR(9, s_nsurv);
    NMS_TR(11, s_longest);
    NMS_TR(1, nms_ns());
#endif
}

// K2b: sequential fallback for a tile flagged by the NMS kernel (pathological inputs only); run by
// one thread of the tile's tile_kernel CTA before it reads the list
__device__ void nms_fallback_tile(const NmsArgs &a, int *parent_all, int t, int b)
{
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
struct GatherArgs
{
    const uint32_t *tile_out;
    const int *tile_out_count;
    int *retry, *error;
    int *tile_count, *tile_overflow; // zeroed for an image that goes into the lowered-threshold pass
    int *tiles_done;                 // [batch] tile CTAs finished (the last one gathers)
    FeatDev *feats;
    int n_tiles, tile_cap, rows, cols;
    int border;      // 28 = BRIEF filter, 0 = none
    int pass;        // 0 first, 1 lowered threshold
    int retry_below; // LVT_CORNERS_LOW_TH, or 0 to disable the retry
};

struct TileArgs
{
    uint32_t *tile_list, *tile_aux, *tile_out;
    const int *tile_count;
    int *tile_out_count;
    const int *retry;
    TileGrid grid;
    int tile_cap, n_tiles, max_per_cell;
    NmsArgs nms; // for the sequential NMS fallback
    int *parent;
    GatherArgs gather; // the image's last tile CTA concatenates the tiles (gather_image)
    int grid_min;      // corners per tile from which the suppression radii use the cell grid
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
    int nL = 0, nR = 0;
    for (int base = 0; base < m; base += 32)
    {
        const int p = base + lane, pr = m - 1 - p; // from the left / from the right
        const bool isL = p < m && (a[lo + p] >> 24) <= pv;
        const bool isR = pr >= 0 && (a[lo + pr] >> 24) >= pv;
        const uint32_t bl = __ballot_sync(0xffffffffu, isL), br = __ballot_sync(0xffffffffu, isR);
        if (isL)
            posL[nL + __popc(bl & ((1u << lane) - 1u))] = (uint16_t)p;
        if (isR)
            posR[nR + __popc(br & ((1u << lane) - 1u))] = (uint16_t)pr;
        nL += __popc(bl);
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
__device__ void block_introsort(uint32_t *a, int n, isort::LevelRange *q0, isort::LevelRange *q1, int *s_cnt /* [3] */,
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
        s_cnt[2] = 0;
    }
    __syncthreads();
    isort::LevelRange *cur = q0, *nxt = q1;
    // counters rotate: level L reads s_cnt[L % 3], fills s_cnt[(L + 1) % 3] and clears s_cnt[(L + 2) % 3]
    // (last read one barrier ago), so a level costs a single barrier
    for (int which = 0;; which = (which + 1) % 3)
    {
        const int ncur = s_cnt[which];
        if (ncur == 0)
            break;
        int *cnt_next = &s_cnt[(which + 1) % 3];
        if (threadIdx.x == 0)
            s_cnt[(which + 2) % 3] = 0;
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
                const int slot = atomicAdd(cnt_next, 2);
                nxt[slot] = isort::LevelRange{r.first, cut, r.depth - 1};
                nxt[slot + 1] = isort::LevelRange{cut, r.last, r.depth - 1};
            }
        }
        __syncthreads();
        isort::LevelRange *t = cur;
        cur = nxt;
        nxt = t;
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

// Big candidates of one detection tile (queue_big_candidate), settled by the tile's whole CTA.
// For each candidate c = (x, y, s): breadth-first flood of its 4-connected corner component through the
// score map (one layer per step, all threads; visited = one bit per tile pixel in shared memory, the
// members (raster << 8 | score) in a shared list whose layers are contiguous).  A larger score anywhere
// suppresses c.  A unique maximum survives.  With several pixels at the maximum the members are sorted
// to raster order and one thread replays OpenCV's merge sequence (the decisions of replay_component);
// c survives iff it is the root.  Survivors are appended to the tile's list.  Components beyond
// kBigMemberCap pixels send the tile to the one-thread fallback (returns false).
constexpr int kBigMemberCap = 8192;

__device__ bool tile_big_candidates(const NmsArgs &a, int t, int b, int n_big, uint32_t *s_members /* [kBigMemberCap] */,
                                    uint32_t *s_bm /* [tw * th / 32 + 1] */, int16_t *s_parent /* [kBigMemberCap] */,
                                    int *s_ctl /* [4] */)
{
    const uint8_t *sm = a.score + (size_t)b * a.rows * a.pitch;
    const int tx = t % a.grid.nx, ty = t / a.grid.nx;
    const int x0 = tx * a.grid.cell, y0 = ty * a.grid.cell, tw = a.grid.tile_w(tx), th = a.grid.tile_h(ty);
    const int nw = (tw * th + 31) >> 5;
    uint32_t *list = a.tile_list + ((size_t)b * a.n_tiles + t) * a.tile_cap;
    auto score_at = [&](int lx, int ly) -> int {
        if (lx < 3 || ly < 3 || lx > tw - 4 || ly > th - 4)
            return 0; // a tile is an image of its own: no corners within 3 px of its border
        const int v = sm[(size_t)(y0 + ly) * a.pitch + x0 + lx];
        return v >= a.threshold ? v : 0;
    };
    for (int k = 0; k < n_big; k++)
    {
        const uint32_t c = a.big_list[((size_t)b * a.n_tiles + t) * kBigCandCap + k];
        const int s = (int)(c >> 24), cx = (int)(c & 0xFFFu) - x0, cy = (int)((c >> 12) & 0xFFFu) - y0;
        const uint32_t self = ((uint32_t)(cy * tw + cx) << 8) | (uint32_t)s;
        for (int i = threadIdx.x; i < nw; i += blockDim.x)
            s_bm[i] = 0;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            s_members[0] = self;
            s_bm[(cy * tw + cx) >> 5] = 1u << ((cy * tw + cx) & 31);
            s_ctl[0] = 1; // members
            s_ctl[1] = 0; // beaten
            s_ctl[2] = 0; // maximum shared
        }
        __syncthreads();
        int lo = 0;
        for (;;)
        {
            const int hi = min(s_ctl[0], kBigMemberCap);
            __syncthreads(); // everybody has read the layer bounds
            for (int i = lo + threadIdx.x; i < hi; i += blockDim.x)
            {
                const int r = (int)(s_members[i] >> 8), py = r / tw, px = r - py * tw;
                int vv[4];
#pragma unroll
                for (int d = 0; d < 4; d++) // independent loads, issued together
                    vv[d] = score_at(px + (d == 0) - (d == 1), py + (d == 2) - (d == 3));
#pragma unroll
                for (int d = 0; d < 4; d++)
                {
                    const int v = vv[d];
                    if (v == 0)
                        continue;
                    if (v > s)
                    {
                        s_ctl[1] = 1;
                        continue;
                    }
                    const int qr = (py + (d == 2) - (d == 3)) * tw + px + (d == 0) - (d == 1);
                    const uint32_t bit = 1u << (qr & 31);
                    if (atomicOr(&s_bm[qr >> 5], bit) & bit)
                        continue; // seen
                    if (v == s)
                        s_ctl[2] = 1;
                    const int pos = atomicAdd(&s_ctl[0], 1);
                    if (pos < kBigMemberCap)
                        s_members[pos] = ((uint32_t)qr << 8) | (uint32_t)v;
                }
            }
            __syncthreads();
            if (s_ctl[1] || s_ctl[0] == hi || s_ctl[0] > kBigMemberCap)
                break;
            lo = hi;
        }
        const int n = s_ctl[0];
        const bool beaten = s_ctl[1] != 0, tie = s_ctl[2] != 0;
        __syncthreads();
        if (beaten)
            continue;
        if (n > kBigMemberCap)
            return false;
        bool survives = true;
        if (tie)
        {
            int P = 1;
            while (P < n)
                P <<= 1;
            for (int i = n + threadIdx.x; i < P; i += blockDim.x)
                s_members[i] = 0xFFFFFFFFu;
            __syncthreads();
            block_bitonic_sort(s_members, P); // raster order (keys are raster << 8 | score)
            if (threadIdx.x == 0)
            {
                auto raster = [&](int i) { return (int)(s_members[i] >> 8); };
                auto sc = [&](int i) { return (int)(s_members[i] & 0xFFu); };
                auto find = [&](int i) {
                    while (s_parent[i] != -1)
                        i = s_parent[i];
                    return i;
                };
                for (int i = 0; i < n; i++)
                    s_parent[i] = -1;
                for (int m = 0; m < n; m++)
                {
                    const int p = raster(m);
                    // member directly above: raster p - tw, somewhere in [0, m)
                    int above = -1;
                    if (p >= tw)
                    {
                        int l = 0, h = m - 1;
                        while (l < h)
                        {
                            const int mid = (l + h) >> 1;
                            if (raster(mid) < p - tw)
                                l = mid + 1;
                            else
                                h = mid;
                        }
                        if (m > 0 && raster(l) == p - tw)
                            above = l;
                    }
                    if (above >= 0)
                    {
                        const int w = find(above);
                        if (sc(m) < sc(w))
                            s_parent[m] = (int16_t)w;
                        else
                            s_parent[w] = (int16_t)m;
                    }
                    if (m != 0 && raster(m - 1) + 1 == p && p % tw != 0)
                    {
                        const int pa = s_parent[m], l = find(m - 1);
                        if (pa == -1)
                        {
                            if (l != m)
                            {
                                if (sc(m) < sc(l))
                                    s_parent[m] = (int16_t)l;
                                else
                                    s_parent[l] = (int16_t)m;
                            }
                        }
                        else if (l != pa)
                        {
                            if (sc(pa) < sc(l))
                            {
                                s_parent[pa] = (int16_t)l;
                                s_parent[m] = (int16_t)l;
                            }
                            else
                            {
                                s_parent[l] = (int16_t)pa;
                                s_parent[m] = (int16_t)pa;
                            }
                        }
                    }
                }
                int root = 0;
                for (int i = 0; i < n; i++)
                    if (s_parent[i] == -1)
                        root = i;
                s_ctl[3] = s_members[root] == self;
            }
            __syncthreads();
            survives = s_ctl[3] != 0;
        }
        if (survives && threadIdx.x == 0)
        {
            const int pos = atomicAdd(&a.tile_count[b * a.n_tiles + t], 1);
            if (pos < a.tile_cap)
                list[pos] = self;
            else
                *a.error = LVTK_ERR_CAPACITY;
        }
        __syncthreads();
    }
    return true;
}

#ifdef LVT_NMS_STATS
__device__ long long g_tile_trace[2048][12];
#define TILE_TR(k, v)                                                                                                 \
    if (threadIdx.x == 0)                                                                                             \
    g_tile_trace[blockIdx.y * gridDim.x + blockIdx.x][k] = (long long)(v)
#define TILE_PHASE(k)                                                                                                 \
    __syncthreads();                                                                                                  \
    tcn = clock64();                                                                                                  \
    TILE_TR(k, tcn - tck);                                                                                            \
    tck = tcn
#else
#define TILE_TR(k, v)
#define TILE_PHASE(k)
#endif

// ---------------------------------------------------------------------------------------------
// K4: gather tiles -> image keypoint list (+ BRIEF border filter, + retry decision)
// ---------------------------------------------------------------------------------------------

// One CTA of 1024 threads per image: run by the LAST tile CTA of the image to finish (tile_kernel), so the
// concatenation costs no launch of its own.  s_pref: 1025 ints, s_scan: 34 ints of shared memory.
__device__ void gather_image(const GatherArgs &a, int b, int *s_pref, int *s_scan)
{
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
        {
            for (int t = threadIdx.x; t < nt; t += blockDim.x)
                a.tile_count[b * nt + t] = a.tile_overflow[b * nt + t] = 0;
            if (threadIdx.x == 0)
                a.tiles_done[b] = 0; // the lowered-threshold pass counts its tiles again
            return;
        }
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


constexpr int kTileGridMin = 1024;      // corners per tile from which the suppression radii use the cell grid
constexpr int kTileGridMinPrefix = 192; // ... for corners with at least this many stronger corners (shorter prefixes: direct scan)
constexpr int kTileBitmapBits = 65536; // raster ranks by bitmap for tiles of up to 256 x 256 pixels

__device__ void tile_body(const TileArgs &a);

__global__ void __launch_bounds__(kTileThreads, 1) tile_kernel(TileArgs a)
{
    // Wait for the previous kernel, but do NOT release the next one yet: the kernel behind this one is the
    // second NMS pass, hundreds of CTAs with 42 KB of shared memory each, which would sit on every SM
    // (waiting for this grid to finish) for as long as the slowest tile takes and keep the other
    // extraction pipelines of lvt_track_pool off the machine.  It is released after the tile work.
    cudaGridDependencySynchronize();
    const int b = blockIdx.y;
    if (a.retry && !a.retry[b])
        return; // lowered-threshold pass of an image that does not need it
    tile_body(a);
    cudaTriggerProgrammaticLaunchCompletion();
    // the image's last tile CTA to get here concatenates the tiles (what used to be a kernel of its own)
    __shared__ int s_last;
    __shared__ int s_pref[1025];
    __shared__ int s_gscan[34];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&a.gather.tiles_done[b], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last)
    {
        __threadfence();
        gather_image(a.gather, b, s_pref, s_gscan);
    }
}

__device__ void tile_body(const TileArgs &a)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t *s_keys = s_dyn, *s_rad = s_dyn + kTileSmemCap, *s_perm = s_dyn + 2 * kTileSmemCap;
    isort::LevelRange *s_q0 = reinterpret_cast<isort::LevelRange *>(s_dyn + 3 * kTileSmemCap), *s_q1 = s_q0 + kTileRanges;
    uint16_t *s_posL = reinterpret_cast<uint16_t *>(s_q1 + kTileRanges), *s_posR = s_posL + kTileSmemCap;
    uint32_t *s_sorted = reinterpret_cast<uint32_t *>(s_posL); // the partition scratch is free once the order is known
    __shared__ int s_cnt[3];
    __shared__ int s_hist[512];
    __shared__ int s_scan[34];
    __shared__ uint32_t s_prefix;
    __shared__ int s_k;

    const int t = blockIdx.x, b = blockIdx.y;
    if (a.retry && !a.retry[b])
        return;
#ifdef LVT_NMS_STATS
    long long tck = clock64(), tcn;
    TILE_TR(0, nms_ns());
#endif
    {
        __shared__ int s_big[4];
        int ov = a.nms.tile_overflow[b * a.n_tiles + t];
        const int n_big = ov & (kTileSequential - 1);
        const int tpx = a.grid.tile_w(t % a.grid.nx) * a.grid.tile_h(t / a.grid.nx);
        if (!(ov & kTileSequential) && n_big > 0)
        {
            // local maxima whose components outgrew the NMS kernel's buffers: flooded here by the whole CTA
            if (tpx > kTileBitmapBits ||
                !tile_big_candidates(a.nms, t, b, n_big, s_keys, s_rad, reinterpret_cast<int16_t *>(s_perm), s_big))
                ov |= kTileSequential;
            __threadfence_block();
        }
        if (ov & kTileSequential)
        {
            if (threadIdx.x == 0)
                nms_fallback_tile(a.nms, a.parent, t, b);
            __syncthreads();
            TILE_TR(11, 1);
        }
        else
        {
            TILE_TR(11, n_big ? 2 : 0);
        }
    }
    const int tx = t % a.grid.nx, ty = t / a.grid.nx;
    const int x0 = tx * a.grid.cell, y0 = ty * a.grid.cell, tw = a.grid.tile_w(tx), th = a.grid.tile_h(ty);
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
    for (int i = threadIdx.x; i < 257; i += blockDim.x)
        s_hist[i] = 0;

    // ---- raster order: keys[rank] = (ly << 20 | lx << 8 | response), perm[rank] = (response << 24 | rank)
    if (small && tw * th <= kTileBitmapBits)
    {
        // every pixel appears at most once: rank = number of set bits below it in the tile's bitmap
        uint32_t *bm = s_rad, *wpref = s_rad + kTileBitmapBits / 32;
        const int nw = (tw * th + 31) >> 5;
        for (int i = threadIdx.x; i < nw; i += blockDim.x)
            bm[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint32_t r = a.tile_list[base + i] >> 8;
            atomicOr(&bm[r >> 5], 1u << (r & 31));
        }
        __syncthreads();
        {
            const int w0 = 2 * threadIdx.x; // nw <= 2048 = 2 * blockDim.x
            const int c0 = w0 < nw ? __popc(bm[w0]) : 0, c1 = w0 + 1 < nw ? __popc(bm[w0 + 1]) : 0;
            int total;
            const int ex = block_exclusive_scan(c0 + c1, s_scan, &total);
            if (w0 < nw)
                wpref[w0] = (uint32_t)ex;
            if (w0 + 1 < nw)
                wpref[w0 + 1] = (uint32_t)(ex + c0);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint32_t key = a.tile_list[base + i], r = key >> 8, resp = key & 0xFFu;
            const int rank = (int)wpref[r >> 5] + __popc(bm[r >> 5] & ((1u << (r & 31)) - 1u));
            const int ly = (int)r / tw, lx = (int)r - ly * tw;
            keys[rank] = ((uint32_t)ly << 20) | ((uint32_t)lx << 8) | resp;
            perm[rank] = (resp << 24) | (uint32_t)rank;
        }
        __syncthreads();
    }
    else
    {
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
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint32_t key = keys[i];
            const int raster = (int)(key >> 8), ly = raster / tw, lx = raster - ly * tw;
            keys[i] = ((uint32_t)ly << 20) | ((uint32_t)lx << 8) | (key & 0xFFu);
            perm[i] = ((key & 0xFFu) << 24) | (uint32_t)i;
        }
        __syncthreads();
    }
    const uint32_t origin = ((uint32_t)y0 << 20) | ((uint32_t)x0 << 8);
    auto pack_out = [&](uint32_t key) -> uint32_t { return key + origin; };
    TILE_PHASE(2);
    TILE_TR(10, n);

    if (n <= a.max_per_cell)
    {
        // lvt_image_features_handler.cpp:144-151: raster order, tile offset added
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            out[i] = pack_out(keys[i]);
        if (threadIdx.x == 0)
            a.tile_out_count[b * a.n_tiles + t] = n;
        return;
    }
    // ---- ANMS (lvt_image_features_handler.cpp:34-83) ------------------------------------------
    // :38-41, std::sort's exact permutation
    if (small)
        block_introsort(perm, n, s_q0, s_q1, s_cnt, s_posL, s_posR);
    else if (threadIdx.x == 0)
        isort::sort(perm, n); // more survivors than fit in shared memory (pathological): sequential replay
    __syncthreads();
    TILE_PHASE(3);
    // corners in emission order; s_hist[v] = number of corners with response >= v.
    // byte_xy: tile-local coordinates fit a byte each -> squared distances by vabsdiff4 + dp4a on
    // (lx | ly << 8) words, which take over perm's place
    const uint32_t *sorted = small ? s_sorted : nullptr;
    const bool byte_xy = small && tw <= 256 && th <= 256;
    uint32_t *xyb = s_perm;
    auto sorted_key = [&](int p) -> uint32_t { return sorted ? sorted[p] : keys[perm[p] & 0xFFFFFFu]; };
    for (int p = threadIdx.x; p < n; p += blockDim.x)
    {
        const uint32_t k = keys[perm[p] & 0xFFFFFFu];
        if (small)
            s_sorted[p] = k;
        if (byte_xy)
            xyb[p] = ((k >> 8) & 0xFFu) | ((k >> 20) << 8); // this thread's own perm[p] was read above
        atomicAdd(&s_hist[k & 0xFFu], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        // suffix sums over the 256 response bins: lane l owns bins 255-8l .. 248-8l
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
        int acc = incl - mine;
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            acc += s_hist[255 - 8 * lane - u];
            s_hist[255 - 8 * lane - u] = acc;
        }
    }
    __syncthreads();
    TILE_PHASE(4);
    // Tiles with many corners: a uniform grid over the tile (at most 32 x 32 cells of 2^gs >= 8 pixels; cell
    // lists of sort positions, any order -- a minimum does not care) turns the search for the nearest
    // stronger corner into a walk over rings of cells around the corner's own cell, which stops as soon as
    // the best distance found cannot be beaten from outside the rings visited so far.  The memory of the
    // introsort schedule is free by now: start[1025] | cursor[1024] | items u16[kTileSmemCap].
    const bool use_grid = small && n >= a.grid_min;
    int gs = 3; // cells of 8 pixels, or the smallest power of two that keeps the grid within 32 x 32 cells
    while ((tw >> gs) >= 32 || (th >> gs) >= 32)
        gs++;
    const int gx = (tw >> gs) + 1, gy = (th >> gs) + 1, ncell = gx * gy;
    int *g_start = reinterpret_cast<int *>(s_q0), *g_cursor = g_start + 1025;
    uint16_t *g_items = reinterpret_cast<uint16_t *>(g_cursor + 1024);
    if (use_grid)
    {
        auto cell_of = [&](uint32_t k) { return (int)((k >> 20) >> gs) * gx + (int)(((k >> 8) & 0xFFFu) >> gs); };
        for (int c = threadIdx.x; c < 1024; c += blockDim.x)
            g_cursor[c] = 0;
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += blockDim.x)
            atomicAdd(&g_cursor[cell_of(sorted_key(p))], 1);
        __syncthreads();
        {
            // ncell <= 1024 == blockDim.x: one cell per thread
            const int mine = (int)threadIdx.x < ncell ? g_cursor[threadIdx.x] : 0;
            int total;
            const int ex = block_exclusive_scan(mine, s_scan, &total);
            if ((int)threadIdx.x < ncell)
            {
                g_start[threadIdx.x] = ex;
                g_cursor[threadIdx.x] = ex;
            }
            if (threadIdx.x == 0)
                g_start[ncell] = total;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += blockDim.x)
            g_items[atomicAdd(&g_cursor[cell_of(sorted_key(p))], 1)] = (uint16_t)p;
        __syncthreads();
    }
    // :52-64  radius^2 = min squared distance to any corner with response > 1.11f * own.  In the sorted
    // order those corners are a prefix; eight lanes share one corner's prefix.
    for (int w = threadIdx.x; w < 8 * n; w += blockDim.x) // blockDim.x is a multiple of 8: whole groups stay together
    {
        const int p = w >> 3, part = w & 7;
        const uint32_t ki = sorted_key(p);
        // response_j > 1.11f * response_i in fp32  <=>  response_j >= floor(thr) + 1 (integers)
        const float thr = __fmul_rn((float)(ki & 0xFFu), 1.11f);
        const int need = (int)floorf(thr) + 1;
        const int prefix_len = need > 255 ? 0 : s_hist[need];
        const int yi = (int)(ki >> 20), xi = (int)((ki >> 8) & 0xFFFu);
        uint32_t best = 0xFFFFFFFFu; // FLT_MAX
        const unsigned group8 = 0xFFu << ((threadIdx.x & 31) & ~7);
        if (use_grid && prefix_len > kTileGridMinPrefix)
        {
            const int cxi = xi >> gs, cyi = yi >> gs;
            for (int r = 0;; r++)
            {
                // ring r = border of the square of cells [cxi - r, cxi + r] x [cyi - r, cyi + r]; the lanes
                // of the group take its cells in turn
                const int side = 2 * r + 1, ring_cells = r == 0 ? 1 : 8 * r;
                for (int c = part; c < ring_cells; c += 8)
                {
                    int ox = 0, oy = 0;
                    if (r > 0)
                    {
                        if (c < side)
                            ox = c - r, oy = -r;
                        else if (c < 2 * side)
                            ox = c - side - r, oy = r;
                        else if (c < 3 * side - 2)
                            ox = -r, oy = c - 2 * side - r + 1;
                        else
                            ox = r, oy = c - (3 * side - 2) - r + 1;
                    }
                    const int cx = cxi + ox, cy = cyi + oy;
                    if ((unsigned)cx >= (unsigned)gx || (unsigned)cy >= (unsigned)gy)
                        continue;
                    const int t1 = g_start[cy * gx + cx + 1];
                    for (int t = g_start[cy * gx + cx]; t < t1; t++)
                    {
                        const int j = g_items[t];
                        if (j >= prefix_len)
                            continue; // not stronger by the 1.11 margin
                        const uint32_t kj = sorted_key(j);
                        const int dx = xi - (int)((kj >> 8) & 0xFFFu), dy = yi - (int)(kj >> 20);
                        best = min(best, (uint32_t)(dx * dx + dy * dy));
                    }
                }
                uint32_t gbest = min(best, __shfl_xor_sync(group8, best, 1));
                gbest = min(gbest, __shfl_xor_sync(group8, gbest, 2));
                gbest = min(gbest, __shfl_xor_sync(group8, gbest, 4));
                // a corner outside the rings visited so far is more than r cells away along x or y
                const uint32_t reach = (uint32_t)r << gs;
                if (gbest <= reach * reach)
                    break;
                if (cxi - r <= 0 && cyi - r <= 0 && cxi + r >= gx - 1 && cyi + r >= gy - 1)
                    break; // the whole grid has been visited
            }
        }
        else if (byte_xy)
        {
            const uint32_t me = xyb[p];
            auto dist2 = [me](uint32_t other) -> uint32_t {
                const uint32_t d = __vabsdiffu4(me, other);
                return __dp4a(d, d, 0u);
            };
            const int full = prefix_len >> 2;
            for (int g = part; g < full; g += 8)
            {
                const uint4 v = reinterpret_cast<const uint4 *>(xyb)[g];
                best = min(min(best, min(dist2(v.x), dist2(v.y))), min(dist2(v.z), dist2(v.w)));
            }
            for (int q = 4 * full + part; q < prefix_len; q += 8)
                best = min(best, dist2(xyb[q]));
        }
        else
            for (int q = part; q < prefix_len; q += 8)
            {
                const uint32_t kj = sorted_key(q);
                const int dx = xi - (int)((kj >> 8) & 0xFFFu), dy = yi - (int)(kj >> 20);
                best = min(best, (uint32_t)(dx * dx + dy * dy));
            }
        const unsigned group = 0xFFu << ((threadIdx.x & 31) & ~7);
        best = min(best, __shfl_xor_sync(group, best, 1));
        best = min(best, __shfl_xor_sync(group, best, 2));
        best = min(best, __shfl_xor_sync(group, best, 4));
        if (part == 0)
            rad[p] = best;
    }
    __syncthreads();

    TILE_PHASE(5);
    // :66-71  decision = radiiSorted[num_to_keep]  (descending): MSB-first radix select.  With byte
    // coordinates a radius is below 2^17 or FLT_MAX, i.e. an 18-bit key -> two 9-bit passes.
    {
        const int bits = byte_xy ? 9 : 8, npass = byte_xy ? 2 : 4, nbins = 1 << bits, per = nbins / 32;
        auto keyof = [&](uint32_t v) -> uint32_t { return byte_xy ? min(v, 0x3FFFFu) : v; };
        if (threadIdx.x == 0)
        {
            s_prefix = 0;
            s_k = a.max_per_cell;
        }
        for (int pass = 0; pass < npass; pass++)
        {
            const int shift = (npass - 1 - pass) * bits;
            for (int i = threadIdx.x; i < nbins; i += blockDim.x)
                s_hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + bits));
            for (int i = threadIdx.x; i < n; i += blockDim.x)
            {
                const uint32_t kv = keyof(rad[i]);
                if ((kv & mask) == prefix)
                    atomicAdd(&s_hist[(kv >> shift) & (nbins - 1)], 1);
            }
            __syncthreads();
            if (threadIdx.x < 32)
            {
                // walk the bins from the top: lane l owns bins nbins-1-per*l .. nbins-per*(l+1)
                const int lane = threadIdx.x, top = nbins - 1 - per * lane;
                int mine = 0;
                for (int u = 0; u < per; u++)
                    mine += s_hist[top - u];
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
                    int k = k0 - (incl - mine), d = top;
                    for (int u = 0; u < per; u++, d--)
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
        if (byte_xy && threadIdx.x == 0 && s_prefix == 0x3FFFFu)
            s_prefix = 0xFFFFFFFFu;
        __syncthreads();
    }
    const uint32_t decision = s_prefix;
    TILE_PHASE(6);

    // :72-80  keep radius >= decision, in sorted order
    int running = 0;
    for (int p0 = 0; p0 < n; p0 += blockDim.x)
    {
        const int p = p0 + threadIdx.x;
        uint32_t key = 0;
        int keep = 0;
        if (p < n)
        {
            key = sorted_key(p);
            keep = rad[p] >= decision;
        }
        int total;
        const int pos = block_exclusive_scan(keep, s_scan, &total);
        if (keep)
            out[running + pos] = pack_out(key);
        running += total;
    }
    if (threadIdx.x == 0)
        a.tile_out_count[b * a.n_tiles + t] = running;
    TILE_PHASE(7);
    TILE_TR(1, nms_ns());
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
static int tile_grid_min()
{
    static const int v = std::getenv("LVT_B200_TILE_GRID_MIN") ? std::atoi(std::getenv("LVT_B200_TILE_GRID_MIN")) : kTileGridMin;
    return v;
}

int launch_detect(const ImagePool &pool, const DetectWorkspace &ws, const DetectParams &dp, const int *d_slots,
                  int n_images, FeatDev *d_feats, int border, int nonmax, cudaStream_t stream)
{
    if (n_images > ws.batch || dp.grid.count() != ws.n_tiles || ws.n_tiles > 1024)
        return LVTK_ERR_ARG;
    const int nt = ws.n_tiles;
    static DeviceOnce once;
    if (int rc = once.run([](int) {
            LVT_CUDA_TRY(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
            return (int)LVTK_OK;
        }))
        return rc;
    const bool allow_retry = dp.threshold_low < dp.threshold;

    ScoreArgs sa{d_slots, ws.score, dp.grid, dp.pitch, dp.rows, dp.cols, allow_retry ? dp.threshold_low : dp.threshold,
                 ws.tile_count, ws.tile_overflow, ws.tiles_done, nt};
    dim3 sgrid((dp.cols + kScoreTileW - 1) / kScoreTileW, (dp.rows + kScoreTileH - 1) / kScoreTileH, n_images);
    LVT_TIMED(stream, K_SCORE, launch_chained(score_kernel, sgrid, dim3(256), 0, stream, pool.tmap_score, sa));
    LVT_LAUNCH_CHECK(stream, "score_kernel");

    // pass 1 (threshold halved, :161-169) is launched unconditionally and exits at once for images
    // whose first pass found enough corners: the decision is on the device, the host never waits
    for (int pass = 0; pass < (allow_retry ? 2 : 1); pass++)
    {
        const int *retry = pass ? ws.retry : nullptr;
        const int th = pass ? dp.threshold_low : dp.threshold;
        NmsArgs na{ws.score, ws.tile_list, ws.big_list, ws.tile_count, ws.tile_overflow, ws.error, retry,      dp.grid,
                   dp.pitch, dp.rows,      dp.cols,       ws.tile_cap,      nt,       th,         nonmax};
        dim3 ngrid((dp.cols + kNmsTile - 1) / kNmsTile, (dp.rows + kNmsTile - 1) / kNmsTile, n_images);
        LVT_TIMED(stream, K_NMS, launch_chained(nms_tile_kernel, ngrid, dim3(256), 0, stream, na));
        LVT_LAUNCH_CHECK(stream, "nms_tile_kernel");
        GatherArgs ga{ws.tile_out, ws.tile_out_count, ws.retry, ws.error, ws.tile_count, ws.tile_overflow, ws.tiles_done, d_feats,
                      nt, ws.tile_cap, dp.rows, dp.cols, border, pass, allow_retry ? kCornersLowTh : 0};
        TileArgs ta{ws.tile_list, ws.tile_aux, ws.tile_out, ws.tile_count,   ws.tile_out_count, retry,
                    dp.grid,      ws.tile_cap, nt,          dp.max_per_cell, na,                ws.parent, ga, tile_grid_min()};
        LVT_TIMED(stream, K_TILE, launch_chained(tile_kernel, dim3(nt, n_images), dim3(kTileThreads), kTileSmemBytes, stream, ta));
        LVT_LAUNCH_CHECK(stream, "tile_kernel");
    }
    LVT_CUDA_TRY(cudaGetLastError());
    return LVTK_OK;
}

} // namespace lvtb

#ifdef LVT_NMS_STATS
extern "C" __attribute__((visibility("default"))) int lvt_debug_tile_trace(long long *out /* [2048][12] */)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, lvtb::g_tile_trace, sizeof(long long) * 2048 * 12);
}
extern "C" __attribute__((visibility("default"))) int lvt_debug_nms_trace(long long *out /* [4096][12] */)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, lvtb::g_nms_trace, sizeof(long long) * 4096 * 12);
}
#endif

