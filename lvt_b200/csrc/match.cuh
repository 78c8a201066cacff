// Greedy Hamming matchers, bit-exact with the reference's sequential, order-dependent loops.
// Header-only device code shared by the seam kernels and the per-frame tracking kernel (track.cu).
//
// The reference matches map points one after another; a feature taken by an earlier point is
// masked for all later ones (lvt/src/lvt_local_map.cpp:149-170,
// lvt/src/lvt_image_features_struct.cpp:85-101), and stereo row matching does the same over left
// features (lvt/src/lvt_image_features_handler.cpp:305-322).  The exact sequential result is the
// unique fixed point of
//     choice[q] = ratio_test( best2( candidates(q) \ { f : exists p < q, choice[p] == f } ) )
// (induction on q), so it is computed in two phases:
//   1. candidates (parallel over the whole GPU, one warp per query): the features inside the
//      search window, each with its 256-bit Hamming distance (8 x __popc), packed as
//      (distance << 20 | index) keys == knnMatch's (distance, index) order, sorted per query by a
//      warp rank-sort and stored as a fixed-capacity list.  Nothing here depends on the order.
//   2. rounds (one thread per query; one CTA, or the CTAs of a cluster for the map pass: RoundsTeam):
//      every query takes the first two keys of its list whose feature is not owned by an earlier
//      query, applies the ratio test, and publishes its choice (atomicMin(owner[f], q), or a store
//      into every CTA's replica of the choices); rounds repeat until one changes nothing.  Round r
//      fixes at least queries 0..r; 6-7 rounds on the benchmark stream.
// Queries whose window holds more candidates than the list capacity and the radius x2 retry pass
// take the list-free path: one warp re-scans the window each round.
#pragma once
#include "extract.cuh"
#include <cooperative_groups.h>

namespace lvtb
{

constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kFree = 0x7FFFFFFF; // owner of a feature nobody has chosen
constexpr int kTaken = -1;        // owner of a feature that was marked before the pass started
constexpr int kMapCandCap = 64;   // keys kept per map point
constexpr int kRowCandCap = 128;  // keys kept per left feature

struct CandLists
{
    uint32_t *keys; // [n_queries][cap], ascending; nullptr = no lists (every query re-scans)
    int *count;     // [n_queries] true candidate count (> cap: list unusable, re-scan)
    int cap;
};

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
    int *items;   // [2 * cap] work lists of the rounds
};

// ---------------------------------------------------------------------------------------------
// window scans (one warp, all lanes call).  visit(key) is called by the lane that owns it.
// ---------------------------------------------------------------------------------------------
// find_match_index's candidate set (struct.cpp:71-101) minus the marks: features of the hash
// cells [hy-r, hy+r] x [hx-r, hx+r] strictly inside the radius
template <class Visit>
__device__ __forceinline__ void scan_projected_window(const FeatDev &f, const CamParams &cam, float2 p, float r2,
                                                      const uint4 q0, const uint4 q1, int lane, Visit &&visit)
{
    const float cell = (float)kHashCell;
    const int hy = (int)floorf(__fdiv_rn(p.y, cell)), hx = (int)floorf(__fdiv_rn(p.x, cell));
    const int sy = max(hy - cam.cell_search_radius, 0), ey = min(hy + cam.cell_search_radius + 1, cam.cells_y);
    const int sx = max(hx - cam.cell_search_radius, 0), ex = min(hx + cam.cell_search_radius + 1, cam.cells_x);
    if (sx >= ex)
        return;
    const int n_rows = ey - sy;
    if (n_rows <= 8)
    {
        // The cells of one grid row are contiguous in the CSR; the rows of the window are scanned as ONE
        // concatenated candidate list, so the dependent loads (row bounds -> item -> position ->
        // descriptor) are paid once per 32 candidates instead of once per grid row.
        int s = 0, len = 0;
        if (lane < n_rows)
        {
            s = f.cell_start[(sy + lane) * cam.cells_x + sx];
            len = f.cell_start[(sy + lane) * cam.cells_x + ex] - s;
        }
        int incl = len; // inclusive prefix of the row lengths over lanes 0 .. 7
#pragma unroll
        for (int o = 1; o < 8; o <<= 1)
        {
            const int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += nb;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 7);
        int r_s[8], r_end[8];
#pragma unroll
        for (int r = 0; r < 8; r++)
        {
            r_s[r] = __shfl_sync(0xffffffffu, s, r);
            r_end[r] = __shfl_sync(0xffffffffu, incl, r);
        }
        for (int base = 0; base < total; base += 32)
        {
            const int idx = base + lane;
            bool ok = false;
            int j = 0;
            if (idx < total)
            {
                int pos = r_s[0] + idx;
#pragma unroll
                for (int r = 1; r < 8; r++)
                    if (idx >= r_end[r - 1])
                        pos = r_s[r] + (idx - r_end[r - 1]);
                j = f.cell_items[pos];
                const float2 k = f.xy[j];
                const float dx = __fsub_rn(k.x, p.x), dy = __fsub_rn(k.y, p.y);
                ok = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2;
            }
            visit(ok, ok ? (((uint32_t)hamming256(q0, q1, f.desc + 8 * (size_t)j) << 20) | (uint32_t)j) : kNoKey);
        }
        return;
    }
    for (int cy = sy; cy < ey; cy++)
    {
        // cells of one grid row are contiguous in the CSR
        const int s = f.cell_start[cy * cam.cells_x + sx], e = f.cell_start[cy * cam.cells_x + ex];
        for (int base = s; base < e; base += 32)
        {
            const int pos = base + lane;
            bool ok = false;
            int j = 0;
            if (pos < e)
            {
                j = f.cell_items[pos];
                const float2 k = f.xy[j];
                const float dx = __fsub_rn(k.x, p.x), dy = __fsub_rn(k.y, p.y);
                ok = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2;
            }
            visit(ok, ok ? (((uint32_t)hamming256(q0, q1, f.desc + 8 * (size_t)j) << 20) | (uint32_t)j) : kNoKey);
        }
    }
}

// row_match's candidate set (struct.cpp:124-137) minus the marks: right features with
// start_y <= y <= end_y, no x constraint
template <class Visit>
__device__ __forceinline__ void scan_row_band(const FeatDev &fr, const CamParams &cam, float2 p, const uint4 q0,
                                              const uint4 q1, int lane, Visit &&visit)
{
    const int start_y = max((int)p.y - kRowSearchRadius, 0);
    const int end_y = min((int)p.y + kRowSearchRadius, cam.img_h);
    if (start_y > end_y)
        return;
    // bins floor(y) in [start_y, end_y] are contiguous in the row CSR
    const int s = fr.row_start[start_y], e = fr.row_start[end_y + 1];
    for (int base = s; base < e; base += 32)
    {
        const int pos = base + lane;
        bool ok = false;
        int j = 0;
        if (pos < e)
        {
            j = fr.row_items[pos];
            const float yj = fr.xy[j].y;
            ok = yj >= (float)start_y && yj <= (float)end_y;
        }
        visit(ok, ok ? (((uint32_t)hamming256(q0, q1, fr.desc + 8 * (size_t)j) << 20) | (uint32_t)j) : kNoKey);
    }
}

// warp-aggregated append of the lanes' keys into buf[cap] (shared memory); n counts all of them
struct WarpCollector
{
    uint32_t *buf;
    int cap, n, lane;
    __device__ __forceinline__ void operator()(bool ok, uint32_t key)
    {
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok)
        {
            const int slot = n + __popc(m & ((1u << lane) - 1u));
            if (slot < cap)
                buf[slot] = key;
        }
        n += __popc(m);
    }
};

// rank-sort the first min(n, 32*KPL) keys of buf (unique keys) and store them ascending
template <int KPL>
__device__ __forceinline__ void warp_sort_store(const uint32_t *buf, int n, uint32_t *out, int lane)
{
    __syncwarp();
    uint32_t k[KPL];
    int rank[KPL];
#pragma unroll
    for (int a = 0; a < KPL; a++)
    {
        k[a] = (lane + 32 * a < n) ? buf[lane + 32 * a] : kNoKey;
        rank[a] = 0;
    }
    for (int src = 0; src < 32; src++)
    {
#pragma unroll
        for (int b = 0; b < KPL; b++)
        {
            if (32 * b >= n)
                break;
            const uint32_t other = __shfl_sync(0xffffffffu, k[b], src);
#pragma unroll
            for (int a = 0; a < KPL; a++)
                rank[a] += other < k[a];
        }
    }
#pragma unroll
    for (int a = 0; a < KPL; a++)
        if (k[a] != kNoKey)
            out[rank[a]] = k[a];
    __syncwarp();
}

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

// fire-and-forget reductions into the shared memory of CTA `rank` of the cluster (the same offset as
// `local`): red.shared::cluster does not wait for the remote CTA's answer the way an atomic with a return
// value does; the cluster barrier that ends a round makes them visible
__device__ __forceinline__ uint32_t dsmem_addr(const void *local, unsigned rank)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ int dsmem_load(const int *local, unsigned rank)
{
    int v;
    // volatile keeps it in program order with the cluster barrier (also a volatile asm); no memory clobber, so
    // that independent loads of one thread are in flight together
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(dsmem_addr(local, rank)));
    return v;
}
__device__ __forceinline__ void dsmem_red_min(int *local, unsigned rank, int v)
{
    asm volatile("red.relaxed.cluster.shared::cluster.min.s32 [%0], %1;" ::"r"(dsmem_addr(local, rank)), "r"(v));
}
__device__ __forceinline__ void dsmem_red_add(int *local, unsigned rank, int v)
{
    asm volatile("red.relaxed.cluster.shared::cluster.add.s32 [%0], %1;" ::"r"(dsmem_addr(local, rank)), "r"(v) : "memory");
}
__device__ __forceinline__ void dsmem_red_or(int *local, unsigned rank, int v)
{
    asm volatile("red.relaxed.cluster.shared::cluster.or.b32 [%0], %1;" ::"r"(dsmem_addr(local, rank)), "r"(v) : "memory");
}

// The CTAs that share one pass of rounds.  nranks == 1: a single CTA (the seam kernels, track_b).
// nranks > 1: the CTAs of a thread-block cluster split the queries (32 consecutive queries per
// warp-slot, dealt round-robin).  Every CTA reads owners from a full replica of the owner array in its
// own shared memory; the array a round WRITES is cut into nranks slices, each at home in one CTA: a
// choice is published with one fire-and-forget red.min into the home slice of its feature
// (distributed shared memory), a cluster barrier ends the round, and every CTA refreshes its replica
// from the home slices (coalesced remote loads).  Three generations of home slices rotate, so the
// slices of the next round are cleared while the current round runs and one barrier per round
// suffices.  The "changed" flag / match count of a round are collected in rank 0's shared memory.
//
// Replica mode (rep != nullptr, the default of track_a_kernel): instead of home slices every CTA keeps a
// replica of ALL queries' current choices (rep[q], shared memory).  A query that changes its choice
// stores the new one into every CTA's replica (one coalesced distributed-shared-memory store per warp and
// CTA; after the first round only a handful change), one cluster barrier ends the round, and every CTA
// rebuilds its own owner array from its replica with local shared-memory atomics -- no refresh of remote
// slices, no flags relayed through rank 0.  Stores of the next round that arrive early only make a replica
// fresher; a round that changes no choice anywhere has seen exactly the final choices in every CTA, and
// query k is final after round k as before, so the fixed point and the termination test are unchanged.
struct RoundsTeam
{
    int rank = 0, nranks = 1;
    int *team_flags = nullptr; // [3][2] in shared memory (every CTA has the array; rank 0's copy is used)
    int *rep = nullptr;        // [rep_cap] shared memory, same offset in every CTA: choice of every query (replica mode)
    int rep_cap = 0;
};
__device__ __forceinline__ void dsmem_store(int *local, unsigned rank, int v)
{
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(dsmem_addr(local, rank)), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// phase 2: the rounds.  All threads of the CTA (of every CTA of the team) call.
//   n_q queries, active(q) says whether query q takes part; lists may be empty (keys == nullptr).
//   slow(q, cur, b1, b2): warp-cooperative window scan restricted to features with cur[f] >= q.
//   marks: initial marks of the n_f features (nullptr = all clear).
// On return choice[q] = feature or -1 for active queries, owner_a[f] != kFree <=> f is marked.
// ---------------------------------------------------------------------------------------------
template <class Active, class Slow>
__device__ inline int block_rounds(const CandLists &L, int n_q, int n_f, Active active, Slow slow, float ratio_th,
                                   float dist_th, const uint8_t *marks, int *choice, int *items /* [2 * n_q] */,
                                   int *owner_a, int *owner_b, int *s_flag /* [8] */, float *out_d1, float *out_d2,
                                   int *rounds_out, long long *dbg = nullptr, uint32_t *skeys = nullptr,
                                   int skey_cap = 0, const RoundsTeam team = RoundsTeam())
{
    namespace cgr = cooperative_groups;
    const int rank = team.rank, nranks = team.nranks;
    // queries of this CTA: local slot j <-> query ((j / 32) * nranks + rank) * 32 + j % 32
    const int n_loc = 32 * (((n_q + 31) / 32 + nranks - 1) / nranks);
    auto query_of = [&](int j) { return ((j >> 5) * nranks + rank) * 32 + (j & 31); };
#define LVT_RDBG(k)                                                                                                   \
    if (dbg && threadIdx.x == 0)                                                                                      \
    dbg[k] = clock64()
    LVT_RDBG(0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int *cur = owner_a, *nxt = owner_b;
    // team: three generations of this CTA's home slice live where the single-CTA path keeps `nxt`
    // (slice length: a power of two >= 32, so that the home of a feature is a shift)
    int slice_log2 = 5;
    while ((nranks << slice_log2) < n_f)
        slice_log2++;
    const int slice = 1 << slice_log2, home0 = rank * slice;
    auto home = [&](int generation) { return owner_b + (generation % 3) * slice; };
    int *home_now = owner_b; // this round's generation (hoisted out of the per-query path)
    int *fast = items + (size_t)rank * n_loc, *slow_items = items + (size_t)(nranks + rank) * n_loc;
    if (threadIdx.x == 0)
    {
        s_flag[2] = 0, s_flag[3] = 0, s_flag[4] = 0;
        if (nranks > 1)
            for (int k = 0; k < 6; k++)
                team.team_flags[k] = 0;
    }
    const bool replica = nranks > 1 && team.rep != nullptr && n_q <= team.rep_cap;
    int *rep = team.rep;
    for (int j = threadIdx.x; j < n_f; j += blockDim.x)
    {
        const int v = (marks && marks[j]) ? kTaken : kFree;
        cur[j] = v;
        if (nranks > 1 && !replica && j >= home0 && j < home0 + slice)
            home(0)[j - home0] = v; // the first round finds its home slices clean (later ones: cleared a round ahead)
    }
    if (replica)
        for (int q = threadIdx.x; q < n_q; q += blockDim.x)
            rep[q] = -1;
    LVT_RDBG(22);
    if (nranks > 1)
        cgr::this_cluster().sync(); // every replica initialised (and every CTA running) before the first remote store
    else
        __syncthreads();
    LVT_RDBG(23);
    // work lists, built once: queries served from their key list / by a warp re-scan (order is irrelevant)
    // (warp-aggregated appends: one shared-memory atomic per warp and list)
    for (int j0 = 0; j0 < n_loc; j0 += blockDim.x)
    {
        const int q = query_of(j0 + threadIdx.x);
        int kind = 0, cnt = 0; // 1 fast, 2 slow
        if (q < n_q)
        {
            choice[q] = -1;
            if (active(q))
            {
                kind = 2;
                if (L.keys)
                {
                    cnt = L.count[q];
                    kind = cnt <= L.cap ? 1 : 2;
                }
            }
        }
#pragma unroll
        for (int which = 1; which <= 2; which++)
        {
            const uint32_t m = __ballot_sync(0xffffffffu, kind == which);
            if (m == 0)
                continue;
            int base = 0;
            if (lane == __ffs(m) - 1)
                base = atomicAdd(&s_flag[1 + which], __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (kind == which) // fast entries carry the list length: (length << 20 | query)
                (which == 1 ? fast : slow_items)[base + __popc(m & ((1u << lane) - 1u))] = which == 1 ? ((cnt << 20) | q) : q;
        }
    }
    __syncthreads();
    const int n_fast = s_flag[2], n_slow = s_flag[3];
    LVT_RDBG(1);

    int count = 0, rounds = 0;
    // feature c is wanted by query q: the smallest q wins (team: in the home slice of c, this round's generation)
    auto claim = [&](int c, int q) {
        if (nranks == 1)
            atomicMin(&nxt[c], q);
        else
        {
            const int h = c >> slice_log2;
            dsmem_red_min(home_now + (c & (slice - 1)), (unsigned)h, q);
        }
    };
    // `prev` = the query's choice of the previous round (a register copy for the cached queries)
    // replica mode: a changed choice goes into every CTA's replica (consecutive lanes hold consecutive queries:
    // one 128-byte store per warp and CTA)
    auto broadcast_choice = [&](int q, int c) {
        for (int r = 0; r < nranks; r++)
            dsmem_store(rep + q, (unsigned)r, c);
    };
    auto publish_cached = [&](int q, uint32_t b1, uint32_t b2, int &my_count, int &prev) {
        const int c = accept_match(b1, b2, ratio_th, dist_th);
        if (c != prev)
        {
            prev = c;
            choice[q] = c;
            s_flag[0] = 1;
            if (replica)
                broadcast_choice(q, c);
        }
        if (c >= 0)
        {
            if (!replica)
                claim(c, q);
            my_count++;
            if (out_d1)
            {
                out_d1[q] = (float)(b1 >> 20);
                out_d2[q] = b2 != kNoKey ? (float)(b2 >> 20) : -1.0f;
            }
        }
    };
    auto publish = [&](int q, uint32_t b1, uint32_t b2, int &my_count) {
        const int c = accept_match(b1, b2, ratio_th, dist_th);
        if (c != choice[q])
        {
            choice[q] = c;
            s_flag[0] = 1;
            if (replica)
                broadcast_choice(q, c);
        }
        if (c >= 0)
        {
            if (!replica)
                claim(c, q);
            my_count++;
            if (out_d1)
            {
                out_d1[q] = (float)(b1 >> 20);
                out_d2[q] = b2 != kNoKey ? (float)(b2 >> 20) : -1.0f;
            }
        }
    };

    // the first 4 x blockDim queries of the fast list stay with their thread across the rounds: query
    // id, candidate count and last choice in registers, the sorted key list in shared memory (skeys,
    // a private slice per query, so no barrier is needed between the copy and the reads) -- a round
    // then touches only shared memory.  A list that does not fit in skeys is read from global memory.
    constexpr int U = 4;
    int rq[U], rcnt[U], rprev[U], roff[U];
    uint4 rk4[U]; // the first four keys of the list
    int max_cnt = 0;
#pragma unroll
    for (int u = 0; u < U; u++)
    {
        const int it = threadIdx.x + u * blockDim.x;
        const int e = it < n_fast ? fast[it] : -1;
        rq[u] = e < 0 ? -1 : (e & 0xFFFFF);
        rcnt[u] = e < 0 ? 0 : (int)((uint32_t)e >> 20);
        rprev[u] = -1;
        max_cnt = max(max_cnt, rcnt[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++) // in flight while the slices are allocated and copied
        rk4[u] = rcnt[u] > 0 ? *reinterpret_cast<const uint4 *>(L.keys + (size_t)rq[u] * L.cap)
                             : make_uint4(kNoKey, kNoKey, kNoKey, kNoKey);
#pragma unroll
    for (int u = 0; u < U; u++)
    {
        // slices are multiples of 4 keys (16-byte copies); one shared-memory atomic per warp
        const int pad = (rcnt[u] + 3) & ~3;
        int incl = pad;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += nb;
        }
        const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 31 && warp_total > 0 && skeys)
            base = atomicAdd(&s_flag[4], warp_total);
        base = __shfl_sync(0xffffffffu, base, 31);
        const int off = base + incl - pad;
        roff[u] = (skeys && pad > 0 && off + pad <= skey_cap) ? off : -1;
    }
    // asynchronous 16-byte copies (LDGSTS): every chunk of every list is in flight at once
#pragma unroll
    for (int u = 0; u < U; u++)
        if (roff[u] >= 0)
        {
            const uint32_t *src = L.keys + (size_t)rq[u] * L.cap;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(skeys + roff[u]);
            for (int k = 4; k < rcnt[u]; k += 4) // the first chunk lives in registers
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 4u * k), "l"(src + k) : "memory");
        }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    LVT_RDBG(31);
    if (dbg)
    {
        int sum = 0, mx = max_cnt, nglob = 0;
#pragma unroll
        for (int u = 0; u < U; u++)
            sum += rcnt[u], nglob += (rq[u] >= 0 && roff[u] < 0 && rcnt[u] > 0);
        if (threadIdx.x == 0)
            dbg[16] = n_fast, dbg[17] = n_slow, dbg[19] = 0, dbg[20] = 0, dbg[21] = 0;
        __syncthreads();
        atomicAdd((unsigned long long *)&dbg[19], (unsigned long long)sum);
        atomicMax((unsigned long long *)&dbg[20], (unsigned long long)mx);
        atomicAdd((unsigned long long *)&dbg[21], (unsigned long long)nglob);
        __syncthreads();
        if (threadIdx.x == 0)
            dbg[18] = s_flag[4];
    }
    // The first two keys of a sorted list whose feature is not marked / taken by an earlier query.  Keys
    // are unique and ascending, so these are the two SMALLEST open keys: every key goes through two
    // min/max updates, closed ones as kNoKey -- no data-dependent branches inside a chunk.  A chunk is
    // four keys (one 16-byte load, four independent owner look-ups); once two open keys are known no
    // later chunk can change them.
    auto best2 = [](const uint4 first, const uint32_t *keys, int cnt, const int *cur_owner, int q, uint32_t &b1,
                    uint32_t &b2) {
        b1 = kNoKey;
        b2 = kNoKey;
        uint4 c = first;
        for (int k = 0;;)
        {
            const uint32_t key[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                // keys past cnt are stale slots of the list row: never looked up
                const int owner = (k + i < cnt) ? cur_owner[key[i] & 0xFFFFFu] : -2;
                const uint32_t t = owner >= q ? key[i] : kNoKey;
                b2 = min(b2, max(b1, t));
                b1 = min(b1, t);
            }
            k += 4;
            if (b2 != kNoKey || k >= cnt)
                break;
            c = *reinterpret_cast<const uint4 *>(keys + k);
        }
    };

    for (;; rounds++)
    {
        // single CTA: clear this round's target array.  Team: this round's home slices are clean already;
        // clear the next round's generation meanwhile (last read two rounds ago, before the previous
        // barrier; nobody stores into it before the next one).
        if (nranks == 1)
            for (int j = threadIdx.x; j < n_f; j += blockDim.x)
                nxt[j] = cur[j] == kTaken ? kTaken : kFree;
        else if (!replica)
        {
            home_now = home(rounds);
            int *next_gen = home(rounds + 1);
            for (int j = threadIdx.x; j < slice && home0 + j < n_f; j += blockDim.x)
                next_gen[j] = cur[home0 + j] == kTaken ? kTaken : kFree;
        }
        if (!replica || rounds == 0)
        {
            if (threadIdx.x == 0)
                s_flag[0] = 0, s_flag[1] = 0;
            __syncthreads();
        } // replica mode, later rounds: reset behind the rebuild of the owners, whose barrier ends the previous round
        if (rounds == 0)
            LVT_RDBG(2);
        if (rounds == 2)
            LVT_RDBG(24);
        int my_count = 0;
        // one thread per query; the first two keys not owned by an earlier query decide
#pragma unroll
        for (int u = 0; u < U; u++)
        {
            if (rq[u] < 0)
                continue;
            uint32_t b1, b2;
            if (roff[u] >= 0)
                best2(rk4[u], skeys + roff[u], rcnt[u], cur, rq[u], b1, b2);
            else
                best2(rk4[u], L.keys + (size_t)rq[u] * L.cap, rcnt[u], cur, rq[u], b1, b2);
            if (rounds == 2 && u == 0)
            {
                asm volatile("" ::"r"(b1), "r"(b2));
                LVT_RDBG(29);
            }
            publish_cached(rq[u], b1, b2, my_count, rprev[u]);
            if (rounds == 2 && u == 0)
            {
                asm volatile("" ::: "memory");
                LVT_RDBG(30);
            }
        }
        // queries beyond the register-resident ones (maps of more than 4 x blockDim points): entries, first
        // chunks and previous choices of four queries are fetched together, so a round pays the L2
        // latency once per four queries
        for (int base = threadIdx.x + U * blockDim.x; base < n_fast; base += U * blockDim.x)
        {
            int e[U], prev[U];
            uint4 k4[U];
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                const int it = base + u * blockDim.x;
                e[u] = it < n_fast ? fast[it] : -1;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                const bool on = e[u] >= 0;
                const int q = e[u] & 0xFFFFF;
                k4[u] = on && (e[u] >> 20) > 0 ? *reinterpret_cast<const uint4 *>(L.keys + (size_t)q * L.cap)
                                               : make_uint4(kNoKey, kNoKey, kNoKey, kNoKey);
                prev[u] = on ? choice[q] : -1;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (e[u] < 0)
                    continue;
                const int q = e[u] & 0xFFFFF, cnt = (int)((uint32_t)e[u] >> 20);
                uint32_t b1, b2;
                best2(k4[u], L.keys + (size_t)q * L.cap, cnt, cur, q, b1, b2);
                publish_cached(q, b1, b2, my_count, prev[u]);
            }
        }
        for (int it = warp; it < n_slow; it += nwarps)
        {
            const int q = slow_items[it];
            uint32_t b1, b2;
            slow(q, cur, b1, b2);
            if (lane == 0)
                publish(q, b1, b2, my_count);
        }
        if (rounds == 2)
            LVT_RDBG(25);
        if (!replica) // (replica mode counts the matches once, at the end)
        {
            // one shared-memory atomic per warp: a thousand threads adding to one address serialise
            const int warp_count = __reduce_add_sync(0xffffffffu, my_count);
            if (lane == 0 && warp_count)
                atomicAdd(&s_flag[1], warp_count);
        }
        __syncthreads();
        if (rounds == 2)
            LVT_RDBG(26);
        int changed;
        if (nranks == 1)
        {
            changed = s_flag[0];
            count = s_flag[1];
            int *t = cur;
            cur = nxt;
            nxt = t;
            __syncthreads();
        }
        else if (replica)
        {
            // "somebody changed a choice in this round": slot rounds % 3 of EVERY CTA's flags (a plain store of 1);
            // slot (rounds + 2) % 3 is cleared locally after the barrier (its last readers: two rounds ago, its next
            // writers: past the next barrier, which needs this CTA's arrival)
            cgr::cluster_group cl = cgr::this_cluster();
            int *flags = team.team_flags;
            if ((int)threadIdx.x < nranks && s_flag[0])
                dsmem_store(flags + (rounds % 3), threadIdx.x, 1);
            if (rounds == 2)
                LVT_RDBG(27);
            cl.sync(); // every changed choice of the round is in every replica
            if (rounds == 2)
                LVT_RDBG(28);
            changed = flags[rounds % 3];
            if (changed)
            {
                // the owners of the next round, rebuilt from the replica with local atomics
                if (marks)
                {
                    for (int j = threadIdx.x; j < n_f; j += blockDim.x)
                        if (cur[j] != kTaken)
                            cur[j] = kFree;
                }
                else // nothing was marked before the pass: a plain fill, 16 bytes per store (cur is 16-byte aligned)
                    for (int j = 4 * threadIdx.x; j < n_f; j += 4 * blockDim.x)
                        *reinterpret_cast<int4 *>(cur + j) = make_int4(kFree, kFree, kFree, kFree);
                if (threadIdx.x == 0)
                    flags[(rounds + 2) % 3] = 0, s_flag[0] = 0, s_flag[1] = 0;
                __syncthreads();
                for (int q = threadIdx.x; q < n_q; q += blockDim.x)
                {
                    const int c = rep[q];
                    if (c >= 0)
                        atomicMin(&cur[c], q);
                }
                __syncthreads();
            }
        }
        else
        {
            // this CTA's share of the round goes to slot rounds % 3 of rank 0's team_flags.  Rank 0 clears
            // the NEXT round's slot here: its last readers (two rounds ago) are past the previous barrier,
            // its next writers come after the barrier below.
            cgr::cluster_group cl = cgr::this_cluster();
            int *slot = team.team_flags + 2 * (rounds % 3);
            if (threadIdx.x == 0)
            {
                // one reduction per CTA and round: (CTAs that saw a change) << 24 | matches
                const int word = (s_flag[0] ? (1 << 24) : 0) + s_flag[1];
                if (word)
                    dsmem_red_add(slot, 0, word);
                if (rank == 0)
                    team.team_flags[2 * ((rounds + 1) % 3)] = 0;
            }
            if (rounds == 2)
                LVT_RDBG(27);
            cl.sync(); // every claim of the round has landed in its home slice
            if (rounds == 2)
                LVT_RDBG(28);
            // the round's flags and the refresh of the replica (warps read 32 consecutive owners of one home
            // slice at a time) travel together
            int word = 0;
            if (threadIdx.x == 0)
                word = dsmem_load(slot, 0);
            for (int f = threadIdx.x; f < n_f; f += blockDim.x)
                cur[f] = dsmem_load(home_now + (f & (slice - 1)), (unsigned)(f >> slice_log2));
            if (threadIdx.x == 0)
                s_flag[5] = word;
            __syncthreads();
            changed = s_flag[5] >> 24;
            count = s_flag[5] & 0xFFFFFF;
        }
        if (rounds < 12)
            LVT_RDBG(3 + rounds);
        if (!changed)
            break;
    }
    if (replica)
    {
        // matches of the pass: the replica holds every query's final choice
        int mine = 0;
        for (int q = threadIdx.x; q < n_q; q += blockDim.x)
            mine += rep[q] >= 0;
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (threadIdx.x == 0)
            s_flag[1] = 0;
        __syncthreads();
        if (lane == 0 && mine)
            atomicAdd(&s_flag[1], mine);
        __syncthreads();
        count = s_flag[1];
    }
    if (nranks > 1)
        cgr::this_cluster().sync(); // rank 0's flags have been read by everybody; no remote access after this
    if (cur != owner_a)
    {
        for (int j = threadIdx.x; j < n_f; j += blockDim.x)
            owner_a[j] = cur[j];
        __syncthreads();
    }
    if (rounds_out && threadIdx.x == 0 && rank == 0)
        *rounds_out = rounds + 1;
    LVT_RDBG(15);
#undef LVT_RDBG
    return count;
}

// One pass of find_match_index over all visible points, in order, greedy.
__device__ inline int block_match_projected(const CandLists &L, const uint32_t *pdesc, const MatchScratch &ms, int m,
                                            const FeatDev &f, int n, const CamParams &cam, float r2, bool use_marks,
                                            int *owner_a, int *owner_b, int *s_flag, float *out_d1, float *out_d2,
                                            int *rounds_out, long long *dbg = nullptr, uint32_t *skeys = nullptr,
                                            int skey_cap = 0, const RoundsTeam team = RoundsTeam())
{
    const int lane = threadIdx.x & 31;
    auto active = [&](int q) { return ms.vis[q] != 0; };
    auto slow = [&](int q, const int *cur, uint32_t &b1, uint32_t &b2) {
        const uint4 q0 = *reinterpret_cast<const uint4 *>(pdesc + 8 * (size_t)q);
        const uint4 q1 = *reinterpret_cast<const uint4 *>(pdesc + 8 * (size_t)q + 4);
        uint32_t k1 = kNoKey, k2 = kNoKey;
        scan_projected_window(f, cam, ms.proj[q], r2, q0, q1, lane, [&](bool ok, uint32_t key) {
            if (ok && cur[key & 0xFFFFFu] >= q)
                top2_insert(key, k1, k2);
        });
        warp_top2(k1, k2, b1, b2);
    };
    return block_rounds(L, m, n, active, slow, cam.tracking_ratio_th, cam.desc_dist_th, use_marks ? f.matched : nullptr,
                        ms.choice, ms.items, owner_a, owner_b, s_flag, out_d1, out_d2, rounds_out, dbg, skeys, skey_cap, team);
}

// Stereo row matching pass (handler.cpp:302-323 + struct.cpp:122-148).  Queries = unmarked left
// features in index order.  On return the marks of both sides are updated and the pairs are
// written in left-index order; returns their count.
__device__ inline int block_row_match(const CandLists &L, const FeatDev &fl, int nl, const FeatDev &fr, int nr,
                                      const CamParams &cam, int *choice, int *items, int *owner_a, int *owner_b,
                                      int *s_flag, int *s_scan, int *out_query, int *out_train, int *rounds_out,
                                      uint32_t *skeys = nullptr, int skey_cap = 0, long long *dbg = nullptr)
{
    const int lane = threadIdx.x & 31;
    auto active = [&](int q) { return fl.matched[q] == 0; }; // tracked from the map this frame: skipped (handler.cpp:307-310)
    auto slow = [&](int q, const int *cur, uint32_t &b1, uint32_t &b2) {
        const uint4 q0 = *reinterpret_cast<const uint4 *>(fl.desc + 8 * (size_t)q);
        const uint4 q1 = *reinterpret_cast<const uint4 *>(fl.desc + 8 * (size_t)q + 4);
        uint32_t k1 = kNoKey, k2 = kNoKey;
        scan_row_band(fr, cam, fl.xy[q], q0, q1, lane, [&](bool ok, uint32_t key) {
            if (ok && cur[key & 0xFFFFFu] >= q)
                top2_insert(key, k1, k2);
        });
        warp_top2(k1, k2, b1, b2);
    };
    block_rounds(L, nl, nr, active, slow, cam.triangulation_ratio_th, cam.desc_dist_th, fr.matched, choice, items,
                 owner_a, owner_b, s_flag, nullptr, nullptr, rounds_out, dbg, skeys, skey_cap);
    for (int j = threadIdx.x; j < nr; j += blockDim.x)
        if (owner_a[j] != kFree)
            fr.matched[j] = 1;
    // pairs in left-index order + left marks (handler.cpp:313-321); the marks read by active() are
    // only written after every thread has read its own
    int running = 0;
    for (int i0 = 0; i0 < nl; i0 += blockDim.x)
    {
        const int i = i0 + threadIdx.x;
        const int c = (i < nl && fl.matched[i] == 0) ? choice[i] : -1;
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
