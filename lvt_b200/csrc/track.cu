// The per-frame tracking chain and the thin kernels behind the seam ABI (lvtk_match_projected /
// lvtk_row_match / lvtk_solve_pose / lvtk_triangulate), all built from the block-level routines
// of match.cuh, pose.cuh, track.cuh.
//
// lvt_system::perform_tracking (lvt/src/lvt_system.cpp:252-306) per frame, all on the device:
//   mapcand_kernel    whole GPU      project the map with the predicted pose, sorted candidate keys per map point
//   track_a_kernel    8-CTA cluster  greedy rounds over the key lists (+ radius x2 retry); rank 0: marks, map
//                                    bookkeeping, solver inputs, LOST decision
//   pose_kernel       2 x 8 CTAs     cluster 0: motion-only BA (fp64 LM, Cauchy), sums exchanged through DSMEM, then
//                                    the motion model's prediction for the next frame; cluster 1: map culling
//   mapcand_kernel    whole GPU      ("stagedcand") the staged points under the new pose, candidate keys
//   track_b_kernel    1 CTA          staged rounds + promotion, triangulation policy, row-matching rounds,
//                                    triangulation, state + result
// The serial parts are instruction-latency bound, so the 1024-thread kernels keep fp64-heavy, register-hungry work
// out (pose has its own kernel); every data-parallel part (candidate generation, projection, Jacobians) is spread
// over many SMs.  No host round trip: all control flow (first frame, lost, retry, policy, refusal for growth) is
// decided on the device through FrameCtl.
//
// The batched engine (context.cu, run_frames) runs stagedcand + track_b of frame t on a side stream next to frame
// t+1's candidate listing and the early part of its map pass (TrackOverlap, track_a_kernel's parts, signal_kernel):
// track_b only appends to the map, and appended points come last in the greedy order.
#include "track.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace lvtb
{

constexpr int kTrackThreads = 1024;
constexpr int kTrackCluster = 8; // CTAs sharing the rounds of the map pass (track_a_kernel)
constexpr int kCandWarps = 8;

// per-frame hand-over between the kernels of the chain (device memory)
struct FrameCtl
{
    int mode; // 0 frame finished by track_a (was lost already), 1 tracking continues, 2 first frame (track_b seeds
              // the map), 3 lost in this frame (track_b, which may read the right image's features, reports),
              // 4 refused: the point stores could overflow in this frame (TrackState::halt), nothing was touched
    int n_matches;
    int inliers;
    int map_n_clean; // map points left by clean_untracked_points, which runs next to the pose solver (pose_kernel's second cluster)
    PoseD pred, opt;
    double opt_W[12]; // world->camera of opt, prepared by pose_kernel for the staged-point projection
    lvt_frame_info info;
    long long cyc[8];
    int rounds[8]; // as FrameResult::rounds
    // the early part of the map pass (track_a_kernel part 1, batched engine): map points it matched (0: it did not
    // run), their matches, its rounds
    int a1_m, a1_count, a1_rounds;
    long long amark[4]; // as FrameResult::amark
    // what the motion model predicts for the NEXT frame from this frame's pose (pose_kernel, once per frame, as soon as
    // the pose is known): the next frame's candidate listing and track_a take it from here instead of each running
    // the same slerp again -- lvt_motion_model::predict_next_pose on the state track_a left
    PoseD next_pred;
    MotionState next_motion;
    double next_W[12];
};

struct TrackArgs
{
    TrackState *st;
    FrameCtl *ctl;
    FrameResult *result;
    PointStore map, staged;
    const FeatDev *feats; // [2] left (or gray), right
    TrackParams tp;
    TrackScratch sc;
    CandLists row_cand; // candidate keys of this frame's row matching (rowcand_kernel, extraction stream)
    int owner_cap;      // ints per owner array in dynamic shared memory
    int key_cap;        // candidate keys that fit behind the two owner arrays
    long long *dbg;     // clock64() trace of the map pass (LVT_B200_SYNC builds of the launch only)
    int *error;         // the context's sticky error flag (the one the host fetches with every result)
    const FrameCtl *ctl_prev; // the previous frame's hand-over block (the early parts read its mode / pose / culled map size)
    int part;                 // track_a_kernel: 0 the whole map pass, 1 its early part, 2 the rest (see there)
    int wait_seq;             // part 2: TrackState::rest_seq to wait for before anything else
};

struct TrackShared
{
    double W[24]; // world->camera of the left [0..11] and right [12..23] camera
    PoseD pose;
    int flag[8];
    int scan[68]; // block_exclusive_scan (34 ints) / block_exclusive_scan4 (33 uint64)
    int ctrl[4];
};

// phase marks: the GPU-wide nanosecond timer (clock64 is per SM: the kernels of a frame run on different SMs)
__device__ __forceinline__ long long phase_clock()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define LVT_PHASE(k)                                                                                                  \
    if (threadIdx.x == 0)                                                                                             \
    a.ctl->cyc[k] = phase_clock()
// finer marks inside track_b (FrameResult::dbg, read back by lvt_debug_frame_marks)
#define LVT_BMARK(k)                                                                                                  \
    if (threadIdx.x == 0)                                                                                             \
    a.result->dbg[k] = phase_clock()

// ---------------------------------------------------------------------------------------------
// phase 1 of the matchers: candidate keys, one warp per query, the whole GPU
// ---------------------------------------------------------------------------------------------
struct MapCandArgs
{
    TrackState *st; // nullptr: explicit pose / m (seam)
    const FrameCtl *ctl;
    int which;            // 0: map points under the predicted pose, 1: staged points under ctl->opt,
                          // 2: map points for the coming frame, EARLY (see below), 3: the points track_b appended,
                          //    for the coming frame
    int staged_threshold;
    PoseD pose;
    int m;
    const double *xyz;
    const uint32_t *desc;
    const FeatDev *feat;
    CamParams cam;
    float2 *proj;
    uint8_t *vis;
    CandLists L;
    const FrameCtl *ctl_prev; // which == 2
};

// one warp (all lanes call): is_point_visible + the sorted candidate keys of map / staged point q
__device__ __forceinline__ void warp_list_candidates(int q, const double *W /* smem, 12 */, const CamParams &cam, float r2,
                                                     const double *xyz, const uint32_t *desc, const FeatDev &f, float2 *proj,
                                                     uint8_t *vis, const CandLists &L, uint32_t *buf /* smem, kMapCandCap */,
                                                     int lane, int slot = -1 /* row of L; default: q */)
{
    if (slot < 0)
        slot = q;
    double u, v;
    const bool ok = point_visible(W, cam, xyz[3 * q], xyz[3 * q + 1], xyz[3 * q + 2], &u, &v);
    const float2 p = ok ? make_float2((float)u, (float)v) : make_float2(0.f, 0.f);
    if (lane == 0)
    {
        vis[q] = ok;
        proj[q] = p;
    }
    int n = 0;
    if (ok)
    {
        const uint4 q0 = *reinterpret_cast<const uint4 *>(desc + 8 * (size_t)q);
        const uint4 q1 = *reinterpret_cast<const uint4 *>(desc + 8 * (size_t)q + 4);
        WarpCollector col{buf, kMapCandCap, 0, lane};
        scan_projected_window(f, cam, p, r2, q0, q1, lane, col);
        n = col.n;
        if (n <= kMapCandCap)
            warp_sort_store<kMapCandCap / 32>(buf, n, L.keys + (size_t)slot * kMapCandCap, lane);
    }
    if (lane == 0)
        L.count[slot] = n;
}

// which == 2, the batched engine: the map pass of the COMING frame is listed as soon as the previous frame's pose
// is known -- its features are extracted, the pose it is predicted from (ctl_prev->opt) and the culled map
// (ctl_prev->map_n_clean points; pose_kernel's second cluster) are final -- while that frame's map maintenance
// (track_b_kernel, another stream) is still running.  Nothing that track_b writes is read here; what it appends
// to the map is listed by track_a_kernel (TrackState::cand_done).  Only behind a frame that goes on tracking.
__global__ void __launch_bounds__(kCandWarps * 32) mapcand_kernel(MapCandArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    __shared__ double W[12];
    __shared__ uint32_t buf[kCandWarps][kMapCandCap];
    int m = a.m, q_first = 0;
    if (a.st)
    {
        if (a.which == 0)
        {
            const bool on = a.st->state == 2;
            if (blockIdx.x == 0 && threadIdx.x == 0)
                a.st->cand_done = on ? a.st->map_n : 0;
            if (!on)
                return; // first frame / lost: no projection matching
            m = a.st->map_n;
        }
        else if (a.which == 2)
        {
            const bool on = a.ctl_prev->mode == 1;
            m = on ? a.ctl_prev->map_n_clean : 0;
            if (blockIdx.x == 0 && threadIdx.x == 0)
                a.st->cand_done = m;
            if (!on)
                return;
        }
        else if (a.which == 3)
        {
            // behind track_b, on its stream: the points it appended, listed for the coming frame
            const bool on = a.ctl->mode == 1;
            m = on ? a.st->map_n : 0;
            q_first = on ? a.ctl->map_n_clean : 0;
            if (blockIdx.x == 0 && threadIdx.x == 0)
                a.st->tail_done = m;
            if (!on || q_first >= m)
                return;
        }
        else
        {
            if (a.ctl->mode != 1 || a.staged_threshold <= 0)
                return;
            m = a.st->staged_n;
        }
        // prepared by a kernel in front: track_b / reset (prediction for the map pass), pose_kernel (the staged pass; the
        // prediction for the early listing of the coming frame)
        const double *src = a.which == 1 ? a.ctl->opt_W : (a.which == 2 ? a.ctl_prev->next_W : a.st->pred_W);
        if (threadIdx.x < 12)
            W[threadIdx.x] = src[threadIdx.x];
    }
    else if (threadIdx.x == 0)
        world_to_camera(a.pose, W);
    __syncthreads();
    const FeatDev f = *a.feat;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int R = a.cam.tracking_radius;
    const float r2 = (float)(R * R);
    for (int q = q_first + blockIdx.x * kCandWarps + warp; q < m; q += gridDim.x * kCandWarps)
        warp_list_candidates(q, W, a.cam, r2, a.xyz, a.desc, f, a.proj, a.vis, a.L, buf[warp], lane);
}

struct RowCandArgs
{
    const FeatDev *feats; // [2] left, right
    CamParams cam;
    CandLists L;
};

__global__ void __launch_bounds__(kCandWarps * 32) rowcand_kernel(RowCandArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    __shared__ uint32_t buf[kCandWarps][kRowCandCap];
    const FeatDev fl = a.feats[0], fr = a.feats[1];
    const int nl = *fl.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = blockIdx.x * kCandWarps + warp; q < nl; q += gridDim.x * kCandWarps)
    {
        const uint4 q0 = *reinterpret_cast<const uint4 *>(fl.desc + 8 * (size_t)q);
        const uint4 q1 = *reinterpret_cast<const uint4 *>(fl.desc + 8 * (size_t)q + 4);
        WarpCollector col{buf[warp], kRowCandCap, 0, lane};
        scan_row_band(fr, a.cam, fl.xy[q], q0, q1, lane, col);
        if (col.n <= kRowCandCap)
            warp_sort_store<kRowCandCap / 32>(buf[warp], col.n, a.L.keys + (size_t)q * kRowCandCap, lane);
        if (lane == 0)
            a.L.count[q] = col.n;
    }
}

// ---------------------------------------------------------------------------------------------
// track_a: find_matches (lvt/src/lvt_local_map.cpp:136-229) + the LOST decision
// ---------------------------------------------------------------------------------------------
__device__ void write_result(const TrackArgs &a, TrackState &S, const PoseD &pose, int state, int map_n, int staged_n)
{
    FrameCtl &c = *a.ctl;
    S.frame_number += 1;
    c.info.frame_number = S.frame_number;
    c.info.state = state;
    c.info.map_points_after = map_n;
    c.info.staged_after = staged_n;
    a.result->pose = pose;
    a.result->info = c.info;
    c.cyc[7] = phase_clock();
    for (int k = 0; k < 8; k++)
        a.result->cycles[k] = c.cyc[k];
    for (int k = 0; k < 8; k++)
        a.result->rounds[k] = c.rounds[k];
    for (int k = 0; k < 4; k++)
        a.result->amark[k] = c.amark[k];
}

// one thread: world->camera of the pose predicted for the next frame (see TrackState::pred_W)
__device__ void prepare_prediction(TrackState &S)
{
    MotionState mm = S.motion; // a copy: track_a performs the real update
    const PoseD pose = motion_predict(mm, S.last_pose);
    world_to_camera(pose, S.pred_W);
}

// The map pass in the batched engine is cut in two (TrackArgs::part), because the points a frame appends to the map
// come LAST in the greedy order: the choices of the points that were there before cannot depend on them.
//   part 1 (early): the rounds over the points the previous frame's culling left, launched behind the early
//           candidate listing, i.e. while track_b_kernel of the previous frame is still running on another stream.
//           Reads nothing track_b writes (no TrackState::map_n / last_pose; sizes and pose come from the previous
//           frame's FrameCtl), writes only scratch, the marks of the COMING frame's features and ctl->a1_*.
//   part 2: after track_b -- everything else: refusal / lost / first frame, the motion-model update, the candidate
//           lists and rounds of the appended points (the early part's matches are marks for them), the radius x2
//           retry, book-keeping, the solver's inputs.
//   part 0: both in one go (the blocking calls, the first frame of a batch).
__global__ void __launch_bounds__(kTrackThreads, 1) track_a_kernel(TrackArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    // Launched as a cluster of kTrackCluster CTAs: all of them share the greedy rounds of the map pass
    // (RoundsTeam, match.cuh); everything else is rank 0's.
    extern __shared__ int s_owner[];
    __shared__ TrackShared sh;
    __shared__ int s_team_flags[6];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
    int *owner_a = s_owner, *owner_b = s_owner + a.owner_cap;
    uint32_t *skeys = reinterpret_cast<uint32_t *>(s_owner + 2 * a.owner_cap);
    // replica of all queries' choices (match.cuh, RoundsTeam): in the second owner array when the queries fit there,
    // else at the end of the key cache; passes beyond that fall back to the home-slice exchange
    auto make_team = [&](int n_queries, int &key_cap) {
        RoundsTeam team;
        team.rank = rank;
        team.nranks = nranks;
        team.team_flags = s_team_flags;
        const int need = (n_queries + 3) & ~3;
        if (need <= a.owner_cap)
            team.rep = owner_b;
        else if (need <= key_cap)
        {
            key_cap -= need;
            team.rep = reinterpret_cast<int *>(skeys) + key_cap;
        }
        team.rep_cap = team.rep ? need : 0;
        return team;
    };

    TrackState &S = *a.st;
    FrameCtl &ctl = *a.ctl;
    const TrackParams &tp = a.tp;
    const FeatDev fl = a.feats[0];
    const int R = tp.cam.tracking_radius;
    if (a.part == 1)
    {
        // ---- the early part (see above); uniform over the cluster
        const bool on = a.ctl_prev->mode == 1;
        const int M1 = on ? S.cand_done : 0; // = the previous frame's map_n_clean, set by the early candidate listing
        if (threadIdx.x == 0 && rank == 0)
        {
            ctl.amark[0] = phase_clock();
            ctl.amark[1] = ctl.amark[2] = ctl.amark[3] = 0;
        }
        if (M1 <= 0)
        {
            if (threadIdx.x == 0 && rank == 0)
                ctl.a1_m = ctl.a1_count = ctl.a1_rounds = 0;
            return;
        }
        const int nl1 = min(*fl.n, a.owner_cap);
        int key_cap = a.key_cap;
        const RoundsTeam team = make_team(M1, key_cap);
        const int count = block_match_projected(a.sc.map_cand, a.map.desc, a.sc.ms, M1, fl, nl1, tp.cam, (float)(R * R), false, owner_a,
                                                owner_b, sh.flag, nullptr, nullptr, &ctl.a1_rounds, rank == 0 ? a.dbg : nullptr, skeys,
                                                key_cap, team);
        if (rank != 0)
            return;
        for (int j = threadIdx.x; j < nl1; j += blockDim.x)
            fl.matched[j] = owner_a[j] != kFree;
        if (threadIdx.x == 0)
        {
            ctl.a1_m = M1;
            ctl.a1_count = count;
            ctl.amark[1] = phase_clock();
        }
        return;
    }
    if (a.part == 2)
    {
        // the previous frame's map maintenance runs on another stream: wait for its sequence number (signal_kernel).
        // Everything it depends on was submitted before this kernel, so it cannot be queued behind us; the bound
        // only keeps a lost signal from hanging the device (the frame then fails loudly).
        if (threadIdx.x == 0)
        {
            const long long t0 = phase_clock();
            for (;;)
            {
                int seen;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(&S.rest_seq) : "memory");
                if ((int)((unsigned)seen - (unsigned)a.wait_seq) >= 0) // sequence numbers wrap
                    break;
                if (phase_clock() - t0 > 2000000000ll)
                {
                    *a.error = S.error = LVTK_ERR_CUDA;
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    const int state0 = S.state;
    const int nl = state0 == 3 ? 0 : min(*fl.n, a.owner_cap);
    const int map_n = S.map_n, staged_n = S.staged_n;
    const PoseD last_pose = S.last_pose;
    // The reference's point vectors grow without bound (lvt/src/lvt_local_map.cpp:331-353); the stores here
    // have a capacity.  A frame adds at most nl points to either store and promotes at most staged_n, so a frame
    // that could overflow is refused BEFORE anything is touched (this kernel is the first of the chain that
    // writes state); the host grows the stores and runs the frame again.  Uniform over the cluster: rank 0
    // writes S.halt only where this predicate already holds.
    const bool refused = S.halt != 0 || (state0 != 3 && ((long)map_n + staged_n + nl > a.map.cap || (long)staged_n + nl > a.staged.cap));
    __syncthreads();
    if (refused)
    {
        if (threadIdx.x == 0 && rank == 0)
        {
            if (S.halt == 0)
                S.halt = S.frame_number + 1;
            lvt_frame_info z = {};
            ctl.info = z; // state 0: not processed
            ctl.mode = 4;
            a.result->info = z;
            a.result->pose = last_pose;
        }
        return;
    }
    if (threadIdx.x == 0 && rank == 0)
    {
        lvt_frame_info z = {};
        ctl.info = z;
        ctl.info.n_features_left = nl;
        ctl.info.n_features_right = 0; // track_b: the right image may still be in extraction on another stream
        for (int k = 0; k < 8; k++)
            ctl.cyc[k] = 0;
        for (int k = 0; k < 8; k++)
            ctl.rounds[k] = 0;
        ctl.n_matches = 0;
        ctl.inliers = 0;
        ctl.map_n_clean = -1;
        ctl.cyc[0] = phase_clock();
    }
    if (state0 == 3)
    {
        // lost: the last pose, nothing else (lvt/src/lvt_system.cpp:159-166)
        if (threadIdx.x == 0 && rank == 0)
        {
            ctl.mode = 0;
            write_result(a, S, last_pose, 3, map_n, staged_n);
        }
        return;
    }
    if (state0 == 1)
    {
        if (threadIdx.x == 0 && rank == 0)
            ctl.mode = 2; // track_b seeds the map at the identity pose (lvt/src/lvt_system.cpp:185-193)
        return;
    }

    // ---- predict (lvt/src/lvt_system.cpp:196); the candidate listings projected with the same prediction
    if (threadIdx.x == 0 && rank == 0)
    {
        if (a.ctl_prev && a.ctl_prev->mode == 1)
        {
            // pose_kernel of the previous frame ran the motion model already (same state, same pose)
            S.motion = a.ctl_prev->next_motion;
            sh.pose = a.ctl_prev->next_pred;
        }
        else
            sh.pose = motion_predict(S.motion, last_pose);
        ctl.pred = sh.pose;
        ctl.info.map_points_before = map_n;
        ctl.info.staged_before = staged_n;
    }
    __syncthreads();
    const int M = map_n;
    const CandLists no_lists{nullptr, nullptr, 0};
    // Candidate lists exist for the points [0, cand_done): all of them behind mapcand_kernel; behind the early
    // listing, the points the previous frame's culling left -- what track_b appended since (promoted and
    // triangulated points, a few hundred at most; a whole seeded map behind a first frame) is listed here,
    // a warp per point over the cluster, under the same prediction (TrackState::pred_W, written by track_b).
    const int cand_done = min(max(S.cand_done, 0), M);
    // the early part's points keep their choices; their features are marks for the points behind them
    const int M1 = a.part == 2 ? min(max(ctl.a1_m, 0), M) : 0;
    // ... and behind the early part also [cand_done, tail_done): listed behind track_b on its stream
    const int listed = (a.part == 2 && S.tail_done > cand_done) ? min(S.tail_done, M) : cand_done;
    if (listed < M) // uniform over the cluster
    {
        if (threadIdx.x < 12)
            sh.W[threadIdx.x] = S.pred_W[threadIdx.x];
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        uint32_t *buf = reinterpret_cast<uint32_t *>(s_owner) + warp * kMapCandCap; // owner arrays / key cache: not in use yet
        for (int q = listed + rank * nwarps + warp; q < M; q += nranks * nwarps)
            warp_list_candidates(q, sh.W, tp.cam, (float)(R * R), a.map.xyz, a.map.desc, fl, a.sc.ms.proj, a.sc.ms.vis, a.sc.map_cand, buf,
                                 lane);
        __threadfence();
        cluster.sync(); // the lists are read by whichever CTA the rounds deal the point to
    }
    bool owners_current = true; // owner_a describes the marks of the frame's features (else: fl.matched does already)
    int count, key_cap = a.key_cap;
    if (M1 == 0)
    {
        const RoundsTeam team = make_team(M, key_cap);
        count = block_match_projected(a.sc.map_cand, a.map.desc, a.sc.ms, M, fl, nl, tp.cam, (float)(R * R), false, owner_a, owner_b,
                                      sh.flag, nullptr, nullptr, &ctl.rounds[0], rank == 0 ? a.dbg : nullptr, skeys, key_cap, team);
    }
    else if (M1 < M)
    {
        const MatchScratch tms{a.sc.ms.proj + M1, a.sc.ms.vis + M1, a.sc.ms.choice + M1, a.sc.ms.items};
        const CandLists tl{a.sc.map_cand.keys + (size_t)M1 * a.sc.map_cand.cap, a.sc.map_cand.count + M1, a.sc.map_cand.cap};
        // A few hundred appended points at most, and when the early part alone has enough matches the retry cannot
        // fire whatever they do: rank 0 settles them on its own, without the cluster's barriers.
        const bool alone = M - M1 <= kTrackThreads && ctl.a1_count >= kNMatchesTh;
        if (alone && rank != 0)
            return;
        const RoundsTeam team = alone ? RoundsTeam() : make_team(M - M1, key_cap);
        count = ctl.a1_count + block_match_projected(tl, a.map.desc + 8 * (size_t)M1, tms, M - M1, fl, nl, tp.cam, (float)(R * R), true,
                                                     owner_a, owner_b, sh.flag, nullptr, nullptr, &ctl.rounds[0], nullptr, skeys, key_cap,
                                                     team);
        if (threadIdx.x == 0 && rank == 0)
            ctl.rounds[0] += ctl.a1_rounds;
    }
    else
    {
        count = ctl.a1_count; // nothing was appended: the early part was the whole pass
        owners_current = false;
        if (threadIdx.x == 0 && rank == 0)
            ctl.rounds[0] = ctl.a1_rounds;
    }
    int retried = 0;
    if (count < kNMatchesTh)
    {
        retried = 1; // marks reset, radius doubled, cell window unchanged (lvt_local_map.cpp:173-199)
        key_cap = a.key_cap;
        const RoundsTeam team = make_team(M, key_cap);
        count = block_match_projected(no_lists, a.map.desc, a.sc.ms, M, fl, nl, tp.cam, (float)((2 * R) * (2 * R)), false,
                                      owner_a, owner_b, sh.flag, nullptr, nullptr, &ctl.rounds[1], nullptr, nullptr, 0, team);
        owners_current = true;
    }
    if (rank != 0)
        return; // the rounds are over (their last cluster barrier made every choice visible): the rest is rank 0's
    if (owners_current)
        for (int j = threadIdx.x; j < nl; j += blockDim.x)
            fl.matched[j] = owner_a[j] != kFree;
    LVT_PHASE(1);
    // bookkeeping (:201-224) + the solver's inputs, in map order.  Four points per thread, interleaved
    // (coalesced), their loads in flight together, and one packed block scan per 4 x blockDim points: the
    // chain through L2 is what this phase costs, not the work.
    int n_matches = 0;
    for (int base = 0; base < M; base += 4 * (int)blockDim.x)
    {
        int idx[4], c[4];
        unsigned long long packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            idx[k] = base + k * (int)blockDim.x + (int)threadIdx.x;
            c[k] = -3;
            if (idx[k] < M)
                c[k] = a.sc.ms.vis[idx[k]] ? a.sc.ms.choice[idx[k]] : -2;
        }
        int cnt[4], age[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (idx[k] < M)
            {
                cnt[k] = a.map.counter[idx[k]];
                age[k] = a.map.age[idx[k]];
            }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if (idx[k] < M)
            {
                a.map.match_idx[idx[k]] = c[k];
                if (c[k] < 0)
                    a.map.counter[idx[k]] = cnt[k] + 1;
                else
                    a.map.age[idx[k]] = age[k] + 1;
            }
            packed |= (unsigned long long)(c[k] >= 0) << (16 * k);
        }
        unsigned long long total;
        const unsigned long long excl = block_exclusive_scan4(packed, reinterpret_cast<unsigned long long *>(sh.scan), &total);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (c[k] >= 0)
            {
                const int d = n_matches + scan4_position(excl, total, k), i = idx[k];
                a.sc.sol_xyz[3 * d] = a.map.xyz[3 * i];
                a.sc.sol_xyz[3 * d + 1] = a.map.xyz[3 * i + 1];
                a.sc.sol_xyz[3 * d + 2] = a.map.xyz[3 * i + 2];
                a.sc.sol_uv[d] = fl.xy[c[k]];
            }
        n_matches += scan4_sum(total);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        ctl.info.tracked = n_matches;
        ctl.info.retried_matching = retried;
        ctl.n_matches = n_matches;
        S.cand_done = S.tail_done = 0; // consumed
        ctl.cyc[2] = phase_clock();
        if (n_matches < tp.min_matches)
        {
            // lost: return the last pose (lvt/src/lvt_system.cpp:267-272,199-204)
            S.state = 3;
            ctl.mode = 3;
        }
        else
        {
            S.last_matches[0] = S.last_matches[1];
            S.last_matches[1] = S.last_matches[2];
            S.last_matches[2] = n_matches;
            ctl.mode = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pose: lvt_pnp_solver::compute_pose on an 8-CTA cluster
// ---------------------------------------------------------------------------------------------
struct PoseArgs
{
    FrameCtl *ctl; // nullptr: explicit inputs (seam)
    const double *xyz;
    const float2 *uv;
    int m;
    PoseD init;
    CamParams cam;
    uint8_t *level, *inlier;
    double *e2;
    PoseD *out;
    int *n_inliers;
    long long *dbg;
    const TrackState *st; // with ctl: where a frame that skips the solver takes its pose from
    EarlyResult *early;   // pose + state for a blocking caller, in mapped host memory (nullptr: not wanted)
    int early_seq;
    // the launch's second cluster (blockIdx.x >= kPoseCluster, frames only): clean_untracked_points
    // (lvt/src/lvt_local_map.cpp:393-413) needs the matches of the frame, not its pose, so the map is culled and
    // compacted on another SM WHILE the solver runs instead of behind it
    PointStore map;
    const FeatDev *feats;
    int untracked_threshold;
};

// dynamic shared memory of pose_kernel: the solver's exchange buffers / the culling CTA's staging (4 x 256 points)
constexpr int kCleanSmemBytes = 4 * kPoseThreads * (4 + 32);
constexpr int kPoseSmemBytes = (int)sizeof(PoseShared) > kCleanSmemBytes ? (int)sizeof(PoseShared) : kCleanSmemBytes;

__global__ void __cluster_dims__(kPoseCluster, 1, 1) __launch_bounds__(kPoseThreads, 1) pose_kernel(PoseArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    extern __shared__ __align__(16) unsigned char pose_smem[];
    PoseShared &s = *reinterpret_cast<PoseShared *>(pose_smem);
    cg::cluster_group cluster = cg::this_cluster();
    if (blockIdx.x >= kPoseCluster)
    {
        // ---- the culling cluster: one CTA works, and only on frames that go on tracking
        if (cluster.block_rank() != 0 || !a.ctl || a.ctl->mode != 1)
            return;
        __shared__ unsigned long long clean_scan[34];
        const long long t0 = clock64();
        const FeatDev fl = a.feats[0];
        const int *cnt = a.map.counter, *midx = a.map.match_idx;
        uint8_t *matched = fl.matched;
        const int th = a.untracked_threshold;
        // a dropped point gives its feature back (:398-404): done where the keep flag is taken
        const int left = block_compact_points(
            a.map, a.st->map_n,
            [cnt, midx, matched, th](int i) {
                if (cnt[i] < th)
                    return true;
                const int mi = midx[i];
                if (mi >= 0)
                    matched[mi] = 0;
                return false;
            },
            reinterpret_cast<int *>(clean_scan), reinterpret_cast<int *>(pose_smem));
        if (threadIdx.x == 0)
        {
            a.ctl->map_n_clean = left;
            a.ctl->rounds[5] = (int)(clock64() - t0);
        }
        return;
    }
    int m = a.m;
    PoseD init = a.init;
    PoseD *out = a.out;
    int *n_inl = a.n_inliers;
    if (a.ctl)
    {
        const int mode = a.ctl->mode;
        if (mode != 1)
        {
            // first frame: identity, tracking; lost: the last pose (lvt/src/lvt_system.cpp:159-166,185-193,199-204)
            if (a.early && cluster.block_rank() == 0 && threadIdx.x == 0)
            {
                PoseD id;
                id.q = Quat{1, 0, 0, 0};
                id.t[0] = id.t[1] = id.t[2] = 0;
                a.early->pose = mode == 2 ? id : a.st->last_pose;
                a.early->state = mode == 2 ? 2 : (mode == 4 ? 0 /* refused: grow the stores, run again */ : 3);
                __threadfence_system();
                *reinterpret_cast<volatile int *>(&a.early->seq) = a.early_seq;
            }
            return; // uniform over the cluster
        }
        m = a.ctl->n_matches;
        init = a.ctl->pred;
        out = &a.ctl->opt;
        n_inl = &a.ctl->inliers;
        if (cluster.block_rank() == 0 && threadIdx.x == 0)
            a.ctl->cyc[3] = phase_clock();
    }
    cluster_solve_pose(cluster, s, a.xyz, a.uv, m, init, a.cam, a.level, a.e2, a.inlier, out, n_inl, a.dbg,
                       a.ctl ? &a.ctl->rounds[4] : nullptr);
    if (a.ctl && cluster.block_rank() == 0 && threadIdx.x == 0)
    {
        a.ctl->cyc[4] = phase_clock();
        world_to_camera(a.ctl->opt, a.ctl->opt_W);
        if (a.early)
        {
            a.early->pose = a.ctl->opt;
            a.early->state = 2;
            __threadfence_system();
            *reinterpret_cast<volatile int *>(&a.early->seq) = a.early_seq;
        }
        // the prediction for the next frame (behind the blocking caller's pose): see FrameCtl::next_pred
        MotionState mm = a.st->motion;
        a.ctl->next_pred = motion_predict(mm, a.ctl->opt);
        a.ctl->next_motion = mm;
        world_to_camera(a.ctl->next_pred, a.ctl->next_W);
    }
}

// ---------------------------------------------------------------------------------------------
// track_b: clean_untracked_points, update_staged_map_points, need_new_triangulation,
// update_with_new_triangulation (lvt/src/lvt_local_map.cpp:331-413, lvt/src/lvt_system.cpp:287-334)
// ---------------------------------------------------------------------------------------------
// lvt_local_map::update_with_new_triangulation for the pose in sh.pose.  Returns (uniformly) the
// number of new points; updates map_n / staged_n (locals).
__device__ int block_new_triangulation(const TrackArgs &a, TrackShared &sh, const FeatDev &fl, int nl, const FeatDev &fr,
                                       int nr, bool dont_stage, int *owner_a, int *owner_b, int &map_n, int &staged_n)
{
    uint32_t *skeys = reinterpret_cast<uint32_t *>(owner_a + 2 * a.owner_cap);
    const TrackParams &tp = a.tp;
    const bool to_map = dont_stage || tp.staged_threshold == 0 || map_n < kNMapPoints;
    const PointStore &dst = to_map ? a.map : a.staged;
    const int base = to_map ? map_n : staged_n;
    int added = 0;

    if (tp.sensor == 1)
    {
        const int np = block_row_match(a.row_cand, fl, nl, fr, nr, tp.cam, a.sc.row_choice, a.sc.bs.items, owner_a,
                                       owner_b, sh.flag, sh.scan, a.sc.pair_query, a.sc.pair_train, &a.ctl->rounds[3], skeys, a.key_cap,
                                       a.dbg ? a.dbg + 60000 : nullptr); // (debug builds of the launch: far behind the triangulated points)
        if (threadIdx.x == 0)
            a.ctl->rounds[7] = (int)(phase_clock() - a.ctl->cyc[5]); // ns into track_b: row matching done
        LVT_BMARK(4);
        if (np == 0)
            return 0;
        if (threadIdx.x == 0)
        {
            world_to_camera(sh.pose, sh.W);
            const PoseD pr = right_pose(sh.pose, (double)tp.cam.baseline);
            world_to_camera(pr, sh.W + 12);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < np; k += blockDim.x)
        {
            double xyz[3];
            const bool ok = triangulate_pair(sh.W, sh.W + 12, tp.cam, fl.xy[a.sc.pair_query[k]],
                                             fr.xy[a.sc.pair_train[k]], xyz);
            a.sc.tri_xyz[3 * k] = xyz[0];
            a.sc.tri_xyz[3 * k + 1] = xyz[1];
            a.sc.tri_xyz[3 * k + 2] = xyz[2];
            a.sc.tri_ok[k] = ok;
        }
        __syncthreads();
        LVT_BMARK(5);
        for (int k0 = 0; k0 < np; k0 += blockDim.x)
        {
            const int k = k0 + threadIdx.x;
            const bool ok = k < np && a.sc.tri_ok[k];
            int total;
            const int pos = block_exclusive_scan(ok, sh.scan, &total);
            if (ok)
            {
                const int d = base + added + pos;
                if (d < dst.cap)
                    copy_point(dst, d, a.sc.tri_xyz + 3 * k, fl.desc + 8 * (size_t)a.sc.pair_query[k], 0, 0, 0);
            }
            added += total;
        }
    }
    else
    {
        // triangulate_rgbd (lvt/src/lvt_local_map.cpp:231-256): fp32 back-projection of EVERY feature
        if (threadIdx.x == 0)
        {
            double Rm[9];
            quat_to_mat(sh.pose.q, Rm);
            for (int i = 0; i < 9; i++)
                sh.W[i] = Rm[i];
        }
        __syncthreads();
        const float inv_fx = __fdiv_rn(1.0f, tp.cam.fx), inv_fy = __fdiv_rn(1.0f, tp.cam.fy);
        for (int i = threadIdx.x; i < nl; i += blockDim.x)
        {
            const int d = base + i;
            if (d >= dst.cap)
                continue;
            const float2 p = fl.xy[i];
            const float z = fl.depth[i];
            const float x = __fmul_rn(__fmul_rn(__fsub_rn(p.x, tp.cam.cx), z), inv_fx);
            const float y = __fmul_rn(__fmul_rn(__fsub_rn(p.y, tp.cam.cy), z), inv_fy);
            double w[3];
            for (int r = 0; r < 3; r++)
                w[r] = (sh.W[3 * r] * (double)x + sh.W[3 * r + 1] * (double)y + sh.W[3 * r + 2] * (double)z) + sh.pose.t[r];
            copy_point(dst, d, w, fl.desc + 8 * (size_t)i, 0, 0, 0);
        }
        added = nl;
    }
    if (base + added > dst.cap)
    {
        if (threadIdx.x == 0) // unreachable since track_a refuses frames that could overflow; reported if it ever is
            *a.error = a.st->error = LVTK_ERR_CAPACITY;
        added = dst.cap - base;
    }
    if (to_map)
        map_n += added;
    else
        staged_n += added;
    __syncthreads();
    return added;
}

__global__ void __launch_bounds__(kTrackThreads, 1) track_b_kernel(TrackArgs a)
{
    LVT_GRID_DEP_SYNC(); // nothing of the previous kernel's output is touched before this
    extern __shared__ int s_owner[];
    __shared__ TrackShared sh;
    int *owner_a = s_owner, *owner_b = s_owner + a.owner_cap;

    TrackState &S = *a.st;
    FrameCtl &ctl = *a.ctl;
    const TrackParams &tp = a.tp;
    const int mode = ctl.mode;
    if (mode == 0 || mode == 4)
        return; // track_a finished the frame (lost) or refused it (TrackState::halt)
    const FeatDev fl = a.feats[0], fr = a.feats[1];
    const int nl = min(*fl.n, a.owner_cap);
    const int nr = tp.sensor == 1 ? min(*fr.n, a.owner_cap) : 0;
    int map_n = S.map_n, staged_n = S.staged_n;
    __syncthreads();
    LVT_PHASE(5);
    if (threadIdx.x == 0)
    {
        for (int k = 0; k < 8; k++)
            a.result->dbg[k] = 0;
        a.result->dbg[0] = phase_clock();
    }
    if (threadIdx.x == 0)
        ctl.info.n_features_right = nr;
    if (mode == 3)
    {
        if (threadIdx.x == 0)
            write_result(a, S, S.last_pose, 3, map_n, staged_n);
        return;
    }

    if (mode == 2)
    {
        // first frame: identity pose, seed the map (lvt/src/lvt_system.cpp:185-193)
        if (threadIdx.x == 0)
        {
            sh.pose.q = Quat{1, 0, 0, 0};
            sh.pose.t[0] = sh.pose.t[1] = sh.pose.t[2] = 0;
        }
        __syncthreads();
        const int added = block_new_triangulation(a, sh, fl, nl, fr, nr, true, owner_a, owner_b, map_n, staged_n);
        if (threadIdx.x == 0)
        {
            S.state = 2;
            S.map_n = map_n;
            S.staged_n = staged_n;
            S.last_matches[0] = map_n;
            prepare_prediction(S);
            ctl.info.triangulated = 1;
            ctl.info.new_points = added;
            write_result(a, S, sh.pose, 2, map_n, staged_n);
        }
        return;
    }

    if (threadIdx.x == 0)
    {
        sh.pose = ctl.opt;
        ctl.info.inliers = ctl.inliers;
    }
    __syncthreads();
    // clean_untracked_points (lvt/src/lvt_local_map.cpp:393-413) has run next to the pose solver (pose_kernel's
    // second cluster): the map is compacted already
    map_n = ctl.map_n_clean;
    LVT_PHASE(6);

    // ---- update_staged_map_points (lvt/src/lvt_local_map.cpp:355-391); stagedcand projected the
    //      staged points with the optimised pose and listed their candidates
    if (tp.staged_threshold > 0 && staged_n > 0)
    {
        const int Sn = staged_n;
        const int R = tp.cam.tracking_radius;
        block_match_projected(a.sc.staged_cand, a.staged.desc, a.sc.bs, Sn, fl, nl, tp.cam, (float)(R * R), true, owner_a,
                              owner_b, sh.flag, nullptr, nullptr, &ctl.rounds[2], nullptr,
                              reinterpret_cast<uint32_t *>(s_owner + 2 * a.owner_cap), a.key_cap);
        LVT_BMARK(1);
        for (int j = threadIdx.x; j < nl; j += blockDim.x)
            if (owner_a[j] != kFree)
                fl.matched[j] = 1;
        // hit -> counter++; promote when counter == staged_threshold or the map is still below 250
        // points.  Sequentially the map grows with every promotion, which is equivalent to:
        // map_n + (#hits before this one) < 250.
        const int map0 = map_n;
        int hits = 0, promoted = 0;
        uint8_t *flag = a.sc.level; // 0 erase, 1 keep staged, 2 promoted
        for (int i0 = 0; i0 < Sn; i0 += blockDim.x)
        {
            const int i = i0 + threadIdx.x;
            const bool hit = i < Sn && a.sc.bs.vis[i] && a.sc.bs.choice[i] >= 0;
            int tot_h;
            const int h = block_exclusive_scan(hit, sh.scan, &tot_h);
            bool prom = false;
            int cnt_new = 0;
            if (hit)
            {
                cnt_new = a.staged.counter[i] + 1;
                a.staged.counter[i] = cnt_new;
                prom = (cnt_new == tp.staged_threshold) || (map0 + hits + h < kNMapPoints);
            }
            int tot_p;
            const int pp = block_exclusive_scan(prom, sh.scan, &tot_p);
            if (i < Sn)
                flag[i] = prom ? 2 : (hit ? 1 : 0);
            if (prom)
            {
                const int d = map0 + promoted + pp;
                if (d < a.map.cap)
                    copy_point(a.map, d, a.staged.xyz + 3 * i, a.staged.desc + 8 * (size_t)i, cnt_new, a.staged.age[i],
                               a.staged.match_idx[i]);
            }
            hits += tot_h;
            promoted += tot_p;
        }
        __syncthreads();
        map_n = map0 + promoted;
        if (map_n > a.map.cap)
        {
            if (threadIdx.x == 0)
                *a.error = S.error = LVTK_ERR_CAPACITY;
            map_n = a.map.cap;
        }
        LVT_BMARK(2);
        staged_n = block_compact_points(a.staged, Sn, [flag](int i) { return flag[i] == 1; }, sh.scan, s_owner);
    }
    LVT_BMARK(3);

    // ---- need_new_triangulation (lvt/src/lvt_system.cpp:308-334)
    if (threadIdx.x == 0)
    {
        ctl.rounds[6] = (int)(phase_clock() - ctl.cyc[5]); // ns into track_b: staged points done
        int need;
        if (tp.triangulation_policy == 2)
            need = 1;
        else if (tp.triangulation_policy == 3)
            need = map_n < 1000;
        else
        {
            need = 1;
            const float ratio = 0.99f;
            for (int i = 2; i > 0; --i)
                if ((float)S.last_matches[i] > __fmul_rn(ratio, (float)S.last_matches[i - 1]))
                    need = 0;
        }
        sh.ctrl[0] = need;
    }
    __syncthreads();
    if (sh.ctrl[0])
    {
        const int added = block_new_triangulation(a, sh, fl, nl, fr, nr, false, owner_a, owner_b, map_n, staged_n);
        if (threadIdx.x == 0)
        {
            ctl.info.triangulated = 1;
            ctl.info.new_points = added;
        }
    }
    __syncthreads();
    LVT_BMARK(6);
    if (threadIdx.x == 0)
    {
        S.map_n = map_n;
        S.staged_n = staged_n;
        S.last_pose = sh.pose;
        for (int k = 0; k < 12; k++) // the prediction for the next frame: pose_kernel has it
            S.pred_W[k] = ctl.next_W[k];
        a.result->dbg[7] = phase_clock();
        write_result(a, S, sh.pose, 2, map_n, staged_n);
    }
}

__global__ void reset_state_kernel(TrackState *st)
{
    // lvt_system::reset (lvt/src/lvt_system.cpp:44-68)
    TrackState &S = *st;
    S.state = 1;
    S.frame_number = 0;
    S.map_n = S.staged_n = 0;
    S.last_matches[0] = S.last_matches[1] = S.last_matches[2] = 0x7FFFFFFF;
    S.error = 0;
    S.halt = 0;
    S.cand_done = S.tail_done = 0;
    S.last_pose.q = Quat{1, 0, 0, 0};
    S.last_pose.t[0] = S.last_pose.t[1] = S.last_pose.t[2] = 0;
    motion_reset(S.motion);
}

__global__ void clear_halt_kernel(TrackState *st) { st->halt = 0; }

// last kernel of a frame on the side stream (batched engine): see TrackState::rest_seq
__global__ void signal_kernel(TrackState *st, int seq)
{
    LVT_GRID_DEP_SYNC(); // the kernels in front (track_b, the listing of the appended points) are complete and visible
    if (threadIdx.x == 0)
    {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(&st->rest_seq), "r"(seq) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// seam kernels
// ---------------------------------------------------------------------------------------------
struct MatchSeamArgs
{
    const uint32_t *pdesc;
    int m;
    const FeatDev *feat;
    CamParams cam;
    int retry_below;
    MatchScratch ms;
    CandLists lists;
    int *match_idx;
    float *d1, *d2;
    int *count_retried; // [2]
    int owner_cap, key_cap;
};

__global__ void __launch_bounds__(kTrackThreads, 1) match_seam_kernel(MatchSeamArgs a)
{
    extern __shared__ int s_owner[];
    __shared__ int s_flag[8];
    int *owner_a = s_owner, *owner_b = s_owner + a.owner_cap;
    const FeatDev f = *a.feat;
    const int n = min(*f.n, a.owner_cap);
    const int R = a.cam.tracking_radius;
    const CandLists no_lists{nullptr, nullptr, 0};
    int count = block_match_projected(a.lists, a.pdesc, a.ms, a.m, f, n, a.cam, (float)(R * R), true, owner_a, owner_b,
                                      s_flag, a.d1, a.d2, nullptr, nullptr,
                                      reinterpret_cast<uint32_t *>(s_owner + 2 * a.owner_cap), a.key_cap);
    int retried = 0;
    if (count < a.retry_below)
    {
        retried = 1;
        count = block_match_projected(no_lists, a.pdesc, a.ms, a.m, f, n, a.cam, (float)((2 * R) * (2 * R)), false,
                                      owner_a, owner_b, s_flag, a.d1, a.d2, nullptr);
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        f.matched[j] = owner_a[j] != kFree;
    for (int i = threadIdx.x; i < a.m; i += blockDim.x)
    {
        const int c = a.ms.vis[i] ? a.ms.choice[i] : -2;
        a.match_idx[i] = c;
        if (c < 0 && a.d1)
        {
            a.d1[i] = 0.f;
            a.d2[i] = 0.f;
        }
    }
    if (threadIdx.x == 0)
    {
        a.count_retried[0] = count;
        a.count_retried[1] = retried;
    }
}

struct RowSeamArgs
{
    const FeatDev *feats; // [2]
    CamParams cam;
    CandLists lists;
    int *choice, *items, *query, *train, *count;
    int owner_cap, key_cap;
};

__global__ void __launch_bounds__(kTrackThreads, 1) row_seam_kernel(RowSeamArgs a)
{
    extern __shared__ int s_owner[];
    __shared__ int s_flag[8];
    __shared__ int s_scan[34];
    const FeatDev fl = a.feats[0], fr = a.feats[1];
    const int nl = min(*fl.n, a.owner_cap), nr = min(*fr.n, a.owner_cap);
    const int np = block_row_match(a.lists, fl, nl, fr, nr, a.cam, a.choice, a.items, s_owner, s_owner + a.owner_cap,
                                   s_flag, s_scan, a.query, a.train, nullptr,
                                   reinterpret_cast<uint32_t *>(s_owner + 2 * a.owner_cap), a.key_cap);
    if (threadIdx.x == 0)
        *a.count = np;
}

struct TriSeamArgs
{
    PoseD pose;
    CamParams cam;
    const float2 *uvl, *uvr;
    int n;
    double *xyz;
    uint8_t *ok;
};

__global__ void tri_seam_kernel(TriSeamArgs a)
{
    __shared__ double W[24];
    if (threadIdx.x == 0)
    {
        world_to_camera(a.pose, W);
        const PoseD pr = right_pose(a.pose, (double)a.cam.baseline);
        world_to_camera(pr, W + 12);
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n)
        return;
    double xyz[3];
    a.ok[k] = triangulate_pair(W, W + 12, a.cam, a.uvl[k], a.uvr[k], xyz);
    a.xyz[3 * k] = xyz[0];
    a.xyz[3 * k + 1] = xyz[1];
    a.xyz[3 * k + 2] = xyz[2];
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
// dynamic shared memory of the single-CTA kernels: two owner arrays + as many candidate keys as fit
// next to them (block_rounds keeps the key lists of its queries there).  Function attributes are per
// device: set once per device to the most a launch can ask for; what a context actually uses is in its
// TrackLaunchCfg (no process-wide mutable state: handles on any device / thread are independent).
constexpr int kTrackStaticSmem = 2048; // static shared memory of the tracking kernels, rounded up

static size_t track_smem_bytes(const TrackLaunchCfg &cfg) { return (2 * (size_t)cfg.owner_cap + (size_t)cfg.key_cap) * sizeof(int); }

static int ensure_smem()
{
    static DeviceOnce once;
    return once.run([](int dev) {
        int optin = 0;
        LVT_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        const int bytes = optin - kTrackStaticSmem;
        LVT_CUDA_TRY(cudaFuncSetAttribute(track_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        LVT_CUDA_TRY(cudaFuncSetAttribute(track_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        LVT_CUDA_TRY(cudaFuncSetAttribute(match_seam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        LVT_CUDA_TRY(cudaFuncSetAttribute(row_seam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        LVT_CUDA_TRY(cudaFuncSetAttribute(pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoseSmemBytes));
        return (int)LVTK_OK;
    });
}

int track_configure(int owner_cap, TrackLaunchCfg *cfg)
{
    if (int rc = ensure_smem())
        return rc;
    int dev = 0, optin = 0;
    LVT_CUDA_TRY(cudaGetDevice(&dev));
    LVT_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int avail = optin - kTrackStaticSmem - 2 * owner_cap * (int)sizeof(int);
    if (avail < 0)
    {
        set_last_error(__FILE__, __LINE__, "the owner arrays do not fit in shared memory");
        return LVTK_ERR_ARG;
    }
    cfg->owner_cap = owner_cap;
    cfg->key_cap = (avail / (int)sizeof(int)) & ~3;
    cfg->cluster = kTrackCluster;
    if (const char *e = std::getenv("LVT_B200_TRACK_CLUSTER")) // 1 or 8 (tuning aid; the home slices of the
        cfg->cluster = std::atoi(e) >= 8 ? 8 : 1;              // team are sized for 8 CTAs)
    return LVTK_OK;
}

size_t frame_ctl_bytes() { return sizeof(FrameCtl); }

int launch_rowcand(const FeatDev *d_feats, const CamParams &cam, const CandLists &L, cudaStream_t stream)
{
    RowCandArgs a{d_feats, cam, L};
    LVT_TIMED(stream, K_ROWCAND, launch_chained(rowcand_kernel, dim3(148 * 2), dim3(kCandWarps * 32), 0, stream, a));
    LVT_LAUNCH_CHECK(stream, "rowcand_kernel");
    return LVTK_OK;
}

int launch_track_frame(TrackState *st, void *ctl_v, FrameResult *result, const PointStore &map, const PointStore &staged,
                       const FeatDev *d_feats, const TrackParams &tp, const TrackScratch &sc, const CandLists &row_cand,
                       const TrackLaunchCfg &cfg, int *d_error, cudaStream_t stream, cudaEvent_t right_ready, int parts,
                       EarlyResult *early, int early_seq, const TrackOverlap *ov, const void *ctl_prev_v)
{
    // parts: 1 = up to the pose (mapcand, track_a, pose), 2 = the rest (stagedcand, track_b), 3 = the whole frame
    // ov (the batched engine, parts == 3): the rest runs on ov->side behind the pose; with ov->early the map pass is
    // cut in two around the previous frame's rest (track_a_kernel) and starts from that frame's FrameCtl
    FrameCtl *ctl = static_cast<FrameCtl *>(ctl_v);
    const FrameCtl *ctl_prev = static_cast<const FrameCtl *>(ctl_prev_v); // the frame launched before this one
    const bool early_pass = ov && ov->early;
    const size_t smem = track_smem_bytes(cfg);
    TrackArgs a{st, ctl, result, map, staged, d_feats, tp, sc, row_cand, cfg.owner_cap, cfg.key_cap,
                debug_sync_enabled() ? reinterpret_cast<long long *>(sc.tri_xyz) + 64 : nullptr, d_error, ctl_prev, 0, 0};
    if (parts & 1)
    {
    MapCandArgs mc{st, ctl, early_pass ? 2 : 0, tp.staged_threshold, PoseD{}, 0, map.xyz, map.desc, d_feats, tp.cam, sc.ms.proj, sc.ms.vis,
                   sc.map_cand, ctl_prev};
    LVT_TIMED(stream, early_pass ? K_MAPCAND_EARLY : K_MAPCAND,
              launch_chained(mapcand_kernel, dim3(148 * 2), dim3(kCandWarps * 32), 0, stream, mc));
    LVT_LAUNCH_CHECK(stream, "mapcand_kernel");
    if (early_pass)
    {
        TrackArgs a1 = a;
        a1.part = 1;
        LVT_TIMED(stream, K_TRACK_A_EARLY,
                  launch_chained_cluster(track_a_kernel, dim3(cfg.cluster), dim3(kTrackThreads), smem, stream, cfg.cluster, a1));
        LVT_LAUNCH_CHECK(stream, "track_a_kernel (early part)");
        // The previous frame's map maintenance (side stream) has to be through before anything else of this frame:
        // the rest of track_a -- ONE CTA, resident behind the early part -- waits for its sequence number on the device.
        // (Parameter sets that send every new point straight to the map -- no staging, RGB-D -- append hundreds to
        // thousands of points per frame: there the rest keeps the cluster for their rounds.)
        a.part = 2;
        a.wait_seq = (int)((unsigned)ov->seq - 1u);
        const int rest_ctas = (tp.staged_threshold == 0 || tp.sensor == 2) ? cfg.cluster : 1;
        LVT_TIMED(stream, K_TRACK_A,
                  launch_chained_cluster(track_a_kernel, dim3(rest_ctas), dim3(kTrackThreads), smem, stream, rest_ctas, a));
    }
    else
        LVT_TIMED(stream, K_TRACK_A,
                  launch_chained_cluster(track_a_kernel, dim3(cfg.cluster), dim3(kTrackThreads), smem, stream, cfg.cluster, a));
    LVT_LAUNCH_CHECK(stream, "track_a_kernel");
    PoseArgs pa{ctl, sc.sol_xyz, sc.sol_uv, 0, PoseD{}, tp.cam, sc.level, sc.inlier, sc.e2, nullptr, nullptr,
                debug_sync_enabled() ? reinterpret_cast<long long *>(sc.tri_xyz) : nullptr, st, early, early_seq,
                map, d_feats, tp.untracked_threshold};
    // two clusters: the solver and, next to it, the culling of the map
    LVT_TIMED(stream, K_POSE, launch_chained(pose_kernel, dim3(2 * kPoseCluster), dim3(kPoseThreads), kPoseSmemBytes, stream, pa));
    LVT_LAUNCH_CHECK(stream, "pose_kernel");
    if (pa.dbg && std::getenv("LVT_B200_POSEDBG"))
    {
        static int calls = 0;
        if (++calls == 8)
        {
            long long h[32];
            cudaMemcpy(h, pa.dbg, sizeof(h), cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "pose pass: compute %lld | warp+cta reduce %lld | cluster.sync %lld | dsmem reduce %lld | boss LM step %lld cycles\n",
                         h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4]);
            std::fprintf(stderr, "pose LM block %lld cycles (bracketed) | LM end -> past the CTA barrier %lld\n", h[6], h[7] - h[5]);
            {
                int last = 10;
                while (last + 1 < 31 && h[last + 1] > h[last])
                    last++;
                const long long base = h[last - 1]; // end of the evaluation before the last one
                std::fprintf(stderr, "pose last iteration, cycles since the previous evaluation returned: LM end %lld | past barrier %lld | "
                                     "eval start %lld | computed %lld | cta sums %lld | cluster sums in %lld | totals %lld | returned %lld\n",
                             h[5] - base, h[7] - base, h[0] - base, h[1] - base, h[2] - base, h[3] - base, h[4] - base, h[last] - base);
            }
            std::fprintf(stderr, "pose solve: setup %lld | evaluations", h[9] - h[8]);
            for (int k = 10; k < 31 && h[k] > h[k - 1]; k++)
                std::fprintf(stderr, " %lld", h[k] - h[k - 1]);
            std::fprintf(stderr, " | total %lld cycles\n", h[31] - h[8]);
        }
    }
    if (a.dbg && std::getenv("LVT_B200_TRACKDBG"))
    {
        static int calls = 0;
        static const int at = std::getenv("LVT_B200_DBG_FRAME") ? std::atoi(std::getenv("LVT_B200_DBG_FRAME")) : 8;
        if (calls == 0)
            cudaMemset(a.dbg, 0, 32 * sizeof(long long)), cudaMemset(pa.dbg, 0, 32 * sizeof(long long));
        if (++calls == at)
        {
            long long h[32];
            cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "map pass round 2: wipe+sync %lld | evaluate %lld | count+sync %lld | forward flags %lld | cluster barrier %lld | read+rotate %lld\n",
                         h[24] - h[4], h[25] - h[24], h[26] - h[25], h[27] - h[26], h[28] - h[27], h[5] - h[28]);
            std::fprintf(stderr, "map pass round 2, thread 0: best2 of its first query %lld | publish %lld | rest of evaluate %lld\n", h[29] - h[24],
                         h[30] - h[29], h[25] - h[30]);
            std::fprintf(stderr, "map pass: n_fast %lld n_slow %lld smem keys %lld sum counts %lld max count %lld from global %lld\n", h[16],
                         h[17], h[18], h[19], h[20], h[21]);
            std::fprintf(stderr, "map pass set-up: owners+replica %lld | first cluster barrier %lld | work lists %lld | lists into shared memory %lld | "
                                 "(trace) + flags + barrier %lld\n", h[22] - h[0], h[23] - h[22], h[1] - h[23], h[31] - h[1], h[2] - h[31]);
            std::fprintf(stderr, "map pass: lists %lld | cache+reset %lld | rounds", h[1] - h[0], h[2] - h[1]);
            for (int r = 0; r < 12 && h[3 + r] > h[2] && h[3 + r] < h[15]; r++)
                std::fprintf(stderr, " %lld", h[3 + r] - (r ? h[2 + r] : h[2]));
            std::fprintf(stderr, " | total %lld cycles\n", h[15] - h[0]);
            // the row-matching pass of the frame before (track_b, single CTA): same marks, second half of the buffer
            cudaMemcpy(h, a.dbg + 60000, sizeof(h), cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "row pass (previous frame that triangulated): n_fast %lld n_slow %lld smem keys %lld sum counts %lld max count %lld | lists %lld | cache+reset %lld | rounds",
                         h[16], h[17], h[18], h[19], h[20], h[1] - h[0], h[2] - h[1]);
            for (int r = 0; r < 12 && h[3 + r] > h[2] && h[3 + r] < h[15]; r++)
                std::fprintf(stderr, " %lld", h[3 + r] - (r ? h[2 + r] : h[2]));
            std::fprintf(stderr, " | total %lld cycles\n", h[15] - h[0]);
            std::fprintf(stderr, "row pass set-up: owners %lld | barrier %lld | work lists %lld | lists into shared memory %lld | (trace) + flags + barrier %lld\n",
                         h[22] - h[0], h[23] - h[22], h[1] - h[23], h[31] - h[1], h[2] - h[31]);
            std::fprintf(stderr, "row pass round 2: wipe+sync %lld | evaluate %lld | count+sync %lld | swap+sync %lld\n", h[24] - h[4], h[25] - h[24],
                         h[26] - h[25], h[5] - h[26]);
        }
    }
    }
    if (!(parts & 2))
        return LVTK_OK;
    if (ov)
    {
        // the rest of the frame on the side stream: the next frame's candidate listing and early map pass follow the
        // pose on `stream` at once
        LVT_CUDA_TRY(cudaEventRecord(ov->pose_done, stream));
        LVT_CUDA_TRY(cudaStreamWaitEvent(ov->side, ov->pose_done, 0));
        stream = ov->side;
    }
    MapCandArgs sc2{st, ctl, 1, tp.staged_threshold, PoseD{}, 0, staged.xyz, staged.desc, d_feats, tp.cam, sc.bs.proj, sc.bs.vis,
                    sc.staged_cand, nullptr};
    LVT_TIMED(stream, K_STAGEDCAND, launch_chained(mapcand_kernel, dim3(148), dim3(kCandWarps * 32), 0, stream, sc2));
    LVT_LAUNCH_CHECK(stream, "stagedcand_kernel");
    // everything up to here needs the left image only; the right image's features and the row-matching
    // candidates (extracted on another stream by the blocking stereo path) join here
    if (right_ready)
        LVT_CUDA_TRY(cudaStreamWaitEvent(stream, right_ready, 0));
    LVT_TIMED(stream, K_TRACK_B, launch_chained(track_b_kernel, dim3(1), dim3(kTrackThreads), smem, stream, a));
    LVT_LAUNCH_CHECK(stream, "track_b_kernel");
    if (ov && ov->next_feats)
    {
        // the points this frame appended to the map, listed for the next frame while its early map pass is running
        MapCandArgs tc{st, ctl, 3, tp.staged_threshold, PoseD{}, 0, map.xyz, map.desc, ov->next_feats, tp.cam, sc.ms.proj, sc.ms.vis,
                       sc.map_cand, nullptr};
        LVT_TIMED(stream, K_TAILCAND, launch_chained(mapcand_kernel, dim3(148), dim3(kCandWarps * 32), 0, stream, tc));
        LVT_LAUNCH_CHECK(stream, "mapcand_kernel (appended points)");
    }
    if (ov)
    {
        count_launch();
        LVT_CUDA_TRY(launch_chained(signal_kernel, dim3(1), dim3(32), 0, stream, st, ov->seq));
        LVT_LAUNCH_CHECK(stream, "signal_kernel");
        LVT_CUDA_TRY(cudaEventRecord(ov->rest_done, stream));
    }
    return LVTK_OK;
}

int launch_clear_halt(TrackState *st, cudaStream_t stream)
{
    clear_halt_kernel<<<1, 1, 0, stream>>>(st);
    LVT_LAUNCH_CHECK(stream, "clear_halt_kernel");
    return LVTK_OK;
}

int launch_reset_state(TrackState *st, cudaStream_t stream)
{
    reset_state_kernel<<<1, 1, 0, stream>>>(st);
    LVT_LAUNCH_CHECK(stream, "reset_state_kernel");
    return LVTK_OK;
}

int launch_match_seam(const double *d_xyz, const uint32_t *d_pdesc, int m, const PoseD &pose, const FeatDev *d_feat,
                      const CamParams &cam, int retry_below, const MatchScratch &ms, const CandLists &lists,
                      int *d_match_idx, float *d_d1, float *d_d2, int *d_count_retried, const TrackLaunchCfg &cfg,
                      cudaStream_t stream)
{
    MapCandArgs mc{nullptr, nullptr, 0, 0, pose, m, d_xyz, d_pdesc, d_feat, cam, ms.proj, ms.vis, lists, nullptr};
    mapcand_kernel<<<148 * 2, kCandWarps * 32, 0, stream>>>(mc);
    LVT_LAUNCH_CHECK(stream, "mapcand_kernel");
    MatchSeamArgs a{d_pdesc, m, d_feat, cam, retry_below, ms, lists, d_match_idx, d_d1, d_d2, d_count_retried, cfg.owner_cap, cfg.key_cap};
    match_seam_kernel<<<1, kTrackThreads, track_smem_bytes(cfg), stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "match_seam_kernel");
    return LVTK_OK;
}

int launch_row_seam(const FeatDev *d_feats, const CamParams &cam, const CandLists &lists, int *d_choice, int *d_items,
                    int *d_query, int *d_train, int *d_count, const TrackLaunchCfg &cfg, cudaStream_t stream)
{
    if (int rc = launch_rowcand(d_feats, cam, lists, stream))
        return rc;
    RowSeamArgs a{d_feats, cam, lists, d_choice, d_items, d_query, d_train, d_count, cfg.owner_cap, cfg.key_cap};
    row_seam_kernel<<<1, kTrackThreads, track_smem_bytes(cfg), stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "row_seam_kernel");
    return LVTK_OK;
}

int launch_pose_seam(const double *d_xyz, const float2 *d_uv, int m, const PoseD &init, const CamParams &cam,
                     uint8_t *d_level, uint8_t *d_inlier, double *d_e2, PoseD *d_out, int *d_n_inliers, cudaStream_t stream)
{
    if (int rc = ensure_smem())
        return rc;
    PoseArgs a{nullptr, d_xyz, d_uv, m, init, cam, d_level, d_inlier, d_e2, d_out, d_n_inliers, nullptr, nullptr, nullptr, 0,
               PointStore{}, nullptr, 0};
    pose_kernel<<<kPoseCluster, kPoseThreads, kPoseSmemBytes, stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "pose_kernel");
    return LVTK_OK;
}

int launch_tri_seam(const PoseD &pose, const CamParams &cam, const float2 *d_uvl, const float2 *d_uvr, int n,
                    double *d_xyz, uint8_t *d_ok, cudaStream_t stream)
{
    if (n <= 0)
        return LVTK_OK;
    TriSeamArgs a{pose, cam, d_uvl, d_uvr, n, d_xyz, d_ok};
    tri_seam_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "tri_seam_kernel");
    return LVTK_OK;
}

} // namespace lvtb
