// The per-frame tracking kernel (one persistent CTA per sequence and frame) and the thin
// kernels behind the seam ABI (lvtk_match_projected / lvtk_row_match / lvtk_solve_pose /
// lvtk_triangulate), all built from the block-level routines of match.cuh, pose.cuh, track.cuh.
#include "track.cuh"

namespace lvtb
{

constexpr int kTrackThreads = 512;

struct TrackArgs
{
    TrackState *st;
    FrameResult *result;
    PointStore map, staged;
    const FeatDev *feats; // [2] left (or gray), right
    TrackParams tp;
    TrackScratch sc;
    int owner_cap; // ints per owner array in dynamic shared memory
};

struct TrackShared
{
    PoseShared pose;
    double W[24]; // world->camera of the left [0..11] and right [12..23] camera
    PoseD pred, opt;
    lvt_frame_info info;
    int flag[2];
    int scan[34];
    int ctrl[4];
};

// lvt_local_map::update_with_new_triangulation (lvt/src/lvt_local_map.cpp:331-353) for the pose
// in sh.opt.  Returns (uniformly) the number of new points; updates *map_n / *staged_n (locals).
__device__ int block_new_triangulation(const TrackArgs &a, TrackShared &sh, const FeatDev &fl, int nl, const FeatDev &fr,
                                       int nr, bool dont_stage, int *owner_a, int *owner_b, int &map_n, int &staged_n)
{
    const TrackParams &tp = a.tp;
    const bool to_map = dont_stage || tp.staged_threshold == 0 || map_n < kNMapPoints;
    const PointStore &dst = to_map ? a.map : a.staged;
    const int base = to_map ? map_n : staged_n;
    int added = 0;

    if (tp.sensor == 1)
    {
        const int np = block_row_match(fl, nl, fr, nr, tp.cam, a.sc.row_choice, owner_a, owner_b, sh.flag, sh.scan,
                                       a.sc.pair_query, a.sc.pair_train);
        if (np == 0)
            return 0;
        if (threadIdx.x == 0)
        {
            world_to_camera(sh.opt, sh.W);
            const PoseD pr = right_pose(sh.opt, (double)tp.cam.baseline);
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
            double R[9];
            quat_to_mat(sh.opt.q, R);
            for (int i = 0; i < 9; i++)
                sh.W[i] = R[i];
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
                w[r] = (sh.W[3 * r] * (double)x + sh.W[3 * r + 1] * (double)y + sh.W[3 * r + 2] * (double)z) + sh.opt.t[r];
            copy_point(dst, d, w, fl.desc + 8 * (size_t)i, 0, 0, 0);
        }
        added = nl;
    }
    if (base + added > dst.cap)
    {
        if (threadIdx.x == 0)
            a.st->error = LVTK_ERR_CAPACITY;
        added = dst.cap - base;
    }
    if (to_map)
        map_n += added;
    else
        staged_n += added;
    __syncthreads();
    return added;
}

__global__ void __launch_bounds__(kTrackThreads, 1) track_frame_kernel(TrackArgs a)
{
    extern __shared__ int s_owner[];
    __shared__ TrackShared sh;
    int *owner_a = s_owner, *owner_b = s_owner + a.owner_cap;

    TrackState &S = *a.st;
    const TrackParams &tp = a.tp;
    const FeatDev fl = a.feats[0], fr = a.feats[1];
    const int nl = min(*fl.n, a.owner_cap);
    const int nr = tp.sensor == 1 ? min(*fr.n, a.owner_cap) : 0;
    const int state0 = S.state;
    int map_n = S.map_n, staged_n = S.staged_n;
    const PoseD last_pose = S.last_pose;

    if (threadIdx.x == 0)
    {
        lvt_frame_info z = {};
        sh.info = z;
        sh.info.n_features_left = nl;
        sh.info.n_features_right = nr;
        sh.info.map_points_before = 0;
    }
    __syncthreads();

    PoseD out_pose = last_pose;
    int new_state = state0;
    bool accepted = false; // computed pose becomes m_last_pose

    if (state0 == 1)
    {
        // first frame: identity pose, seed the map (lvt/src/lvt_system.cpp:185-193)
        if (threadIdx.x == 0)
        {
            sh.opt.q = Quat{1, 0, 0, 0};
            sh.opt.t[0] = sh.opt.t[1] = sh.opt.t[2] = 0;
        }
        __syncthreads();
        const int added = block_new_triangulation(a, sh, fl, nl, fr, nr, true, owner_a, owner_b, map_n, staged_n);
        new_state = 2;
        out_pose = sh.opt;
        if (threadIdx.x == 0)
        {
            S.last_matches[0] = map_n;
            sh.info.triangulated = 1;
            sh.info.new_points = added;
        }
    }
    else if (state0 == 2)
    {
        // ---- predict + find_matches (lvt/src/lvt_system.cpp:196, lvt/src/lvt_local_map.cpp:136-229)
        if (threadIdx.x == 0)
        {
            sh.pred = motion_predict(S.motion, last_pose);
            world_to_camera(sh.pred, sh.W);
            sh.info.map_points_before = map_n;
            sh.info.staged_before = staged_n;
        }
        __syncthreads();
        const int M = map_n;
        block_project(a.map.xyz, M, sh.W, tp.cam, a.sc.ms);
        const int R = tp.cam.tracking_radius;
        int count = block_match_projected(a.map.desc, a.sc.ms, M, fl, nl, tp.cam, (float)(R * R), false, owner_a,
                                          owner_b, sh.flag, nullptr, nullptr);
        int retried = 0;
        if (count < kNMatchesTh)
        {
            retried = 1; // marks reset, radius doubled, cell window unchanged (lvt_local_map.cpp:173-199)
            count = block_match_projected(a.map.desc, a.sc.ms, M, fl, nl, tp.cam, (float)((2 * R) * (2 * R)), false,
                                          owner_a, owner_b, sh.flag, nullptr, nullptr);
        }
        for (int j = threadIdx.x; j < nl; j += blockDim.x)
            fl.matched[j] = owner_a[j] != kFree;
        // bookkeeping (:201-224) + the solver's inputs, in map order
        int n_matches = 0;
        for (int i0 = 0; i0 < M; i0 += blockDim.x)
        {
            const int i = i0 + threadIdx.x;
            int c = -3;
            if (i < M)
            {
                c = a.sc.ms.vis[i] ? a.sc.ms.choice[i] : -2;
                a.map.match_idx[i] = c;
                if (c < 0)
                    a.map.counter[i] += 1;
                else
                    a.map.age[i] += 1;
            }
            int total;
            const int pos = block_exclusive_scan(c >= 0, sh.scan, &total);
            if (c >= 0)
            {
                const int d = n_matches + pos;
                a.sc.sol_xyz[3 * d] = a.map.xyz[3 * i];
                a.sc.sol_xyz[3 * d + 1] = a.map.xyz[3 * i + 1];
                a.sc.sol_xyz[3 * d + 2] = a.map.xyz[3 * i + 2];
                a.sc.sol_uv[d] = fl.xy[c];
            }
            n_matches += total;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            sh.info.tracked = n_matches;
            sh.info.retried_matching = retried;
        }
        if (n_matches < tp.min_matches)
        {
            new_state = 3; // lost: return the last pose (lvt/src/lvt_system.cpp:267-272,199-204)
        }
        else
        {
            if (threadIdx.x == 0)
            {
                S.last_matches[0] = S.last_matches[1];
                S.last_matches[1] = S.last_matches[2];
                S.last_matches[2] = n_matches;
            }
            // ---- pose (lvt/src/lvt_pnp_solver.cpp:60-128)
            const int inl = block_solve_pose(sh.pose, a.sc.sol_xyz, a.sc.sol_uv, n_matches, sh.pred, tp.cam, a.sc.level,
                                             a.sc.e2, a.sc.inlier, &sh.opt, sh.scan);
            if (threadIdx.x == 0)
                sh.info.inliers = inl;
            // ---- clean_untracked_points (lvt/src/lvt_local_map.cpp:393-413)
            const int th = tp.untracked_threshold;
            for (int i = threadIdx.x; i < M; i += blockDim.x)
                if (a.map.counter[i] >= th && a.map.match_idx[i] >= 0)
                    fl.matched[a.map.match_idx[i]] = 0;
            __syncthreads();
            const int *cnt = a.map.counter;
            map_n = block_compact_points(a.map, M, [cnt, th](int i) { return cnt[i] < th; }, sh.scan);

            // ---- update_staged_map_points (lvt/src/lvt_local_map.cpp:355-391)
            if (tp.staged_threshold > 0 && staged_n > 0)
            {
                const int Sn = staged_n;
                if (threadIdx.x == 0)
                    world_to_camera(sh.opt, sh.W);
                __syncthreads();
                block_project(a.staged.xyz, Sn, sh.W, tp.cam, a.sc.ms);
                block_match_projected(a.staged.desc, a.sc.ms, Sn, fl, nl, tp.cam, (float)(R * R), true, owner_a, owner_b,
                                      sh.flag, nullptr, nullptr);
                for (int j = threadIdx.x; j < nl; j += blockDim.x)
                    if (owner_a[j] != kFree)
                        fl.matched[j] = 1;
                // hit -> counter++; promote when counter == staged_threshold or the map is still
                // below 250 points.  Sequentially the map grows with every promotion, which is
                // equivalent to: map_n + (#hits before this one) < 250.
                const int map0 = map_n;
                int hits = 0, promoted = 0;
                uint8_t *flag = a.sc.level; // 0 erase, 1 keep staged, 2 promoted
                for (int i0 = 0; i0 < Sn; i0 += blockDim.x)
                {
                    const int i = i0 + threadIdx.x;
                    const bool hit = i < Sn && a.sc.ms.vis[i] && a.sc.ms.choice[i] >= 0;
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
                            copy_point(a.map, d, a.staged.xyz + 3 * i, a.staged.desc + 8 * (size_t)i, cnt_new,
                                       a.staged.age[i], a.staged.match_idx[i]);
                    }
                    hits += tot_h;
                    promoted += tot_p;
                }
                __syncthreads();
                map_n = map0 + promoted;
                if (map_n > a.map.cap)
                {
                    if (threadIdx.x == 0)
                        S.error = LVTK_ERR_CAPACITY;
                    map_n = a.map.cap;
                }
                staged_n = block_compact_points(a.staged, Sn, [flag](int i) { return flag[i] == 1; }, sh.scan);
            }

            // ---- need_new_triangulation (lvt/src/lvt_system.cpp:308-334)
            if (threadIdx.x == 0)
            {
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
                    sh.info.triangulated = 1;
                    sh.info.new_points = added;
                }
            }
            out_pose = sh.opt;
            accepted = true;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        S.frame_number += 1;
        S.state = new_state;
        S.map_n = map_n;
        S.staged_n = staged_n;
        if (accepted)
            S.last_pose = out_pose;
        sh.info.frame_number = S.frame_number;
        sh.info.state = new_state;
        sh.info.map_points_after = map_n;
        sh.info.staged_after = staged_n;
        a.result->pose = out_pose;
        a.result->info = sh.info;
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
    S.last_pose.q = Quat{1, 0, 0, 0};
    S.last_pose.t[0] = S.last_pose.t[1] = S.last_pose.t[2] = 0;
    motion_reset(S.motion);
}

// ---------------------------------------------------------------------------------------------
// seam kernels
// ---------------------------------------------------------------------------------------------
struct MatchSeamArgs
{
    const double *xyz;
    const uint32_t *pdesc;
    int m;
    PoseD pose;
    const FeatDev *feat;
    CamParams cam;
    int retry_below;
    MatchScratch ms;
    int *match_idx;
    float *d1, *d2;
    int *count_retried; // [2]
    int owner_cap;
};

__global__ void __launch_bounds__(kTrackThreads, 1) match_seam_kernel(MatchSeamArgs a)
{
    extern __shared__ int s_owner[];
    __shared__ double W[12];
    __shared__ int s_flag[2];
    int *owner_a = s_owner, *owner_b = s_owner + a.owner_cap;
    const FeatDev f = *a.feat;
    const int n = min(*f.n, a.owner_cap);
    if (threadIdx.x == 0)
        world_to_camera(a.pose, W);
    __syncthreads();
    block_project(a.xyz, a.m, W, a.cam, a.ms);
    const int R = a.cam.tracking_radius;
    int count = block_match_projected(a.pdesc, a.ms, a.m, f, n, a.cam, (float)(R * R), true, owner_a, owner_b, s_flag,
                                      a.d1, a.d2);
    int retried = 0;
    if (count < a.retry_below)
    {
        retried = 1;
        count = block_match_projected(a.pdesc, a.ms, a.m, f, n, a.cam, (float)((2 * R) * (2 * R)), false, owner_a,
                                      owner_b, s_flag, a.d1, a.d2);
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
    int *choice, *query, *train, *count;
    int owner_cap;
};

__global__ void __launch_bounds__(kTrackThreads, 1) row_seam_kernel(RowSeamArgs a)
{
    extern __shared__ int s_owner[];
    __shared__ int s_flag[2];
    __shared__ int s_scan[34];
    const FeatDev fl = a.feats[0], fr = a.feats[1];
    const int nl = min(*fl.n, a.owner_cap), nr = min(*fr.n, a.owner_cap);
    const int np = block_row_match(fl, nl, fr, nr, a.cam, a.choice, s_owner, s_owner + a.owner_cap, s_flag, s_scan,
                                   a.query, a.train);
    if (threadIdx.x == 0)
        *a.count = np;
}

struct PoseSeamArgs
{
    const double *xyz;
    const float2 *uv;
    int m;
    PoseD init;
    CamParams cam;
    uint8_t *level, *inlier;
    double *e2;
    PoseD *out;
};

__global__ void __launch_bounds__(kTrackThreads, 1) pose_seam_kernel(PoseSeamArgs a)
{
    __shared__ PoseShared s;
    __shared__ int s_scan[34];
    __shared__ PoseD s_out;
    block_solve_pose(s, a.xyz, a.uv, a.m, a.init, a.cam, a.level, a.e2, a.inlier, &s_out, s_scan);
    if (threadIdx.x == 0)
        *a.out = s_out;
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
static int ensure_smem(int owner_cap)
{
    static int configured = 0;
    const int bytes = 2 * owner_cap * (int)sizeof(int);
    if (bytes > configured)
    {
        LVT_CUDA_TRY(cudaFuncSetAttribute(track_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        LVT_CUDA_TRY(cudaFuncSetAttribute(match_seam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        LVT_CUDA_TRY(cudaFuncSetAttribute(row_seam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        configured = bytes;
    }
    return LVTK_OK;
}

int launch_track_frame(TrackState *st, FrameResult *result, const PointStore &map, const PointStore &staged,
                       const FeatDev *d_feats, const TrackParams &tp, const TrackScratch &sc, int owner_cap,
                       cudaStream_t stream)
{
    if (int rc = ensure_smem(owner_cap))
        return rc;
    TrackArgs a{st, result, map, staged, d_feats, tp, sc, owner_cap};
    track_frame_kernel<<<1, kTrackThreads, 2 * owner_cap * sizeof(int), stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "track_frame_kernel");
    return LVTK_OK;
}

int launch_reset_state(TrackState *st, cudaStream_t stream)
{
    reset_state_kernel<<<1, 1, 0, stream>>>(st);
    LVT_LAUNCH_CHECK(stream, "reset_state_kernel");
    return LVTK_OK;
}

int launch_match_seam(const double *d_xyz, const uint32_t *d_pdesc, int m, const PoseD &pose, const FeatDev *d_feat,
                      const CamParams &cam, int retry_below, const MatchScratch &ms, int *d_match_idx, float *d_d1,
                      float *d_d2, int *d_count_retried, int owner_cap, cudaStream_t stream)
{
    if (int rc = ensure_smem(owner_cap))
        return rc;
    MatchSeamArgs a{d_xyz, d_pdesc, m, pose, d_feat, cam, retry_below, ms, d_match_idx, d_d1, d_d2, d_count_retried, owner_cap};
    match_seam_kernel<<<1, kTrackThreads, 2 * owner_cap * sizeof(int), stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "match_seam_kernel");
    return LVTK_OK;
}

int launch_row_seam(const FeatDev *d_feats, const CamParams &cam, int *d_choice, int *d_query, int *d_train, int *d_count,
                    int owner_cap, cudaStream_t stream)
{
    if (int rc = ensure_smem(owner_cap))
        return rc;
    RowSeamArgs a{d_feats, cam, d_choice, d_query, d_train, d_count, owner_cap};
    row_seam_kernel<<<1, kTrackThreads, 2 * owner_cap * sizeof(int), stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "row_seam_kernel");
    return LVTK_OK;
}

int launch_pose_seam(const double *d_xyz, const float2 *d_uv, int m, const PoseD &init, const CamParams &cam,
                     uint8_t *d_level, uint8_t *d_inlier, double *d_e2, PoseD *d_out, cudaStream_t stream)
{
    PoseSeamArgs a{d_xyz, d_uv, m, init, cam, d_level, d_inlier, d_e2, d_out};
    pose_seam_kernel<<<1, kTrackThreads, 0, stream>>>(a);
    LVT_LAUNCH_CHECK(stream, "pose_seam_kernel");
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
