// Device-resident tracking state of one sequence and the block-level map maintenance steps.
// Header-only device code; the kernels are in track.cu.
//
// The reference keeps the local map in host vectors (lvt/src/lvt_local_map.h:64-85) and walks
// them with scalar loops.  Here the map lives in HBM as structure-of-arrays and the whole of
// lvt_system::perform_tracking (lvt/src/lvt_system.cpp:252-306) runs as a chain of kernels per
// frame (track.cu): projection matching, pose solve, culling, staging, triangulation -- no host
// round trip between the stages, only the pose and the per-frame counters leave the GPU.
#pragma once
#include "match.cuh"
#include "pose.cuh"

namespace lvtb
{

// map points or staged points (lvt_local_map::lvt_map_point, lvt/src/lvt_local_map.h:64-73)
struct PointStore
{
    double *xyz;    // [cap][3] world position
    uint32_t *desc; // [cap][8]
    int *counter;   // staged: frames tracked while staged; map: frames it failed tracking
    int *age;
    int *match_idx;
    int cap;
};

struct MotionState // lvt/src/lvt_motion_model.h:41-45
{
    Quat last_q, ang_vel;
    double last_pos[3], lin_vel[3];
};

struct TrackState
{
    int state;        // lvt_system::eState 1/2/3
    int frame_number; // lvt_system::m_frame_number
    int map_n, staged_n;
    int last_matches[3]; // lvt/src/lvt_system.cpp:37, N_MATCHES_WINDOWS = 3
    int error;           // sticky LVTK_ERR_*
    int halt;            // != 0: frame number (1-based) of the first frame refused because the point stores could
                         // overflow; that frame and every later one leave the state untouched until the host has
                         // grown the stores and cleared this (lvtk_ctx: ctx_grow_points)
    PoseD last_pose;
    MotionState motion;
    // world->camera of the pose the motion model predicts for the NEXT frame, prepared by the kernel
    // that finishes a frame so that mapcand_kernel starts projecting at once (same arithmetic as
    // track_a's own prediction: motion_predict on a copy of the motion state, world_to_camera)
    double pred_W[12];
    // candidate lists of the map pass exist for the map points [0, cand_done): all of them behind the regular
    // mapcand_kernel; behind the EARLY one (launched as soon as the previous frame's pose is known, while that
    // frame's map maintenance is still running) the points its culling left.  track_a_kernel lists the rest
    // -- what track_b has appended since -- itself and clears this.
    int cand_done;
    // ... and for [cand_done, tail_done): the points track_b appended, listed behind it on its own stream for the
    // coming frame (mapcand_kernel, which == 3) while that frame's early map pass is running
    int tail_done;
    // sequence number of the last frame whose map maintenance (track_b and what follows it on the side stream) is
    // through -- written by signal_kernel; the rest of the next frame's track_a waits for it ON THE DEVICE
    // (TrackArgs::wait_seq): its CTA is resident behind the early part already, so it starts within a microsecond
    // instead of competing for an empty SM with the extraction kernels after a stream-level wait
    int rest_seq;
};

struct FrameResult
{
    PoseD pose;
    lvt_frame_info info;
    long long cycles[8]; // clock64() at the phase boundaries of the tracking kernel (profiling aid)
    int rounds[8];       // fixed-point rounds: map pass, retry pass, staged pass, row matching; [4] = evaluations of the
                         // pose solver (passes over the correspondences: linearisations + LM trials), [5] = clock cycles of
                         // the map culling next to the solver, [6] / [7] = ns into track_b when the staged points / the
                         // row matching were done (profiling aids)
    long long dbg[8];    // nanosecond marks inside track_b (profiling aid, lvt_debug_frame_marks)
    long long amark[4];  // nanosecond marks of the early map pass (track_a_kernel part 1): start, end; [2..3] spare
};

// what a blocking caller waits for: available as soon as the pose solver is through, while the map
// maintenance of the frame (track_b) is still running
struct EarlyResult
{
    PoseD pose;
    int state;
    int seq; // written last (after a system-wide fence): the caller polls it in pinned host memory
};

struct TrackParams
{
    CamParams cam;
    int sensor;              // 1 stereo, 2 rgbd
    int min_matches;         // min_num_matches_for_tracking
    int untracked_threshold; // untracked_threshold
    int staged_threshold;
    int triangulation_policy;
};

// shared-memory layout of the tracking kernels for one context (track_configure: per-device function
// attributes set once, thread-safe; nothing process-wide is consulted at launch time)
struct TrackLaunchCfg
{
    int owner_cap = 0; // ints per owner array (= feature capacity)
    int key_cap = 0;   // candidate keys that fit behind the two owner arrays
    int cluster = 1;   // CTAs of track_a_kernel's cluster
};

// The batched engine overlaps a frame's map maintenance (stagedcand + track_b, on `side`) with the next frame's
// candidate listing and early map pass (track.cu, track_a_kernel): the events tie the two streams together.
struct TrackOverlap
{
    cudaStream_t side;          // where stagedcand + track_b of this frame run
    cudaEvent_t pose_done;      // recorded behind this frame's pose_kernel (the side stream waits for it)
    cudaEvent_t rest_done;      // recorded behind this frame's track_b_kernel
    cudaEvent_t prev_rest_done; // the previous frame's rest_done (early only)
    bool early;                 // the previous frame was launched with an overlap too: list / match early
    int seq;                    // this frame's sequence number (TrackState::rest_seq once its rest is through)
    const FeatDev *next_feats;  // device: the left features of the frame behind this one when they are extracted (the
                                // points track_b appends are listed for it right behind track_b), else nullptr
};
// scratch for one tracking CTA (global memory, sized for the point / feature capacities)
struct TrackScratch
{
    MatchScratch ms;   // [pcap]
    double *sol_xyz;   // [pcap][3]  matched map points, in map order
    float2 *sol_uv;    // [pcap]     their matched keypoints
    uint8_t *level;    // [pcap]
    uint8_t *inlier;   // [pcap]
    double *e2;        // [pcap]
    CandLists map_cand; // [pcap][kMapCandCap] candidate keys of the map pass (mapcand_kernel)
    // track_b (staged points, row matching) has its own copies of everything the map pass of the NEXT frame
    // writes: in the batched engine the two run at the same time (context.cu, run_frames)
    MatchScratch bs;       // [pcap] / items [2 * pcap + 1024]
    CandLists staged_cand; // [pcap][kMapCandCap]
    int *row_choice;   // [fcap]
    int *pair_query;   // [fcap]
    int *pair_train;   // [fcap]
    double *tri_xyz;   // [fcap][3]
    uint8_t *tri_ok;   // [fcap]
};

// ---- motion model (lvt/src/lvt_motion_model.cpp:34-65), one thread -----------------------------
__device__ inline Quat quat_slerp(const Quat &a, double t, const Quat &b)
{
    const double one = 1.0 - 2.220446049250313e-16;
    const double d = quat_dot(a, b), abs_d = fabs(d);
    double s0, s1;
    if (abs_d >= one)
    {
        s0 = 1.0 - t;
        s1 = t;
    }
    else
    {
        const double theta = acos(abs_d), sin_theta = sin(theta);
        s0 = sin((1.0 - t) * theta) / sin_theta;
        s1 = sin(t * theta) / sin_theta;
    }
    if (d < 0)
        s1 = -s1;
    Quat r;
    r.w = s0 * a.w + s1 * b.w;
    r.x = s0 * a.x + s1 * b.x;
    r.y = s0 * a.y + s1 * b.y;
    r.z = s0 * a.z + s1 * b.z;
    return r;
}

__device__ inline void motion_reset(MotionState &m)
{
    m.last_q = Quat{1, 0, 0, 0};
    m.ang_vel = Quat{1, 0, 0, 0};
    for (int i = 0; i < 3; i++)
        m.last_pos[i] = m.lin_vel[i] = 0.0;
}

__device__ inline PoseD motion_predict(MotionState &m, const PoseD &cur)
{
    double nl[3];
    for (int i = 0; i < 3; i++)
        nl[i] = ((cur.t[i] - m.last_pos[i]) + m.lin_vel[i]) * 0.5;
    const Quat diff = quat_mul(cur.q, quat_inverse(m.last_q));
    const Quat nav = quat_normalized(quat_slerp(diff, 0.5, m.ang_vel));
    m.last_q = cur.q;
    m.ang_vel = nav;
    PoseD out;
    for (int i = 0; i < 3; i++)
    {
        m.last_pos[i] = cur.t[i];
        m.lin_vel[i] = nl[i];
        out.t[i] = m.last_pos[i] + m.lin_vel[i];
    }
    out.q = quat_normalized(quat_mul(cur.q, nav));
    return out;
}

// ---- triangulation (lvt/src/lvt_local_map.cpp:258-329), one thread per row match ---------------
// minimum-norm least squares of A[:, 0:3] x = -A[:, 3] by one-sided Jacobi SVD (stands in for
// Eigen::JacobiSVD(...).solve at :292)
__device__ inline void solve_ls_4x3(const double Ain[4][4], double x[3])
{
    double a[4][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, b[4];
    for (int i = 0; i < 4; i++)
    {
        for (int j = 0; j < 3; j++)
            a[i][j] = Ain[i][j];
        b[i] = -Ain[i][3];
    }
    for (int sweep = 0; sweep < 30; sweep++)
    {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; pq++)
        {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2; // (0,1) (0,2) (1,2)
            double alpha = 0, beta = 0, gamma = 0;
            for (int i = 0; i < 4; i++)
            {
                alpha += a[i][p] * a[i][p];
                beta += a[i][q] * a[i][q];
                gamma += a[i][p] * a[i][q];
            }
            if (gamma == 0.0 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta))
                continue;
            rotated = true;
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
            for (int i = 0; i < 4; i++)
            {
                const double ap = a[i][p], aq = a[i][q];
                a[i][p] = c * ap - s * aq;
                a[i][q] = s * ap + c * aq;
            }
            for (int i = 0; i < 3; i++)
            {
                const double vp = V[i][p], vq = V[i][q];
                V[i][p] = c * vp - s * vq;
                V[i][q] = s * vp + c * vq;
            }
        }
        if (!rotated)
            break;
    }
    double s2[3], ab[3], s2max = 0;
    for (int j = 0; j < 3; j++)
    {
        s2[j] = 0;
        ab[j] = 0;
        for (int i = 0; i < 4; i++)
        {
            s2[j] += a[i][j] * a[i][j];
            ab[j] += a[i][j] * b[i];
        }
        s2max = fmax(s2max, s2[j]);
    }
    x[0] = x[1] = x[2] = 0;
    const double thr = 4.0 * 2.220446049250313e-16;
    for (int j = 0; j < 3; j++)
    {
        if (s2[j] <= thr * thr * s2max)
            continue;
        const double w = ab[j] / s2[j];
        for (int i = 0; i < 3; i++)
            x[i] += V[i][j] * w;
    }
}

// Wl, Wr: world->camera of the left / right camera (12 doubles each)
__device__ inline bool triangulate_pair(const double *Wl, const double *Wr, const CamParams &cam, float2 u1, float2 u2,
                                        double out[3])
{
    const double cx = cam.cx, cy = cam.cy, inv_fx = 1.0 / (double)cam.fx, inv_fy = 1.0 / (double)cam.fy;
    const double u1x = ((double)u1.x - cx) * inv_fx, u1y = ((double)u1.y - cy) * inv_fy;
    const double u2x = ((double)u2.x - cx) * inv_fx, u2y = ((double)u2.y - cy) * inv_fy;
    double A[4][4];
    for (int c = 0; c < 4; c++)
    {
        A[0][c] = u1x * Wl[8 + c] - Wl[c];
        A[1][c] = u1y * Wl[8 + c] - Wl[4 + c];
        A[2][c] = u2x * Wr[8 + c] - Wr[c];
        A[3][c] = u2y * Wr[8 + c] - Wr[4 + c];
    }
    solve_ls_4x3(A, out);
    double ul, vl, ur, vr;
    if (!point_visible(Wl, cam, out[0], out[1], out[2], &ul, &vl) ||
        !point_visible(Wr, cam, out[0], out[1], out[2], &ur, &vr))
        return false;
    {
        const double ex = ul - (double)u1.x, ey = vl - (double)u1.y;
        if ((ex * ex + ey * ey) > kReprojectionTh2)
            return false;
    }
    {
        const double ex = ur - (double)u2.x, ey = vr - (double)u2.y;
        if ((ex * ex + ey * ey) > kReprojectionTh2)
            return false;
    }
    return true;
}

// right camera pose (lvt/src/lvt_pose.cpp:28-34)
__device__ inline PoseD right_pose(const PoseD &l, double baseline)
{
    double R[9];
    quat_to_mat(l.q, R);
    PoseD r = l;
    r.t[0] = R[0] * baseline + l.t[0];
    r.t[1] = R[3] * baseline + l.t[1];
    r.t[2] = R[6] * baseline + l.t[2];
    return r;
}

__device__ inline void copy_point(const PointStore &dst, int d, const double *xyz, const uint32_t *desc, int counter,
                                  int age, int match_idx)
{
    dst.xyz[3 * d] = xyz[0];
    dst.xyz[3 * d + 1] = xyz[1];
    dst.xyz[3 * d + 2] = xyz[2];
    const uint4 a = *reinterpret_cast<const uint4 *>(desc), b = *reinterpret_cast<const uint4 *>(desc + 4);
    *reinterpret_cast<uint4 *>(dst.desc + 8 * (size_t)d) = a;
    *reinterpret_cast<uint4 *>(dst.desc + 8 * (size_t)d + 4) = b;
    dst.counter[d] = counter;
    dst.age[d] = age;
    dst.match_idx[d] = match_idx;
}

// In-place, order-preserving compaction of a PointStore by keep(i).  Four points per thread, interleaved
// (point j * blockDim + t of the chunk: coalesced accesses), one packed block scan per 4 x blockDim points.
// A chunk is moved field by field through SHARED memory -- every kept element of the chunk is read into the
// staging buffer (flat, coalesced: the 24-byte positions as a flat array of doubles), a barrier, then
// written to its place -- so destinations never run ahead of sources and nothing is staged in registers
// (at 1024 threads a thread has 64 of them: register staging spills to local memory, i.e. goes through L2).
// Chunks in front of the first dropped point do not move at all.
// s_scan: >= 33 uint64 of shared memory; smem: >= 4 x blockDim x (4 + 32) bytes of shared memory, 16-byte aligned.
template <class Keep>
__device__ inline int block_compact_points(const PointStore &ps, int n, Keep keep, int *s_scan, int *smem,
                                           long long *dbg = nullptr)
{
#define LVT_CDBG(k)                                                                                                   \
    if (dbg && threadIdx.x == 0)                                                                                      \
    dbg[k] = clock64()
    constexpr int K = 4;
    const int T = (int)blockDim.x, t = (int)threadIdx.x;
    unsigned long long *scratch = reinterpret_cast<unsigned long long *>(s_scan);
    int *s_map = smem; // source index of the kept points of the chunk, in order
    double *buf_d = reinterpret_cast<double *>(smem + K * T);
    uint4 *buf_q = reinterpret_cast<uint4 *>(smem + K * T);
    int *buf_i = smem + K * T;
    int running = 0;
    for (int base = 0; base < n; base += K * T)
    {
        int idx[K];
        bool k[K];
        unsigned long long packed = 0;
#pragma unroll
        for (int j = 0; j < K; j++)
        {
            idx[j] = base + j * T + t;
            k[j] = idx[j] < n && keep(idx[j]);
            packed |= (unsigned long long)k[j] << (16 * j);
        }
        LVT_CDBG(1);
        unsigned long long total;
        const unsigned long long excl = block_exclusive_scan4(packed, scratch, &total); // syncs
        LVT_CDBG(2);
        const int kept = scan4_sum(total), chunk = min(K * T, n - base);
        if (running != base || kept != chunk) // uniform: something in or before this chunk was dropped
        {
#pragma unroll
            for (int j = 0; j < K; j++)
                if (k[j])
                    s_map[scan4_position(excl, total, j)] = idx[j];
            __syncthreads();
            // positions: 3 x kept doubles.  Asynchronous copies (LDGSTS): every element of the chunk is in
            // flight at once, no register in between
            for (int e = t; e < 3 * kept; e += T)
            {
                const int p = e / 3;
                cp_async<8>(buf_d + e, ps.xyz + 3 * (size_t)s_map[p] + (e - 3 * p));
            }
            cp_async_wait_all();
            __syncthreads();
            for (int e = t; e < 3 * kept; e += T)
                ps.xyz[3 * (size_t)running + e] = buf_d[e];
            __syncthreads();
            LVT_CDBG(3);
            // descriptors: 2 x kept 16-byte halves
            for (int e = t; e < 2 * kept; e += T)
                cp_async<16>(buf_q + e, ps.desc + 8 * (size_t)s_map[e >> 1] + 4 * (e & 1));
            cp_async_wait_all();
            __syncthreads();
            for (int e = t; e < 2 * kept; e += T)
                *reinterpret_cast<uint4 *>(ps.desc + 8 * (size_t)running + 4 * (size_t)e) = buf_q[e];
            __syncthreads();
            LVT_CDBG(4);
            // counter, age, match_idx
            for (int e = t; e < kept; e += T)
            {
                const int src = s_map[e];
                cp_async<4>(buf_i + e, ps.counter + src);
                cp_async<4>(buf_i + kept + e, ps.age + src);
                cp_async<4>(buf_i + 2 * kept + e, ps.match_idx + src);
            }
            cp_async_wait_all();
            __syncthreads();
            for (int e = t; e < kept; e += T)
            {
                ps.counter[running + e] = buf_i[e];
                ps.age[running + e] = buf_i[kept + e];
                ps.match_idx[running + e] = buf_i[2 * kept + e];
            }
        }
        running += kept;
        __syncthreads();
        LVT_CDBG(5);
    }
#undef LVT_CDBG
    return running;
}

} // namespace lvtb
