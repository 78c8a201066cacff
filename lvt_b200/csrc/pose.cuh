// Block-level motion-only bundle adjustment (one CTA), header-only device code.
//
// Replaces the g2o graph built and solved per frame at lvt/src/lvt_pnp_solver.cpp:60-128:
// one free SBACam vertex, M fixed points, M EdgeProjectP2MC with a Cauchy kernel
// (delta^2 = 5.991), Levenberg-Marquardt on the single 6x6 block (BlockSolver_6_3 +
// LinearSolverPCG == one preconditioned-CG step == one exact solve), two passes of optimize(5),
// edges with chi2 > 5.991 demoted to level 1 after each pass.
//
// Per evaluation every thread takes correspondences i = tid, tid + T, ...: fp64 projection,
// residual, robust weight and (for the linearisation pass) the 2x6 Jacobian; the 21 + 6 + 1
// partial sums of H = sum w J^T J, b = -sum w J^T e and the robust cost are reduced with warp
// shuffles, then across warps through shared memory in a fixed order (deterministic).  Thread 0
// runs the 6x6 solve and the LM accept / reject logic between evaluations.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace lvtb
{

struct CamState
{
    double t[3];
    Quat r;
    double w2n[12]; // [R^T | -R^T t], row-major 3x4
    double fx, fy, cx, cy;
};

// SBACam::setTransform (g2o types/sba/sbacam.h)
__device__ inline void cam_refresh(CamState &c)
{
    double R[9];
    quat_to_mat(c.r, R);
    for (int i = 0; i < 3; i++)
    {
        const double a = R[i], b = R[3 + i], d = R[6 + i];
        c.w2n[4 * i + 0] = a;
        c.w2n[4 * i + 1] = b;
        c.w2n[4 * i + 2] = d;
        c.w2n[4 * i + 3] = -(a * c.t[0] + b * c.t[1] + d * c.t[2]);
    }
}

// SBACam::update: t += dt ; r = normalize(r * (dv, sqrt(1 - |dv|^2)))
__device__ inline void cam_update(CamState &c, const double u[6])
{
    c.t[0] += u[0];
    c.t[1] += u[1];
    c.t[2] += u[2];
    Quat qr;
    qr.x = u[3];
    qr.y = u[4];
    qr.z = u[5];
    qr.w = sqrt(1.0 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]));
    c.r = quat_normalized(quat_mul(c.r, qr));
    cam_refresh(c);
}

// LU with partial pivoting, 6x6
__device__ inline bool solve6(const double *A_in /* 36 */, const double *b_in, double *x)
{
    double A[6][7];
    for (int i = 0; i < 6; i++)
    {
        for (int j = 0; j < 6; j++)
            A[i][j] = A_in[6 * i + j];
        A[i][6] = b_in[i];
    }
    for (int c = 0; c < 6; c++)
    {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(A[r][c]) > fabs(A[piv][c]))
                piv = r;
        if (A[piv][c] == 0.0)
            return false;
        if (piv != c)
            for (int j = 0; j < 7; j++)
            {
                const double tmp = A[piv][j];
                A[piv][j] = A[c][j];
                A[c][j] = tmp;
            }
        for (int r = c + 1; r < 6; r++)
        {
            const double fct = A[r][c] / A[c][c];
            for (int j = c; j < 7; j++)
                A[r][j] -= fct * A[c][j];
        }
    }
    for (int i = 5; i >= 0; i--)
    {
        double s = A[i][6];
        for (int j = i + 1; j < 6; j++)
            s -= A[i][j] * x[j];
        x[i] = s / A[i][i];
    }
    return true;
}

// LDL^T solve of (H + lambda I) x = b, fully unrolled so that everything stays in registers: no
// square roots, one reciprocal per pivot.  false when a pivot is not positive (the caller then
// falls back to LU with partial pivoting)
__device__ __forceinline__ bool chol_solve6(const double *H, double lambda, const double *b, double *x)
{
    double L[6][6], D[6], Dinv[6]; // A = L D L^T, L unit lower triangular
#pragma unroll
    for (int j = 0; j < 6; j++)
    {
        double dj = H[7 * j] + lambda;
#pragma unroll
        for (int k = 0; k < j; k++)
            dj -= L[j][k] * L[j][k] * D[k];
        if (!(dj > 0.0))
            return false;
        D[j] = dj;
        Dinv[j] = 1.0 / dj;
#pragma unroll
        for (int i = j + 1; i < 6; i++)
        {
            double v = H[6 * i + j];
#pragma unroll
            for (int k = 0; k < j; k++)
                v -= L[i][k] * L[j][k] * D[k];
            L[i][j] = v * Dinv[j];
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; k++)
            v -= L[i][k] * y[k];
        y[i] = v;
    }
#pragma unroll
    for (int i = 5; i >= 0; i--)
    {
        double v = y[i] * Dinv[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++)
            v -= L[k][i] * x[k];
        x[i] = v;
    }
    return true;
}

constexpr int kPoseSums = 29;   // 21 (upper H) + 6 (b) + robust chi2 + number of active edges
constexpr int kPoseCluster = 8; // CTAs (SMs) sharing the correspondences of one solve
constexpr int kPoseThreads = 256;
constexpr int kPoseMaxRanks = 16;
constexpr int kPoseCached = 2; // edges per thread kept in registers across the evaluations
constexpr int kPoseAccStride = kPoseThreads + kPoseThreads / 32; // one pad per 32 entries: conflict-free both ways

struct PoseShared
{
    CamState cam;                          // every CTA keeps its own copy of the camera under evaluation
    double acc[kPoseSums][kPoseAccStride]; // per-thread partial sums, one padded row per sum
    double part[32];                       // scratch of the inlier count
    double gather[2][kPoseMaxRanks][32];   // [parity][rank][sum]: every CTA's sums, pushed here by their owners (DSMEM)
    double sums[32];                       // cluster totals (every CTA computes the same values)
    double sys[27];                        // thread 0's LM state between evaluations: the system of the last accepted
    double x[6];                           // state (upper H, b), the step on trial, the camera before it -- kept here
    CamState backup;                       // rather than in registers, which all 256 threads would have to pay for
    int cont;                              // 0 evaluate again, 1 end of pass
};

// One edge of a thread: fixed for the whole solve, so position, measurement, level and the last error
// live in registers (the first kPoseCached edges of a thread; the rest stays in global memory).
struct PoseEdge
{
    double X, Y, Z, zx, zy, e2;
    int level; // -1: no edge
};

// One evaluation at s.cam over the active edges owned by this CTA: errors, robust cost, and the
// linearisation (H, b) -- g2o recomputes both at every accepted state, so an accepted trial's
// evaluation doubles as the next iteration's buildSystem.  Edge i is owned by thread
// (i / blockDim) % nranks == rank, the same one in every evaluation.  Every CTA ends up with the
// cluster totals in s.sums after one cluster barrier; the summation order is fixed (deterministic).
template <class Cluster>
__device__ inline void pose_evaluate(Cluster &cluster, PoseShared &s, PoseEdge (&edge)[kPoseCached], const double *xyz,
                                     const float2 *uv, const uint8_t *level, double *e2, int m, int rank, int nranks,
                                     int parity, long long *dbg = nullptr)
{
#define LVT_PDBG(k)                                                                                                   \
    if (dbg && rank == 0 && threadIdx.x == 0)                                                                         \
    dbg[k] = clock64()
    LVT_PDBG(0);
    const CamState &c = s.cam;
    const double dsqr = kReprojectionTh2, dsqr_reci = 1.0 / kReprojectionTh2; // delta = sqrt(5.991)
    double acc[kPoseSums];
#pragma unroll
    for (int k = 0; k < kPoseSums; k++)
        acc[k] = 0.0;
    const double w0 = c.w2n[0], w1 = c.w2n[1], w2 = c.w2n[2], w3 = c.w2n[3], w4 = c.w2n[4], w5 = c.w2n[5], w6 = c.w2n[6],
                 w7 = c.w2n[7], w8 = c.w2n[8], w9 = c.w2n[9], w10 = c.w2n[10], w11 = c.w2n[11];
    const double t0 = c.t[0], t1 = c.t[1], t2 = c.t[2], fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;

    // returns chi2 of the edge
    auto one_edge = [&](double X, double Y, double Z, double zx, double zy) -> double {
        const double px = fma(w0, X, fma(w1, Y, fma(w2, Z, w3)));
        const double py = fma(w4, X, fma(w5, Y, fma(w6, Z, w7)));
        const double pz = fma(w8, X, fma(w9, Y, fma(w10, Z, w11)));
        // EdgeProjectP2MC::computeError: (K w2n p).head<2>() / z - measurement
        const double ipz = 1.0 / pz;
        const double ex = fma(fma(fx, px, cx * pz), ipz, -zx);
        const double ey = fma(fma(fy, py, cy * pz), ipz, -zy);
        const double chi = fma(ex, ex, ey * ey);
        const double aux = fma(dsqr_reci, chi, 1.0);
        acc[27] = fma(dsqr, log(aux), acc[27]); // RobustKernelCauchy rho[0]
        acc[28] += 1.0;
        const double w = 1.0 / aux; // rho[1]
        // EdgeProjectP2MC::linearizeOplus, camera block
        const double ipz2 = ipz * ipz;
        const double ipz2fx = ipz2 * fx, ipz2fy = ipz2 * fy;
        const double pw0 = X - t0, pw1 = Y - t1, pw2 = Z - t2;
        double J0[6], J1[6];
        J0[0] = (px * w8 - pz * w0) * ipz2fx;
        J0[1] = (px * w9 - pz * w1) * ipz2fx;
        J0[2] = (px * w10 - pz * w2) * ipz2fx;
        J1[0] = (py * w8 - pz * w4) * ipz2fy;
        J1[1] = (py * w9 - pz * w5) * ipz2fy;
        J1[2] = (py * w10 - pz * w6) * ipz2fy;
        // dRd{x,y,z} * (p - t), with dRidx = [0 0 0; 0 0 2; 0 -2 0] etc. applied to R^T
        const double r0 = fma(w0, pw0, fma(w1, pw1, w2 * pw2));
        const double r1 = fma(w4, pw0, fma(w5, pw1, w6 * pw2));
        const double r2 = fma(w8, pw0, fma(w9, pw1, w10 * pw2));
        const double q[3][3] = {{0.0, 2 * r2, -2 * r1}, {-2 * r2, 0.0, 2 * r0}, {2 * r1, -2 * r0, 0.0}};
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            J0[3 + k] = (pz * q[k][0] - px * q[k][2]) * ipz2fx;
            J1[3 + k] = (pz * q[k][1] - py * q[k][2]) * ipz2fy;
        }
        const double g0 = -ex * w, g1 = -ey * w;
        int idx = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
        {
            const double J0w = J0[a] * w, J1w = J1[a] * w;
#pragma unroll
            for (int bcol = a; bcol < 6; bcol++)
            {
                acc[idx] = fma(J0w, J0[bcol], fma(J1w, J1[bcol], acc[idx]));
                idx++;
            }
        }
#pragma unroll
        for (int a = 0; a < 6; a++)
            acc[21 + a] = fma(J0[a], g0, fma(J1[a], g1, acc[21 + a]));
        return chi;
    };
#pragma unroll
    for (int k = 0; k < kPoseCached; k++)
        if (edge[k].level == 0)
            edge[k].e2 = one_edge(edge[k].X, edge[k].Y, edge[k].Z, edge[k].zx, edge[k].zy);
    for (int i = (kPoseCached * nranks + rank) * (int)blockDim.x + (int)threadIdx.x; i < m; i += nranks * (int)blockDim.x)
    {
        if (level[i])
            continue;
        const float2 z = uv[i];
        e2[i] = one_edge(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], (double)z.x, (double)z.y);
    }
    LVT_PDBG(1);
    // CTA: every thread parks its sums in shared memory (row k = sum k), then all 256 threads add
    // them up -- thread t takes row t / 8, columns (t % 8) * 32 .. + 32 in four independent chains,
    // and the 8 threads of a row meet in three shuffle steps.  Cluster: the row leaders push the CTA's
    // sums into every CTA's gather[parity][rank] (DSMEM stores); one cluster barrier; local reads.
    const int slot = threadIdx.x + (threadIdx.x >> 5);
#pragma unroll
    for (int k = 0; k < kPoseSums; k++)
        s.acc[k][slot] = acc[k];
    __syncthreads();
    {
        constexpr int kSlices = kPoseThreads / 32; // 8
        const int k = threadIdx.x / kSlices, w = threadIdx.x % kSlices;
        double v = 0.0;
        if (k < kPoseSums)
        {
            const double *src = &s.acc[k][w * 33];
            double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
            {
                c0 += src[j];
                c1 += src[j + 1];
                c2 += src[j + 2];
                c3 += src[j + 3];
            }
            v = (c0 + c1) + (c2 + c3);
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        // every lane of the row's group holds the row total: lane w serves destination CTA w, w + 8, ...
        // (one DSMEM store per thread instead of nranks stores by the row leader)
        if (k < kPoseSums)
        {
            if (nranks == 1)
            {
                if (w == 0)
                    s.sums[k] = v;
            }
            else
                for (int r = w; r < nranks; r += kSlices)
                    *cluster.map_shared_rank(&s.gather[parity][rank][k], r) = v;
        }
    }
    LVT_PDBG(2);
    if (nranks == 1)
    {
        __syncthreads();
        LVT_PDBG(3);
        LVT_PDBG(4);
        return;
    }
    cluster.sync();
    LVT_PDBG(3);
    if (threadIdx.x < kPoseSums)
    {
        double v = 0.0;
        for (int r = 0; r < nranks; r++)
            v += s.gather[parity][r][threadIdx.x];
        s.sums[threadIdx.x] = v;
    }
    __syncthreads();
    LVT_PDBG(4);
}

// lvt_pnp_solver::compute_pose on a thread-block cluster.  Every thread of every CTA calls it.
// level / e2 / inlier: per-edge scratch (global).  Rank 0 / thread 0 writes *pose_out and
// *n_inliers_out.  Every CTA reads the same cluster totals and its thread 0 runs the same LM step on
// them (bitwise identical results), so one cluster barrier per evaluation is all the CTAs exchange;
// control flow is uniform across the cluster.
template <class Cluster>
__device__ inline void cluster_solve_pose(Cluster &cluster, PoseShared &s, const double *xyz, const float2 *uv, int m,
                                          const PoseD &init, const CamParams &cp, uint8_t *level, double *e2,
                                          uint8_t *inlier, PoseD *pose_out, int *n_inliers_out, long long *dbg = nullptr,
                                          int *n_evals_out = nullptr)
{
    int n_evals = 0; // passes over the correspondences (uniform over the cluster)
#define LVT_SDBG(k)                                                                                                   \
    if (dbg && rank == 0 && threadIdx.x == 0 && (k) < 32)                                                             \
    dbg[k] = clock64()
    const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
    LVT_SDBG(8);
    if (threadIdx.x == 0)
    {
        CamState &c = s.cam; // the same arithmetic in every CTA: identical copies
        c.fx = cp.fx;
        c.fy = cp.fy;
        c.cx = cp.cx;
        c.cy = cp.cy;
        c.t[0] = init.t[0];
        c.t[1] = init.t[1];
        c.t[2] = init.t[2];
        Quat q = init.q; // SE3Quat(r, t): normalizeRotation()
        if (q.w < 0)
        {
            q.w = -q.w;
            q.x = -q.x;
            q.y = -q.y;
            q.z = -q.z;
        }
        c.r = quat_normalized(q);
        cam_refresh(c);
    }
    PoseEdge edge[kPoseCached];
#pragma unroll
    for (int k = 0; k < kPoseCached; k++)
    {
        const int i = (k * nranks + rank) * (int)blockDim.x + (int)threadIdx.x;
        edge[k].level = -1;
        edge[k].e2 = 0.0;
        if (i < m)
        {
            const float2 z = uv[i];
            edge[k].X = xyz[3 * i];
            edge[k].Y = xyz[3 * i + 1];
            edge[k].Z = xyz[3 * i + 2];
            edge[k].zx = (double)z.x;
            edge[k].zy = (double)z.y;
            edge[k].level = 0;
        }
    }
    for (int i = (kPoseCached * nranks + rank) * (int)blockDim.x + (int)threadIdx.x; i < m; i += nranks * (int)blockDim.x)
    {
        level[i] = 0;
        inlier[i] = 1;
        e2[i] = 0.0;
    }
    // every CTA of the cluster must be executing before anybody stores into its shared memory
    // (pose_evaluate pushes the CTA sums through DSMEM)
    if (nranks > 1)
        cluster.sync();
    else
        __syncthreads();

    LVT_SDBG(9);
    // LM state of rank 0 / thread 0 (OptimizationAlgorithmLevenberg::solve)
    double lambda = 0, ni = 2, current_chi = 0, rho = 0;
    int qmax = 0, it = 0;
    const bool boss = threadIdx.x == 0; // in every CTA
    int parity = 0;
    auto load_system = [&]() {
#pragma unroll
        for (int k = 0; k < 27; k++)
            s.sys[k] = s.sums[k];
    };
    // one block-Jacobi preconditioned CG step on the single 6x6 block (LinearSolverPCG):
    // d = (H + lambda I)^-1 b, x = (b.d / d.Ad) d; then SBACam::update
    auto propose = [&]() {
        s.backup = s.cam;
        double H[36], bvec[6], d[6], x[6];
        {
            int idx = 0;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int bcol = a; bcol < 6; bcol++)
                {
                    const double v = s.sys[idx++];
                    H[6 * a + bcol] = v;
                    H[6 * bcol + a] = v;
                }
#pragma unroll
            for (int a = 0; a < 6; a++)
                bvec[a] = s.sys[21 + a];
        }
#pragma unroll
        for (int j = 0; j < 6; j++)
            x[j] = 0.0;
        bool solved = chol_solve6(H, lambda, bvec, d);
        if (!solved)
        {
            double A[36];
            for (int i = 0; i < 36; i++)
                A[i] = H[i];
            for (int j = 0; j < 6; j++)
                A[7 * j] += lambda;
            solved = solve6(A, bvec, d);
        }
        // LinearSolverPCG's single step scales d by alpha = b.d / d.Ad, which is 1 up to rounding for
        // an exact block solve; it is dropped here (1e-16 relative, far below the parity tolerance)
        if (solved)
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                x[i] = d[i];
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
            s.x[i] = x[i];
        cam_update(s.cam, x);
    };

    for (int pass = 0; pass < 2; pass++)
    {
        // errors + linearisation at the starting state of this optimize()
        pose_evaluate(cluster, s, edge, xyz, uv, level, e2, m, rank, nranks, parity);
        parity ^= 1;
        n_evals++;
        LVT_SDBG(9 + n_evals);
        if (boss)
        {
            if (s.sums[28] == 0.0)
                s.cont = 1; // initializeOptimization(0) with an empty active set: optimize() does nothing
            else
            {
                current_chi = s.sums[27];
                load_system();
                double max_diag = 0;
                {
                    // diagonal of H inside the packed upper triangle: 0, 6, 11, 15, 18, 20
                    const int diag[6] = {0, 6, 11, 15, 18, 20};
#pragma unroll
                    for (int j = 0; j < 6; j++)
                        max_diag = fmax(fabs(s.sys[diag[j]]), max_diag);
                }
                lambda = 1e-5 * max_diag; // computeLambdaInit, _tau = 1e-5
                ni = 2;
                it = 0;
                rho = 0;
                qmax = 0;
                propose();
                s.cont = 0;
            }
        }
        while (true)
        {
            __syncthreads();
            if (s.cont != 0)
                break;
            pose_evaluate(cluster, s, edge, xyz, uv, level, e2, m, rank, nranks, parity, dbg); // the trial state
            parity ^= 1;
            n_evals++;
            LVT_SDBG(9 + n_evals);
            if (boss)
            {
                const double temp_chi = s.sums[27];
                double scale = 0;
#pragma unroll
                for (int j = 0; j < 6; j++)
                    scale += s.x[j] * (lambda * s.x[j] + s.sys[21 + j]);
                scale += 1e-3;
                rho = (current_chi - temp_chi) / scale;
                if (rho > 0 && isfinite(temp_chi))
                {
                    const double tt = 2 * rho - 1;
                    double alpha = 1. - tt * tt * tt;
                    alpha = fmin(alpha, 2. / 3.);
                    lambda *= fmax(1. / 3., alpha);
                    ni = 2;
                    current_chi = temp_chi;
                    load_system(); // the system of the next iteration
                }
                else
                {
                    lambda *= ni;
                    ni *= 2;
                    s.cam = s.backup; // the edges keep the rejected trial's error
                }
                qmax++;
                if (rho < 0 && qmax < 10)
                {
                    propose(); // another trial of the same iteration
                    s.cont = 0;
                }
                else if (qmax == 10 || rho == 0 || it == 4)
                    s.cont = 1; // Terminate, or optimize(5) is through
                else
                {
                    it++;
                    rho = 0;
                    qmax = 0;
                    propose();
                    s.cont = 0;
                }
                if (dbg && rank == 0)
                    dbg[5] = clock64();
            }
        }
        // lvt_pnp_solver.cpp:109-116
#pragma unroll
        for (int k = 0; k < kPoseCached; k++)
            if (edge[k].level == 0 && edge[k].e2 > kReprojectionTh2)
                edge[k].level = 1;
        for (int i = (kPoseCached * nranks + rank) * (int)blockDim.x + (int)threadIdx.x; i < m; i += nranks * (int)blockDim.x)
        {
            if (e2[i] > kReprojectionTh2)
            {
                level[i] = 1;
                inlier[i] = 0;
            }
        }
        __syncthreads();
    }
    // the register-resident edges report their marks (and last errors) once
#pragma unroll
    for (int k = 0; k < kPoseCached; k++)
    {
        const int i = (k * nranks + rank) * (int)blockDim.x + (int)threadIdx.x;
        if (i < m)
        {
            level[i] = (uint8_t)edge[k].level;
            inlier[i] = edge[k].level == 0;
            e2[i] = edge[k].e2;
        }
    }
    // inlier count: CTA sums exchanged through DSMEM
    {
        double cnt = 0;
#pragma unroll
        for (int k = 0; k < kPoseCached; k++)
            cnt += edge[k].level == 0;
        for (int i = (kPoseCached * nranks + rank) * (int)blockDim.x + (int)threadIdx.x; i < m; i += nranks * (int)blockDim.x)
            cnt += inlier[i];
        cnt = warp_sum(cnt);
        __syncthreads();
        if ((threadIdx.x & 31) == 0)
            s.part[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            double v = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++)
                v += s.part[w];
            *cluster.map_shared_rank(&s.gather[parity][rank][0], 0) = v;
        }
        cluster.sync();
        if (rank == 0 && threadIdx.x == 0)
        {
            double v = 0;
            for (int r = 0; r < nranks; r++)
                v += s.gather[parity][r][0];
            *n_inliers_out = (int)v;
            if (n_evals_out)
                *n_evals_out = n_evals;
            pose_out->q = s.cam.r;
            pose_out->t[0] = s.cam.t[0];
            pose_out->t[1] = s.cam.t[1];
            pose_out->t[2] = s.cam.t[2];
        }
        cluster.sync(); // nobody leaves while another CTA may still read its shared memory
    }
    LVT_SDBG(31);
#undef LVT_SDBG
}

} // namespace lvtb
