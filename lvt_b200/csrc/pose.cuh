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

constexpr int kPoseSums = 28; // 21 (upper H) + 6 (b) + 1 (robust chi2)

struct PoseShared
{
    CamState cam;
    double partial[32][kPoseSums]; // per-warp partial sums
    double sums[kPoseSums];
    int cont;
};

// One evaluation at s.cam over the active edges.  with_system: also accumulate H and b.
// Writes e2[i] (the edge's chi2 as last computed) and leaves the totals in s.sums.
template <bool kWithSystem>
__device__ inline void pose_evaluate(PoseShared &s, const double *xyz, const float2 *uv, const uint8_t *level, double *e2,
                                     int m)
{
    const CamState &c = s.cam;
    const double dsqr = kReprojectionTh2, dsqr_reci = 1.0 / kReprojectionTh2; // delta = sqrt(5.991)
    double acc[kWithSystem ? kPoseSums : 1];
#pragma unroll
    for (int k = 0; k < (kWithSystem ? kPoseSums : 1); k++)
        acc[k] = 0.0;

    for (int i = threadIdx.x; i < m; i += blockDim.x)
    {
        if (level[i])
            continue;
        const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        const double px = c.w2n[0] * X + c.w2n[1] * Y + c.w2n[2] * Z + c.w2n[3];
        const double py = c.w2n[4] * X + c.w2n[5] * Y + c.w2n[6] * Z + c.w2n[7];
        const double pz = c.w2n[8] * X + c.w2n[9] * Y + c.w2n[10] * Z + c.w2n[11];
        // EdgeProjectP2MC::computeError: (K w2n p).head<2>() / z - measurement
        const float2 z = uv[i];
        const double ex = (c.fx * px + c.cx * pz) / pz - (double)z.x;
        const double ey = (c.fy * py + c.cy * pz) / pz - (double)z.y;
        const double chi = ex * ex + ey * ey;
        e2[i] = chi;
        const double aux = dsqr_reci * chi + 1.0;
        acc[kWithSystem ? 27 : 0] += dsqr * log(aux); // RobustKernelCauchy rho[0]
        if (kWithSystem)
        {
            const double w = 1.0 / aux; // rho[1]
            // EdgeProjectP2MC::linearizeOplus, camera block
            const double ipz2 = 1.0 / (pz * pz);
            const double ipz2fx = ipz2 * c.fx, ipz2fy = ipz2 * c.fy;
            const double pw[3] = {X - c.t[0], Y - c.t[1], Z - c.t[2]};
            double J0[6], J1[6];
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                const double d0 = -c.w2n[k], d1 = -c.w2n[4 + k], d2 = -c.w2n[8 + k];
                J0[k] = (pz * d0 - px * d2) * ipz2fx;
                J1[k] = (pz * d1 - py * d2) * ipz2fy;
            }
            // dRd{x,y,z} * (p - t), with dRidx = [0 0 0; 0 0 2; 0 -2 0] etc. applied to R^T
            const double r0 = c.w2n[0] * pw[0] + c.w2n[1] * pw[1] + c.w2n[2] * pw[2];
            const double r1 = c.w2n[4] * pw[0] + c.w2n[5] * pw[1] + c.w2n[6] * pw[2];
            const double r2 = c.w2n[8] * pw[0] + c.w2n[9] * pw[1] + c.w2n[10] * pw[2];
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
#pragma unroll
                for (int bcol = a; bcol < 6; bcol++)
                    acc[idx++] += (J0[a] * J0[bcol] + J1[a] * J1[bcol]) * w;
            }
#pragma unroll
            for (int a = 0; a < 6; a++)
                acc[21 + a] += J0[a] * g0 + J1[a] * g1;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (kWithSystem)
    {
#pragma unroll
        for (int k = 0; k < kPoseSums; k++)
        {
            const double v = warp_sum(acc[k]);
            if (lane == 0)
                s.partial[warp][k] = v;
        }
    }
    else
    {
        const double v = warp_sum(acc[0]);
        if (lane == 0)
            s.partial[warp][27] = v;
    }
    __syncthreads();
    if (threadIdx.x < kPoseSums && (kWithSystem || threadIdx.x == 27))
    {
        double v = 0.0;
        for (int w = 0; w < nwarps; w++)
            v += s.partial[w][threadIdx.x];
        s.sums[threadIdx.x] = v;
    }
    __syncthreads();
}

// lvt_pnp_solver::compute_pose.  All threads of the block call it.  level / e2: per-edge scratch
// (global).  On return (thread-uniform) pose_out holds the optimised pose, inlier[i] the marks.
// Returns the inlier count.  blockDim.x must be a multiple of 32, <= 1024.
__device__ inline int block_solve_pose(PoseShared &s, const double *xyz, const float2 *uv, int m, const PoseD &init,
                                       const CamParams &cp, uint8_t *level, double *e2, uint8_t *inlier, PoseD *pose_out,
                                       int *s_scan)
{
    if (threadIdx.x == 0)
    {
        CamState &c = s.cam;
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
    for (int i = threadIdx.x; i < m; i += blockDim.x)
    {
        level[i] = 0;
        inlier[i] = 1;
        e2[i] = 0.0;
    }
    __syncthreads();

    // thread-0 LM state
    double lambda = 0, ni = 2, current_chi = 0, rho = 0;
    double H[36], bvec[6], x[6];
    CamState backup;
    int qmax = 0;

    for (int pass = 0; pass < 2; pass++)
    {
        // any active edge?  (initializeOptimization(0) with an empty active set does nothing)
        int mine = 0;
        for (int i = threadIdx.x; i < m; i += blockDim.x)
            mine += (level[i] == 0);
        int n_active;
        block_exclusive_scan(mine, s_scan, &n_active);
        if (n_active > 0)
        {
            for (int it = 0; it < 5; it++)
            {
                pose_evaluate<true>(s, xyz, uv, level, e2, m);
                if (threadIdx.x == 0)
                {
                    current_chi = s.sums[27];
                    int idx = 0;
                    for (int a = 0; a < 6; a++)
                        for (int bcol = a; bcol < 6; bcol++)
                        {
                            H[6 * a + bcol] = s.sums[idx];
                            H[6 * bcol + a] = s.sums[idx];
                            idx++;
                        }
                    for (int a = 0; a < 6; a++)
                        bvec[a] = s.sums[21 + a];
                    if (it == 0)
                    {
                        double max_diag = 0;
                        for (int j = 0; j < 6; j++)
                            max_diag = fmax(fabs(H[7 * j]), max_diag);
                        lambda = 1e-5 * max_diag; // computeLambdaInit, _tau = 1e-5
                        ni = 2;
                    }
                    rho = 0;
                    qmax = 0;
                }
                bool stop_iterations = false;
                while (true) // trials of OptimizationAlgorithmLevenberg::solve
                {
                    if (threadIdx.x == 0)
                    {
                        backup = s.cam;
                        double A[36];
                        for (int i = 0; i < 36; i++)
                            A[i] = H[i];
                        for (int j = 0; j < 6; j++)
                            A[7 * j] += lambda;
                        // one block-Jacobi preconditioned CG step on a single block
                        double d[6];
                        for (int j = 0; j < 6; j++)
                            x[j] = 0.0;
                        if (solve6(A, bvec, d))
                        {
                            double dn = 0, dq = 0;
                            for (int i = 0; i < 6; i++)
                            {
                                dn += bvec[i] * d[i];
                                double Ad = 0;
                                for (int j = 0; j < 6; j++)
                                    Ad += A[6 * i + j] * d[j];
                                dq += d[i] * Ad;
                            }
                            if (!(dn <= 1e-6 * dn))
                            {
                                const double alpha = dn / dq;
                                for (int i = 0; i < 6; i++)
                                    x[i] = alpha * d[i];
                            }
                        }
                        cam_update(s.cam, x);
                    }
                    __syncthreads();
                    pose_evaluate<false>(s, xyz, uv, level, e2, m);
                    if (threadIdx.x == 0)
                    {
                        const double temp_chi = s.sums[27];
                        double scale = 0;
                        for (int j = 0; j < 6; j++)
                            scale += x[j] * (lambda * x[j] + bvec[j]);
                        scale += 1e-3;
                        rho = (current_chi - temp_chi) / scale;
                        if (rho > 0 && isfinite(temp_chi))
                        {
                            double alpha = 1. - pow((2 * rho - 1), 3);
                            alpha = fmin(alpha, 2. / 3.);
                            lambda *= fmax(1. / 3., alpha);
                            ni = 2;
                            current_chi = temp_chi;
                        }
                        else
                        {
                            lambda *= ni;
                            ni *= 2;
                            s.cam = backup; // the edges keep the rejected trial's error
                        }
                        qmax++;
                        const bool again = (rho < 0 && qmax < 10);
                        // 0: next iteration, 1: another trial, 2: terminate this optimize()
                        s.cont = again ? 1 : ((qmax == 10 || rho == 0) ? 2 : 0);
                    }
                    __syncthreads();
                    const int cont = s.cont;
                    if (cont == 1)
                        continue;
                    stop_iterations = (cont == 2);
                    break;
                }
                if (stop_iterations)
                    break;
            }
        }
        // lvt_pnp_solver.cpp:109-116
        for (int i = threadIdx.x; i < m; i += blockDim.x)
        {
            if (e2[i] > kReprojectionTh2)
            {
                level[i] = 1;
                inlier[i] = 0;
            }
        }
        __syncthreads();
    }
    int mine = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        mine += inlier[i];
    int n_inliers;
    block_exclusive_scan(mine, s_scan, &n_inliers);
    if (threadIdx.x == 0)
    {
        pose_out->q = s.cam.r;
        pose_out->t[0] = s.cam.t[0];
        pose_out->t[1] = s.cam.t[1];
        pose_out->t[2] = s.cam.t[2];
    }
    __syncthreads();
    return n_inliers;
}

} // namespace lvtb
