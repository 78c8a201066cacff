/*
 * CPU ORACLE (test infrastructure) -- motion-only bundle adjustment.
 *
 * Restates lvt/src/lvt_pnp_solver.cpp:42-128: one free g2o::VertexCam, M fixed points, M
 * g2o::EdgeProjectP2MC with RobustKernelCauchy(delta = sqrt(5.991)), Levenberg-Marquardt over
 * BlockSolver_6_3 + LinearSolverPCG, two passes of optimize(5) with chi2 > 5.991 demotion.
 *
 * g2o (tag 20170730_git, README.md:13) is not in /root/reference nor in this image; its
 * published algorithm is restated from: types/sba/sbacam.h (SBACam::update, setTransform,
 * setProjection, setDr), types/sba/types_sba.{h,cpp} (EdgeProjectP2MC::computeError,
 * ::linearizeOplus), core/base_binary_edge.hpp (constructQuadraticForm),
 * core/robust_kernel_impl.cpp (RobustKernelCauchy::robustify),
 * core/optimization_algorithm_levenberg.cpp (solve, computeLambdaInit, computeScale),
 * core/sparse_optimizer.cpp (optimize), solvers/pcg/linear_solver_pcg.hpp.
 * PARITY: unpinned against g2o itself; pinned by finite-difference and convergence tests.
 */
#include "lvto.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>

namespace lvto
{

namespace
{

struct Cam
{
    Vec3 t;
    Quat r;
    double w2n[3][4];
    double w2i[3][4];
    double dR[3][3][3]; /* dRdx, dRdy, dRdz */
    double fx, fy, cx, cy;

    /* SBACam::setTransform / setProjection / setDr */
    void refresh()
    {
        const Mat3 R = qmat(r);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                w2n[i][j] = R.m[j][i];
        for (int i = 0; i < 3; i++)
            w2n[i][3] = -(w2n[i][0] * t.x + w2n[i][1] * t.y + w2n[i][2] * t.z);
        for (int c = 0; c < 4; c++)
        {
            w2i[0][c] = fx * w2n[0][c] + cx * w2n[2][c];
            w2i[1][c] = fy * w2n[1][c] + cy * w2n[2][c];
            w2i[2][c] = w2n[2][c];
        }
        for (int c = 0; c < 3; c++)
        {
            /* dRidx = [0 0 0; 0 0 2; 0 -2 0] */
            dR[0][0][c] = 0;
            dR[0][1][c] = 2 * w2n[2][c];
            dR[0][2][c] = -2 * w2n[1][c];
            /* dRidy = [0 0 -2; 0 0 0; 2 0 0] */
            dR[1][0][c] = -2 * w2n[2][c];
            dR[1][1][c] = 0;
            dR[1][2][c] = 2 * w2n[0][c];
            /* dRidz = [0 2 0; -2 0 0; 0 0 0] */
            dR[2][0][c] = 2 * w2n[1][c];
            dR[2][1][c] = -2 * w2n[0][c];
            dR[2][2][c] = 0;
        }
    }
    /* SBACam::update */
    void update(const double u[6])
    {
        t.x += u[0];
        t.y += u[1];
        t.z += u[2];
        Quat qr;
        qr.x = u[3];
        qr.y = u[4];
        qr.z = u[5];
        qr.w = std::sqrt(1.0 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]));
        r = qnormalized(qmul(r, qr));
        refresh();
    }
};

/* solve A x = b (6x6) by LU with partial pivoting; false if singular */
bool solve6(const double Ain[6][6], const double bin[6], double x[6])
{
    double A[6][7];
    for (int i = 0; i < 6; i++)
    {
        for (int j = 0; j < 6; j++)
            A[i][j] = Ain[i][j];
        A[i][6] = bin[i];
    }
    for (int c = 0; c < 6; c++)
    {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (std::fabs(A[r][c]) > std::fabs(A[piv][c]))
                piv = r;
        if (A[piv][c] == 0.0)
            return false;
        if (piv != c)
            for (int j = 0; j < 7; j++)
                std::swap(A[piv][j], A[c][j]);
        for (int r = c + 1; r < 6; r++)
        {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j < 7; j++)
                A[r][j] -= f * A[c][j];
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

} // namespace

Pose solve_pose(const lvt_params_c &prm, const Pose &init, const std::vector<Vec3> &pts, const std::vector<float> &uv,
                std::vector<uint8_t> *inlier_marks)
{
    const int M = (int)pts.size();
    const double TH2 = 5.991; /* LVT_REPROJECTION_TH2 */
    const double delta = std::sqrt(TH2);
    const double dsqr = delta * delta, dsqr_reci = 1.0 / dsqr;

    Cam cam;
    cam.fx = prm.fx;
    cam.fy = prm.fy;
    cam.cx = prm.cx;
    cam.cy = prm.cy;
    cam.t = init.p;
    cam.r = init.q;
    /* SE3Quat(r, t) -> normalizeRotation() */
    if (cam.r.w < 0)
    {
        cam.r.w = -cam.r.w;
        cam.r.x = -cam.r.x;
        cam.r.y = -cam.r.y;
        cam.r.z = -cam.r.z;
    }
    cam.r = qnormalized(cam.r);
    cam.refresh();

    std::vector<int> level(M, 0);
    std::vector<uint8_t> marks(M, 1);
    std::vector<double> ex(M, 0.0), ey(M, 0.0); /* _error of every edge, as last computed */

    /* computeActiveErrors + activeRobustChi2 */
    auto compute_errors = [&]() -> double {
        double chi = 0;
        for (int i = 0; i < M; i++)
        {
            if (level[i] != 0)
                continue;
            const Vec3 &p = pts[i];
            const double p0 = cam.w2i[0][0] * p.x + cam.w2i[0][1] * p.y + cam.w2i[0][2] * p.z + cam.w2i[0][3];
            const double p1 = cam.w2i[1][0] * p.x + cam.w2i[1][1] * p.y + cam.w2i[1][2] * p.z + cam.w2i[1][3];
            const double p2 = cam.w2i[2][0] * p.x + cam.w2i[2][1] * p.y + cam.w2i[2][2] * p.z + cam.w2i[2][3];
            ex[i] = p0 / p2 - (double)uv[2 * i];
            ey[i] = p1 / p2 - (double)uv[2 * i + 1];
            const double e2 = ex[i] * ex[i] + ey[i] * ey[i];
            chi += dsqr * std::log(dsqr_reci * e2 + 1.0);
        }
        return chi;
    };

    /* buildSystem: linearizeOplus + constructQuadraticForm over the active edges */
    auto build_system = [&](double H[6][6], double b[6]) {
        for (int i = 0; i < 6; i++)
        {
            b[i] = 0;
            for (int j = 0; j < 6; j++)
                H[i][j] = 0;
        }
        for (int i = 0; i < M; i++)
        {
            if (level[i] != 0)
                continue;
            const Vec3 &p = pts[i];
            const double px = cam.w2n[0][0] * p.x + cam.w2n[0][1] * p.y + cam.w2n[0][2] * p.z + cam.w2n[0][3];
            const double py = cam.w2n[1][0] * p.x + cam.w2n[1][1] * p.y + cam.w2n[1][2] * p.z + cam.w2n[1][3];
            const double pz = cam.w2n[2][0] * p.x + cam.w2n[2][1] * p.y + cam.w2n[2][2] * p.z + cam.w2n[2][3];
            const double ipz2 = 1.0 / (pz * pz);
            const double ipz2fx = ipz2 * cam.fx, ipz2fy = ipz2 * cam.fy;
            const double pwt[3] = {p.x - cam.t.x, p.y - cam.t.y, p.z - cam.t.z};
            double J[2][6];
            for (int k = 0; k < 3; k++)
            {
                /* d/dt_k : dp = -w2n.col(k) */
                const double d0 = -cam.w2n[0][k], d1 = -cam.w2n[1][k], d2 = -cam.w2n[2][k];
                J[0][k] = (pz * d0 - px * d2) * ipz2fx;
                J[1][k] = (pz * d1 - py * d2) * ipz2fy;
                /* d/dq_k : dp = dRd{k} * (pw - t) */
                const double q0 = cam.dR[k][0][0] * pwt[0] + cam.dR[k][0][1] * pwt[1] + cam.dR[k][0][2] * pwt[2];
                const double q1 = cam.dR[k][1][0] * pwt[0] + cam.dR[k][1][1] * pwt[1] + cam.dR[k][1][2] * pwt[2];
                const double q2 = cam.dR[k][2][0] * pwt[0] + cam.dR[k][2][1] * pwt[1] + cam.dR[k][2][2] * pwt[2];
                J[0][3 + k] = (pz * q0 - px * q2) * ipz2fx;
                J[1][3 + k] = (pz * q1 - py * q2) * ipz2fy;
            }
            const double e2 = ex[i] * ex[i] + ey[i] * ey[i];
            const double rho1 = 1.0 / (dsqr_reci * e2 + 1.0);
            const double r0 = -ex[i] * rho1, r1 = -ey[i] * rho1; /* omega_r * rho[1], omega = I */
            for (int a = 0; a < 6; a++)
            {
                b[a] += J[0][a] * r0 + J[1][a] * r1;
                for (int c = 0; c < 6; c++)
                    H[a][c] += (J[0][a] * J[0][c] + J[1][a] * J[1][c]) * rho1;
            }
        }
    };

    static const bool trace = std::getenv("LVTO_LM_TRACE") != nullptr; /* one line per solve: A accept, r reject, | pass */
    for (int pass = 0; pass < 2 /* N_PASSES */; pass++)
    {
        int n_active = 0;
        for (int i = 0; i < M; i++)
            n_active += (level[i] == 0);
        if (n_active > 0)
        {
            double lambda = 0, ni = 2;
            for (int it = 0; it < 5; it++)
            {
                double current_chi = compute_errors();
                double H[6][6], b[6];
                build_system(H, b);
                if (it == 0)
                {
                    double max_diag = 0;
                    for (int j = 0; j < 6; j++)
                        max_diag = std::max(std::fabs(H[j][j]), max_diag);
                    lambda = 1e-5 * max_diag; /* _tau */
                    ni = 2;
                }
                double rho = 0;
                int qmax = 0;
                do
                {
                    const Cam backup = cam; /* push() */
                    double A[6][6];
                    for (int i = 0; i < 6; i++)
                        for (int j = 0; j < 6; j++)
                            A[i][j] = H[i][j] + (i == j ? lambda : 0.0);
                    /* LinearSolverPCG on a single 6x6 block, block-Jacobi preconditioner = A^-1:
                     * d = A^-1 b ; a = (b.d)/(d.Ad) ; x = a d  (second iteration stops on dn <= d0) */
                    double d[6], x[6] = {0, 0, 0, 0, 0, 0};
                    if (solve6(A, b, d))
                    {
                        double dn = 0, dq = 0;
                        for (int i = 0; i < 6; i++)
                        {
                            dn += b[i] * d[i];
                            double Ad = 0;
                            for (int j = 0; j < 6; j++)
                                Ad += A[i][j] * d[j];
                            dq += d[i] * Ad;
                        }
                        if (!(dn <= 1e-6 * dn))
                        {
                            const double a = dn / dq;
                            for (int i = 0; i < 6; i++)
                                x[i] = a * d[i];
                        }
                    }
                    cam.update(x);
                    double temp_chi = compute_errors();
                    double scale = 0;
                    for (int j = 0; j < 6; j++)
                        scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho = (current_chi - temp_chi) / scale;
                    if (rho > 0 && std::isfinite(temp_chi))
                    {
                        double alpha = 1. - std::pow((2 * rho - 1), 3);
                        alpha = (std::min)(alpha, 2. / 3.);
                        const double scale_factor = (std::max)(1. / 3., alpha);
                        lambda *= scale_factor;
                        ni = 2;
                        current_chi = temp_chi;
                        if (trace)
                            std::fputc('A', stderr);
                    }
                    else
                    {
                        lambda *= ni;
                        ni *= 2;
                        cam = backup; /* pop(); the edges keep the rejected trial's _error */
                        if (trace)
                            std::fputc('r', stderr);
                    }
                    qmax++;
                } while (rho < 0 && qmax < 10);
                if (qmax == 10 || rho == 0)
                    break; /* Terminate */
            }
        }
        if (trace)
            std::fputc(pass == 0 ? '|' : '\n', stderr);
        for (int k = 0; k < M; k++)
        {
            if (ex[k] * ex[k] + ey[k] * ey[k] > TH2)
            {
                level[k] = 1;
                marks[k] = 0;
            }
        }
    }

    if (inlier_marks)
        *inlier_marks = marks;
    Pose out;
    out.p = cam.t;
    out.q = cam.r;
    return out;
}

} // namespace lvto
