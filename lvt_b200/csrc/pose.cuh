// Motion-only bundle adjustment on a thread-block cluster, header-only device code.
//
// Replaces the g2o graph built and solved per frame at lvt/src/lvt_pnp_solver.cpp:60-128:
// one free SBACam vertex, M fixed points, M EdgeProjectP2MC with a Cauchy kernel
// (delta^2 = 5.991), Levenberg-Marquardt on the single 6x6 block (BlockSolver_6_3 +
// LinearSolverPCG == one preconditioned-CG step == one exact solve), two passes of optimize(5),
// edges with chi2 > 5.991 demoted to level 1 after each pass.
//
// Per evaluation every thread takes its correspondences (kept in registers across the evaluations):
// fp64 projection, residual, robust weight and the 2x6 Jacobian; the 21 + 6 + 1 + 1 partial sums of
// H = sum w J^T J, b = -sum w J^T e, the robust cost and the edge count are reduced inside the warp
// (transposed butterfly), across the warps of a CTA through shared memory, and across the CTAs of
// the cluster by counted st.async pushes into every CTA's gather buffer -- always in the same order
// (deterministic, identical in every CTA).  Warp 0 of EVERY CTA then runs the same 6x6 solve and LM
// accept / reject logic in registers, so nothing but the sums travels (cluster_solve_pose below).
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

// Solve of (H + lambda I) x = b for the 6x6 SPD system of the LM step, H as packed upper triangle
// (rows 0..5: 0-5, 6-10, 11-14, 15-17, 18-19, 20), by 3x3 blocks with closed-form (adjugate) inverses:
//   [A B; B^T C] : x2 = S^-1 (b2 - B^T A^-1 b1),  x1 = A^-1 b1 - (A^-1 B) x2,  S = C - B^T A^-1 B.
// This is the step every thread of the cluster waits for, and its cost is the LENGTH of its dependent
// fp64 chain: two reciprocals and ~40 dependent operations here against six reciprocals and ~100 for an
// unrolled LDL^T.  false when A or S is not positive definite in floating point (the caller falls back
// to LU with partial pivoting).
struct Sym3
{
    double a, b, c, d, e, f; // [[a b c], [b d e], [c e f]]
};
__device__ __forceinline__ bool sym3_adjugate(const Sym3 &m, Sym3 &adj, double &inv_det)
{
    adj.a = fma(m.d, m.f, -(m.e * m.e));
    adj.b = fma(m.c, m.e, -(m.b * m.f));
    adj.c = fma(m.b, m.e, -(m.c * m.d));
    adj.d = fma(m.a, m.f, -(m.c * m.c));
    adj.e = fma(m.b, m.c, -(m.a * m.e));
    adj.f = fma(m.a, m.d, -(m.b * m.b));
    const double det = fma(m.a, adj.a, fma(m.b, adj.b, m.c * adj.c));
    inv_det = 1.0 / det;
    return m.a > 0.0 && adj.f > 0.0 && det > 0.0 && isfinite(inv_det); // leading principal minors
}
__device__ __forceinline__ void sym3_mul(const Sym3 &m, double x, double y, double z, double &ox, double &oy, double &oz)
{
    ox = fma(m.a, x, fma(m.b, y, m.c * z));
    oy = fma(m.b, x, fma(m.d, y, m.e * z));
    oz = fma(m.c, x, fma(m.e, y, m.f * z));
}
__device__ __forceinline__ bool spd_solve6_blocks(const double (&h)[27], double lambda, double (&x)[6])
{
    const Sym3 A{h[0] + lambda, h[1], h[2], h[6] + lambda, h[7], h[11] + lambda};
    const Sym3 C{h[15] + lambda, h[16], h[17], h[18] + lambda, h[19], h[20] + lambda};
    const double B[3][3] = {{h[3], h[4], h[5]}, {h[8], h[9], h[10]}, {h[12], h[13], h[14]}};
    Sym3 adjA;
    double idA;
    if (!sym3_adjugate(A, adjA, idA))
        return false;
    // Y = A^-1 B (columns), u = A^-1 b1
    double Y[3][3], u[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        double y0, y1, y2;
        sym3_mul(adjA, B[0][j], B[1][j], B[2][j], y0, y1, y2);
        Y[0][j] = y0 * idA;
        Y[1][j] = y1 * idA;
        Y[2][j] = y2 * idA;
    }
    {
        double u0, u1, u2;
        sym3_mul(adjA, h[21], h[22], h[23], u0, u1, u2);
        u[0] = u0 * idA;
        u[1] = u1 * idA;
        u[2] = u2 * idA;
    }
    // S = C - B^T Y (symmetric), z = b2 - B^T u
    auto bty = [&](int i, int j) { return fma(B[0][i], Y[0][j], fma(B[1][i], Y[1][j], B[2][i] * Y[2][j])); };
    const Sym3 S{C.a - bty(0, 0), C.b - bty(0, 1), C.c - bty(0, 2), C.d - bty(1, 1), C.e - bty(1, 2), C.f - bty(2, 2)};
    double z[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        z[i] = h[24 + i] - fma(B[0][i], u[0], fma(B[1][i], u[1], B[2][i] * u[2]));
    Sym3 adjS;
    double idS;
    if (!sym3_adjugate(S, adjS, idS))
        return false;
    double x3, x4, x5;
    sym3_mul(adjS, z[0], z[1], z[2], x3, x4, x5);
    x[3] = x3 * idS;
    x[4] = x4 * idS;
    x[5] = x5 * idS;
#pragma unroll
    for (int i = 0; i < 3; i++)
        x[i] = u[i] - fma(Y[i][0], x[3], fma(Y[i][1], x[4], Y[i][2] * x[5]));
    return true;
}

// the same system by LU with partial pivoting: only when the block solve declines (a system that is not
// positive definite in floating point: never on sane input).  Out of line: its local arrays and divisions
// stay out of the instruction stream of the LM step.
static __device__ __noinline__ bool solve6_lu_packed(const double *sys /* 27 */, double lambda, double *d /* 6 */)
{
    double A[36], bvec[6];
    int idx = 0;
    for (int a = 0; a < 6; a++)
        for (int bcol = a; bcol < 6; bcol++)
        {
            const double v = sys[idx++];
            A[6 * a + bcol] = v;
            A[6 * bcol + a] = v;
        }
    for (int j = 0; j < 6; j++)
    {
        A[7 * j] += lambda;
        bvec[j] = sys[21 + j];
    }
    return solve6(A, bvec, d);
}

constexpr int kPoseSums = 29;   // 21 (upper H) + 6 (b) + robust chi2 + number of active edges
constexpr int kPoseCluster = 8; // CTAs (SMs) sharing the correspondences of one solve
constexpr int kPoseThreads = 256;
constexpr int kPoseWarps = kPoseThreads / 32;
constexpr int kPoseMaxRanks = 16;
constexpr int kPoseCached = 2; // edges per thread kept in registers across the evaluations

// LM state between evaluations (OptimizationAlgorithmLevenberg::solve): written by lane 0 of warp 0, read
// by every lane of warp 0 at the next step -- in shared memory so that the 256 threads of the CTA do not
// all pay its registers during the Jacobian pass
struct PoseLM
{
    double sys[27];          // upper H (21) and b (6) of the last accepted state
    double x[6];             // the step on trial
    double bt[3];            // the camera before it (SBACam push/pop): position ...
    Quat br;                 // ... and rotation
    double lambda, ni, current_chi;
    int qmax, it;
};

struct PoseShared
{
    CamState cam;                             // the camera under evaluation; every CTA keeps an identical copy
    double wpart[kPoseWarps][32];             // per-warp totals of the 29 sums
    double gather[2][kPoseMaxRanks][32];      // [parity][rank][sum]: every CTA's totals, pushed here by their owners
                                              // (st.async through distributed shared memory, counted by mbar[parity])
    double sums[32];                          // cluster totals (every CTA computes the same values)
    double part[32];                          // scratch of the inlier count
    PoseLM lm;
    unsigned long long mbar[2];
    int cont;                                 // 0 evaluate again, 1 end of pass
};

// One edge of a thread: fixed for the whole solve, so position, measurement, level and the last error
// live in registers (the first kPoseCached edges of a thread; the rest stays in global memory).
struct PoseEdge
{
    double X, Y, Z, zx, zy, e2;
    int level; // -1: no edge
};

// value -> shared memory of CTA `rank` of the cluster at the offset of `local`, counted (8 bytes) by that
// CTA's mbarrier at the offset of `local_bar`: no barrier instruction, the receiver waits on its own mbarrier
__device__ __forceinline__ void dsmem_store_counted(double *local, unsigned long long *local_bar, unsigned rank, double v)
{
    uint32_t addr, bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr),
                 "l"(__double_as_longlong(v)), "r"(bar)
                 : "memory");
}

// Sums of 32 per-lane values over the warp, all 32 at once: after the call lane l holds the warp total of
// v[l].  Five exchange levels of 16, 8, 4, 2, 1 values (31 shuffles of a double instead of 32 x 5); the
// order of the additions is fixed.
__device__ __forceinline__ double warp_transpose_sum(double (&v)[32], int lane)
{
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1)
    {
        const bool up = (lane & n) != 0;
#pragma unroll
        for (int i = 0; i < n; i++)
        {
            const double send = up ? v[i] : v[i + n];
            const double keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}

// One evaluation at s.cam over the active edges owned by this CTA: errors, robust cost, and the
// linearisation (H, b) -- g2o recomputes both at every accepted state, so an accepted trial's
// evaluation doubles as the next iteration's buildSystem.  Edge i is owned by thread
// (i / blockDim) % nranks == rank, the same one in every evaluation.  Reduction: warp (shuffles) -> CTA
// (shared memory, warp 0) -> cluster: warp 0 pushes the CTA's totals into every CTA's gather[parity]
// and waits on its own mbarrier for the 29 x nranks values addressed to it.  On return the lanes of
// WARP 0 hold the cluster totals (lane k: sum k; s.sums has them too, valid for warp 0); the other
// warps return at once.  The summation order is fixed (deterministic, identical in every CTA).
template <class Cluster>
__device__ inline double pose_evaluate(Cluster &cluster, PoseShared &s, PoseEdge (&edge)[kPoseCached], const double *xyz,
                                       const float2 *uv, const uint8_t *level, double *e2, int m, int rank, int nranks,
                                       int n_eval, long long *dbg = nullptr)
{
#define LVT_PDBG(k)                                                                                                   \
    if (dbg && rank == 0 && threadIdx.x == 0)                                                                         \
    dbg[k] = clock64()
    LVT_PDBG(0);
    const CamState &c = s.cam;
    const double dsqr = kReprojectionTh2, dsqr_reci = 1.0 / kReprojectionTh2; // delta = sqrt(5.991)
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; k++)
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int parity = n_eval & 1;
    s.wpart[warp][lane] = warp_transpose_sum(acc, lane);
    __syncthreads();
    LVT_PDBG(2);
    if (warp != 0)
        return 0.0;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kPoseWarps; w++)
        v += s.wpart[w][lane];
    if (lane < kPoseSums)
        for (int r = 0; r < nranks; r++)
            dsmem_store_counted(&s.gather[parity][rank][lane], &s.mbar[parity], (unsigned)r, v);
    // use (n_eval >> 1) of this parity's barrier: armed right here for the values of THIS evaluation (the
    // counted stores of faster CTAs may have arrived already: the transaction count just goes negative first)
    if (lane == 0)
        mbar_arrive_expect_tx(reinterpret_cast<uint64_t *>(&s.mbar[parity]), (uint32_t)(nranks * kPoseSums * sizeof(double)));
    mbar_wait(reinterpret_cast<uint64_t *>(&s.mbar[parity]), (uint32_t)((n_eval >> 1) & 1));
    LVT_PDBG(3);
    double tot = 0.0;
    for (int r = 0; r < nranks; r++)
        tot += s.gather[parity][r][lane];
    s.sums[lane] = tot;
    __syncwarp();
    LVT_PDBG(4);
    return tot;
}

// lvt_pnp_solver::compute_pose on a thread-block cluster.  Every thread of every CTA calls it.
// level / e2 / inlier: per-edge scratch (global).  Rank 0 / thread 0 writes *pose_out and
// *n_inliers_out.  Every CTA receives the same cluster totals and its warp 0 runs the same LM step on
// them (bitwise identical results), so the values pushed through distributed shared memory are all
// the CTAs exchange -- no cluster barrier inside the iteration; control flow is uniform across the cluster.
//
// The LM step runs on all lanes of warp 0 redundantly, entirely in registers (the 27 sums arrive by
// shared-memory broadcast, one store of the new camera at the end): the dependent chain through the 6x6
// factorisation and the quaternion update is what the other 255 threads wait for.
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
    const int lane = threadIdx.x & 31;
    const bool lm_warp = threadIdx.x < 32;
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
        mbar_init(reinterpret_cast<uint64_t *>(&s.mbar[0]), 1);
        mbar_init(reinterpret_cast<uint64_t *>(&s.mbar[1]), 1);
        mbar_fence_init();
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
    // every CTA of the cluster must be executing, with its mbarriers initialised, before anybody stores
    // into its shared memory (pose_evaluate pushes the CTA sums through DSMEM)
    cluster.sync();
    LVT_SDBG(9);

    // ---- the LM step (warp 0, all lanes redundantly, registers only) --------------------------------
    // propose: one block-Jacobi preconditioned CG step on the single 6x6 block (LinearSolverPCG):
    // d = (H + lambda I)^-1 b, x = (b.d / d.Ad) d; then SBACam::update.  sys: packed upper H + b.
    auto propose = [&](const double (&sys)[27], double lambda, const double (&t_in)[3], const Quat &r_in) {
        double d[6];
        if (!spd_solve6_blocks(sys, lambda, d))
        {
            double tmp_sys[27], tmp_d[6];
            for (int k = 0; k < 27; k++)
                tmp_sys[k] = sys[k];
            const bool ok = solve6_lu_packed(tmp_sys, lambda, tmp_d);
            for (int k = 0; k < 6; k++)
                d[k] = ok ? tmp_d[k] : 0.0;
        }
        // LinearSolverPCG's single step scales d by alpha = b.d / d.Ad, which is 1 up to rounding for
        // an exact block solve; it is dropped here (1e-16 relative, far below the parity tolerance)
        // SBACam::update: t += dt ; r = normalize(r * (dv, sqrt(1 - |dv|^2))) ; setTransform.  Explicit
        // fused multiply-adds and one reciprocal square root: this chain is what the whole cluster waits for.
        const double t0 = t_in[0] + d[0], t1 = t_in[1] + d[1], t2 = t_in[2] + d[2];
        const double vx = d[3], vy = d[4], vz = d[5];
        const double vw = sqrt(1.0 - fma(vx, vx, fma(vy, vy, vz * vz)));
        Quat q; // r_in * (vw, vx, vy, vz)
        q.w = fma(r_in.w, vw, -fma(r_in.x, vx, fma(r_in.y, vy, r_in.z * vz)));
        q.x = fma(r_in.w, vx, fma(r_in.x, vw, fma(r_in.y, vz, -(r_in.z * vy))));
        q.y = fma(r_in.w, vy, fma(r_in.y, vw, fma(r_in.z, vx, -(r_in.x * vz))));
        q.z = fma(r_in.w, vz, fma(r_in.z, vw, fma(r_in.x, vy, -(r_in.y * vx))));
        const double inv_n = rsqrt(fma(q.w, q.w, fma(q.x, q.x, fma(q.y, q.y, q.z * q.z))));
        q.w *= inv_n;
        q.x *= inv_n;
        q.y *= inv_n;
        q.z *= inv_n;
        // SBACam::setTransform: w2n = [R^T | -R^T t]
        double w2n[12];
        {
            const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
            const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
            const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
            const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
            const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx,
                                 txz - twy,       tyz + twx, 1 - (txx + tyy)};
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
                const double a = R[i], b = R[3 + i], c3 = R[6 + i];
                w2n[4 * i + 0] = a;
                w2n[4 * i + 1] = b;
                w2n[4 * i + 2] = c3;
                w2n[4 * i + 3] = -fma(a, t0, fma(b, t1, c3 * t2));
            }
        }
        if (lane == 0)
        {
            s.lm.bt[0] = t_in[0];
            s.lm.bt[1] = t_in[1];
            s.lm.bt[2] = t_in[2];
            s.lm.br = r_in;
#pragma unroll
            for (int i = 0; i < 6; i++)
                s.lm.x[i] = d[i];
            s.cam.t[0] = t0;
            s.cam.t[1] = t1;
            s.cam.t[2] = t2;
            s.cam.r = q;
#pragma unroll
            for (int i = 0; i < 12; i++)
                s.cam.w2n[i] = w2n[i];
        }
    };

    for (int pass = 0; pass < 2; pass++)
    {
        // One loop body for both kinds of evaluation (one copy of the code: the loop is bound by instruction
        // fetch as much as by arithmetic): `first` = errors + linearisation at the starting state of this
        // optimize(); afterwards every evaluation is an LM trial.
        bool first = true;
        while (true)
        {
            pose_evaluate(cluster, s, edge, xyz, uv, level, e2, m, rank, nranks, n_evals, dbg);
            n_evals++;
#ifdef LVT_POSE_EXPERIMENT
            if (dbg) // experiment: the marginal cost of one more evaluation at the same state
            {
                LVT_SDBG(9 + n_evals);
                __syncthreads();
                pose_evaluate(cluster, s, edge, xyz, uv, level, e2, m, rank, nranks, n_evals, dbg);
                n_evals++;
            }
#endif
            LVT_SDBG(9 + n_evals);
            if (lm_warp)
            {
                long long lm_t0 = 0;
                if (dbg)
                {
                    lm_t0 = clock64();
                    asm volatile("" ::: "memory");
                }
                // state of the iteration (uniform over the warp)
                double lambda = s.lm.lambda, ni = s.lm.ni, current_chi = s.lm.current_chi;
                int qmax = s.lm.qmax, it = s.lm.it;
                const double temp_chi = s.sums[27];
                double rho = 0;
                bool accepted, step;
                int cont;
                if (first)
                {
                    // computeLambdaInit (_tau = 1e-5) on the diagonal of H (packed upper: 0, 6, 11, 15, 18, 20)
                    const double max_diag = fmax(fmax(fmax(fabs(s.sums[0]), fabs(s.sums[6])), fmax(fabs(s.sums[11]), fabs(s.sums[15]))),
                                                 fmax(fabs(s.sums[18]), fabs(s.sums[20])));
                    lambda = 1e-5 * max_diag;
                    ni = 2;
                    current_chi = temp_chi;
                    it = 0;
                    qmax = 0;
                    accepted = true; // the starting state is the state to linearise at
                    // initializeOptimization(0) with an empty active set: optimize() does nothing
                    step = s.sums[28] != 0.0;
                    cont = step ? 0 : 1;
                }
                else
                {
                    double scale = 0;
#pragma unroll
                    for (int j = 0; j < 6; j++)
                    {
                        const double xj = s.lm.x[j];
                        scale += xj * (lambda * xj + s.lm.sys[21 + j]);
                    }
                    scale += 1e-3;
                    rho = (current_chi - temp_chi) / scale;
                    accepted = rho > 0 && isfinite(temp_chi);
                    if (accepted)
                    {
                        const double tt = 2 * rho - 1;
                        double alpha = 1. - tt * tt * tt;
                        alpha = fmin(alpha, 2. / 3.);
                        lambda *= fmax(1. / 3., alpha);
                        ni = 2;
                        current_chi = temp_chi;
                    }
                    else
                    {
                        lambda *= ni;
                        ni *= 2;
                    }
                    qmax++;
                    if (rho < 0 && qmax < 10)
                        step = true, cont = 0; // another trial of the same iteration
                    else if (qmax == 10 || rho == 0 || it == 4)
                        step = false, cont = 1; // Terminate, or optimize(5) is through
                    else
                    {
                        it++;
                        qmax = 0;
                        step = true, cont = 0;
                    }
                }
                // accepted: the state just evaluated, with its own linearisation (the system of the next
                // iteration); rejected: back to the state before the trial (pop), the edges keep the trial's errors
                double sys[27], t_in[3];
                Quat r_in;
                {
                    const double *sys_src = accepted ? s.sums : s.lm.sys;
                    const double *t_src = accepted ? s.cam.t : s.lm.bt;
                    const Quat *r_src = accepted ? &s.cam.r : &s.lm.br;
#pragma unroll
                    for (int k = 0; k < 27; k++)
                        sys[k] = sys_src[k];
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        t_in[k] = t_src[k];
                    r_in = *r_src;
                }
                __syncwarp(); // every lane has read the shared state before lane 0 replaces it
                if (step)
                    propose(sys, lambda, t_in, r_in);
                else if (!accepted && lane == 0)
                {
                    CamState oc;
                    oc.t[0] = t_in[0], oc.t[1] = t_in[1], oc.t[2] = t_in[2];
                    oc.r = r_in;
                    cam_refresh(oc);
                    s.cam.t[0] = oc.t[0], s.cam.t[1] = oc.t[1], s.cam.t[2] = oc.t[2];
                    s.cam.r = oc.r;
                    for (int i = 0; i < 12; i++)
                        s.cam.w2n[i] = oc.w2n[i];
                }
                if (accepted && lane < 27)
                    s.lm.sys[lane] = sys[lane];
                if (lane == 0)
                {
                    s.lm.lambda = lambda;
                    s.lm.ni = ni;
                    s.lm.current_chi = current_chi;
                    s.lm.qmax = qmax;
                    s.lm.it = it;
                    s.cont = cont;
                }
                if (dbg)
                {
                    asm volatile("" ::: "memory"); // the stores above are issued before the clock is read
                    __syncwarp();
                    const long long lm_t1 = clock64();
                    if (rank == 0 && lane == 0)
                    {
                        dbg[5] = lm_t1;
                        dbg[6] = lm_t1 - lm_t0;
                    }
                }
            }
            first = false;
            __syncthreads();
            LVT_SDBG(7);
            if (s.cont != 0)
                break;
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
            *cluster.map_shared_rank(&s.part[16 + rank], 0) = v;
        }
        cluster.sync();
        if (rank == 0 && threadIdx.x == 0)
        {
            double v = 0;
            for (int r = 0; r < nranks; r++)
                v += s.part[16 + r];
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
