"""Independent cross-checks (CPU, no GPU) of the oracle's numerics whose third-party originals are not available
offline (Eigen's JacobiSVD solve, g2o's Levenberg-Marquardt with a Cauchy kernel).  They do not
make the oracle bit-pinned against Eigen / g2o -- nothing offline can -- but they show that the restated algorithms
compute what the reference's expressions mean, against LAPACK (numpy) and scipy.optimize:

* triangulation: the minimum-norm least-squares solution of the 4x3 system of lvt_local_map.cpp:285-292
  (A.leftCols<3>().jacobiSvd(...).solve(-A.col(3))) equals numpy.linalg.lstsq (LAPACK gelsd) on the same matrix,
  for arbitrary camera poses, down to the conditioning of the system;
* pose: the two-pass motion-only bundle adjustment of lvt_pnp_solver.cpp:60-128 ends at the minimiser of the cost its
  graph states -- sum over the inlier edges of delta^2 log(1 + |e|^2 / delta^2), delta^2 = 5.991 -- found here by
  scipy.optimize.least_squares from the same start, over the same SE3 parametrisation-free unknowns (rotation vector +
  translation).
"""
import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from helpers import configs

TH2 = 5.991  # LVT_REPROJECTION_TH2 (lvt/src/lvt_definitions.h:29)


def _world_to_camera(q_wxyz, t):
    """lvt_pose_utils::compute_world_to_camera_transform (lvt/src/lvt_pose.cpp:36-43): [R^T | -R^T t]"""
    R = Rotation.from_quat([q_wxyz[1], q_wxyz[2], q_wxyz[3], q_wxyz[0]]).as_matrix()
    return np.hstack([R.T, (-R.T @ np.asarray(t, float))[:, None]])


def _right_pose(q_wxyz, t, baseline):
    """lvt_pose_utils::compute_right_camera_pose (lvt/src/lvt_pose.cpp:28-34): the baseline along the camera's x axis"""
    R = Rotation.from_quat([q_wxyz[1], q_wxyz[2], q_wxyz[3], q_wxyz[0]]).as_matrix()
    return q_wxyz, np.asarray(t, float) + R[:, 0] * baseline


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_triangulation_least_squares_equals_lapack(oracle, seed):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    rng = np.random.default_rng(seed)
    rot = Rotation.from_rotvec(rng.normal(0, 0.3, 3))
    q = np.roll(rot.as_quat(), 1)  # (x, y, z, w) -> (w, x, y, z)
    t = rng.normal(0, 2.0, 3)
    # points in front of the (posed) left camera, from 3 m to 120 m (disparity down to ~3 px), pixel noise on both sides
    n = 400
    pc = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.2, 0.2, n), np.ones(n)], 1) * rng.uniform(3, 120, n)[:, None]
    pw = pc @ rot.as_matrix().T + t
    ul = np.stack([p.fx * pc[:, 0] / pc[:, 2] + p.cx, p.fy * pc[:, 1] / pc[:, 2] + p.cy], 1)
    ur = np.stack([p.fx * (pc[:, 0] - p.baseline) / pc[:, 2] + p.cx, ul[:, 1]], 1)
    ul = (ul + rng.normal(0, 0.4, ul.shape)).astype(np.float32)
    ur = (ur + rng.normal(0, 0.4, ur.shape)).astype(np.float32)
    xyz, ok = ctx.triangulate(q, t, ul, ur)
    assert ok.sum() > 200
    cml = _world_to_camera(q, t)
    cmr = _world_to_camera(*_right_pose(q, t, p.baseline))
    inv_fx, inv_fy = 1.0 / p.fx, 1.0 / p.fy
    worst = 0.0
    for i in np.flatnonzero(ok):
        u1x, u1y = (float(ul[i, 0]) - p.cx) * inv_fx, (float(ul[i, 1]) - p.cy) * inv_fy
        u2x, u2y = (float(ur[i, 0]) - p.cx) * inv_fx, (float(ur[i, 1]) - p.cy) * inv_fy
        A = np.stack([u1x * cml[2] - cml[0], u1y * cml[2] - cml[1], u2x * cmr[2] - cmr[0], u2y * cmr[2] - cmr[1]])
        x, _, rank, sv = np.linalg.lstsq(A[:, :3], -A[:, 3], rcond=None)
        assert rank == 3
        # both are backward-stable solvers of the same system: they agree to cond(A) * eps * |x|
        bound = 50 * (sv[0] / sv[-1]) * np.finfo(float).eps * max(1.0, np.linalg.norm(x))
        err = np.linalg.norm(xyz[i] - x)
        assert err <= bound, (i, err, bound)
        worst = max(worst, err)
    assert worst < 1e-8
    # and the accepted points reproject within the gate on both sides (lvt_local_map.cpp:303-320)
    good = xyz[ok > 0]
    for cm, uv in ((cml, ul), (cmr, ur)):
        c = good @ cm[:, :3].T + cm[:, 3]
        pr = np.stack([p.fx * c[:, 0] / c[:, 2] + p.cx, p.fy * c[:, 1] / c[:, 2] + p.cy], 1)
        assert (np.sum((pr - uv[ok > 0]) ** 2, 1) <= TH2 + 1e-9).all()


def _reproj(params, pts, uv, p):
    """e = K R^T (X - t) dehomogenised - z, camera pose (R, t) from a rotation vector + translation"""
    R = Rotation.from_rotvec(params[:3]).as_matrix()
    c = (pts - params[3:]) @ R
    return np.stack([p.fx * c[:, 0] / c[:, 2] + p.cx, p.fy * c[:, 1] / c[:, 2] + p.cy], 1) - uv


def _cauchy_minimiser(x0, pts, uv, p):
    """argmin sum delta^2 log(1 + |e_i|^2 / delta^2).  g2o robustifies the edge's chi2 = |e|^2, not its components:
    the residual vector of an edge is e * sqrt(rho(|e|^2) / |e|^2) (smooth at e = 0, squared norm = rho)"""
    def res(x):
        e = _reproj(x, pts, uv, p)
        z = np.sum(e * e, 1) / TH2
        w = np.sqrt(np.where(z > 1e-12, np.log1p(z) / np.maximum(z, 1e-300), 1.0 - 0.5 * z))
        return (e * w[:, None]).ravel()
    return least_squares(res, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000).x


@pytest.mark.parametrize("seed,outliers", [(3, 0), (4, 25), (5, 60)])
def test_pose_solver_ends_at_the_minimiser_of_its_cost(oracle, seed, outliers):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    rng = np.random.default_rng(seed)
    n = 400
    pts = np.stack([rng.uniform(-10, 10, n), rng.uniform(-2, 2, n), rng.uniform(5, 40, n)], 1)
    rot = Rotation.from_rotvec(rng.normal(0, 0.01, 3))
    t_true = rng.normal(0, 0.15, 3)
    c = (pts - t_true) @ rot.as_matrix()
    uv = np.stack([p.fx * c[:, 0] / c[:, 2] + p.cx, p.fy * c[:, 1] / c[:, 2] + p.cy], 1) + rng.normal(0, 0.5, (n, 2))
    uv[:outliers] += rng.uniform(15, 50, (outliers, 2)) * rng.choice([-1, 1], (outliers, 2))
    uv = uv.astype(np.float32)
    q, t, marks = ctx.solve_pose(pts, uv, [1, 0, 0, 0], [0, 0, 0])
    # the reference's two passes: optimise over everything, demote chi2 > 5.991, optimise over the rest
    x1 = _cauchy_minimiser(np.zeros(6), pts, uv.astype(float), p)
    inl = np.sum(_reproj(x1, pts, uv.astype(float), p) ** 2, 1) <= TH2
    x2 = _cauchy_minimiser(x1, pts[inl], uv[inl].astype(float), p)
    assert marks[:outliers].sum() == 0
    # the same edges survive both demotions (after the first pass and at the end, lvt_pnp_solver.cpp:100-121)
    assert np.array_equal(marks.astype(bool), inl)
    assert np.array_equal(marks.astype(bool), np.sum(_reproj(x2, pts, uv.astype(float), p) ** 2, 1) <= TH2)
    R_o = Rotation.from_quat([q[1], q[2], q[3], q[0]])
    dR = (R_o.inv() * Rotation.from_rotvec(x2[:3])).magnitude()
    # five LM iterations per pass from 0.15 m away end within 1e-7 m / 1e-8 rad of the minimiser (measured: 6e-8 / 4e-9)
    assert np.linalg.norm(t - x2[3:]) < 1e-6 and dR < 1e-7, (np.linalg.norm(t - x2[3:]), dR)
