"""Rectification in front of track() (examples/euroc/euroc_example.cpp:106-107,142-143)
and the dataset drivers / trajectory writers of the three example executables.
CPU part: the oracle against the genuine OpenCV outputs (tests/golden/rectify_cv2.npz, made by
tools/make_golden_rectify.py with cv2 4.13) and the drivers over the oracle.
GPU part (-m gpu): the CUDA path against the oracle and the same golden file, through the C ABI."""
import hashlib
import os

import numpy as np
import pytest

from helpers import GOLDEN, configs, make_stream
from lvt_b200 import capi, datasets, euroc_calib


def _golden():
    return np.load(os.path.join(GOLDEN, "rectify_cv2.npz"))


def _test_image(seed, w, h):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    return cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 1.2)


def _check_against_cv2(lib, full=True):
    g = _golden()
    p = configs.make_params("euroc_synth")
    for name in [str(n) for n in g["names"]]:
        img = g[name + "_img"]
        h, w = img.shape
        ctx = lib.context(configs.make_params("euroc_synth", img_width=w, img_height=h))
        r = capi.Rectify.make(g[name + "_K"], g[name + "_D"], g[name + "_R"], g[name + "_P"])
        mx, my = ctx.rectify_maps(r, h, w)
        if name == "identity":
            # cv2's AVX2 build contracts x*fx + cx into one FMA: exact zeros come out as -1.1e-15 there
            assert np.abs(mx - g[name + "_map_x"]).max() < 1e-12 and np.abs(my - g[name + "_map_y"]).max() < 1e-12
        else:
            assert np.array_equal(mx, g[name + "_map_x"]), name
            assert np.array_equal(my, g[name + "_map_y"]), name
        assert np.array_equal(ctx.rectify(img, r), g[name + "_out"]), name
        ctx.destroy()
    if full:
        w, h = euroc_calib.IMG_SIZE
        ctx = lib.context(p)
        img = _test_image(int(g["euroc_full_img_seed"]), w, h)
        for side, (K, D, R, P) in zip(("left", "right"), euroc_calib.rectify_args()):
            r = capi.Rectify.make(K, D, R, P)
            mx, my = ctx.rectify_maps(r, h, w)
            got = [hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest() for a in (mx, my, ctx.rectify(img, r))]
            assert got == [str(x) for x in g["euroc_full_%s_sha1" % side]], side
        ctx.destroy()


def test_oracle_rectify_equals_cv2_fixtures(oracle):
    _check_against_cv2(oracle)


def test_oracle_rectify_live_cv2(oracle):
    """random calibrations, live against cv2 where it is importable"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    w, h = 120, 90
    ctx = oracle.context(configs.make_params("euroc_synth", img_width=w, img_height=h))
    for _ in range(6):
        K = np.array([[rng.uniform(60, 140), 0, rng.uniform(40, 80)], [0, rng.uniform(60, 140), rng.uniform(30, 60)], [0, 0, 1]])
        D = np.array([rng.uniform(-0.4, 0.2), rng.uniform(-0.1, 0.2), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01),
                      rng.uniform(-0.05, 0.05)])
        rv = rng.uniform(-0.05, 0.05, 3)
        R = cv2.Rodrigues(rv)[0]
        P = K.copy()
        P[0, 0] *= rng.uniform(0.8, 1.1)
        P[1, 1] *= rng.uniform(0.8, 1.1)
        img = _test_image(int(rng.integers(1 << 30)), w, h)
        m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
        r = capi.Rectify.make(K, D, R, P)
        mx, my = ctx.rectify_maps(r, h, w)
        assert np.array_equal(mx, m1) and np.array_equal(my, m2)
        assert np.array_equal(ctx.rectify(img, r), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))


def _degenerate_calibrations():
    """maps that leave every comfortable range: rays with w <= 0 (|map| up to 1e13 -> the short-range
    saturation of cv::remap's fixed-point coordinates), distortion polynomials of thousands of pixels, and a
    projection shifted 5000 px away (every tap outside)"""
    K = np.array([[80.0, 0, 60], [0, 80.0, 45], [0, 0, 1]])
    a = 1.45
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    return [(K, np.array([-0.2, 0.05, 0, 0, 0.0]), Ry, K),
            (K, np.array([5.0, 20.0, 0.1, 0.1, 50.0]), np.eye(3), K),
            (K, np.zeros(5), np.eye(3), np.array([[80.0, 0, -5000], [0, 80.0, 45], [0, 0, 1]]))]


def test_oracle_rectify_degenerate_maps_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    w, h = 120, 90
    ctx = oracle.context(configs.make_params("euroc_synth", img_width=w, img_height=h))
    img = _test_image(3, w, h)
    for K, D, R, P in _degenerate_calibrations():
        m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
        # (map values of 1e3 .. 1e13 differ from cv2's FMA-contracted ones in the last bits; the image must not)
        assert np.array_equal(ctx.rectify(img, capi.Rectify.make(K, D, R, P)), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    ctx.destroy()


def test_rectify_bad_arguments(oracle):
    ctx = oracle.context(configs.make_params("euroc_synth"))
    vo = oracle.create(configs.make_params("euroc_synth"))
    sing = capi.Rectify.make(np.eye(3), np.zeros(5), np.eye(3), np.zeros((3, 3)))
    ok = capi.Rectify.make(np.eye(3), np.zeros(5), np.eye(3), np.eye(3))
    assert oracle.lib.lvt_set_rectification(vo.h, sing, sing) != 0      # P R is singular
    assert oracle.lib.lvt_set_rectification(vo.h, ok, None) != 0        # one camera only
    assert oracle.lib.lvt_set_rectification(vo.h, ok, ok) == 0
    assert oracle.lib.lvt_set_rectification(vo.h, None, None) == 0
    vo.destroy()
    ctx.destroy()


MILD = dict(D=[-0.03, 0.01, 0.0004, -0.0003, 0.0])


def _mild_rectification(p):
    """a mild warp shared by both cameras: rows stay aligned, so the stereo stream keeps tracking"""
    K = np.array([[p.fx, 0, p.cx], [0, p.fy, p.cy], [0, 0, 1.0]])
    a = 0.004
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    r = capi.Rectify.make(K, MILD["D"], R, K)
    return r, r


def _run_rectified(lib, n):
    p = configs.make_params("euroc_synth", max_keypoints_per_cell=300)
    st = make_stream("euroc_synth", n)
    vo = lib.create(p)
    vo.set_rectification(*_mild_rectification(p))
    out = []
    for t in range(n):
        a, b = st.frame(t)
        R, tt = vo.track(a, b)
        xy, desc = vo.features(0)
        out.append((R, tt, vo.frame_info(), xy, desc))
    vo.destroy()
    return out


def test_oracle_tracks_through_rectification(oracle):
    out = _run_rectified(oracle, 4)
    assert out[-1][2]["state"] == capi.STATE_TRACKING and out[-1][2]["tracked"] > 100


def _write_png(path, img):
    cv2 = pytest.importorskip("cv2")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    assert cv2.imwrite(path, img)


def _yaml(path, p):
    keys = [k for k, _ in capi.Params._fields_]
    with open(path, "w") as f:
        f.write("%YAML:1.0\n")
        for k in keys:
            f.write("%s: %r\n" % (k, getattr(p, k)))
    return path


def test_dataset_drivers_and_writers(oracle, tmp_path):
    pytest.importorskip("cv2")
    n = 3
    # KITTI layout: sequences/NN/image_{0,1}/%06d.png, calib yml with camera_matrix + baseline
    p = configs.make_params("kitti_stock")
    st = make_stream("kitti_stock", n)
    for t in range(n):
        a, b = st.frame(t)
        _write_png(str(tmp_path / "seq" / "07" / "image_0" / ("%06d.png" % t)), a)
        _write_png(str(tmp_path / "seq" / "07" / "image_1" / ("%06d.png" % t)), b)
    cfg = _yaml(str(tmp_path / "vo.yaml"), p)
    calib = tmp_path / "07.yml"
    calib.write_text("%%YAML:1.0\n\ncamera_matrix: !!opencv-matrix\n  rows: 3\n  cols: 3\n  dt: d\n  data: [ %r, 0, %r, 0, %r,\n %r, 0, 0, 1 ]\n\nbaseline: %r\n"
                     % (float(p.fx), float(p.cx), float(p.fy), float(p.cy), float(p.baseline)))
    poses = datasets.run_kitti(oracle, str(tmp_path / "seq"), 7, cfg, str(calib), str(tmp_path / "07.txt"))
    vo = oracle.create(p)
    ref = [vo.track(*st.frame(t)) for t in range(n)]
    vo.destroy()
    for (R, t), (R2, t2) in zip(poses, ref):
        assert np.array_equal(R, R2) and np.array_equal(t, t2)
    rows = np.loadtxt(str(tmp_path / "07.txt")).reshape(n, 12)
    assert np.allclose(rows[:, [3, 7, 11]], [t for _, t in ref], atol=1e-9)

    # EuRoC layout: root/NAME/mav0/cam{0,1}/data/<stamp>.png + stamps/NAME.txt; raw frames in, body poses out
    pe = configs.make_params("euroc_synth", max_keypoints_per_cell=200)
    se = make_stream("euroc_synth", n)
    stamps = [1403636579763555584 + 50000000 * t for t in range(n)]
    (tmp_path / "stamps").mkdir()
    (tmp_path / "stamps" / "MH_01.txt").write_text("".join("%d\n" % s for s in stamps))
    for t in range(n):
        a, b = se.frame(t)
        _write_png(str(tmp_path / "euroc" / "MH_01" / "mav0" / "cam0" / "data" / ("%d.png" % stamps[t])), a)
        _write_png(str(tmp_path / "euroc" / "MH_01" / "mav0" / "cam1" / "data" / ("%d.png" % stamps[t])), b)
    cfg_e = _yaml(str(tmp_path / "euroc.yaml"), pe)
    poses_e, ts = datasets.run_euroc(oracle, str(tmp_path / "euroc"), str(tmp_path / "stamps"), "MH_01", cfg_e,
                                     str(tmp_path / "MH_01.txt"))
    assert np.allclose(ts, [s / 1e9 for s in stamps])
    # frame 0: identity camera pose -> the body pose is T_BS itself
    assert np.allclose(poses_e[0][0], euroc_calib.T_BS[:3, :3]) and np.allclose(poses_e[0][1], euroc_calib.T_BS[:3, 3])
    rows = np.loadtxt(str(tmp_path / "MH_01.txt")).reshape(n, 8)
    q = rows[0, 4:8]
    assert abs(np.linalg.norm(q) - 1) < 1e-6
    vo = oracle.create(pe)
    (Kl, Dl, Rl, Pl), (Kr, Dr, Rr, Pr) = euroc_calib.rectify_args()
    vo.set_rectification(capi.Rectify.make(Kl, Dl, Rl, Pl), capi.Rectify.make(Kr, Dr, Rr, Pr))
    for t in range(n):
        R, tt = vo.track(*se.frame(t))
    body = euroc_calib.T_BS @ np.block([[R, tt[:, None]], [np.zeros((1, 3)), np.ones((1, 1))]])
    assert np.allclose(poses_e[-1][1], body[:3, 3])
    vo.destroy()

    # TUM layout: root/NAME/{rgb,depth}/*.png + associations/NAME.txt, depth in 1/5000 m
    pt = configs.make_params("tum_synth", max_keypoints_per_cell=400)
    stt = make_stream("tum_synth", n)
    lines = []
    for t in range(n):
        g, d = stt.frame(t)
        _write_png(str(tmp_path / "tum" / "fr3" / "rgb" / ("%d.png" % t)), np.repeat(g[:, :, None], 3, 2))
        _write_png(str(tmp_path / "tum" / "fr3" / "depth" / ("%d.png" % t)), np.round(d * 5000).astype(np.uint16))
        lines.append("%f rgb/%d.png %f depth/%d.png\n" % (1.0 + t, t, 1.0 + t, t))
    (tmp_path / "assoc").mkdir()
    (tmp_path / "assoc" / "fr3.txt").write_text("".join(lines))
    cfg_t = _yaml(str(tmp_path / "tum.yaml"), pt)
    poses_t, ts_t = datasets.run_tum_rgbd(oracle, str(tmp_path / "tum"), str(tmp_path / "assoc"), "fr3", cfg_t,
                                          str(tmp_path / "fr3.txt"))
    assert len(poses_t) == n and ts_t == [1.0, 2.0, 3.0]
    assert np.loadtxt(str(tmp_path / "fr3.txt")).shape == (n, 8)


def test_quat_from_matrix_branches():
    for axis, ang in (((0, 0, 1), 0.3), ((1, 0, 0), 3.0), ((0, 1, 0), 3.1), ((0, 0, 1), 3.1)):
        a = np.array(axis, float)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        x, y, z, w = datasets.quat_from_matrix(R)
        R2 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(R, R2, atol=1e-12)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_rectify_equals_cv2_fixtures(cuda):
    _check_against_cv2(cuda)


@pytest.mark.gpu
def test_gpu_rectify_equals_oracle_random(cuda, oracle):
    rng = np.random.default_rng(9)
    p = configs.make_params("euroc_synth")
    h, w = p.img_height, p.img_width
    cg, co = cuda.context(p), oracle.context(p)
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    for k in range(4):
        K = np.array([[rng.uniform(300, 500), 0, rng.uniform(300, 450)], [0, rng.uniform(300, 500), rng.uniform(200, 280)], [0, 0, 1]])
        D = np.array([rng.uniform(-0.4, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01),
                      rng.uniform(-0.1, 0.1)])
        a, b = rng.uniform(-0.1, 0.1, 2)
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ np.array(
            [[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        P = K * np.array([[rng.uniform(0.5, 1.5)], [rng.uniform(0.5, 1.5)], [1.0]])
        if k == 3:
            P[0, 2] += 4000  # everything maps outside the raw image
        r = capi.Rectify.make(K, D, R, P)
        mg, mo = cg.rectify_maps(r, h, w), co.rectify_maps(r, h, w)
        assert np.array_equal(mg[0], mo[0]) and np.array_equal(mg[1], mo[1])
        assert np.array_equal(cg.rectify(img, r), co.rectify(img, r))
    for K, D, R, P in _degenerate_calibrations():
        r = capi.Rectify.make(K, D, R, P)
        mg, mo = cg.rectify_maps(r, h, w), co.rectify_maps(r, h, w)
        assert np.array_equal(mg[0], mo[0], equal_nan=True) and np.array_equal(mg[1], mo[1], equal_nan=True)
        assert np.array_equal(cg.rectify(img, r), co.rectify(img, r))
    cg.destroy()
    co.destroy()


@pytest.mark.gpu
def test_gpu_tracks_through_rectification(cuda, oracle):
    """raw frames in, rectified on the device inside lvt_track: everything downstream bit-exact"""
    n = 6
    got, ref = _run_rectified(cuda, n), _run_rectified(oracle, n)
    for t, ((R, tt, fi, xy, desc), (R2, tt2, fi2, xy2, desc2)) in enumerate(zip(got, ref)):
        assert np.array_equal(xy, xy2) and np.array_equal(desc, desc2), "frame %d features" % t
        assert fi == fi2, "frame %d: %s vs %s" % (t, fi, fi2)
        assert np.abs(tt - tt2).max() < 1e-6 and np.abs(R - R2).max() < 1e-6
    assert got[-1][2]["state"] == capi.STATE_TRACKING


@pytest.mark.gpu
def test_gpu_pool_rectifies_on_upload(cuda, oracle):
    n = 4
    p = configs.make_params("euroc_synth", max_keypoints_per_cell=300)
    st = make_stream("euroc_synth", n)
    res = []
    for lib in (cuda, oracle):
        vo = lib.create(p)
        vo.set_rectification(*_mild_rectification(p))
        vo.pool_reserve(n)
        for t in range(n):
            vo.pool_upload(t, *st.frame(t))
        poses, infos = vo.track_pool(0, n)
        res.append((poses, infos))
        vo.destroy()
    assert res[0][1] == res[1][1]
    assert np.abs(res[0][0] - res[1][0]).max() < 1e-6


def _make_euroc_tree(tmp_path, n, p):
    se = make_stream("euroc_synth", n)
    stamps = [1403636579763555584 + 50000000 * t for t in range(n)]
    (tmp_path / "stamps").mkdir(exist_ok=True)
    (tmp_path / "stamps" / "V1_01.txt").write_text("".join("%d\n" % s for s in stamps))
    for t in range(n):
        a, b = se.frame(t)
        _write_png(str(tmp_path / "euroc" / "V1_01" / "mav0" / "cam0" / "data" / ("%d.png" % stamps[t])), a)
        _write_png(str(tmp_path / "euroc" / "V1_01" / "mav0" / "cam1" / "data" / ("%d.png" % stamps[t])), b)
    return _yaml(str(tmp_path / "euroc.yaml"), p)


@pytest.mark.gpu
def test_gpu_dataset_drivers_equal_oracle(cuda, oracle, tmp_path):
    """the EuRoC and TUM drivers over the CUDA library give the oracle's trajectories (raw EuRoC frames are
    rectified on the device with the genuine calibration)"""
    pytest.importorskip("cv2")
    n = 4
    pe = configs.make_params("euroc_synth", max_keypoints_per_cell=200)
    cfg_e = _make_euroc_tree(tmp_path, n, pe)
    out = []
    for lib, tag in ((cuda, "g"), (oracle, "o")):
        poses, ts = datasets.run_euroc(lib, str(tmp_path / "euroc"), str(tmp_path / "stamps"), "V1_01", cfg_e,
                                       str(tmp_path / ("V1_01_%s.txt" % tag)))
        out.append(np.array([np.concatenate([R.ravel(), t]) for R, t in poses]))
    assert np.abs(out[0] - out[1]).max() < 1e-6
    assert np.allclose(np.loadtxt(str(tmp_path / "V1_01_g.txt")), np.loadtxt(str(tmp_path / "V1_01_o.txt")), atol=1e-6)

    pt = configs.make_params("tum_synth", max_keypoints_per_cell=400)
    stt = make_stream("tum_synth", n)
    lines = []
    for t in range(n):
        g, d = stt.frame(t)
        _write_png(str(tmp_path / "tum" / "fr3" / "rgb" / ("%d.png" % t)), np.repeat(g[:, :, None], 3, 2))
        _write_png(str(tmp_path / "tum" / "fr3" / "depth" / ("%d.png" % t)), np.round(d * 5000).astype(np.uint16))
        lines.append("%f rgb/%d.png %f depth/%d.png\n" % (1.0 + t, t, 1.0 + t, t))
    (tmp_path / "assoc").mkdir()
    (tmp_path / "assoc" / "fr3.txt").write_text("".join(lines))
    cfg_t = _yaml(str(tmp_path / "tum.yaml"), pt)
    out = []
    for lib in (cuda, oracle):
        poses, _ = datasets.run_tum_rgbd(lib, str(tmp_path / "tum"), str(tmp_path / "assoc"), "fr3", cfg_t)
        out.append(np.array([np.concatenate([R.ravel(), t]) for R, t in poses]))
    assert np.abs(out[0] - out[1]).max() < 1e-6


# ---- RGB-D keypoint undistortion (lvt_image_features_handler.cpp:268-294, cv::undistortPoints) ----
TUM_FR1_LIKE = dict(k1=0.2312, k2=-0.7849, p1=-0.0033, p2=-0.0001, k3=0.9172)


def _undistorted_features(lib):
    p = configs.make_params("tum_synth", max_keypoints_per_cell=400, **TUM_FR1_LIKE)
    gray, depth = make_stream("tum_synth", 1, seed=2).frame(0)
    vo = lib.create(p, capi.SENSOR_RGBD)
    vo.track_rgbd(gray, depth)
    xy, _ = vo.features(0)
    vo.destroy()
    ctx = lib.context(p)
    k, _ = ctx.extract(gray)
    ctx.destroy()
    return np.stack([k["x"], k["y"]], 1).astype(np.float32), xy, p


def test_oracle_undistortion_equals_cv2_fixture(oracle):
    """the keypoints of an RGB-D frame after the reference's cv::undistortPoints call == the genuine OpenCV output
    (tests/golden/undistort_cv2.npz, tools/make_golden_rectify.py --undistort)"""
    g = np.load(os.path.join(GOLDEN, "undistort_cv2.npz"))
    detected, xy, _ = _undistorted_features(oracle)
    assert np.array_equal(detected, g["detected"])
    assert np.array_equal(xy, g["undistorted"])
    assert np.abs(xy - detected).max() > 1.0  # the distortion is not a no-op on this frame


def test_oracle_undistortion_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    detected, xy, p = _undistorted_features(oracle)
    K = np.array([[p.fx, 0, p.cx], [0, p.fy, p.cy], [0, 0, 1]], np.float64)
    dist = np.array([p.k1, p.k2, p.p1, p.p2, p.k3], np.float64)
    assert np.array_equal(xy, cv2.undistortPoints(detected.reshape(-1, 1, 2), K, dist, None, None, K).reshape(-1, 2))


@pytest.mark.gpu
def test_gpu_rgbd_with_lens_distortion(cuda, oracle):
    """RGB-D frames with a distorted camera model: undistorted keypoints (== the cv2 fixture), visibility bounds from
    the undistorted image corners, and everything downstream agree with the oracle"""
    g = np.load(os.path.join(GOLDEN, "undistort_cv2.npz"))
    detected, xy, p = _undistorted_features(cuda)
    assert np.array_equal(detected, g["detected"]) and np.array_equal(xy, g["undistorted"])
    n = 5
    st = make_stream("tum_synth", n, seed=2)
    vg, vo = cuda.create(p, capi.SENSOR_RGBD), oracle.create(p, capi.SENSOR_RGBD)
    for t in range(n):
        gray, depth = st.frame(t)
        Rg, tg = vg.track_rgbd(gray, depth)
        Ro, to = vo.track_rgbd(gray, depth)
        fx, fd = vg.features(0)
        ox, od = vo.features(0)
        assert np.array_equal(fx, ox) and np.array_equal(fd, od), t
        assert vg.frame_info() == vo.frame_info(), (t, vg.frame_info(), vo.frame_info())
        assert np.abs(tg - to).max() < 1e-6 and np.abs(Rg - Ro).max() < 1e-6
    assert vg.frame_info()["tracked"] > 100
    vg.destroy()
    vo.destroy()
