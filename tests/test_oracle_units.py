"""Host-side behaviour of the oracle that the reference fixes: tile grid, ANMS, BRIEF properties,
matcher rules, pose solver, triangulation, state machine, parameters."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, capi, configs, kp_array, make_stream, synth


@pytest.fixture(scope="module")
def frame0():
    return make_stream("kitti_synth", 2).frame(0)


def test_golden_trajectory(oracle):
    for name in ("kitti_synth", "euroc_synth"):
        g = np.load(GOLDEN + "/track_%s.npz" % name)
        st = make_stream(name, len(g["poses"]))
        vo = oracle.create(configs.make_params(name), 1)
        for t in range(len(g["poses"])):
            R, tt = vo.track(*st.frame(t))
            fi = vo.frame_info()
            assert [fi[k] for k in sorted(fi)] == list(g["infos"][t])
            assert np.allclose(np.concatenate([R.ravel(), tt]), g["poses"][t], atol=1e-9)


def test_detect_tiles_and_anms(oracle, frame0):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    k = ctx.detect(frame0[0])
    # 10 tiles, each keeps >= k+1 when it had more than k corners (handler.cpp:66-82)
    tx = (k["x"] // 250).astype(int) + 5 * (k["y"] // 250).astype(int)
    counts = np.bincount(tx, minlength=10)
    assert counts.sum() == len(k) and (counts >= 151).all()
    assert np.all(np.diff(tx) >= 0)  # tiles are concatenated in rect order
    # no corner within 3 px of a tile edge
    lx, ly = k["x"] % 250, k["y"] % 250
    assert lx.min() >= 3 and ly.min() >= 3


def test_low_corner_retry(oracle):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    img = np.full((375, 1242), 100, np.uint8)
    rng = np.random.default_rng(0)
    for _ in range(60):  # faint corners: contrast 20 < threshold 25, found only by the retry at 13
        x, y = rng.integers(40, 1200), rng.integers(40, 330)
        img[y:y + 6, x:x + 6] = 120
    k = ctx.detect(img)
    assert 0 < len(k) < 200
    assert k["response"].max() < 25


def test_brief_properties(oracle, frame0):
    ctx = oracle.context(configs.make_params("kitti_synth"))
    img = frame0[0]
    kps = ctx.detect(img)
    out, desc = ctx.brief(img, kps)
    inside = (kps["x"] >= 28) & (kps["x"] < 1242 - 28) & (kps["y"] >= 28) & (kps["y"] < 375 - 28)
    assert np.array_equal(kp_array(out), kp_array(kps[inside]))
    # direct 9x9 box sums reproduce the bits
    pairs = []
    for line in open(os.path.join(ROOT, "oracle", "brief_pairs.inc")):
        if line.startswith("{"):
            pairs.append([int(v) for v in line[1:line.index("}")].split(",")])
    pairs = np.array(pairs)
    assert pairs.shape == (256, 4) and np.abs(pairs).max() <= 24
    im = img.astype(np.int64)
    for i in (0, 7, len(out) // 2, len(out) - 1):
        X, Y = int(out["x"][i]), int(out["y"][i])
        bits = []
        for dy1, dx1, dy2, dx2 in pairs:
            s1 = im[Y + dy1 - 4:Y + dy1 + 5, X + dx1 - 4:X + dx1 + 5].sum()
            s2 = im[Y + dy2 - 4:Y + dy2 + 5, X + dx2 - 4:X + dx2 + 5].sum()
            bits.append(int(s1 < s2))
        assert np.array_equal(np.packbits(bits), desc[i])
    # an integer translation of the image translates the keypoints and leaves the bits unchanged
    shifted = np.roll(img, (5, 7), axis=(0, 1))
    mid = out[(out["x"] > 60) & (out["x"] < 1100) & (out["y"] > 60) & (out["y"] < 300)].copy()
    _, d0 = ctx.brief(img, mid)
    mid["x"] += 7
    mid["y"] += 5
    _, d1 = ctx.brief(shifted, mid)
    assert np.array_equal(d0, d1)


def test_match_rules(oracle):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    rng = np.random.default_rng(3)
    kps = np.zeros(3, capi.KP_DTYPE)
    kps["x"], kps["y"] = [100, 110, 400], [100, 100, 300]
    desc = rng.integers(0, 256, (3, 32)).astype(np.uint8)
    Z = 10.0
    pt = np.array([[(100.0 - p.cx) / p.fx * Z, (100.0 - p.cy) / p.fy * Z, Z]])
    q = desc[0].copy()
    r = ctx.match_projected(pt, q, [1, 0, 0, 0], [0, 0, 0], kps, desc, retry_below=0)
    assert r["idx"][0] == 0 and r["count"] == 1 and r["matched"].tolist() == [1, 0, 0]
    # a second, equally good candidate breaks the ratio test
    desc2 = desc.copy()
    desc2[1] = desc[0]
    assert ctx.match_projected(pt, q, [1, 0, 0, 0], [0, 0, 0], kps, desc2, retry_below=0)["idx"][0] == -1
    # behind the camera / outside the image -> -2
    r = ctx.match_projected(np.array([[0, 0, -5.0], [1e4, 0, 1.0]]), np.stack([q, q]), [1, 0, 0, 0], [0, 0, 0], kps, desc,
                            retry_below=0)
    assert r["idx"].tolist() == [-2, -2]
    # greedy: the first point takes the feature, the second identical point finds nothing
    r = ctx.match_projected(np.repeat(pt, 2, 0), np.stack([q, q]), [1, 0, 0, 0], [0, 0, 0], kps[:1], desc[:1], retry_below=0)
    assert r["idx"].tolist() == [0, -1]


def test_pose_solver_recovers_motion(oracle):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    rng = np.random.default_rng(11)
    n = 300
    pts = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(6, 30, n)], 1)
    t_true = np.array([0.4, -0.05, 0.1])
    ang = 0.02
    q_true = np.array([np.cos(ang / 2), 0, np.sin(ang / 2), 0])
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    pc = (pts - t_true) @ R  # R^T (p - t)
    uv = np.stack([p.fx * pc[:, 0] / pc[:, 2] + p.cx, p.fy * pc[:, 1] / pc[:, 2] + p.cy], 1)
    uv += rng.normal(0, 0.3, uv.shape)
    uv[:30] += rng.uniform(20, 60, (30, 2))  # gross outliers
    q, t, marks = ctx.solve_pose(pts, uv.astype(np.float32), [1, 0, 0, 0], [0, 0, 0])
    assert np.linalg.norm(t - t_true) < 5e-3 and np.abs(q - q_true).max() < 2e-3
    assert marks[:30].sum() == 0 and marks[30:].sum() > 250
    # noiseless data: exact to solver precision
    uv0 = np.stack([p.fx * pc[:, 0] / pc[:, 2] + p.cx, p.fy * pc[:, 1] / pc[:, 2] + p.cy], 1).astype(np.float32)
    q, t, marks = ctx.solve_pose(pts, uv0, [1, 0, 0, 0], [0, 0, 0])
    assert np.linalg.norm(t - t_true) < 1e-4 and marks.all()


def test_triangulation_roundtrip(oracle):
    p = configs.make_params("kitti_synth")
    ctx = oracle.context(p)
    rng = np.random.default_rng(5)
    n = 200
    pts = np.stack([rng.uniform(-6, 6, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 40, n)], 1)
    ul = np.stack([p.fx * pts[:, 0] / pts[:, 2] + p.cx, p.fy * pts[:, 1] / pts[:, 2] + p.cy], 1)
    ur = np.stack([p.fx * (pts[:, 0] - p.baseline) / pts[:, 2] + p.cx, ul[:, 1]], 1)
    xyz, ok = ctx.triangulate([1, 0, 0, 0], [0, 0, 0], ul.astype(np.float32), ur.astype(np.float32))
    vis = (ul[:, 0] > 0) & (ul[:, 0] < 1242) & (ur[:, 0] > 0) & (ul[:, 1] > 0) & (ul[:, 1] < 375)
    assert ok[vis].all()
    assert np.abs(xyz[vis] - pts[vis]).max() < 5e-2  # float32 pixel rounding at 40 m
    bad = ur.copy()
    bad[:, 1] += 8.0  # 4 px of reprojection error on each side -> 16 > 5.991
    _, ok2 = ctx.triangulate([1, 0, 0, 0], [0, 0, 0], ul.astype(np.float32), bad.astype(np.float32))
    assert ok2.sum() == 0


def test_state_machine_lost_and_reset(oracle):
    p = configs.make_params("kitti_synth")
    st = make_stream("kitti_synth", 3)
    vo = oracle.create(p, 1)
    assert vo.get_state() == capi.STATE_NOT_INITIALIZED
    R, t = vo.track(*st.frame(0))
    assert vo.get_state() == capi.STATE_TRACKING and np.array_equal(R, np.eye(3)) and not t.any()
    R1, t1 = vo.track(*st.frame(1))
    blank = np.full((375, 1242), 128, np.uint8)
    R2, t2 = vo.track(blank, blank)  # no features -> fewer than 10 matches -> lost
    assert vo.get_state() == capi.STATE_LOST and np.array_equal(t2, t1)
    R3, t3 = vo.track(*st.frame(2))  # stays lost, keeps returning the last pose
    assert vo.get_state() == capi.STATE_LOST and np.array_equal(t3, t1)
    vo.reset()
    assert vo.get_state() == capi.STATE_NOT_INITIALIZED
    R, t = vo.track(*st.frame(0))
    assert vo.get_state() == capi.STATE_TRACKING and not t.any()


def test_params_yaml(oracle, tmp_path):
    f = tmp_path / "vo.yaml"
    f.write_text("%YAML:1.0\n\nfx: 718.856\nfy: 718.856\ncx: 607.1928\ncy: 185.2157\nbaseline: 0.537\nimg_width: 1242\n"
                 "img_height: 375\nnear_plane_distance: 0.01\nfar_plane_distance: 500.0\ntracking_radius: 25\n"
                 "agast_threshold: 25\ndetection_cell_size: 250\nmax_keypoints_per_cell: 150\nhashing_cell_size: 99\n"
                 "triangulation_policy: 1\nenable_logging: 0\n")
    p = oracle.params_from_file(str(f))
    assert p.img_width == 1242 and abs(p.fx - 718.856) < 1e-3 and p.tracking_radius == 25
    assert p.staged_threshold == 0 and p.k1 == 0.0  # absent keys read as 0 (cv::FileNode)
    assert oracle.create_from_file(str(tmp_path / "missing.yaml")) is None
    d = oracle.default_params()
    assert d.max_keypoints_per_cell == 150 and d.staged_threshold == 2 and abs(d.descriptor_matching_threshold - 30) < 1e-6
    vo = oracle.create_from_file(str(f))
    assert vo is not None and vo.get_state() == 1
    assert oracle.lib.lvt_create(str(f).encode(), 3) is None  # bad sensor type


def test_external_corners_and_rgbd(oracle):
    p = configs.make_params("kitti_synth")
    st = make_stream("kitti_synth", 3)
    ctx = oracle.context(p)
    vo = oracle.create(p, 1)
    for t in range(3):
        L, Rr = st.frame(t)
        kl, kr = ctx.detect(L), ctx.detect(Rr)
        R, tt = vo.track_with_external_corners(L, Rr, np.stack([kl["x"], kl["y"]], 1), np.stack([kr["x"], kr["y"]], 1))
    assert vo.get_state() == capi.STATE_TRACKING and 0.5 < tt[0] < 1.2
    name = "tum_synth"
    pr = configs.make_params(name)
    sr = make_stream(name, 4, seed=1)
    vr = oracle.create(pr, 2)
    for t in range(4):
        R, tt = vr.track_rgbd(*sr.frame(t))
    fi = vr.frame_info()
    assert vr.get_state() == capi.STATE_TRACKING and fi["tracked"] > 300
    assert abs(tt[0] - 3 * 8 * 2.0 / pr.fx) < 5e-3
