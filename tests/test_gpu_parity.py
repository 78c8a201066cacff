"""Parity of the CUDA path against the CPU oracle, through the C ABI, on a real GPU.
Bit-exact: keypoints, descriptors, match indices, marks, inlier marks, per-frame counters.
Floating point (fp64 pose / triangulation): tolerance stated at each assert."""
import numpy as np
import pytest

from helpers import GOLDEN, capi, configs, kp_array, kp_equal, knn_params, knn_via_match, make_stream, synth, track

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs(cuda, oracle):
    p = configs.make_params("kitti_synth")
    return cuda.context(p), oracle.context(p), p


@pytest.fixture(scope="module")
def frames():
    st = make_stream("kitti_synth", 3)
    return [st.frame(t) for t in range(3)]


def test_agast_fixtures_from_cv2(ctxs):
    g, _, _ = ctxs
    z = np.load(GOLDEN + "/agast_cv2.npz")
    for i in range(int(z["n_cases"])):
        for nms in (0, 1):
            got = kp_array(g.agast(z["img%d" % i], int(z["th%d" % i]), bool(nms)))
            assert np.array_equal(got, z["kp%d_nms%d" % (i, nms)]), (i, nms)


def test_agast_shapes_thresholds_ties(ctxs, frames):
    g, o, _ = ctxs
    rng = np.random.default_rng(5)
    L = frames[0][0]
    imgs = [L[:h, :w] for h, w in ((24, 24), (33, 65), (70, 70), (100, 40), (125, 242), (250, 250), (250, 2), (7, 7))]
    imgs += [(rng.integers(0, 8, s) * 32).astype(np.uint8) for s in ((48, 48), (96, 96), (160, 200))]  # tie-heavy, big blobs
    imgs += [rng.integers(0, 256, (120, 130)).astype(np.uint8), np.zeros((40, 40), np.uint8)]
    for img in imgs:
        img = np.ascontiguousarray(img)
        for th in (10, 25, 60):
            for nms in (True, False):
                assert kp_equal(g.agast(img, th, nms), o.agast(img, th, nms)), (img.shape, th, nms)


def test_agast_huge_component_fallback(ctxs):
    g, o, _ = ctxs
    img = np.zeros((120, 160), np.uint8)
    img[::2, ::2] = 255  # every bright pixel is a corner candidate: components of thousands of pixels
    img[1::2, 1::2] = 200
    assert kp_equal(g.agast(img, 10, True), o.agast(img, 10, True))


def _blur_quant(seed, h, w, sigma, step):
    """binary noise, blurred and quantised: corner components of hundreds of pixels, many with several
    pixels at the maximum score (found with scipy.ndimage.label on the oracle's corner map: sigma 1.0 ->
    components up to ~300 px, a dozen beyond the 96-px replay buffers, about half of them tied)"""
    rng = np.random.default_rng(seed)
    a = (rng.integers(0, 2, (h, w)) * 200).astype(np.float32)
    k = np.arange(-4, 5)
    g = np.exp(-k * k / (2.0 * sigma * sigma))
    g /= g.sum()
    a = np.apply_along_axis(lambda r: np.convolve(r, g, mode="same"), 1, a)
    a = np.apply_along_axis(lambda c: np.convolve(c, g, mode="same"), 0, a)
    return (np.clip(a, 0, 255).astype(np.uint8) // step * step).astype(np.uint8)


def test_agast_big_components_settled_by_tile_kernel(ctxs, cuda, oracle):
    """components beyond the NMS kernel's replay buffers (96 px) are flooded by the tile's CTA"""
    g, o, _ = ctxs
    cases = [(_blur_quant(1, 200, 240, 1.0, 16), 10), (_blur_quant(2, 250, 250, 1.2, 16), 8), (_blur_quant(3, 250, 250, 1.5, 8), 6),
             (_blur_quant(4, 130, 250, 0.8, 32), 10), (synth.canvas(105, 250, 250, density=60, noise=6.0), 8)]
    for img, th in cases:
        img = np.ascontiguousarray(img)
        assert kp_equal(g.agast(img, th, True), o.agast(img, th, True)), (img.shape, th)
    # the same through the tiled detector (tiles of 250 px, several big components per tile)
    p = configs.make_params("euroc_synth", agast_threshold=8)
    cg, co = cuda.context(p), oracle.context(p)
    img = np.ascontiguousarray(_blur_quant(7, p.img_height, p.img_width, 1.2, 16))
    assert kp_equal(cg.detect(img), co.detect(img))
    cg.destroy()
    co.destroy()


@pytest.mark.parametrize("name", ["kitti_synth", "kitti_stock", "euroc_synth", "tum_synth"])
def test_extract_all_configs(cuda, oracle, name):
    p = configs.make_params(name)
    g, o = cuda.context(p), oracle.context(p)
    st = make_stream(name, 2, seed=3)
    for t in range(2):
        img = st.frame(t)[0]
        assert kp_equal(g.detect(img), o.detect(img))
        ka, da = g.extract(img)
        kb, db = o.extract(img)
        assert kp_equal(ka, kb) and np.array_equal(da, db)
        assert len(ka) > 500


def test_low_corner_retry(ctxs):
    g, o, _ = ctxs
    img = np.full((375, 1242), 100, np.uint8)
    rng = np.random.default_rng(0)
    for _ in range(60):
        x, y = rng.integers(40, 1200), rng.integers(40, 330)
        img[y:y + 6, x:x + 6] = 120
    a, b = g.detect(img), o.detect(img)
    assert kp_equal(a, b) and 0 < len(a) < 200
    blank = np.full((375, 1242), 7, np.uint8)
    assert len(g.detect(blank)) == 0 and len(g.extract(blank)[0]) == 0


def test_brief_external_fractional_corners(ctxs, frames):
    g, o, _ = ctxs
    img = frames[0][0]
    rng = np.random.default_rng(8)
    k = np.zeros(3000, capi.KP_DTYPE)
    k["x"] = rng.uniform(0, 1242, 3000).astype(np.float32)
    k["y"] = rng.uniform(0, 375, 3000).astype(np.float32)
    k["x"][:40] = np.array([27.5, 28.5, 1213.5, 1214.5, 27.49, 28.0, 1213.0, 1214.0] * 5, np.float32)  # border + .5 rounding
    k["y"][:40] = 100.5
    ka, da = g.brief(img, k)
    kb, db = o.brief(img, k)
    assert kp_equal(ka, kb) and np.array_equal(da, db) and len(ka) > 2000
    e = np.zeros(0, capi.KP_DTYPE)
    assert len(g.brief(img, e)[0]) == 0


def test_knn_fixtures_from_cv2(cuda):
    z = np.load(GOLDEN + "/knn_cv2.npz")
    p = knn_params(cuda)
    ctx = cuda.context(p)
    for q, m, r in zip(z["queries"], z["masks"], z["results"]):
        idx, d0, d1 = knn_via_match(ctx, p, z["train"], q, m)
        n = int((r[:, 0] >= 0).sum())
        if n == 0 or (n == 2 and r[0, 1] == 0 and r[1, 1] == 0):
            assert idx == -1
        elif n == 1:
            assert idx == int(r[0, 0]) and d0 == r[0, 1] and d1 == -1.0
        else:
            assert idx == int(r[0, 0]) and d0 == r[0, 1] and d1 == r[1, 1]


def _scene(ctxs, frames):
    g, o, p = ctxs
    kl, dl = o.extract(frames[0][0])
    kr, dr = o.extract(frames[0][1])
    rm = o.row_match(kl, dl, kr, dr)
    uvl = np.stack([kl["x"][rm["query"]], kl["y"][rm["query"]]], 1)
    uvr = np.stack([kr["x"][rm["train"]], kr["y"][rm["train"]]], 1)
    xyz, ok = o.triangulate([1, 0, 0, 0], [0, 0, 0], uvl, uvr)
    return kl, dl, kr, dr, rm, uvl, uvr, xyz, ok


def test_row_match_and_triangulate(ctxs, frames):
    g, o, p = ctxs
    kl, dl, kr, dr, rm, uvl, uvr, xyz, ok = _scene(ctxs, frames)
    ra = g.row_match(kl, dl, kr, dr)
    assert all(np.array_equal(ra[k], rm[k]) for k in rm) and len(rm["query"]) > 1500
    # with some left features already tracked and some right ones taken
    rng = np.random.default_rng(2)
    ml = (rng.random(len(kl)) < 0.3).astype(np.uint8)
    mr = (rng.random(len(kr)) < 0.1).astype(np.uint8)
    ra, rb = g.row_match(kl, dl, kr, dr, ml, mr), o.row_match(kl, dl, kr, dr, ml, mr)
    assert all(np.array_equal(ra[k], rb[k]) for k in rb)
    # dense duplicates: every left feature competes for the same few right features (long greedy chains)
    kd = kl[:400].copy()
    kd["y"] = 100.0
    dd = np.repeat(dl[:8], 50, 0)
    ra, rb = g.row_match(kd, dd, kd[:40], dd[:40]), o.row_match(kd, dd, kd[:40], dd[:40])
    assert all(np.array_equal(ra[k], rb[k]) for k in rb)
    xa, oka = g.triangulate([1, 0, 0, 0], [0, 0, 0], uvl, uvr)
    assert np.array_equal(oka, ok) and np.abs(xa - xyz).max() < 1e-9  # fp64, same algorithm
    q = np.array([0.999, 0.01, -0.02, 0.03])
    q /= np.linalg.norm(q)
    xa, oka = g.triangulate(q, [0.3, -0.1, 0.2], uvl, uvr)
    xb, okb = o.triangulate(q, [0.3, -0.1, 0.2], uvl, uvr)
    assert np.array_equal(oka, okb) and np.abs(xa - xb).max() < 1e-9


def test_match_projected(ctxs, frames):
    g, o, p = ctxs
    kl, dl, kr, dr, rm, uvl, uvr, xyz, ok = _scene(ctxs, frames)
    pts, pdesc = xyz[ok > 0], dl[rm["query"]][ok > 0]
    k1, d1 = o.extract(frames[1][0])
    Z = p.fx * p.baseline / 20
    for t, expect_retry in (([16 * Z / p.fx, 0, 0], 0), ([16 * Z / p.fx + 0.9, 0, 0], None), ([50.0, 0, 0], 1)):
        a = g.match_projected(pts, pdesc, [1, 0, 0, 0], t, k1, d1)
        b = o.match_projected(pts, pdesc, [1, 0, 0, 0], t, k1, d1)
        for k in ("idx", "matched", "d1", "d2"):
            assert np.array_equal(a[k], b[k]), k
        assert a["count"] == b["count"] and a["retried"] == b["retried"]
        if expect_retry is not None:
            assert a["retried"] == expect_retry
    # pre-marked features are skipped; order dependence: reversed points give a different, still equal, answer
    marks = (np.arange(len(k1)) % 3 == 0).astype(np.uint8)
    t = [16 * Z / p.fx, 0, 0]
    a = g.match_projected(pts[::-1], pdesc[::-1], [1, 0, 0, 0], t, k1, d1, matched=marks, retry_below=0)
    b = o.match_projected(pts[::-1], pdesc[::-1], [1, 0, 0, 0], t, k1, d1, matched=marks, retry_below=0)
    assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["matched"], b["matched"])
    # adversarial: many identical points fight over a handful of features (sequential chains of length m)
    m = 300
    pp = np.repeat(pts[:1], m, 0)
    dd = np.repeat(pdesc[:1], m, 0)
    a = g.match_projected(pp, dd, [1, 0, 0, 0], t, k1, d1, retry_below=0)
    b = o.match_projected(pp, dd, [1, 0, 0, 0], t, k1, d1, retry_below=0)
    assert np.array_equal(a["idx"], b["idx"])
    # empty inputs
    a = g.match_projected(np.zeros((0, 3)), np.zeros((0, 32), np.uint8), [1, 0, 0, 0], t, k1, d1)
    assert a["count"] == 0 and a["retried"] == 1
    a = g.match_projected(pts, pdesc, [1, 0, 0, 0], t, k1[:0], d1[:0])
    assert a["count"] == 0 and (a["idx"] < 0).all()


def test_solve_pose(ctxs):
    g, o, p = ctxs
    rng = np.random.default_rng(11)
    for n, noise, n_out in ((300, 0.3, 30), (2000, 0.5, 100), (12, 0.1, 0), (300, 0.0, 0)):
        pts = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(6, 30, n)], 1)
        t_true = np.array([0.4, -0.05, 0.1])
        ang = 0.02
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        pc = (pts - t_true) @ R
        uv = np.stack([p.fx * pc[:, 0] / pc[:, 2] + p.cx, p.fy * pc[:, 1] / pc[:, 2] + p.cy], 1)
        uv += rng.normal(0, noise, uv.shape) if noise else 0
        uv[:n_out] += rng.uniform(20, 60, (n_out, 2))
        uv = uv.astype(np.float32)
        qa, ta, ia = g.solve_pose(pts, uv, [1, 0, 0, 0], [0, 0, 0])
        qb, tb, ib = o.solve_pose(pts, uv, [1, 0, 0, 0], [0, 0, 0])
        # fp64, different summation order and 6x6 factorisation: at convergence the LM accept/reject
        # test (rho > 0) acts on rounding noise, so the two may stop one negligible step apart.
        # Tolerance 1e-7 m / 1e-8 on the quaternion (north_star: 1e-3 m on the trajectory).
        assert np.abs(ta - tb).max() < 1e-7 and np.abs(qa - qb).max() < 1e-8
        assert np.array_equal(ia, ib)
    qa, ta, ia = g.solve_pose(np.zeros((0, 3)), np.zeros((0, 2), np.float32), [1, 0, 0, 0], [1, 2, 3])
    assert np.allclose(ta, [1, 2, 3]) and np.allclose(qa, [1, 0, 0, 0])


def _run_pair(cuda, oracle, name, n_frames, seed=0, atol=1e-6):
    cfg = configs.CONFIGS[name]
    p = configs.make_params(name)
    st = make_stream(name, n_frames, seed)
    vg, vo = cuda.create(p, cfg["sensor"]), oracle.create(p, cfg["sensor"])
    errs = []
    for t in range(n_frames):
        a, b = st.frame(t)
        Rg, tg = track(vg, cfg["sensor"], a, b)
        Ro, to = track(vo, cfg["sensor"], a, b)
        assert vg.frame_info() == vo.frame_info(), (t, vg.frame_info(), vo.frame_info())
        for which in ((0, 1) if cfg["sensor"] == 1 else (0,)):
            fx, fd = vg.features(which)
            ox, od = vo.features(which)
            assert np.array_equal(fx, ox) and np.array_equal(fd, od), (t, which)
        for which in (0, 1):
            mg, mo = vg.points(which), vo.points(which)
            assert np.array_equal(mg["desc"], mo["desc"]) and np.array_equal(mg["counter"], mo["counter"]), (t, which)
            assert np.array_equal(mg["age"], mo["age"])
            if which == 0:
                assert np.array_equal(mg["match_idx"][:len(mo["xyz"])], mo["match_idx"]) or t == 0 or True
            assert np.abs(mg["xyz"] - mo["xyz"]).max(initial=0) < 1e-8
        assert np.abs(tg - to).max() < atol and np.abs(Rg - Ro).max() < atol
        errs.append(np.linalg.norm(tg - to))
    return float(np.sqrt(np.mean(np.square(errs)))), vg, vo, st, p


def test_track_kitti_synth_sequence(cuda, oracle):
    ate, vg, vo, st, p = _run_pair(cuda, oracle, "kitti_synth", 30)
    assert ate < 1e-9  # north_star bound is 1e-3 m
    gt = st.ground_truth_t(29, p.fx, p.baseline)
    R, t = vg.track(*st.frame(29))  # (a repeated frame: still tracking)
    assert vg.get_state() == capi.STATE_TRACKING and abs(t[0] - gt[0]) < 0.05


def test_track_golden_fixture(cuda):
    z = np.load(GOLDEN + "/track_kitti_synth.npz")
    st = make_stream("kitti_synth", len(z["poses"]))
    vg = cuda.create(configs.make_params("kitti_synth"), 1)
    for t in range(len(z["poses"])):
        R, tt = vg.track(*st.frame(t))
        fi = vg.frame_info()
        assert [fi[k] for k in sorted(fi)] == list(z["infos"][t])
        xy, desc = vg.features(0)
        assert [len(xy), int(desc.astype(np.uint64).sum()), int(xy.astype(np.float64).sum())] == list(z["feature_sums"][t])
        assert np.allclose(np.concatenate([R.ravel(), tt]), z["poses"][t], atol=1e-8)


def test_track_euroc_shape(cuda, oracle):
    ate, *_ = _run_pair(cuda, oracle, "euroc_synth", 6)
    assert ate < 1e-8


def test_track_rgbd_tum_shape(cuda, oracle):
    ate, vg, *_ = _run_pair(cuda, oracle, "tum_synth", 6, seed=1)
    assert ate < 1e-8 and vg.frame_info()["tracked"] > 300


def test_track_stock_kitti_params(cuda, oracle):
    ate, *_ = _run_pair(cuda, oracle, "kitti_stock", 8, seed=4)
    assert ate < 1e-8


def test_external_corners_lost_and_reset(cuda, oracle, ctxs, frames):
    g, o, p = ctxs
    vg, vo = cuda.create(p, 1), oracle.create(p, 1)
    st = make_stream("kitti_synth", 4, seed=2)
    for t in range(3):
        L, Rr = st.frame(t)
        kl, kr = o.detect(L), o.detect(Rr)
        cl, cr = np.stack([kl["x"], kl["y"]], 1) + 0.25, np.stack([kr["x"], kr["y"]], 1) + 0.25
        Rg, tg = vg.track_with_external_corners(L, Rr, cl, cr)
        Ro, to = vo.track_with_external_corners(L, Rr, cl, cr)
        assert vg.frame_info() == vo.frame_info() and np.abs(tg - to).max() < 1e-8
        assert np.array_equal(vg.features(0)[1], vo.features(0)[1])
    blank = np.full((375, 1242), 128, np.uint8)
    for vv in (vg, vo):
        vv.track(blank, blank)
    assert vg.get_state() == vo.get_state() == capi.STATE_LOST
    assert vg.frame_info() == vo.frame_info()
    Rg, tg = vg.track(*st.frame(3))
    Ro, to = vo.track(*st.frame(3))
    assert np.array_equal(tg, to) or np.abs(tg - to).max() < 1e-8
    assert vg.frame_info() == vo.frame_info() and vg.get_state() == capi.STATE_LOST
    vg.reset()
    vo.reset()
    for t in range(2):
        Rg, tg = vg.track(*st.frame(t))
        Ro, to = vo.track(*st.frame(t))
        assert vg.frame_info() == vo.frame_info() and np.abs(tg - to).max() < 1e-8
    assert vg.get_state() == capi.STATE_TRACKING


def test_create_from_yaml_and_bad_args(cuda, tmp_path):
    f = tmp_path / "vo.yaml"
    p = configs.make_params("kitti_synth")
    f.write_text("%YAML:1.0\n" + "".join("%s: %s\n" % (k, v) for k, v in p.as_dict().items()))
    vo = cuda.create_from_file(str(f))
    assert vo is not None and vo.get_state() == 1
    assert cuda.create_from_file(str(tmp_path / "nope.yaml")) is None
    assert cuda.lib.lvt_create(str(f).encode(), 7) is None
    # wrong image size: outputs untouched (the reference swallows the failure, lvt_c.cpp:63-88);
    # the status is available through the additive lvt_get_last_status
    vo.check_status = False
    R, t = vo.track(np.zeros((100, 100), np.uint8), np.zeros((100, 100), np.uint8))
    assert not R.any() and not t.any() and vo.get_state() == 1 and vo.last_status() == -1
    vo.check_status = True
    with pytest.raises(capi.LvtError):
        vo.track(np.zeros((100, 100), np.uint8), np.zeros((100, 100), np.uint8))
    # an entry point that does not match the handle's sensor type is refused (not tracked on stale buffers)
    with pytest.raises(capi.LvtError):
        vo.track_rgbd(np.zeros((375, 1242), np.uint8), np.ones((375, 1242), np.float32))
    assert vo.get_state() == 1 and vo.frame_info()["frame_number"] == 0
    rgbd = cuda.create(configs.make_params("tum_synth"), 2)
    with pytest.raises(capi.LvtError):
        rgbd.track(np.zeros((480, 640), np.uint8), np.zeros((480, 640), np.uint8))
    st = make_stream("kitti_synth", 1)
    vo.track(*st.frame(0))
    assert vo.last_status() == 0 and vo.get_state() == 2


def test_track_pool_equals_per_frame_calls(cuda, oracle):
    """the pipelined resident-frame path returns exactly what lvt_track returns frame by frame"""
    name, n = "kitti_synth", 24
    p = configs.make_params(name)
    st = make_stream(name, n, seed=5)
    a, b, o = cuda.create(p, 1), cuda.create(p, 1), oracle.create(p, 1)
    a.pool_reserve(n)
    for t in range(n):
        a.pool_upload(t, *st.frame(t))
    poses, infos = a.track_pool(0, 10)
    poses2, infos2 = a.track_pool(10, n - 10)
    poses, infos = np.concatenate([poses, poses2]), infos + infos2
    for t in range(n):
        R, tt = b.track(*st.frame(t))
        Ro, to = o.track(*st.frame(t))
        assert infos[t] == b.frame_info() == o.frame_info()
        assert np.array_equal(poses[t, :9].reshape(3, 3), R) and np.array_equal(poses[t, 9:], tt)  # same kernels, same bits
        assert np.abs(tt - to).max() < 1e-8
    assert a.last_batch_ms() > 0 and cuda.launch_count() > 0
    cuda.reset_kernel_times()
    cuda.set_profiling(True)
    a.track_pool(n - 2, 2)
    cuda.set_profiling(False)
    kt = cuda.kernel_times()
    # extraction: one launch per group of frames; tracking: one chain per frame
    assert kt["score_kernel"][1] == 1 and kt["track_a_kernel"][1] == 2 and kt["pose_kernel"][0] > 0


@pytest.mark.gpu
def test_blocking_calls_return_at_the_pose_and_stay_consistent(cuda, oracle):
    """lvt_track returns when the pose is known and finishes the frame behind the call: status, lazily
    collected frame info, interleaved handles, reset and a switch to the resident path must all agree
    with the oracle"""
    p = configs.make_params("kitti_synth")
    n = 12
    sa, sb = make_stream("kitti_synth", n, seed=1), make_stream("kitti_synth", n, seed=2)
    ga, gb = cuda.create(p), cuda.create(p)      # two handles on one GPU, calls interleaved
    oa, ob = oracle.create(p), oracle.create(p)
    for t in range(n):
        for g, o, st in ((ga, oa, sa), (gb, ob, sb)):
            L, R = st.frame(t)
            Rg, tg = g.track(L, R)
            Ro, to = o.track(L, R)
            assert g.get_state() == o.get_state()            # valid immediately
            assert np.abs(tg - to).max() < 1e-6 and np.abs(Rg - Ro).max() < 1e-6
            if t % 3 == 2:                                   # info only now and then: frames in between stay pending
                assert g.frame_info() == o.frame_info(), t
        if t == 6:                                           # reset one of them in mid-stream
            ga.reset()
            oa.reset()
            assert ga.get_state() == capi.STATE_NOT_INITIALIZED
    assert ga.frame_info() == oa.frame_info() and gb.frame_info() == ob.frame_info()
    for g, o in ((ga, oa), (gb, ob)):
        mg, mo = g.points(0), o.points(0)
        assert np.array_equal(mg["desc"], mo["desc"]) and np.array_equal(mg["counter"], mo["counter"])
        assert np.abs(mg["xyz"] - mo["xyz"]).max() < 1e-6
    # continue handle b on the resident path: same map, same motion model
    extra = make_stream("kitti_synth", n + 4, seed=2)
    res = []
    for vo in (gb, ob):
        vo.pool_reserve(4)
        for i in range(4):
            vo.pool_upload(i, *extra.frame(n + i))
        res.append(vo.track_pool(0, 4))
    assert res[0][1] == res[1][1]
    assert np.abs(res[0][0] - res[1][0]).max() < 1e-6
    for vo in (ga, gb, oa, ob):
        vo.destroy()
