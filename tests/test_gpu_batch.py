"""The batched engine (lvt_track_pool on resident frames, lvt_track_batch* on frames in host memory) returns
exactly what the blocking per-frame calls return: same kernels, same bits -- and what the CPU oracle returns
within the usual bounds.  Group sizes that do not divide the batch, pageable and page-locked buffers, RGB-D,
rectified raw frames, growth of the point stores inside a batch."""
import numpy as np
import pytest

from helpers import capi, configs, make_stream, track

pytestmark = pytest.mark.gpu


def _frames(name, n, seed):
    st = make_stream(name, n, seed)
    fr = [st.frame(t) for t in range(n)]
    return [f[0] for f in fr], [f[1] for f in fr]


def _per_frame(lib, name, a, b, **overrides):
    cfg = configs.CONFIGS[name]
    vo = lib.create(configs.make_params(name, **overrides), cfg["sensor"])
    poses, infos = [], []
    for x, y in zip(a, b):
        R, t = track(vo, cfg["sensor"], x, y)
        poses.append(np.concatenate([R.ravel(), t]))
        infos.append(vo.frame_info())
    return np.array(poses), infos, vo


@pytest.mark.parametrize("name,n,seed", [("kitti_synth", 23, 3), ("tum_synth", 14, 1), ("euroc_synth", 9, 2)])
def test_host_batch_equals_blocking_calls(cuda, oracle, name, n, seed):
    cfg = configs.CONFIGS[name]
    a, b = _frames(name, n, seed)
    ref_p, ref_i, ref = _per_frame(cuda, name, a, b)
    orc_p, orc_i, _ = _per_frame(oracle, name, a, b)
    vo = cuda.create(configs.make_params(name), cfg["sensor"])
    # three calls of awkward sizes: groups that do not fill, a single frame, the rest
    cuts = [0, 7, 8, n]
    poses, infos = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        p, i = vo.track_batch(a[lo:hi], b[lo:hi])
        poses.append(p)
        infos += i
    poses = np.concatenate(poses)
    assert infos == ref_i == orc_i
    assert np.array_equal(poses, ref_p)                      # same kernels in the same order: same bits
    assert np.abs(poses - orc_p).max() < 1e-6
    assert vo.get_state() == ref.get_state() and vo.last_status() == 0
    for which in ((0, 1) if cfg["sensor"] == 1 else (0,)):   # the last frame's features / the map are the same too
        assert np.array_equal(vo.features(which)[1], ref.features(which)[1])
    assert np.array_equal(vo.points(0)["desc"], ref.points(0)["desc"])
    # and the sequence continues with a blocking call
    st = make_stream(name, n + 1, seed)
    Rv, tv = track(vo, cfg["sensor"], *st.frame(n))
    Rr, tr = track(ref, cfg["sensor"], *st.frame(n))
    assert np.array_equal(tv, tr) and vo.frame_info() == ref.frame_info()


def test_host_batch_from_page_locked_buffers(cuda):
    """page-locked caller buffers are read by the DMA engine directly (no staging copy): same result"""
    n = 10
    a, b = _frames("kitti_synth", n, 4)
    H, W = a[0].shape
    pinned = cuda.pinned_empty((n, 2, H, W))
    for i in range(n):
        pinned[i, 0], pinned[i, 1] = a[i], b[i]
    p = configs.make_params("kitti_synth")
    v1, v2, v3 = cuda.create(p, 1), cuda.create(p, 1), cuda.create(p, 1)
    p1, i1 = v1.track_batch(a, b)
    p2, i2 = v2.track_batch([pinned[i, 0] for i in range(n)], [pinned[i, 1] for i in range(n)])
    assert np.array_equal(p1, p2) and i1 == i2
    for i in range(n):                                       # the blocking call accepts them as well
        R, t = v3.track(pinned[i, 0], pinned[i, 1])
        assert np.array_equal(np.concatenate([R.ravel(), t]), p1[i])


def test_resident_rgbd_pool(cuda, oracle):
    n = 12
    a, b = _frames("tum_synth", n, 1)
    ref_p, ref_i, _ = _per_frame(oracle, "tum_synth", a, b)
    res = []
    for lib in (cuda, oracle):
        vo = lib.create(configs.make_params("tum_synth"), 2)
        vo.pool_reserve(n)
        for t in range(n):
            vo.pool_upload(t, a[t], b[t])
        p1, i1 = vo.track_pool(0, 5)
        p2, i2 = vo.track_pool(5, n - 5)
        res.append((np.concatenate([p1, p2]), i1 + i2))
    assert res[0][1] == res[1][1] == ref_i
    assert np.abs(res[0][0] - ref_p).max() < 1e-6 and np.abs(res[1][0] - ref_p).max() == 0


def test_host_batch_grows_the_point_stores(cuda, monkeypatch):
    n = 12
    a, b = _frames("kitti_synth", n, 6)
    ref_p, ref_i, _ = _per_frame(cuda, "kitti_synth", a, b)
    monkeypatch.setenv("LVT_B200_POINT_CAP", "1024")
    vo = cuda.create(configs.make_params("kitti_synth"), 1)
    monkeypatch.delenv("LVT_B200_POINT_CAP")
    p, i = vo.track_batch(a, b)
    assert i == ref_i and np.array_equal(p, ref_p) and vo.point_capacity() > 1024


@pytest.mark.parametrize("group", ["1", "8"])
def test_group_size_does_not_change_results(cuda, group, monkeypatch):
    n = 19
    a, b = _frames("kitti_synth", n, 8)
    ref_p, ref_i, _ = _per_frame(cuda, "kitti_synth", a, b)
    monkeypatch.setenv("LVT_B200_GROUP", group)
    vo = cuda.create(configs.make_params("kitti_synth"), 1)
    p, i = vo.track_batch(a, b)
    monkeypatch.delenv("LVT_B200_GROUP")
    assert i == ref_i and np.array_equal(p, ref_p)


def test_bad_arguments(cuda):
    a, b = _frames("kitti_synth", 2, 0)
    vo = cuda.create(configs.make_params("kitti_synth"), 1)
    with pytest.raises(capi.LvtError):
        vo.track_batch(a, [np.ones(a[0].shape, np.float32)] * 2)   # the RGB-D entry point on a stereo handle
    with pytest.raises(capi.LvtError):
        vo.track_batch([x[:100] for x in a], [x[:100] for x in b])  # wrong size
    assert vo.last_status() == -1 and vo.get_state() == 1
    p, i = vo.track_batch(a, b)
    assert i[1]["state"] == 2 and i[1]["frame_number"] == 2


def test_batch_with_lost_frames_and_reset(cuda, oracle):
    """State changes INSIDE a batch: the engine launches a frame's early map pass before the previous frame's map
    maintenance is through, so every transition (first frame -> tracking, tracking -> lost on a blank frame, lost
    frames, reset -> first frame again) has to fall back to the plain order on the device.  Against the oracle and
    the blocking calls, frame by frame."""
    name, n = "kitti_synth", 18
    a, b = _frames(name, n, seed=5)
    blank = np.full_like(a[0], 128)
    a[9], b[9] = blank, blank                      # no corners: lost (lvt_system.cpp:267-272), and stays lost
    ref_p, ref_i, ref = _per_frame(cuda, name, a, b)
    orc_p, orc_i, _ = _per_frame(oracle, name, a, b)
    vo = cuda.create(configs.make_params(name), 1)
    poses, infos = vo.track_batch(a, b)
    assert infos == ref_i == orc_i
    assert [i["state"] for i in infos[8:11]] == [capi.STATE_TRACKING, capi.STATE_LOST, capi.STATE_LOST]
    assert np.array_equal(poses, ref_p) and np.abs(poses - orc_p).max() < 1e-6
    # reset between two batches: the next batch starts with a first frame again
    for v in (vo, ref):
        v.reset()
    p2, i2 = vo.track_batch(a[:6], b[:6])
    r2 = [np.concatenate([x.ravel(), y]) for x, y in (track(ref, 1, l, r) for l, r in zip(a[:6], b[:6]))]
    assert np.array_equal(p2, np.array(r2)) and i2[-1]["state"] == capi.STATE_TRACKING
    assert i2 == ref_i[:6]


def test_batch_retry_pass_with_few_matches(cuda, oracle):
    """fewer than 50 matches in the map pass: the radius x2 retry (lvt_local_map.cpp:173-199) fires inside the rest of
    track_a, after the early part; a tiny map (max_keypoints_per_cell = 6) gets there"""
    name, n = "kitti_synth", 10
    a, b = _frames(name, n, seed=7)
    over = dict(max_keypoints_per_cell=6)
    ref_p, ref_i, _ = _per_frame(cuda, name, a, b, **over)
    orc_p, orc_i, _ = _per_frame(oracle, name, a, b, **over)
    vo = cuda.create(configs.make_params(name, **over), 1)
    poses, infos = vo.track_batch(a, b)
    assert infos == ref_i == orc_i
    assert any(i["retried_matching"] for i in infos) and not all(i["retried_matching"] for i in infos[1:])
    assert np.array_equal(poses, ref_p) and np.abs(poses - orc_p).max() < 1e-6
