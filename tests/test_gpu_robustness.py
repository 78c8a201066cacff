"""Behaviour of the CUDA library around its fixed resources, against the CPU oracle (real GPU):
growth of the point stores (the reference's vectors grow, lvt/src/lvt_local_map.cpp:331-353),
independent handles on independent host threads (lvt/src/lvt_c.cpp:33-148 has no shared state),
a second BRIEF table injected through lvt_set_brief_pairs, and a C++ program linked against
liblvt_b200.so through include/lvt_system.hpp."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from helpers import ROOT, capi, configs, make_stream, track

pytestmark = pytest.mark.gpu


def _trajectory(lib, name, n, seed, **overrides):
    cfg = configs.CONFIGS[name]
    p = configs.make_params(name, **overrides)
    st = make_stream(name, n, seed)
    vo = lib.create(p, cfg["sensor"])
    out = []
    for t in range(n):
        R, tt = track(vo, cfg["sensor"], *st.frame(t))
        out.append((R, tt, vo.frame_info()))
    pts = vo.points(0)
    cap = vo.point_capacity()
    vo.destroy()
    return out, pts, cap


def _same(a, b, tol=1e-6):
    for t, ((Ra, ta, ia), (Rb, tb, ib)) in enumerate(zip(a, b)):
        assert ia == ib, (t, {k: (ia[k], ib[k]) for k in ia if ia[k] != ib[k]})
        assert np.abs(ta - tb).max() < tol and np.abs(Ra - Rb).max() < tol, t


@pytest.mark.parametrize("name,n,seed,cap", [("kitti_synth", 25, 0, 1024), ("tum_synth", 20, 1, 2048)])
def test_point_stores_grow_like_the_reference_vectors(cuda, oracle, name, n, seed, cap, monkeypatch):
    """start with point stores far smaller than the map becomes: frames that could overflow are refused on the
    device, the stores double, the frame runs again -- results identical to the oracle's unbounded vectors"""
    monkeypatch.setenv("LVT_B200_POINT_CAP", str(cap))
    got, gpts, gcap = _trajectory(cuda, name, n, seed)
    monkeypatch.delenv("LVT_B200_POINT_CAP")
    ref, opts, _ = _trajectory(oracle, name, n, seed)
    _same(got, ref)
    assert gcap > cap and gcap >= got[-1][2]["map_points_after"]
    assert np.array_equal(gpts["desc"], opts["desc"]) and np.array_equal(gpts["counter"], opts["counter"])
    assert np.abs(gpts["xyz"] - opts["xyz"]).max() < 1e-6


def test_point_stores_grow_inside_a_resident_batch(cuda, monkeypatch):
    """lvt_track_pool: the refused frame and everything behind it run again after the growth"""
    n = 16
    p = configs.make_params("kitti_synth")
    st = make_stream("kitti_synth", n, seed=7)
    monkeypatch.setenv("LVT_B200_POINT_CAP", "1024")
    a = cuda.create(p, 1)
    monkeypatch.delenv("LVT_B200_POINT_CAP")
    b = cuda.create(p, 1)
    a.pool_reserve(n)
    for t in range(n):
        a.pool_upload(t, *st.frame(t))
    poses, infos = a.track_pool(0, n)
    for t in range(n):
        R, tt = b.track(*st.frame(t))
        assert infos[t] == b.frame_info(), t
        assert np.array_equal(poses[t, :9].reshape(3, 3), R) and np.array_equal(poses[t, 9:], tt)
    assert a.point_capacity() > 1024 and b.point_capacity() >= 32768


def test_distinct_handles_on_distinct_threads(cuda, oracle):
    """four handles, four host threads, different configurations, all at once; each must reproduce the oracle"""
    jobs = [("kitti_synth", 40, 0), ("kitti_synth", 40, 1), ("euroc_synth", 20, 2), ("tum_synth", 30, 1)]
    ref = [_trajectory(oracle, *j)[0] for j in jobs]
    got, errs = [None] * len(jobs), []

    def work(i):
        try:
            got[i] = _trajectory(cuda, *jobs[i])[0]  # ctypes releases the GIL inside every call
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for g, r in zip(got, ref):
        _same(g, r)


def test_batched_engine_on_distinct_threads(cuda, oracle):
    """the batched engine of several handles at once (each with its own tracking / side / extraction streams, a CTA of
    every handle waiting on the device for that handle's map maintenance): every handle reproduces the oracle and its
    own blocking calls, whatever the others are doing"""
    jobs = [("kitti_synth", 36, 4), ("kitti_synth", 36, 5), ("euroc_synth", 18, 3), ("tum_synth", 24, 2), ("kitti_synth", 36, 6)]
    ref = [_trajectory(oracle, *j)[0] for j in jobs]
    got, errs = [None] * len(jobs), []

    def work(i):
        try:
            name, n, seed = jobs[i]
            cfg = configs.CONFIGS[name]
            st = make_stream(name, n, seed)
            fr = [st.frame(t) for t in range(n)]
            vo = cuda.create(configs.make_params(name), cfg["sensor"])
            out = []
            for lo in range(0, n, 12):  # three calls per handle
                a, b = [f[0] for f in fr[lo:lo + 12]], [f[1] for f in fr[lo:lo + 12]]
                poses, infos = vo.track_batch(a, b)  # (depth images select lvt_track_batch_rgbd)
                out += [(poses[k, :9].reshape(3, 3), poses[k, 9:], infos[k]) for k in range(len(infos))]
            got[i] = out
            vo.destroy()
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for g, r in zip(got, ref):
        _same(g, r)


def test_second_brief_table_through_set_brief_pairs(cuda, oracle):
    """the only way a user gets OpenCV-exact descriptor bits is to inject opencv_contrib's table: the
    injection route itself is tested with a second table (other seed) in both libraries, with a handle that
    was created BEFORE the switch (it must pick the table up) and one created after"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_brief_pairs
    rng = np.random.default_rng(99)
    other = np.clip(np.rint(rng.normal(0.0, 48.0 / 5.0, (256, 4))), -24, 24).astype(np.int8)
    assert not np.array_equal(other, np.array(gen_brief_pairs.pairs(), np.int8))
    p = configs.make_params("kitti_synth")
    st = make_stream("kitti_synth", 6, seed=11)
    early_g, early_o = cuda.create(p, 1), oracle.create(p, 1)
    early_g.track(*st.frame(0))
    early_o.track(*st.frame(0))
    d_default = early_g.features(0)[1]
    assert np.array_equal(d_default, early_o.features(0)[1])
    try:
        assert cuda.set_brief_pairs(other) == 0 and oracle.set_brief_pairs(other) == 0
        assert cuda.set_brief_pairs(np.full((256, 4), 25)) != 0  # offsets beyond the 48x48 patch are refused
        early_g.reset()
        early_o.reset()
        late_g, late_o = cuda.create(p, 1), oracle.create(p, 1)
        for t in range(6):
            for g, o in ((early_g, early_o), (late_g, late_o)):
                Rg, tg = g.track(*st.frame(t))
                Ro, to = o.track(*st.frame(t))
                assert g.frame_info() == o.frame_info() and np.abs(tg - to).max() < 1e-8
                for which in (0, 1):
                    assert np.array_equal(g.features(which)[1], o.features(which)[1])
            if t == 0:
                d_other = late_g.features(0)[1]
                assert d_other.shape == d_default.shape and (d_other != d_default).mean() > 0.3  # the bits did change
    finally:
        cuda.set_brief_pairs(None)
        oracle.set_brief_pairs(None)
    fresh = cuda.create(p, 1)
    fresh.track(*st.frame(0))
    assert np.array_equal(fresh.features(0)[1], d_default)


def test_cxx_lvt_system_mirror_links_the_cuda_library(cuda, tmp_path):
    """include/lvt_system.hpp over liblvt_b200.so from a C++11 program (what a maintainer of the reference builds)"""
    exe = str(tmp_path / "cxx_smoke_cuda")
    so_dir = os.path.dirname(cuda.path)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cxx_system_smoke.cpp"), "-o", exe, "-L", so_dir, "-llvt_b200",
                    "-Wl,-rpath," + so_dir], check=True)
    r = subprocess.run([exe, "320", "200"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
