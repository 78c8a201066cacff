#!/usr/bin/env python
"""Development probe (run under gpurun): drives every seam of the CUDA library next to the CPU
oracle and prints where they differ.  Not a test; tests/ holds the asserted versions."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lvt_b200  # noqa: E402
from lvt_b200 import capi, configs, synth  # noqa: E402


def kp_equal(a, b):
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("x", "y", "response"))


def report(name, ok, extra=""):
    print("%-34s %s %s" % (name, "OK  " if ok else "FAIL", extra), flush=True)


def main():
    lib = lvt_b200.load()
    orc = capi.Library(os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so"))
    name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
    cfg = configs.CONFIGS[name]
    p = configs.make_params(name)
    nf = 12
    st = synth.StereoStream(n_frames=nf, seed=0, **cfg["stream"])
    L, R = st.frame(0)
    g, o = lib.context(p), orc.context(p)

    for th, nm in ((25, True), (25, False), (12, True), (60, True)):
        a, b = g.agast(L[:200, :300], th, nm), o.agast(L[:200, :300], th, nm)
        report("agast th=%d nms=%d" % (th, nm), kp_equal(a, b), "n=%d/%d" % (len(a), len(b)))
    rng = np.random.default_rng(5)
    q = (rng.integers(0, 8, (160, 200)) * 32).astype(np.uint8)
    a, b = g.agast(q, 20, True), o.agast(q, 20, True)
    report("agast tie-heavy", kp_equal(a, b), "n=%d/%d" % (len(a), len(b)))
    nz = rng.integers(0, 256, (120, 130)).astype(np.uint8)
    a, b = g.agast(nz, 10, True), o.agast(nz, 10, True)
    report("agast noise", kp_equal(a, b), "n=%d/%d" % (len(a), len(b)))
    if not kp_equal(a, b):
        sa = set(zip(a["x"], a["y"], a["response"]))
        sb = set(zip(b["x"], b["y"], b["response"]))
        print("   only gpu", sorted(sa - sb)[:8], "only cpu", sorted(sb - sa)[:8])

    a, b = g.detect(L), o.detect(L)
    report("detect (tiles+anms)", kp_equal(a, b), "n=%d/%d" % (len(a), len(b)))
    if not kp_equal(a, b):
        sa = set(zip(a["x"], a["y"]))
        sb = set(zip(b["x"], b["y"]))
        print("   same set:", sa == sb, "only gpu", len(sa - sb), "only cpu", len(sb - sa))
        n = min(len(a), len(b))
        bad = [i for i in range(n) if (a[i]["x"], a[i]["y"]) != (b[i]["x"], b[i]["y"])]
        print("   first order mismatch at", bad[:5])
    ka, da = g.brief(L, b)
    kb, db = o.brief(L, b)
    report("brief", kp_equal(ka, kb) and np.array_equal(da, db), "n=%d/%d bitdiff=%s" % (
        len(ka), len(kb), int(np.unpackbits(da ^ db).sum()) if da.shape == db.shape else "shape"))
    ka, da = g.extract(L)
    kb, db = o.extract(L)
    report("extract", kp_equal(ka, kb) and np.array_equal(da, db), "n=%d/%d" % (len(ka), len(kb)))
    kr, dr = o.extract(R)

    # row match
    ra, rb = g.row_match(kb, db, kr, dr), o.row_match(kb, db, kr, dr)
    ok = all(np.array_equal(ra[k], rb[k]) for k in ra)
    report("row_match", ok, "n=%d/%d" % (len(ra["query"]), len(rb["query"])))
    # triangulate
    qi = np.array([1.0, 0, 0, 0])
    ti = np.zeros(3)
    uvl = np.stack([kb["x"][rb["query"]], kb["y"][rb["query"]]], 1)
    uvr = np.stack([kr["x"][rb["train"]], kr["y"][rb["train"]]], 1)
    xa, va = g.triangulate(qi, ti, uvl, uvr)
    xb, vb = o.triangulate(qi, ti, uvl, uvr)
    report("triangulate", np.array_equal(va, vb) and np.abs(xa - xb).max() < 1e-9,
           "valid=%d/%d maxdiff=%.2e" % (va.sum(), vb.sum(), np.abs(xa - xb).max()))
    pts = xb[vb > 0]
    pdesc = db[rb["query"]][vb > 0]
    # projected matching against frame 1
    L1, _ = st.frame(1)
    k1, d1 = o.extract(L1)
    Z = p.fx * p.baseline / st.d
    t1 = np.array([st.s * Z / p.fx, 0, 0])
    ma, mb = g.match_projected(pts, pdesc, qi, t1, k1, d1), o.match_projected(pts, pdesc, qi, t1, k1, d1)
    ok = all(np.array_equal(ma[k], mb[k]) for k in ("idx", "matched")) and ma["count"] == mb["count"]
    report("match_projected", ok, "count=%d/%d idxdiff=%d d1eq=%s" % (
        ma["count"], mb["count"], int((ma["idx"] != mb["idx"]).sum()), np.array_equal(ma["d1"], mb["d1"])))
    # far-off pose: forces the radius x2 retry
    t2 = t1 + np.array([0.9, 0, 0])
    ma2, mb2 = g.match_projected(pts, pdesc, qi, t2, k1, d1), o.match_projected(pts, pdesc, qi, t2, k1, d1)
    ok = np.array_equal(ma2["idx"], mb2["idx"]) and ma2["retried"] == mb2["retried"]
    report("match_projected retry", ok, "count=%d/%d retried=%d/%d" % (ma2["count"], mb2["count"], ma2["retried"], mb2["retried"]))
    # pose
    sel = mb["idx"] >= 0
    uv = np.stack([k1["x"][mb["idx"][sel]], k1["y"][mb["idx"][sel]]], 1)
    qa, ta, ia = g.solve_pose(pts[sel], uv, qi, t1 * 0.8)
    qb, tb, ib = o.solve_pose(pts[sel], uv, qi, t1 * 0.8)
    report("solve_pose", np.abs(ta - tb).max() < 1e-8 and np.array_equal(ia, ib),
           "dt=%.2e dq=%.2e inl=%d/%d" % (np.abs(ta - tb).max(), np.abs(qa - qb).max(), ia.sum(), ib.sum()))

    # end to end
    vg, vo = lib.create(p, cfg["sensor"]), orc.create(p, cfg["sensor"])
    tg = to = 0.0
    for t in range(nf):
        a, b = st.frame(t)
        t0 = time.time()
        Rg, Tg = vg.track(a, b)
        tg += time.time() - t0
        t0 = time.time()
        Ro, To = vo.track(a, b)
        to += time.time() - t0
        fa, fb = vg.frame_info(), vo.frame_info()
        fx, fd = vg.features(0)
        ox, od = vo.features(0)
        mg, mo = vg.points(0), vo.points(0)
        same_map = len(mg["xyz"]) == len(mo["xyz"]) and np.array_equal(mg["desc"], mo["desc"]) and \
            np.array_equal(mg["counter"], mo["counter"]) and np.array_equal(mg["age"], mo["age"])
        mapdiff = np.abs(mg["xyz"] - mo["xyz"]).max() if len(mg["xyz"]) == len(mo["xyz"]) and len(mo["xyz"]) else -1
        report("track frame %d" % t, fa == fb and np.array_equal(fx, ox) and np.array_equal(fd, od) and same_map,
               "dt=%.2e map=%d/%d (xyz %.1e) trk=%d/%d inl=%d/%d tri=%d/%d" % (
                   np.abs(Tg - To).max(), fa["map_points_after"], fb["map_points_after"], mapdiff, fa["tracked"],
                   fb["tracked"], fa["inliers"], fb["inliers"], fa["new_points"], fb["new_points"]))
        if fa != fb:
            print("   gpu", fa)
            print("   cpu", fb)
    print("ms/frame gpu %.3f  oracle %.3f" % (1e3 * tg / nf, 1e3 * to / nf))


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
