import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import capi, configs
lib = lvt_b200.load()
orc = capi.Library(os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so"))
p = configs.make_params("kitti_synth")
g, o = lib.context(p), orc.context(p)
rng = np.random.default_rng(5)
q = (rng.integers(0, 8, (160, 200)) * 32).astype(np.uint8)
for nm in (False, True):
    a, b = g.agast(q, 20, nm), o.agast(q, 20, nm)
    sa = set(zip(a["x"].astype(int), a["y"].astype(int), a["response"].astype(int)))
    sb = set(zip(b["x"].astype(int), b["y"].astype(int), b["response"].astype(int)))
    print("nms", nm, len(a), len(b), "only gpu", len(sa - sb), "only cpu", len(sb - sa))
    print("  only gpu", sorted(sa - sb)[:10])
    print("  only cpu", sorted(sb - sa)[:10])
# shrink: find a small crop that still fails
raw = o.agast(q, 20, False)
for size in (24, 32, 48, 64, 96):
    bad = 0
    for y0 in range(0, 160 - size, size // 2):
        for x0 in range(0, 200 - size, size // 2):
            c = np.ascontiguousarray(q[y0:y0 + size, x0:x0 + size])
            a, b = g.agast(c, 20, True), o.agast(c, 20, True)
            if len(a) != len(b) or not np.array_equal(a["x"], b["x"]) or not np.array_equal(a["y"], b["y"]):
                bad += 1
                if bad == 1:
                    print("size", size, "crop", y0, x0, "n", len(a), len(b))
                    sa = set(zip(a["x"].astype(int), a["y"].astype(int)))
                    sb = set(zip(b["x"].astype(int), b["y"].astype(int)))
                    print("   only gpu", sorted(sa - sb), "only cpu", sorted(sb - sa))
                    r = o.agast(c, 20, False)
                    m = np.zeros((size, size), int)
                    m[r["y"].astype(int), r["x"].astype(int)] = r["response"].astype(int)
                    np.save(os.path.join(ROOT, "gpurun_out", "crop_%d.npy" % size), c)
                    ys = sorted(sb - sa) + sorted(sa - sb)
                    if ys:
                        cx, cy = ys[0]
                        print(m[max(cy - 6, 0):cy + 7, max(cx - 6, 0):cx + 7])
    print("size", size, "failing crops", bad)
