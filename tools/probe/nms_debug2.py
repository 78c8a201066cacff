import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import capi, configs, synth
lib = lvt_b200.load()
orc = capi.Library(os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so"))
p = configs.make_params("kitti_synth")
g, o = lib.context(p), orc.context(p)
rng = np.random.default_rng(5)
q = (rng.integers(0, 8, (160, 200)) * 32).astype(np.uint8)
L, _ = synth.StereoStream(1242, 375, 2).frame(0)
def same(a, b):
    return len(a) == len(b) and np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.array_equal(a["response"], b["response"])
for name, img, th in (("natural", L, 25), ("tie", q, 20)):
    for (h, w) in ((24, 24), (32, 32), (48, 48), (64, 64), (96, 96), (40, 100), (100, 40), (33, 65), (70, 70), (128, 128), (160, 200)):
        c = np.ascontiguousarray(img[:h, :w])
        b = o.agast(c, th, True)
        res = []
        for rep in range(4):
            a = g.agast(c, th, True)
            res.append((same(a, b), len(a)))
        araw, braw = g.agast(c, th, False), o.agast(c, th, False)
        print(name, (h, w), "cpu n", len(b), "gpu", res, "raw same", same(araw, braw), len(araw), len(braw), flush=True)
