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
rng = np.random.default_rng(11)
for n, noise, n_out in ((300, 0.3, 30), (2000, 0.5, 100), (12, 0.1, 0), (300, 0.0, 0)):
    pts = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(6, 30, n)], 1)
    t_true = np.array([0.4, -0.05, 0.1]); ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    pc = (pts - t_true) @ R
    uv = np.stack([p.fx * pc[:, 0] / pc[:, 2] + p.cx, p.fy * pc[:, 1] / pc[:, 2] + p.cy], 1)
    uv += rng.normal(0, noise, uv.shape) if noise else 0
    uv[:n_out] += rng.uniform(20, 60, (n_out, 2))
    uv = uv.astype(np.float32)
    qa, ta, ia = g.solve_pose(pts, uv, [1, 0, 0, 0], [0, 0, 0])
    qb, tb, ib = o.solve_pose(pts, uv, [1, 0, 0, 0], [0, 0, 0])
    print(n, noise, n_out, "dt %.3e dq %.3e inl %d/%d diffmarks %d" % (np.abs(ta - tb).max(), np.abs(qa - qb).max(), ia.sum(), ib.sum(), (ia != ib).sum()), "err true", np.linalg.norm(ta - t_true), np.linalg.norm(tb - t_true))
