"""a short resident batch + host batch (compute-sanitizer target: memcheck / racecheck of the overlapped chain)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
lib = lvt_b200.load()
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
p = configs.make_params(name)
n = 14
st = synth.StereoStream(n_frames=n, seed=0, **configs.CONFIGS[name]["stream"])
vo = lib.create(p, 1)
vo.pool_reserve(n)
for t in range(n):
    vo.pool_upload(t, *st.frame(t))
poses, infos = vo.track_pool(0, n)
print("pool ok", infos[-1]["tracked"], infos[-1]["map_points_after"])
vo2 = lib.create(p, 1)
fr = [st.frame(t) for t in range(n)]
p2, _ = vo2.track_batch([f[0] for f in fr], [f[1] for f in fr])
print("batch ok, same poses:", bool(np.array_equal(poses, p2)))
