"""Device-side timeline of the blocking lvt_track call (LVT_B200_TIMELINE=1)."""
import os, sys, time
os.environ["LVT_B200_TIMELINE"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
p = configs.make_params(name)
n = 160
st = synth.StereoStream(n_frames=n, seed=0, **configs.CONFIGS[name]["stream"])
frames = [tuple(np.ascontiguousarray(a) for a in st.frame(t)) for t in range(n)]
lib = lvt_b200.load()
vo = lib.create(p, 1)
for t in range(10):
    vo.track(*frames[t])
t0 = time.perf_counter()
for t in range(10, n):
    vo.track(*frames[t])
print("blocking lvt_track: %.1f us/frame" % (1e6 * (time.perf_counter() - t0) / (n - 10)))
