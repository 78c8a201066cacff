"""EuRoC-shape raw stereo frames through lvt_track with lvt_set_rectification on (the genuine EuRoC
calibration): wall time per frame and the per-kernel device times incl. rectify_kernel."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import capi, configs, synth, euroc_calib as ec
lib = lvt_b200.load()
p = configs.make_params("euroc_synth", max_keypoints_per_cell=300)
n = 40
st = synth.StereoStream(n_frames=n, seed=0, **configs.CONFIGS["euroc_synth"]["stream"])
frames = [tuple(np.ascontiguousarray(a) for a in st.frame(t)) for t in range(n)]
vo = lib.create(p, 1)
K = np.array([[p.fx, 0, p.cx], [0, p.fy, p.cy], [0, 0, 1.0]])
r = capi.Rectify.make(K, [-0.03, 0.01, 0.0004, -0.0003, 0.0], np.eye(3), K)
vo.set_rectification(r, r)
for t in range(8):
    vo.track(*frames[t])
lib.reset_kernel_times(); lib.set_profiling(True)
for t in range(8, 24):
    vo.track(*frames[t])
lib.set_profiling(False)
for k, (ms, cnt) in lib.kernel_times().items():
    if cnt: print("  %-22s %8.1f us x %d" % (k, 1e3 * ms / cnt, cnt))
t0 = time.perf_counter()
for t in range(24, n):
    vo.track(*frames[t])
print("blocking lvt_track with rectification: %.1f us/frame, state %d" % (1e6 * (time.perf_counter() - t0) / (n - 24), vo.get_state()))
